"""TEST INFRASTRUCTURE — generates tests/golden/folder_case.npz by running the UNMODIFIED reference
`Evaluator.evaluate_from_folder` (PGNR/models/evaluator.py:165-269) on a small synthetic input folder, on the CPU, in the
build container (needs /root/reference; albumentations is stubbed by oracle/ref_import.py with its documented Resize
behaviour on top of the container's cv2).

The fixture holds the input folder (key frames, DAIN frames, OpenPose JSON documents) and the PNG frames the reference
wrote, so that tests/test_gpu_folder.py can rebuild the folder anywhere, run rib.evaluate_from_folder on it and compare
file names and frames.  Model size 64x96, key frames 80x120, DAIN frames 72x100 (both resized by the evaluator), 3 key
frames, 5 poses (2x interpolation), weights = rib.synth.synth_state_dict(seed=0).

    python oracle/make_golden_folder.py
"""
import contextlib
import io
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'render-in-between_b200')):
    if p not in sys.path:
        sys.path.insert(0, p)

MODEL_H, MODEL_W = 64, 96
KEY_HW, DAIN_HW = (80, 120), (72, 100)
N_KEY, RATE = 3, 2


def synth_openpose(rng, h, w, t, n_frames):
    """One OpenPose document: 25 body joints + 21 + 21 hand points in pixel coordinates of an h x w image."""
    base = rng.uniform(0.2, 0.8, size=(25, 2)) * np.array([w, h])
    drift = rng.uniform(-1.0, 1.0, size=(25, 2)) * 1.7
    body = np.concatenate([base + drift * t, np.full((25, 1), 0.9)], axis=1)
    body[7, 2] = 0.0                                               # one undetected joint
    hands = []
    for j in (4, 7):                                               # hands around the wrists
        pts = body[j, :2][None] + rng.uniform(-3, 3, size=(21, 2))
        hands.append(np.concatenate([pts, np.full((21, 1), 0.8)], axis=1))
    small = np.concatenate([rng.uniform(0.4, 0.5, size=(25, 2)) * np.array([w, h]), np.full((25, 1), 0.5)], axis=1)
    return {'version': 1.3, 'people': [
        {'pose_keypoints_2d': small.reshape(-1).tolist(), 'hand_left_keypoints_2d': [0.0] * 63, 'hand_right_keypoints_2d': [0.0] * 63},
        {'pose_keypoints_2d': body.reshape(-1).tolist(), 'hand_left_keypoints_2d': hands[0].reshape(-1).tolist(),
         'hand_right_keypoints_2d': hands[1].reshape(-1).tolist()}]}


def main():
    import torch
    from PIL import Image
    from oracle import ref_import
    from rib.arch import Arch
    from rib.config import default_gen_cfg
    from rib.synth import synth_image, synth_state_dict
    ns = ref_import.load()
    cfg = ns.get_config(ns.config_path)
    cfg.model_height, cfg.model_width = MODEL_H, MODEL_W
    cfg.load_height, cfg.load_width = MODEL_H, MODEL_W
    with contextlib.redirect_stdout(io.StringIO()):
        ev = ns.Evaluator(cfg)
    torch.manual_seed(0)
    gen = ns.Generator(cfg.gen)
    gen.load_state_dict(synth_state_dict(Arch(default_gen_cfg()), seed=0), strict=True)
    gen.eval()

    rng = np.random.default_rng(5)
    t_len = (N_KEY - 1) * RATE + 1

    def u8(x):
        return ((x * 0.5 + 0.5).clamp(0, 1) * 255).round().to(torch.uint8).permute(0, 2, 3, 1).numpy()

    keys = u8(synth_image(N_KEY, KEY_HW[0], KEY_HW[1], seed=81))
    dain = u8(synth_image(t_len, DAIN_HW[0], DAIN_HW[1], seed=82))
    docs = [json.dumps(synth_openpose(np.random.default_rng(7), KEY_HW[0], KEY_HW[1], t, t_len)) for t in range(t_len)]
    sub = 'clip_a'
    with tempfile.TemporaryDirectory() as tmp:
        dirs = {k: os.path.join(tmp, k, sub) for k in ('inputs', 'DAIN', 'Predict_motion')}
        for d in dirs.values():
            os.makedirs(d)
        key_names = ['%05d.png' % (i * RATE) for i in range(N_KEY)]
        dain_names = ['frame_%03d.png' % i for i in range(t_len)]
        pose_names = ['%05d_keypoints.json' % i for i in range(t_len)]
        for n, a in zip(key_names, keys):
            Image.fromarray(a).save(os.path.join(dirs['inputs'], n))
        for n, a in zip(dain_names, dain):
            Image.fromarray(a).save(os.path.join(dirs['DAIN'], n))
        for n, d in zip(pose_names, docs):
            open(os.path.join(dirs['Predict_motion'], n), 'w').write(d)
        save_dir = os.path.join(tmp, 'Generated_frames')
        with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            ev.evaluate_from_folder(gen, os.path.join(tmp, 'inputs'), os.path.join(tmp, 'DAIN'),
                                    os.path.join(tmp, 'Predict_motion'), save_dir, gt_dir=None, gen_vid=False)
        out_names = sorted(os.listdir(os.path.join(save_dir, sub)))
        out = np.stack([np.asarray(Image.open(os.path.join(save_dir, sub, n))) for n in out_names])
    assert out.shape == (t_len, MODEL_H, MODEL_W, 3), out.shape
    path = os.path.join(ROOT, 'tests', 'golden', 'folder_case.npz')
    np.savez_compressed(path, sub=sub, model_hw=np.array([MODEL_H, MODEL_W]), keys=keys, dain=dain,
                        key_names=np.array(key_names), dain_names=np.array(dain_names), pose_names=np.array(pose_names),
                        pose_docs=np.array(docs), out_names=np.array(out_names), out=out)
    print('wrote', path, os.path.getsize(path), 'bytes; frames', out_names)


if __name__ == '__main__':
    main()
