"""Generates the committed golden fixtures under tests/golden/ by running the UNMODIFIED reference
(/root/reference, imported through oracle/ref_import.py) in the build container.

    python oracle/make_golden.py

The reference itself ships no golden vectors; these files are what pins the oracle
and the CUDA path on the GPU box, where /root/reference does not exist.
"""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'render-in-between_b200'))

from oracle import ref_import  # noqa: E402
from rib.arch import Arch  # noqa: E402
from rib.config import default_gen_cfg  # noqa: E402
from rib.synth import synth_image, synth_joints, synth_state_dict  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')


def ref_label(ds, joints, h, w):
    lm = [(float(j[0]), float(j[1])) for j in joints]
    conf = [float(j[2]) for j in joints]
    sk = ds._generate_skeleton(lm, conf, h, w)
    pm = ds._generate_pose_map(lm, conf, h, w)
    t_sk = ds.to_tensor_norm(sk).unsqueeze(0)
    t_pm = torch.from_numpy(pm).float().unsqueeze(0)
    return torch.cat([t_sk, t_pm], dim=1)[0].numpy(), sk


def raster_cases():
    """Seeded joint sets: generic poses, joints hugging corners/borders (reflection + clipping),
    off-image and zero-confidence joints, a tiny image (multi-wrap reflection)."""
    cases = []
    for seed, (h, w) in enumerate([(96, 128), (96, 128), (128, 96), (64, 64), (48, 32)]):
        rng = np.random.default_rng(100 + seed)
        j = np.concatenate([rng.uniform(-6, w + 6, (19, 1)), rng.uniform(-6, h + 6, (19, 1)),
                            (rng.uniform(size=(19, 1)) > 0.1).astype(np.float64)], axis=1)
        if seed % 2 == 1:
            j[0, :2] = rng.uniform(0, 5, 2)
            j[1, :2] = (w - rng.uniform(0.01, 5), h - rng.uniform(0.01, 5))
            j[8, :2] = (rng.uniform(0, 3), h / 2 + rng.uniform())
            j[0, 2] = j[1, 2] = j[8, 2] = 1.0
        cases.append((h, w, j))
    return cases


def main():
    os.makedirs(OUT, exist_ok=True)
    np.random.seed(0)
    # ---- rasteriser: small cases stored in full ----
    for i, (h, w, j) in enumerate(raster_cases()):
        ds = ref_import.make_dataset(h, w)
        label, sk = ref_label(ds, j, h, w)
        np.savez_compressed(os.path.join(OUT, 'raster_%d.npz' % i), joints=j, height=h, width=w, label=label,
                            skeleton_u8=sk)
    # ---- rasteriser: full-size cases stored as hashes ----
    big = []
    for seed, (h, w) in enumerate([(512, 512), (320, 480), (256, 256)]):
        j = synth_joints(3, h, w, seed=seed)
        ds = ref_import.make_dataset(h, w)
        hashes = []
        for t in range(j.shape[0]):
            label, _ = ref_label(ds, j[t], h, w)
            hashes.append(hashlib.sha256(np.ascontiguousarray(label).tobytes()).hexdigest())
        big.append(dict(height=h, width=w, seed=seed, n_frames=3, sha256=hashes))
    import json
    with open(os.path.join(OUT, 'raster_fullsize_sha256.json'), 'w') as f:
        json.dump(big, f, indent=1)
    # ---- generator: reference outputs for the synthetic weights ----
    arch = Arch(default_gen_cfg())
    sd = synth_state_dict(arch, seed=0)
    G = ref_import.make_generator()
    G.load_state_dict(sd, strict=True)
    G.eval()
    h, w, b = 64, 96, 2
    ds = ref_import.make_dataset(h, w)
    joints = synth_joints(b, h, w, seed=7)
    label = torch.from_numpy(np.stack([ref_label(ds, joints[t], h, w)[0] for t in range(b)]))
    fake, prev = synth_image(b, h, w, seed=1), synth_image(b, h, w, seed=2)
    with torch.no_grad():
        img, mask = G(label, None, fake, prev)
        m3 = mask.repeat(1, 3, 1, 1)
        fuse = img * m3 + fake * (1 - m3)
    np.savez_compressed(os.path.join(OUT, 'generator_64x96.npz'), joints=joints, img_final=img.numpy(),
                        mask=mask.numpy(), fuse=fuse.numpy(), weight_seed=0, fake_seed=1, prev_seed=2)
    keys = {k: list(v.shape) for k, v in G.state_dict().items()}
    with open(os.path.join(OUT, 'state_dict_keys.json'), 'w') as f:
        json.dump(keys, f, indent=0)
    print('golden fixtures written to', OUT)


if __name__ == '__main__':
    main()
