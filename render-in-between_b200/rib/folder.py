"""Folder-driven entry of the hot path: `Evaluator.evaluate_from_folder`
(PGNR/models/evaluator.py:165-269, called by PGNR/inference.py:35) with the per-frame host loop replaced by the
on-GPU clip scheduler.

Kept from the reference, line by line:
  * one clip per sub-folder of `pose_dir`, sorted (:169, :171); outputs go to `save_dir/<sub-folder>/` (:174-177)
  * key frames = the jpg/png files of `train_dir/<sub>` sorted, DAIN frames = those of `dain_dir/<sub>`, poses = the
    json files of `pose_dir/<sub>` (:179-181)
  * sample rate = 2 ** int(log2((num_poses - 1) / (num_keyframes - 1))), sequence length = (K - 1) * rate + 1 (:187-191)
  * every image goes through A.Resize(model_h, model_w, INTER_CUBIC) and the key points are scaled with it (:218-220);
    frames are normalised with to_tensor_norm (:223-224); the label of frame i is rasterised from its key points (:222-229)
  * frames with i % rate == 0 are the key frames, passed through; every other frame is generated from its label, its
    DAIN frame and the previous fused frame, and blended with the predicted mask (:240-262)
  * frame i is saved as `<basename of dain_list[i] without extension>.png` through tensor2images (:265-266)
What changed: images are decoded on the host (PIL, a thread pool), everything from the resize on runs on the GPU
(rib.resize_cubic_u8 -> rib.frames_from_u8 -> ClipRenderer), and the PNGs are encoded by a thread pool afterwards.
The reference resizes the key frame of interval k once per frame of the interval; here it is resized once.
"""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from . import io as rio
from . import ops
from .clip import ClipRenderer

_IMG_EXT = ('jpg', 'png')


def _sorted_files(folder, ext):
    return [os.path.join(folder, f) for f in sorted(os.listdir(folder)) if f.endswith(ext)]


def clip_layout(num_keyframes, num_poses):
    """(sample_rate, seq_len) exactly as evaluator.py:187-191 derives them from the file counts."""
    if num_keyframes < 2:
        raise ValueError('a clip needs at least two key frames')
    sample_rate = 2 ** int(np.log2((num_poses - 1) / (num_keyframes - 1)))
    return sample_rate, (num_keyframes - 1) * sample_rate + 1


def _decode(paths, workers):
    """uint8 [N, h, w, 3] from image files (np.asarray(Image.open(p)), evaluator.py:207-211)."""
    from PIL import Image

    def one(p):
        a = np.asarray(Image.open(p))
        if a.ndim != 3 or a.shape[2] != 3:
            raise ValueError('%s: expected an 8-bit RGB image' % p)
        return a

    if workers <= 1:
        arrs = [one(p) for p in paths]
    else:
        with ThreadPoolExecutor(max_workers=workers) as ex:
            arrs = list(ex.map(one, paths))
    if any(a.shape != arrs[0].shape for a in arrs):
        raise ValueError('frames of a clip must share one size')
    return np.stack(arrs)


def render_folder_clip(model, image_list, dain_list, pose_list, height, width, device, workers=8):
    """One clip: returns (uint8 [seq_len, height, width, 3] CPU tensor, the DAIN paths of those frames)."""
    rate, seq_len = clip_layout(len(image_list), len(pose_list))
    if len(dain_list) < seq_len or len(pose_list) < seq_len:
        raise ValueError('need %d DAIN frames and poses, found %d / %d' % (seq_len, len(dain_list), len(pose_list)))
    gen_idx = [i for i in range(seq_len) if i % rate]
    keys = _decode(image_list, workers)
    dain = _decode([dain_list[i] for i in gen_idx], workers)      # a key frame's DAIN image is never read (:240-244)
    h0, w0 = keys.shape[1:3]
    # key points: the reference scales them with the image they are attached to (the key frame, :218-219)
    joints = rio.clip_joints(pose_list[:seq_len], (h0, w0), (height, width))
    with torch.no_grad():
        keys_d = ops.resize_cubic_u8(torch.from_numpy(keys).to(device), height, width)
        dain_d = ops.frames_from_u8(ops.resize_cubic_u8(torch.from_numpy(dain).to(device), height, width))
        out = ClipRenderer(model, sample_rate=rate).render(keys_d, torch.from_numpy(joints).to(device),
                                                           backgrounds=dain_d, want_u8=True, want_fuse=False)
    return out['u8'].cpu(), dain_list[:seq_len]


def evaluate_from_folder(model, train_dir, dain_dir, pose_dir, save_dir, height=320, width=480, device=None,
                         workers=8, verbose=True):
    """Drop-in for `Evaluator.evaluate_from_folder(model, train_dir, dain_dir, pose_dir, save_dir)` with gt_dir=None,
    gen_vid=False (what inference.py calls).  `height`, `width`: the model size the reference takes from its config
    (model_height / model_width, configs/HSM.yaml:192-193).  Returns {sub-folder: [written png paths]}."""
    if device is None:
        device = next(model.parameters()).device
    device = torch.device(device)
    if device.type != 'cuda':
        raise RuntimeError('rib.evaluate_from_folder runs on a CUDA device; there is no CPU path')
    model.eval()
    written = {}
    subfolders = [f for f in sorted(os.listdir(pose_dir)) if os.path.isdir(os.path.join(pose_dir, f))]
    for sub in subfolders:
        if verbose:
            print('Evaluating {} .....'.format(sub))
        frames_dir = os.path.join(save_dir, sub)
        if not os.path.exists(frames_dir):
            if verbose:
                print('Creating directory: {}'.format(frames_dir))
            os.makedirs(frames_dir)
        image_list = _sorted_files(os.path.join(train_dir, sub), _IMG_EXT)
        dain_list = _sorted_files(os.path.join(dain_dir, sub), _IMG_EXT)
        pose_list = _sorted_files(os.path.join(pose_dir, sub), ('json',))
        frames, names = render_folder_clip(model, image_list, dain_list, pose_list, height, width, device, workers)
        written[sub] = rio.save_frames(frames, rio.frame_names(names, frames_dir), workers=workers)
    return written
