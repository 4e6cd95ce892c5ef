"""Host-side input decoding next to the hot path: OpenPose JSON -> the 19 x 3 joint array the
rasteriser consumes, and the key-point side of the evaluator's resize.

`read_keypoints` restates `read_json_keypoint` (PGNR/utils/utils.py:12-60): the person with the largest bounding box over
its first 15 BODY_25 joints (confidence > 0.1, at least 4 of them) is kept; its joints [0..14, 19, 22] are followed by the
mean of the left-hand and of the right-hand key points with positive confidence (needs more than 5 of them, else a zero row).
Pinned against the reference by tests/golden/keypoints_cases.json (oracle/make_golden_keypoints.py).

`scale_keypoints` is the key-point half of `A.Resize` as the evaluator uses it (PGNR/models/evaluator.py:18-26, :218-220):
x * W / w0, y * H / h0.  albumentations is not installed in the build container, so this one line is restated from its
documented behaviour (parity unpinned).

`save_frames` is the output side: the uint8 HWC frames that `rib.composite` / `ClipRenderer` already
produce on the GPU (tensor2images semantics, PGNR/utils/utils.py:122-147) are written as PNGs named after the DAIN inputs,
exactly what `Image.fromarray(tensor2images(fuse)).save(name)` writes per frame (PGNR/models/evaluator.py:265-266), but
encoded by a thread pool (zlib releases the GIL) instead of one frame at a time between generator calls.
"""
import json
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

BODY25_SUBSET = list(range(0, 15)) + [19, 22]


def _hand_mean(pts, thres=0.0):
    pts = np.asarray(pts, dtype=np.float64).reshape(-1, 3)
    valid = pts[:, 2] > thres
    if valid.sum() > 5:
        return np.mean(pts[valid, :], axis=0, keepdims=True)
    return np.zeros((1, 3))


def _largest_person(people, thres=0.1):
    best, best_area = -1, -1
    for i, person in enumerate(people):
        j = np.array(person['pose_keypoints_2d'], dtype=np.float64).reshape((-1, 3))[:15, :]
        valid = j[:, 2] > thres
        if valid.sum() < 4:
            continue
        xs, ys = j[valid, 0], j[valid, 1]
        area = (np.amax(xs) - np.amin(xs)) * (np.amax(ys) - np.amin(ys))
        if area > best_area:
            best, best_area = i, area
    return best


def read_keypoints(src):
    """`src`: path of an OpenPose JSON file, or the decoded dict.  Returns float64 [19, 3] (x, y, confidence)."""
    if isinstance(src, (str, bytes)) or hasattr(src, '__fspath__'):
        with open(src) as f:
            src = json.load(f)
    people = src.get('people', [])
    idx = _largest_person(people) if len(people) > 0 else -1
    if idx == -1:
        return np.zeros((19, 3))
    person = people[idx]
    body = np.array(person['pose_keypoints_2d'], dtype=np.float64).reshape(-1, 3)[BODY25_SUBSET, :]
    return np.concatenate((body, _hand_mean(person['hand_left_keypoints_2d']), _hand_mean(person['hand_right_keypoints_2d'])),
                          axis=0)


def scale_keypoints(joints, src_hw, dst_hw):
    """Key points of an image resized from src_hw = (h0, w0) to dst_hw = (H, W); confidences are untouched."""
    out = np.array(joints, dtype=np.float64, copy=True)
    out[..., 0] *= dst_hw[1] / src_hw[1]
    out[..., 1] *= dst_hw[0] / src_hw[0]
    return out


def clip_joints(json_paths, src_hw, dst_hw):
    """[T, 19, 3] joint array of a clip (one OpenPose file per frame) at the model resolution: the `joints` argument of
    ClipRenderer.render (evaluator.py:213-220 for every frame)."""
    return np.stack([scale_keypoints(read_keypoints(p), src_hw, dst_hw) for p in json_paths])


def frame_names(dain_paths, frames_dir):
    """The reference's output names (evaluator.py:265): basename of the DAIN frame with a .png extension."""
    return [os.path.join(frames_dir, os.path.basename(p))[:-4] + '.png' for p in dain_paths]


def save_frames(frames_u8, paths, workers=8):
    """frames_u8: uint8 [T, H, W, 3] (numpy array or CPU torch tensor, e.g. ClipRenderer.render(...)['u8'].cpu());
    paths: T file names.  Writes lossless PNGs with PIL, `workers` frames at a time; returns the paths."""
    from PIL import Image
    arr = frames_u8.numpy() if hasattr(frames_u8, 'numpy') else np.asarray(frames_u8)
    if arr.dtype != np.uint8 or arr.ndim != 4 or arr.shape[-1] != 3:
        raise ValueError('save_frames expects uint8 [T, H, W, 3]')
    if len(paths) != arr.shape[0]:
        raise ValueError('need one path per frame')

    def one(i):
        Image.fromarray(arr[i]).save(paths[i])
        return paths[i]

    if workers <= 1:
        return [one(i) for i in range(len(paths))]
    with ThreadPoolExecutor(max_workers=workers) as ex:
        return list(ex.map(one, range(len(paths))))
