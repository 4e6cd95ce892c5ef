"""Host-side wrappers of the bandwidth-bound stages (rasterise, warp, composite).

Each function mirrors the reference call it replaces and calls the CUDA kernels through the C ABI
on torch's current stream.  Inputs must already be CUDA tensors; nothing here computes on the CPU.
"""
import ctypes as C

import numpy as np
import torch

from ._lib import check, lib
from .config import HSM_RASTER


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda(t, name, dtype):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == dtype and t.is_contiguous()):
        raise ValueError('%s must be a contiguous CUDA %s tensor' % (name, dtype))


def gaussian_taps(sigma=HSM_RASTER['gauss_sigma'], truncate=4.0):
    """The 41 normalised taps scipy.ndimage.gaussian_filter(sigma=5) uses (computed like scipy does)."""
    radius = int(truncate * float(sigma) + 0.5)
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    return np.ascontiguousarray(phi / phi.sum(), dtype=np.float64)


def rasterize(joints, height, width, skeleton_thres=HSM_RASTER['skeleton_thres'],
              foot_thres=HSM_RASTER['foot_thres']):
    """joints [B,19,3] float64 (x, y, conf) CUDA -> label [B,22,H,W] float32 CUDA.

    Replaces dataset._generate_skeleton + _generate_pose_map + to_tensor_norm + cat
    (PGNR/models/evaluator.py:222-229, :250); bit-exact.
    """
    _require_cuda(joints, 'joints', torch.float64)
    if joints.dim() != 3 or joints.shape[1] != 19 or joints.shape[2] != 3:
        raise ValueError('joints must be [B, 19, 3]')
    b = joints.shape[0]
    label = torch.empty(b, 22, height, width, dtype=torch.float32, device=joints.device)
    taps = gaussian_taps()
    if taps.shape[0] != 41:
        raise ValueError('rasteriser supports sigma=5 (41 taps) only')
    check(lib.rib_rasterize(joints.data_ptr(), b, height, width, taps.ctypes.data_as(C.POINTER(C.c_double)),
                            float(skeleton_thres), float(foot_thres), label.data_ptr(), _stream()), 'rib_rasterize')
    return label


def warp(src, flow):
    """Bilinear resample of src [B,C,H,W] by flow [B,2,H,W] (pixels), border padding (stage A3)."""
    _require_cuda(src, 'src', torch.float32)
    _require_cuda(flow, 'flow', torch.float32)
    b, c, h, w = src.shape
    if tuple(flow.shape) != (b, 2, h, w):
        raise ValueError('flow must be [B, 2, H, W]')
    out = torch.empty_like(src)
    check(lib.rib_warp(src.data_ptr(), flow.data_ptr(), out.data_ptr(), b, c, h, w, _stream()), 'rib_warp')
    return out


def composite(pred_img, pred_mask, dain_img, want_u8=False):
    """fuse = pred*mask + dain*(1-mask) (PGNR/models/evaluator.py:256-258); optionally also the
    uint8 HWC frame of tensor2images (PGNR/utils/utils.py:122-147)."""
    _require_cuda(pred_img, 'pred_img', torch.float32)
    _require_cuda(pred_mask, 'pred_mask', torch.float32)
    _require_cuda(dain_img, 'dain_img', torch.float32)
    b, c, h, w = pred_img.shape
    if c != 3 or tuple(pred_mask.shape) != (b, 1, h, w) or dain_img.shape != pred_img.shape:
        raise ValueError('composite: shape mismatch')
    out = torch.empty_like(pred_img)
    u8 = torch.empty(b, h, w, 3, dtype=torch.uint8, device=pred_img.device) if want_u8 else None
    check(lib.rib_composite(pred_img.data_ptr(), pred_mask.data_ptr(), dain_img.data_ptr(), out.data_ptr(),
                            u8.data_ptr() if want_u8 else None, b, h, w, _stream()), 'rib_composite')
    return (out, u8) if want_u8 else out
