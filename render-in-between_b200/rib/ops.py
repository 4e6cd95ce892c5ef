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


_raster_ws = {}


def _raster_workspace(b, h, w, device):
    """Per-device scratch for the rasteriser (limb tables, stamp flags, heat-map windows, one word per pixel);
    grown on demand."""
    key = (device.type, device.index)
    need = int(lib.rib_rasterize_workspace_bytes(b, h, w))
    ws = _raster_ws.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.uint8, device=device)
        _raster_ws[key] = ws
    return ws


def rasterize(joints, height, width, skeleton_thres=HSM_RASTER['skeleton_thres'],
              foot_thres=HSM_RASTER['foot_thres'], planar_out=None, want_label=True):
    """joints [B,19,3] float64 (x, y, conf) CUDA -> label [B,22,H,W] float32 CUDA.

    Replaces dataset._generate_skeleton + _generate_pose_map + to_tensor_norm + cat
    (PGNR/models/evaluator.py:222-229, :250); bit-exact.

    planar_out: optional device address of a 16-bit [B][4][H][W][8] buffer (Generator.bind) that receives the
    same label in the generator's input layout; with want_label=False only that buffer is written.
    """
    _require_cuda(joints, 'joints', torch.float64)
    if joints.dim() != 3 or joints.shape[1] != 19 or joints.shape[2] != 3:
        raise ValueError('joints must be [B, 19, 3]')
    if not want_label and planar_out is None:
        raise ValueError('rasterize: nothing to write')
    b = joints.shape[0]
    label = torch.empty(b, 22, height, width, dtype=torch.float32, device=joints.device) if want_label else None
    taps = gaussian_taps()
    if taps.shape[0] != 41:
        raise ValueError('rasteriser supports sigma=5 (41 taps) only')
    ws = _raster_workspace(b, height, width, joints.device)
    with torch.cuda.device(joints.device):
        check(lib.rib_rasterize(joints.data_ptr(), b, height, width, taps.ctypes.data_as(C.POINTER(C.c_double)),
                                float(skeleton_thres), float(foot_thres), label.data_ptr() if want_label else None,
                                C.c_void_p(planar_out) if planar_out is not None else None, ws.data_ptr(), ws.numel(),
                                _stream()), 'rib_rasterize')
    return label


def _frames(t, name, tail, dtypes=(torch.float32,)):
    """A float32 CUDA tensor [B, *tail] whose frames are dense but may be strided along dim 0 (t[s::r] views)."""
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype in dtypes):
        raise ValueError('%s must be a CUDA %s tensor' % (name, ' / '.join(str(d) for d in dtypes)))
    if tuple(t.shape[1:]) != tuple(tail):
        raise ValueError('%s: shape mismatch' % name)
    if t.shape[0] > 0 and not t[0].is_contiguous():
        t = t.contiguous()
    return t, (t.stride(0) if t.shape[0] > 1 else 0)


def resize_cubic_u8(frames, height, width):
    """uint8 [B,h,w,3] CUDA frames -> uint8 [B,height,width,3]: cv2.resize(..., interpolation=cv2.INTER_CUBIC), the
    image half of the evaluator's A.Resize (PGNR/models/evaluator.py:18-26, :218-220), on the GPU."""
    if not (isinstance(frames, torch.Tensor) and frames.is_cuda and frames.dtype == torch.uint8 and frames.dim() == 4
            and frames.shape[3] == 3):
        raise ValueError('frames must be a CUDA uint8 tensor [B, h, w, 3]')
    frames = frames.contiguous()
    b, h, w, _ = frames.shape
    out = torch.empty(b, height, width, 3, dtype=torch.uint8, device=frames.device)
    with torch.cuda.device(frames.device):
        check(lib.rib_resize_cubic_u8(frames.data_ptr(), out.data_ptr(), b, h, w, height, width, 0, 0, _stream()),
              'rib_resize_cubic_u8')
    return out


def frames_from_u8(frames, out=None, out_u8=None):
    """uint8 [B,H,W,3] CUDA frames (a decoded PNG) -> float32 [B,3,H,W] in [-1,1]: dataset.to_tensor_norm
    (PGNR/datasets/HSM_auto_dataset.py:73-75) on the GPU, bit-exact.  `frames` / `out` may be strided along dim 0.
    out_u8: optional uint8 [B,H,W,3] destination (may be strided along dim 0) for tensor2images(out), the bytes the
    evaluator saves for a key frame (PGNR/models/evaluator.py:240-244, :265-266)."""
    if not (isinstance(frames, torch.Tensor) and frames.is_cuda and frames.dtype == torch.uint8 and frames.dim() == 4
            and frames.shape[3] == 3):
        raise ValueError('frames must be a CUDA uint8 tensor [B, H, W, 3]')
    b, h, w, _ = frames.shape
    if b > 0 and not frames[0].is_contiguous():
        frames = frames.contiguous()
    if out is None:
        out = torch.empty(b, 3, h, w, dtype=torch.float32, device=frames.device)
    elif not (out.is_cuda and out.dtype == torch.float32 and tuple(out.shape) == (b, 3, h, w) and out[0].is_contiguous()):
        raise ValueError('frames_from_u8: bad out tensor')
    if out_u8 is not None and not (out_u8.is_cuda and out_u8.dtype == torch.uint8 and tuple(out_u8.shape) == (b, h, w, 3)
                                   and out_u8[0].is_contiguous()):
        raise ValueError('frames_from_u8: bad out_u8 tensor')
    with torch.cuda.device(frames.device):
        check(lib.rib_frames_from_u8(frames.data_ptr(), out.data_ptr(), out_u8.data_ptr() if out_u8 is not None else None,
                                     b, h, w, frames.stride(0) if b > 1 else 0, out.stride(0) if b > 1 else 0,
                                     (out_u8.stride(0) if b > 1 else 0) if out_u8 is not None else 0, _stream()),
              'rib_frames_from_u8')
    return out


def warp(src, flow, out=None):
    """Bilinear resample of src [B,C,H,W] by flow [B,2,H,W] (pixels), border padding (stage A3).
    Both inputs may be strided along the frame dimension (e.g. flows[s::r]); `out`: optional dense destination.
    flow may be float16 (converted exactly on load): the result equals warp(src, flow.float())."""
    if src.dim() != 4:
        raise ValueError('src must be [B, C, H, W]')
    b, c, h, w = src.shape
    src, src_bs = _frames(src, 'src', (c, h, w))
    if tuple(flow.shape) != (b, 2, h, w):
        raise ValueError('flow must be [B, 2, H, W]')
    flow, flow_bs = _frames(flow, 'flow', (2, h, w), (torch.float32, torch.float16))
    if out is None:
        out = torch.empty(b, c, h, w, dtype=torch.float32, device=src.device)
    elif not (out.is_cuda and out.dtype == torch.float32 and tuple(out.shape) == (b, c, h, w) and out.is_contiguous()):
        raise ValueError('warp: bad out tensor')
    with torch.cuda.device(src.device):
        check(lib.rib_warp(src.data_ptr(), flow.data_ptr(), 1 if flow.dtype == torch.float16 else 0, out.data_ptr(), b, c, h, w,
                           src_bs, flow_bs, 0, _stream()), 'rib_warp')
    return out


def composite(pred_img, pred_mask, dain_img, want_u8=False, out=None, out_u8=None):
    """fuse = pred*mask + dain*(1-mask) (PGNR/models/evaluator.py:256-258); optionally also the
    uint8 HWC frame of tensor2images (PGNR/utils/utils.py:122-147).

    pred_mask=None: frames pass through unchanged (key frames, evaluator.py:240-244; dain_img is ignored).
    out / out_u8: optional destinations [B,3,H,W] f32 / [B,H,W,3] u8 that may be strided along the frame
    dimension (e.g. clip[s::r]), so a batch lands in its frames of the clip without a scatter copy."""
    b, c, h, w = pred_img.shape
    pred_img, img_bs = _frames(pred_img, 'pred_img', (3, h, w))
    if pred_mask is not None:
        _require_cuda(pred_mask, 'pred_mask', torch.float32)
        _require_cuda(dain_img, 'dain_img', torch.float32)
        if tuple(pred_mask.shape) != (b, 1, h, w) or dain_img.shape != pred_img.shape:
            raise ValueError('composite: shape mismatch')
    if out is None and out_u8 is None:                            # plain call: allocate the results
        out = torch.empty(b, 3, h, w, dtype=torch.float32, device=pred_img.device)
        if want_u8:
            out_u8 = torch.empty(b, h, w, 3, dtype=torch.uint8, device=pred_img.device)
    elif want_u8 and out_u8 is None:
        out_u8 = torch.empty(b, h, w, 3, dtype=torch.uint8, device=pred_img.device)
    f32_bs = u8_bs = 0
    if out is not None:
        if not (out.is_cuda and out.dtype == torch.float32 and tuple(out.shape) == (b, 3, h, w) and out[0].is_contiguous()):
            raise ValueError('composite: bad out tensor')
        f32_bs = out.stride(0) if b > 1 else 0
    if out_u8 is not None:
        if not (out_u8.is_cuda and out_u8.dtype == torch.uint8 and tuple(out_u8.shape) == (b, h, w, 3)
                and out_u8[0].is_contiguous()):
            raise ValueError('composite: bad out_u8 tensor')
        u8_bs = out_u8.stride(0) if b > 1 else 0
    with torch.cuda.device(pred_img.device):
        check(lib.rib_composite(pred_img.data_ptr(), pred_mask.data_ptr() if pred_mask is not None else None,
                                dain_img.data_ptr() if pred_mask is not None else None,
                                out.data_ptr() if out is not None else None,
                                out_u8.data_ptr() if out_u8 is not None else None, b, h, w, img_bs, f32_bs, u8_bs,
                                _stream()), 'rib_composite')
    return (out, out_u8) if (want_u8 or out_u8 is not None) else out
