"""Chunk-planar activation layout of the CUDA kernels: a 16-bit map with C channels (C % 8 == 0) is
stored as [B][C/8][H][W][8] — one 16-byte vector per pixel per 8-channel plane (csrc/conv_gemm.cuh).
These helpers convert to and from NCHW; they are used by the tests and debug hooks only."""
import torch


def to_planar(x_nchw, dtype):
    b, c, h, w = x_nchw.shape
    if c % 8:
        raise ValueError('channel count must be a multiple of 8')
    return x_nchw.reshape(b, c // 8, 8, h, w).permute(0, 1, 3, 4, 2).contiguous().to(dtype)


def from_planar(x_planar):
    b, p, h, w, e = x_planar.shape
    return x_planar.permute(0, 1, 4, 2, 3).reshape(b, p * e, h, w)


def to_parity_planar(x_nchw, dtype):
    """[B][C/8][py][px][H/2][W/2][8]: the layout a stride-2 convolution reads (four dense parity tiles)."""
    b, c, h, w = x_nchw.shape
    if c % 8 or h % 2 or w % 2:
        raise ValueError('need C % 8 == 0 and even H, W')
    x = x_nchw.reshape(b, c // 8, 8, h // 2, 2, w // 2, 2)        # b, plane, e, y', py, x', px
    return x.permute(0, 1, 4, 6, 3, 5, 2).contiguous().to(dtype)   # b, plane, py, px, y', x', e
