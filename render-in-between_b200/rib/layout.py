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
