"""Deterministic synthetic weights and inputs (no datasets or checkpoints are reachable offline).

Weights: every tensor of the reference state-dict (rib.arch.Arch.state_spec) is drawn from a seeded
CPU generator with torch's default conv init scale, and the spectral-norm vectors u, v are
converged by power iteration, so that the eval-mode sigma = u^T W v equals ||W||_2 as it does in a
trained checkpoint (at raw random init sigma is ~75x too small and tanh saturates).

Inputs: joints, key frames and backgrounds as described in SURVEY.md §8(d).
"""
import math

import numpy as np
import torch


def synth_state_dict(arch, seed=0, power_iters=60):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    pending = {}
    for key, shape, kind in arch.state_spec():
        if kind in ('w', 'sn_w'):
            fan_in = shape[1] * shape[2] * shape[3]
            bound = 1.0 / math.sqrt(fan_in)
            sd[key] = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif kind == 'b':
            sd[key] = (torch.rand(shape, generator=g) * 2 - 1) * 0.1
        elif kind == 'in_w':
            sd[key] = 1.0 + 0.2 * (torch.rand(shape, generator=g) * 2 - 1)
        elif kind == 'in_b':
            sd[key] = 0.2 * (torch.rand(shape, generator=g) * 2 - 1)
        elif kind in ('sn_u', 'sn_v'):
            pending[key] = shape
            sd[key] = None
    # converge u, v (torch.nn.utils.spectral_norm semantics: W_mat = W.reshape(Cout, -1))
    for key in list(sd):
        if not key.endswith('.weight_orig'):
            continue
        base = key[:-len('weight_orig')]
        wm = sd[key].reshape(sd[key].shape[0], -1).double()
        u = torch.randn(wm.shape[0], generator=g, dtype=torch.float64)
        u = u / u.norm()
        for _ in range(power_iters):
            v = wm.t() @ u
            v = v / (v.norm() + 1e-12)
            u = wm @ v
            u = u / (u.norm() + 1e-12)
        sd[base + 'weight_u'] = u.float()
        sd[base + 'weight_v'] = v.float()
    return sd


def synth_joints(n_frames, height, width, seed=0, invalid_frac=0.05):
    """[n_frames, 19, 3] float64 (x, y, conf): generic (non-integer) floats, smooth motion."""
    rng = np.random.default_rng(seed)
    base = rng.uniform(0.15, 0.85, size=(19, 2)) * np.array([width, height])
    vel = rng.uniform(-1.0, 1.0, size=(19, 2)) * min(6.0, 0.02 * min(height, width))
    phase = rng.uniform(0, 2 * np.pi, size=(19, 2))
    amp = rng.uniform(0.0, 0.03, size=(19, 2)) * np.array([width, height])
    t = np.arange(n_frames)[:, None, None]
    xy = base[None] + vel[None] * t + amp[None] * np.sin(0.37 * t + phase[None])
    conf = np.ones((n_frames, 19, 1))
    conf[rng.uniform(size=(n_frames, 19, 1)) < invalid_frac] = 0.0
    return np.concatenate([xy, conf], axis=2)


def synth_image(batch, height, width, seed=0):
    """[batch, 3, H, W] float32 in [-1, 1]: low-pass filtered noise (a stand-in for a video frame)."""
    g = torch.Generator().manual_seed(1000 + seed)
    x = torch.rand(batch, 3, height // 8 + 2, width // 8 + 2, generator=g) * 2 - 1
    x = torch.nn.functional.interpolate(x, size=(height, width), mode='bicubic', align_corners=False)
    return x.clamp_(-1, 1).contiguous()


def synth_flow(batch, height, width, seed=0, max_px=8.0):
    """[batch, 2, H, W] float32 smooth flow in pixels, |flow| <= max_px."""
    g = torch.Generator().manual_seed(2000 + seed)
    f = torch.rand(batch, 2, height // 32 + 2, width // 32 + 2, generator=g) * 2 - 1
    f = torch.nn.functional.interpolate(f, size=(height, width), mode='bicubic', align_corners=False)
    return (f.clamp_(-1, 1) * max_px).contiguous()
