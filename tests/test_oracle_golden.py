"""CPU: the oracle restatements against the golden fixtures produced by the unmodified reference."""
import glob
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from oracle import generator_oracle as go
from oracle import raster_oracle as ro
from rib.synth import synth_image, synth_joints


def _cases(golden_dir):
    return sorted(glob.glob(os.path.join(golden_dir, 'raster_[0-9].npz')))


def test_raster_oracle_matches_reference_fixtures(golden_dir):
    files = _cases(golden_dir)
    assert len(files) >= 5
    for f in files:
        z = np.load(f)
        h, w, j = int(z['height']), int(z['width']), z['joints']
        lm, conf = [(a[0], a[1]) for a in j], [a[2] for a in j]
        sk = ro.skeleton(lm, conf, h, w)
        assert np.array_equal(sk, z['skeleton_u8']), f
        lab = ro.label(lm, conf, h, w)
        assert lab.dtype == np.float32 and np.array_equal(lab, z['label']), f   # bit-exact


def test_raster_oracle_fullsize_hash(golden_dir):
    spec = json.load(open(os.path.join(golden_dir, 'raster_fullsize_sha256.json')))
    s = [e for e in spec if e['height'] == 256][0]
    j = synth_joints(s['n_frames'], s['height'], s['width'], seed=s['seed'])
    for t in range(2):
        lab = ro.label([(a[0], a[1]) for a in j[t]], [a[2] for a in j[t]], s['height'], s['width'])
        assert hashlib.sha256(np.ascontiguousarray(lab).tobytes()).hexdigest() == s['sha256'][t]


def test_heatmap_reflection_max_is_at_border():
    # joint 3 px from the left edge: reflection folds mass back, the maximum is not at the joint
    g = ro.heatmap_one(3.2, 40.5, 1.0, 96, 128)
    assert g.max() == 1.0 and g[40, 3] < 1.0 and g[40, 0] == 1.0
    assert ro.heatmap_one(-1.0, 5.0, 1.0, 96, 128).max() == 0.0        # off-image -> all zero
    assert ro.heatmap_one(5.0, 5.0, 0.0, 96, 128).max() == 0.0         # no confidence -> all zero


def test_edge_curve_gaps_and_short_edges():
    cx, cy = ro.edge_curve(100.3, 50.2, 140.9, 61.7)
    assert len(cx) == 40 and cx[0] == 100 and cx[-1] == 140 and len(set(cx)) == 40   # one abscissa skipped
    assert ro.edge_curve(10.2, 10.1, 10.9, 10.5) is None                               # |d| < 1 -> not drawn
    cx, cy = ro.edge_curve(10.2, 80.0, 12.0, 20.0)                                     # y-major
    assert cy[0] == 20 and cy[-1] == 80


def test_generator_oracle_matches_reference_fixture(golden_dir, arch, synth_sd):
    z = np.load(os.path.join(golden_dir, 'generator_64x96.npz'))
    j = z['joints']
    b, h, w = j.shape[0], 64, 96
    label = torch.from_numpy(np.stack([ro.label([(a[0], a[1]) for a in j[t]], [a[2] for a in j[t]], h, w)
                                       for t in range(b)]))
    fake, prev = synth_image(b, h, w, seed=int(z['fake_seed'])), synth_image(b, h, w, seed=int(z['prev_seed']))
    with torch.no_grad():
        img, mask = go.generator_forward(synth_sd, arch, label, fake, prev)
        fuse = go.composite(img, mask, fake)
    # same arithmetic, possibly another CPU: allow last-bit differences of the BLAS kernels
    assert (img - torch.from_numpy(z['img_final'])).abs().max() < 2e-5
    assert (mask - torch.from_numpy(z['mask'])).abs().max() < 2e-5
    assert (fuse - torch.from_numpy(z['fuse'])).abs().max() < 2e-5


def test_to_uint8_truncates():
    x = torch.tensor([[-1.2, -1.0, 0.0, 0.999, 1.0, 3.0]]).view(1, 1, 1, 6).repeat(1, 3, 1, 1)
    u = go.to_uint8(x)
    assert u[0, 0, :, 0].tolist() == [0, 0, 127, 254, 255, 255]
