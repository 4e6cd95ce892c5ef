"""CPU: the bicubic-resize oracle against cv2's committed outputs and, when importable, cv2 itself."""
import os

import numpy as np
import pytest

from oracle import resize_oracle as rz


def _check(got, want, name):
    d = np.abs(got.astype(np.int32) - want.astype(np.int32))
    assert d.max() <= 1, name
    assert (d != 0).mean() <= 5e-4, (name, float((d != 0).mean()))      # near-ties of the final rounding only


def test_resize_oracle_matches_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, 'resize_cases.npz'))
    n = int(z['n'])
    assert n >= 6
    for i in range(n):
        img, want = z['in_%d' % i], z['out_%d' % i]
        _check(rz.resize_cubic_u8(img, want.shape[0], want.shape[1]), want, 'case %d' % i)
    same = z['in_4']
    assert np.array_equal(rz.resize_cubic_u8(same, same.shape[0], same.shape[1]), same)       # identity size: untouched


def test_resize_oracle_matches_cv2_when_present():
    cv2 = pytest.importorskip('cv2')
    if 'IPP' not in ''.join(l for l in cv2.getBuildInformation().splitlines() if 'Intel IPP:' in l and 'NO' not in l):
        pytest.skip('this cv2 build does not use the float (IPP) resize the oracle restates')
    rng = np.random.default_rng(2)
    for (h, w, H, W) in [(120, 160, 128, 128), (90, 70, 160, 240), (200, 300, 64, 96)]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        _check(rz.resize_cubic_u8(img, H, W), cv2.resize(img, (W, H), interpolation=cv2.INTER_CUBIC), (h, w, H, W))


def test_resize_fixed_point_variant_is_within_one_level():
    rng = np.random.default_rng(4)
    img = rng.integers(0, 256, (40, 56, 3), dtype=np.uint8)
    a, b = rz.resize_cubic_u8(img, 64, 96), rz.resize_cubic_u8(img, 64, 96, variant='fixed')
    assert np.abs(a.astype(int) - b.astype(int)).max() <= 1
    g = rz.resize_cubic_u8(img[:, :, 0], 64, 96)
    assert g.shape == (64, 96) and np.array_equal(g, a[:, :, 0])
    with pytest.raises(ValueError):
        rz.resize_cubic_u8(img.astype(np.float32), 8, 8)
