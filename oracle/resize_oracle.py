"""TEST INFRASTRUCTURE — CPU restatement of the image half of the evaluator's resize:
`A.Resize(height, width, interpolation=cv2.INTER_CUBIC)` (PGNR/models/evaluator.py:18-26, applied at :218-220) is
`cv2.resize(img, (W, H), interpolation=cv2.INTER_CUBIC)` on a uint8 HWC image.

Third-party arithmetic: OpenCV (container: opencv-python 4.13.0 built with Intel IPP 2022.2; unpinned by the reference).
Published algorithm: separable 4-tap cubic convolution with A = -0.75, source coordinate (d + 0.5) * scale - 0.5 in
float32, taps clamped to the image (replicated border), result rounded to nearest and saturated to uint8.
Two arithmetic variants exist in OpenCV:
  * 'float' (default here): coefficients and sums in floating point — what the IPP-backed build in this container
    computes.  Pinned against cv2 itself: max |diff| 1, at most 0.05 % of the pixels differ (near-ties of the rounding;
    tests/test_oracle_resize.py, fixtures from oracle/make_golden_resize.py).
  * 'fixed': OpenCV's own (non-IPP) 8-bit path — coefficients quantised to 11 fractional bits
    (`saturate_cast<short>(c * 2048)`), integer horizontal pass, `(sum + 2^21) >> 22` vertical pass.  Differs from the
    float variant by one level on about 4 % of the pixels; kept for builds without IPP (parity of this variant is
    unpinned: no such build is available here).
A GPU kernel for this stage should be held to the float variant within one level.
"""
import numpy as np

_A = -0.75
_SCALE = 2048


def _cubic_coeffs(x):
    """interpolateCubic (float32 arithmetic as in OpenCV), x in [0, 1): 4 taps for offsets -1, 0, 1, 2."""
    x = np.float32(x)
    a = np.float32(_A)
    c0 = ((a * (x + 1) - 5 * a) * (x + 1) + 8 * a) * (x + 1) - 4 * a
    c1 = ((a + 2) * x - (a + 3)) * x * x + 1
    c2 = ((a + 2) * (1 - x) - (a + 3)) * (1 - x) * (1 - x) + 1
    c3 = np.float32(1.0) - c0 - c1 - c2
    return np.array([c0, c1, c2, c3], dtype=np.float32)


def _axis_tables(src, dst, fixed):
    """Per destination index: the 4 clamped source indices and the 4 coefficients (float64, or int64 when quantised)."""
    scale = np.float64(src) / np.float64(dst)
    idx = np.zeros((dst, 4), dtype=np.int64)
    coef = np.zeros((dst, 4), dtype=np.int64 if fixed else np.float64)
    for d in range(dst):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(np.floor(f))
        c = _cubic_coeffs(np.float32(f - s))
        if fixed:
            c = np.clip(np.rint((c * np.float32(_SCALE)).astype(np.float64)), -32768, 32767)   # cvRound: half to even
        coef[d] = c
        idx[d] = np.clip(np.arange(s - 1, s + 3), 0, src - 1)
    return idx, coef


def resize_cubic_u8(img, height, width, variant='float'):
    """uint8 [h, w, c] (or [h, w]) -> uint8 [height, width, c]; see the module docstring for `variant`."""
    img = np.asarray(img)
    if img.dtype != np.uint8:
        raise ValueError('resize_cubic_u8 expects uint8')
    if variant not in ('float', 'fixed'):
        raise ValueError('variant must be "float" or "fixed"')
    squeeze = img.ndim == 2
    if squeeze:
        img = img[:, :, None]
    h, w, _ = img.shape
    if (h, w) == (height, width):
        out = img.copy()
        return out[:, :, 0] if squeeze else out
    fixed = variant == 'fixed'
    xi, xc = _axis_tables(w, width, fixed)
    yi, yc = _axis_tables(h, height, fixed)
    src = img.astype(np.int64 if fixed else np.float64)
    rows = (src[:, xi, :] * xc[None, :, :, None]).sum(axis=2)                  # [h, width, c] horizontal pass
    acc = (rows[yi, :, :] * yc[:, :, None, None]).sum(axis=1)                 # [height, width, c] vertical pass
    if fixed:
        out = np.clip((acc + (1 << 21)) >> 22, 0, 255).astype(np.uint8)
    else:
        out = np.clip(np.rint(acc), 0, 255).astype(np.uint8)
    return out[:, :, 0] if squeeze else out
