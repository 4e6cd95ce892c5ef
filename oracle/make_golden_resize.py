"""Generates tests/golden/resize_cases.npz with the cv2 of this container (the reference's resize IS cv2.resize through
albumentations, PGNR/models/evaluator.py:18-26): small uint8 images and their INTER_CUBIC resizes.

    python oracle/make_golden_resize.py
"""
import os

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = [(48, 64, 64, 96), (37, 53, 64, 64), (60, 80, 32, 48), (30, 20, 64, 96), (64, 96, 64, 96), (25, 25, 80, 16)]


def main():
    rng = np.random.default_rng(5)
    out = {'cv2_version': np.array(cv2.__version__), 'n': np.array(len(CASES))}
    for i, (h, w, H, W) in enumerate(CASES):
        if i % 2 == 0:
            img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        else:
            yy, xx = np.mgrid[0:h, 0:w]
            img = np.stack([127 + 120 * np.sin(xx / 5.0 + yy / 7.0), 127 + 120 * np.cos(xx / 9.0), yy * 255.0 / h], -1).astype(np.uint8)
        out['in_%d' % i] = img
        out['out_%d' % i] = cv2.resize(img, (W, H), interpolation=cv2.INTER_CUBIC)
    path = os.path.join(ROOT, 'tests', 'golden', 'resize_cases.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
