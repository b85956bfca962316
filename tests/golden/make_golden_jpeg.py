"""Generates tests/golden/jpeg_roundtrip.npz: outputs of cv2.imencode / cv2.imdecode (the library call the reference makes
in basicsr/data/degradations.py:876-892) as executed in the build container, for the GPU box and for pinning
oracle/jpeg_oracle.py.  Run: python tests/golden/make_golden_jpeg.py"""
import os

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CASES = [(16, 48, 30), (10, 32, 95), (32, 96, 50), (21, 67, 75), (13, 41, 10), (17, 33, 100), (8, 8, 60), (31, 95, 1)]


def image(rng, h, w, kind):
    if kind == 0:
        return rng.integers(0, 256, (h, w, 3)).astype(np.uint8)
    base = cv2.resize(rng.random((h // 6 + 2, w // 6 + 2, 3)).astype(np.float32), (w, h), interpolation=cv2.INTER_CUBIC)
    return (np.clip(base, 0, 1) * 255).astype(np.uint8)


def main():
    rng = np.random.default_rng(11)
    out = {}
    for i, (h, w, q) in enumerate(CASES):
        img = image(rng, h, w, i % 2)
        enc = cv2.imencode('.jpg', img, [int(cv2.IMWRITE_JPEG_QUALITY), q])[1]
        out[f'in{i}'], out[f'out{i}'], out[f'q{i}'] = img, cv2.imdecode(enc, 1), np.int32(q)
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'jpeg_roundtrip.npz'), n=np.int32(len(CASES)), **out)
    print('wrote', len(CASES), 'cases; cv2', cv2.__version__)


if __name__ == '__main__':
    main()
