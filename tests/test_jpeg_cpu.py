"""Pins oracle/jpeg_oracle.py (the integer restatement of the libjpeg-turbo round trip behind
degradations.add_jpg_compression, degradations.py:876-892): bit-exact against cv2.imencode / cv2.imdecode run here, and
against the committed outputs of the same call (tests/golden/jpeg_roundtrip.npz)."""
import os

import cv2
import numpy as np
import pytest

from oracle import jpeg_oracle as jo

GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'jpeg_roundtrip.npz')


def cv2_roundtrip(img, q):
    return cv2.imdecode(cv2.imencode('.jpg', img, [int(cv2.IMWRITE_JPEG_QUALITY), int(q)])[1], 1)


def test_quant_tables_known_answers():
    ql, qc = jo.quant_tables(50)                   # quality 50 = the Annex K tables themselves
    assert np.array_equal(ql, jo.STD_LUMA) and np.array_equal(qc, jo.STD_CHROMA)
    ql, qc = jo.quant_tables(100)
    assert ql.min() == 1 and ql.max() == 1 and qc.max() == 1
    ql, _ = jo.quant_tables(1)
    assert ql.max() == 255                         # baseline clamp


def test_golden_roundtrip_bit_exact():
    g = np.load(GOLD)
    for i in range(int(g['n'])):
        got = jo.jpeg_roundtrip_u8(g[f'in{i}'], int(g[f'q{i}']))
        assert np.array_equal(got, g[f'out{i}']), (i, g[f'in{i}'].shape, int(g[f'q{i}']))


@pytest.mark.parametrize('h,w', [(10, 32), (16, 48), (23, 71), (32, 96), (1, 1), (9, 17)])
def test_roundtrip_matches_cv2(h, w):
    rng = np.random.default_rng(h * 100 + w)
    for q in (1, 7, 30, 49, 50, 51, 75, 90, 100):
        for kind in range(3):
            if kind == 0:
                img = rng.integers(0, 256, (h, w, 3)).astype(np.uint8)
            elif kind == 1:
                img = np.clip(cv2.resize(rng.random((4, 5, 3)).astype(np.float32), (w, h),
                                         interpolation=cv2.INTER_LINEAR), 0, 1)
                img = (img * 255).astype(np.uint8)
            else:
                img = np.full((h, w, 3), int(rng.integers(0, 256)), np.uint8)
            assert np.array_equal(jo.jpeg_roundtrip_u8(img, q), cv2_roundtrip(img, q)), (h, w, q, kind)


def test_add_jpg_compression_matches_float_call():
    """The reference hands cv2.imencode a float image * 255 (degradations.py:889-891)."""
    rng = np.random.default_rng(5)
    img = rng.random((19, 45, 3)).astype(np.float32) * 1.2 - 0.1
    for q in (30, 63, 99):
        ref = np.clip(img, 0, 1)
        enc = cv2.imencode('.jpg', ref * 255., [int(cv2.IMWRITE_JPEG_QUALITY), q])[1]
        ref = np.float32(cv2.imdecode(enc, 1)) / 255.
        assert np.array_equal(jo.add_jpg_compression(img, q), ref)
