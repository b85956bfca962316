"""CPU tests of the degradation path: blur-kernel taps of the oracle against (a) the reference pyblur imported in the
build container, (b) the known answers recorded in SURVEY.md §8c; and the product's host-side tap builders against the
oracle (taps must be bit-identical)."""
import numpy as np
import pytest

from oracle import pyblur_oracle as po
from oracle import ref_import

DIMS = [7, 9, 11, 13, 15, 17, 19, 21]


def test_known_answers():
    assert np.count_nonzero(po.disk_kernel(7)) == 32
    assert np.count_nonzero(po.disk_kernel(9)) == 60
    assert np.count_nonzero(po.disk_kernel(21)) == 332
    d7 = po.disk_kernel(7)
    assert d7[0].sum() == 0 and d7[:, 0].sum() == 0               # centre is dim/2, not (dim-1)/2
    k = po.line_kernel(7, 45, 'full')
    assert np.array_equal(np.nonzero(k), np.nonzero(np.fliplr(np.eye(7)))) and np.isclose(k.sum(), 1)
    k = po.line_kernel(7, 30, 'full')
    assert sorted(zip(*np.nonzero(k))) == [(1, 6), (2, 4), (2, 5), (3, 3), (4, 1), (4, 2), (5, 0)]
    k = po.line_kernel(21, 90, 'full')
    assert np.count_nonzero(k[:, 10]) == 21
    assert np.allclose(po.box_kernel(9), 1 / 81)


def test_product_taps_equal_oracle_taps():
    from image_restoration_b200 import degradation as dg
    for dim in DIMS:
        assert np.array_equal(dg.BoxKernel(dim), po.box_kernel(dim))
        assert np.array_equal(dg.DiskKernel(dim), po.disk_kernel(dim))
        for ai in range((dim // 2) * 4):
            ang = 180.0 * ai / ((dim // 2) * 4)
            for lt in dg.LINE_TYPES:
                assert np.array_equal(dg.LineKernel(dim, ang, lt), po.line_kernel(dim, ang, lt)), (dim, ang, lt)
    for pid in (0, 17, 99):
        assert np.array_equal(dg.psfDictionary[pid], po.psf_kernel(pid))
    assert len(dg.psfDictionary) == 100


def test_random_draw_order_matches_reference_calls():
    from image_restoration_b200 import degradation as dg
    rng = np.random.RandomState(0)
    kinds = [dg.random_blur_kernel(rng)[1][0] for _ in range(200)]
    assert set(kinds) == {'box', 'disk', 'line', 'psf'}
    ks, sizes, nz = dg.random_degradation_params(4, 128, 384, rng=np.random.RandomState(1))
    assert len(ks) == 4 and all(32 <= s[0] <= 96 and 10 <= s[1] <= 32 for s in sizes)
    assert nz.shape[0] == 4 and nz.dtype == np.float32


@pytest.mark.skipif(not ref_import.available(), reason='/root/reference not present')
def test_oracle_taps_equal_reference_pyblur():
    for dim in DIMS:
        pb = ref_import.load_reference_pyblur()
        import importlib
        box = importlib.import_module('pyblur.BoxBlur')
        defocus = importlib.import_module('pyblur.DefocusBlur')
        assert np.array_equal(box.BoxKernel(dim), po.box_kernel(dim))
        assert np.array_equal(defocus.DiskKernel(dim), po.disk_kernel(dim))
        for ai in range((dim // 2) * 4):
            ang = 180.0 * ai / ((dim // 2) * 4)
            for lt in ('full', 'right', 'left'):
                pb = ref_import.load_reference_pyblur()          # fresh dictionary (reference mutates it)
                lmb = importlib.import_module('pyblur.LinearMotionBlur')
                assert np.array_equal(lmb.LineKernel(dim, ang, lt), po.line_kernel(dim, ang, lt)), (dim, ang, lt)
    psf = importlib.import_module('pyblur.PsfBlur')
    for pid in range(100):
        assert np.array_equal(psf.psfDictionary[pid], po.psf_kernel(pid))


@pytest.mark.skipif(not ref_import.available(), reason='/root/reference not present')
def test_oracle_blur_equals_reference_blur_functions():
    from PIL import Image
    import importlib
    pb = ref_import.load_reference_pyblur()
    rng = np.random.RandomState(3)
    img = rng.randint(0, 256, (32, 96, 3)).astype(np.uint8)
    pil = Image.fromarray(img)
    assert np.array_equal(np.array(pb.BoxBlur(pil, 9)), po.blur_u8(img, po.box_kernel(9)))
    assert np.array_equal(np.array(pb.DefocusBlur(pil, 13)), po.blur_u8(img, po.disk_kernel(13)))
    assert np.array_equal(np.array(pb.LinearMotionBlur(pil, 15, 60, 'full')),
                          po.blur_u8(img, po.line_kernel(15, 60, 'full')))
    assert np.array_equal(np.array(pb.PsfBlur(pil, 42)), po.blur_u8(img, po.psf_kernel(42)))
