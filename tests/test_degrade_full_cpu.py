"""CPU pinning of the full LQ-synthesis path (b200ir_degrade_full):
 * the host mirror of the reference's random draws (image_restoration_b200.degradation.sample_params) + the oracle
   (oracle/degrade_full_oracle.py) against the REFERENCE's own functions imported from /root/reference, same seeds ->
   identical LQ tensors;
 * the oracle against the committed reference outputs (tests/golden/degrade_full.npz) -- runs without the reference;
 * the explicit filter2D sum (what the CUDA kernel evaluates) against cv2.filter2D: OpenCV switches to a DFT for
   kernels >= 11x11, so a direct sum differs from it in the last fp32 bits; bounded here."""
import math
import os
import random

import cv2
import numpy as np
import pytest
import torch

from image_restoration_b200 import degradation as D
from oracle import degrade_full_oracle as dfo
from oracle import ref_import

GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'degrade_full.npz')


def to_u8(x):
    return np.rint((x * 0.5 + 0.5) * 255).astype(np.int32)


def golden_case(g, i):
    k, kmax = int(g['ksize'][i]), g['taps'].shape[1]
    o = (kmax - k) // 2
    lw, lh = int(g['lr_w'][i]), int(g['lr_h'][i])
    kernel = g['taps'][i, o:o + k, o:o + k]
    if not int(g['f64'][i]):
        kernel = kernel.astype(np.float32)         # psf kernels: the reference's convolve2d runs in float32
    return dict(gt=g['gt'][i], mode=int(g['modes'][i]), kernel=kernel, lr_size=(lw, lh),
                noise=g['noise'][i, :lh, :lw], quality=int(g['quality'][i]), jitter=g['jitter'][i], gray=int(g['gray'][i]),
                bsigma=float(g['bsigma'][i]),
                cj=[(int(g['cj_op'][i][k]), float(g['cj_f'][i][k])) for k in range(int(g['cj_n'][i]))])


def test_golden_covers_every_stage():
    g = np.load(GOLD)
    kinds = set(str(k) for k in g['kinds'])
    assert {'iso', 'aniso', 'motion', 'average', 'median', 'bilateral', 'pyblur'} <= kinds, kinds
    assert g['gray'].any() and (g['jitter'] != 0).any() and (g['quality'] > 0).all() and (g['cj_n'] == 4).sum() >= 4


@pytest.mark.parametrize('fname', ['degrade_full.npz', 'degrade_full_floatgt.npz', 'degrade_full_bicubic.npz'])
def test_oracle_reproduces_golden_reference_outputs(fname):
    g = np.load(os.path.join(os.path.dirname(GOLD), fname))
    for i in range(len(g['seeds'])):
        c = golden_case(g, i)
        lib, _ = dfo.degrade_full(c['gt'], c['mode'], c['kernel'], c['lr_size'], c['noise'], c['quality'], c['jitter'],
                                  c['gray'], exact_blur=False, bilateral_sigma=c['bsigma'], cj=c['cj'])
        assert np.array_equal(to_u8(lib), g['out_u8'][i].astype(np.int32)), (i, str(g['kinds'][i]))
        # explicit-sum blur: the same image up to the last-bit difference of OpenCV's DFT path, i.e. rare single codes
        exact, _ = dfo.degrade_full(c['gt'], c['mode'], c['kernel'], c['lr_size'], c['noise'], c['quality'], c['jitter'],
                                    c['gray'], exact_blur=True, bilateral_sigma=c['bsigma'], cj=c['cj'])
        diff = np.abs(to_u8(exact) - g['out_u8'][i].astype(np.int32))
        if c['mode'] in (1, 3, 5):  # pyblur: the explicit summation tree is scipy's; median / bicubic: integer -> identical
            assert diff.max() == 0, (i, diff.max())
        else:
            assert (diff > 0).mean() < 0.02 and diff.max() <= 6, (i, str(g['kinds'][i]), (diff > 0).mean(), diff.max())


def test_explicit_convolve2d_is_bit_exact_against_scipy():
    """Pins oracle.degrade_full_oracle.convolve2d_same_fill (the summation tree the CUDA kernel reproduces) against
    scipy.signal.convolve2d for every pyblur kernel family, in float64 (box / disk / line) and float32 (psf)."""
    from scipy.signal import convolve2d
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (37, 53, 3)).astype(np.uint8)
    img[:20, :30] = 200                         # flat patch: sums land exactly on integers -> truncation ties
    kernels = [D.BoxKernel(d) for d in (7, 9, 15, 21)] + [D.DiskKernel(d) for d in (7, 13, 21)]
    kernels += [D.LineKernel(11, 63, 'right'), D.LineKernel(21, 99, 'left'), D.LineKernel(7, 45, 'full')]
    kernels += [np.asarray(D.psfDictionary[i], dtype=np.float32) for i in (3, 50, 87)]
    kernels += [D.BoxKernel(15).astype(np.float32)]
    for k in kernels:
        k = np.asarray(k)
        a = np.array(img, dtype='float32')
        ref = np.stack([convolve2d(a[:, :, c], k, mode='same', fillvalue=255.0) for c in range(3)], axis=2)
        got = dfo.convolve2d_same_fill(img, k)
        assert got.dtype == ref.dtype and np.array_equal(got, ref), (k.shape, k.dtype)


def test_explicit_median_and_bilateral_against_cv2():
    """Pins oracle median_u8 (exact) and bilateral_u8 (the arithmetic the CUDA kernel runs) against cv2: the median is an
    integer operation and must be identical; the bilateral filter accumulates in fp32 inside OpenCV's SIMD code, whose
    scalar tail differs from its vector body, so a handful of bytes per image may round the other way."""
    rng = np.random.default_rng(8)
    noise = rng.integers(0, 256, (48, 80, 3)).astype(np.uint8)
    smooth = (np.clip(cv2.resize(rng.random((5, 8, 3)).astype(np.float32), (80, 48), interpolation=cv2.INTER_CUBIC), 0, 1)
              * 255).astype(np.uint8)
    for img in (noise, smooth):
        for k in (3, 7, 21):
            assert np.array_equal(dfo.median_u8(img, k), cv2.medianBlur(img, k)), k
        for sigma in (150, 187, 250):
            ref = cv2.bilateralFilter(img, 21, sigma, sigma)
            got = dfo.bilateral_u8(img, 21, sigma)
            d = np.abs(got.astype(int) - ref.astype(int))
            assert d.max() <= 1 and (d > 0).sum() <= 3, (sigma, d.max(), int((d > 0).sum()))
    k = D.bilateral_space_kernel(21, 200)
    assert k.shape == (21, 21) and k[0, 0] == 0 and k[10, 10] == 1 and k[0, 10] > 0 and np.count_nonzero(k) == 317


def test_filter2d_direct_close_to_cv2():
    rng = np.random.default_rng(3)
    img = rng.random((40, 70, 3)).astype(np.float32)
    for ks in (3, 7, 9, 11, 21):
        k = D.bivariate_Gaussian(ks, 2.0, 0.7, 0.4, isotropic=False).astype(np.float32)
        ref = cv2.filter2D(img, -1, k)
        assert np.abs(dfo.filter2d_direct(img, k) - ref).max() <= 2e-6, ks
    k = D.motion_kernel(21, True)
    assert np.abs(dfo.filter2d_direct(img, k) - cv2.filter2D(img, -1, k)).max() <= 2e-6


def test_kernel_builders_known_answers():
    k = D.bivariate_Gaussian(21, 3.0, 3.0, 0, isotropic=True)
    assert k.shape == (21, 21) and abs(k.sum() - 1) < 1e-12 and k.argmax() == 10 * 21 + 10 and np.allclose(k, k.T)
    assert np.count_nonzero(D.motion_kernel(7, True)[3]) == 7 and np.count_nonzero(D.motion_kernel(7, False)[:, 3]) == 7
    assert D.average_kernel(5).dtype == np.float32 and np.allclose(D.average_kernel(5), 0.04)
    assert D.random_mixed_kernel(['bicubic'], [1.0], 21)[0] == 5
    for kind in ('pyblur_motion', 'random_cover'):      # RandomMotion / RandomCover are undefined in the reference's pyblur
        with pytest.raises(NotImplementedError):
            D.random_mixed_kernel([kind], [1.0], 21)


@pytest.mark.skipif(not ref_import.available(), reason='/root/reference not present')
def test_kernel_builders_match_reference():
    deg, _ = ref_import.load_reference_degradations()
    for ks, sx, sy, th in [(21, 0.6, 4.0, 0.3), (7, 5.0, 1.0, -2.0), (13, 2.5, 2.5, 0.0)]:
        assert np.array_equal(D.bivariate_Gaussian(ks, sx, sy, th, isotropic=False),
                              deg.bivariate_Gaussian(ks, sx, sy, th, isotropic=False))
        assert np.array_equal(D.bivariate_Gaussian(ks, sx, sy, th, isotropic=True),
                              deg.bivariate_Gaussian(ks, sx, sy, th, isotropic=True))


@pytest.mark.skipif(not ref_import.available(), reason='/root/reference not present')
def test_generalized_and_plateau_kinds_match_reference_draws():
    """Same seeds -> random_mixed_kernel builds the kernel the reference's random_mixed_kernels applies (checked through
    the blurred image: cv2.filter2D with our kernel == the reference's output)."""
    deg, _ = ref_import.load_reference_degradations()
    rng = np.random.default_rng(4)
    img = rng.random((32, 48, 3)).astype(np.float32)
    for kind in ('generalized_iso', 'generalized_aniso', 'plateau_iso', 'plateau_aniso', 'iso', 'aniso'):
        for seed in (1, 2, 3):
            random.seed(seed)
            np.random.seed(seed)
            ref = deg.random_mixed_kernels(img=img, kernel_list=[kind], kernel_prob=[1.0], kernel_size=21,
                                           sigma_x_range=[0.1, 10], sigma_y_range=[0.1, 10],
                                           rotation_range=[-math.pi, math.pi], noise_range=None, pad_kernel=True,
                                           pad_kernel_size=21)
            mode, k, desc = D.random_mixed_kernel([kind], [1.0], 21, [0.1, 10], [0.1, 10], (-math.pi, math.pi),
                                                  pad_kernel=True, pad_kernel_size=21, py_random=random.Random(seed),
                                                  np_random=np.random.RandomState(seed))
            assert mode == 2 and np.array_equal(cv2.filter2D(img, -1, k), ref), (kind, seed, desc)


@pytest.mark.skipif(not ref_import.available(), reason='/root/reference not present')
def test_host_mirror_and_oracle_equal_reference_chain():
    """Same seeds -> the reference's __getitem__ chain and (sample_params + oracle) give the same LQ tensor."""
    import importlib.util
    spec = importlib.util.spec_from_file_location('mk', os.path.join(os.path.dirname(GOLD), 'make_golden_degrade_full.py'))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    deg, DS = ref_import.load_reference_degradations()
    rng = np.random.default_rng(21)
    H, W = 64, 192
    kinds = set()
    for i in range(16):
        gt = mk.smooth_crop(rng, H, W)
        seed = 500 + i
        ref_import.load_reference_pyblur()
        random.seed(seed)
        np.random.seed(seed)
        torch.manual_seed(seed)
        ref = mk.reference_lq(deg, DS, gt, mk.OPT)
        p = D.sample_params(1, H, W, mk.OPT, py_random=random.Random(seed), np_random=np.random.RandomState(seed),
                            torch_generator=torch.Generator().manual_seed(seed))
        kinds.add(p['desc'][0][0])
        lw, lh = p['sizes'][0]
        got, _ = dfo.degrade_full(gt, p['modes'][0], p['kernels'][0], (lw, lh), p['noise'][0, :lh, :lw], p['quality'][0],
                                  p['jitter'][0], p['gray'][0], exact_blur=False, lib_jpeg=False,
                                  bilateral_sigma=p['bilateral_sigma'][0], cj=p['color_jitter_pt'][0])
        assert np.array_equal(got, ref), (i, p['desc'][0])
    assert len(kinds) >= 5 and {'median', 'bilateral'} <= kinds, kinds


def test_explicit_resize_is_bit_exact_against_cv2():
    """Pins oracle.degrade_full_oracle.resize_linear (the arithmetic the CUDA kernels evaluate) against cv2.resize as it
    runs in this container (float images go to Intel IPP, which differs from OpenCV's own C++ formula by up to ~1e-5)."""
    rng = np.random.default_rng(2)
    for (h, w, H, W) in [(128, 384, 20, 62), (20, 62, 128, 384), (128, 384, 32, 96), (10, 32, 128, 384), (128, 384, 13, 41),
                         (13, 41, 128, 384), (64, 256, 7, 30), (7, 30, 64, 256), (256, 256, 21, 21), (2, 2, 32, 64),
                         (128, 384, 96, 32)]:
        img = rng.random((h, w, 3)).astype(np.float32)
        ref = cv2.resize(img, (W, H), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(dfo.resize_linear(img, (W, H)), ref), (h, w, H, W)


def test_random_mask_oracle_reproduces_golden_reference_outputs():
    """`random_mask: true`: tests/golden/degrade_full_mask.npz holds the REFERENCE's outputs (its own random_mask, regular /
    irregular / half kinds) and the shapes the host mirror drew from the same seeds; oracle(mirror's shapes) == reference."""
    g = np.load(os.path.join(os.path.dirname(GOLD), 'degrade_full_mask.npz'))
    assert {1, 2} <= set(int(m) for m in g['mask_modes'])
    for i in range(len(g['seeds'])):
        c = golden_case(g, i)
        lib, _ = dfo.degrade_full(c['gt'], c['mode'], c['kernel'], c['lr_size'], c['noise'], c['quality'], c['jitter'],
                                  c['gray'], exact_blur=False, bilateral_sigma=c['bsigma'], cj=c['cj'],
                                  mask_mode=int(g['mask_modes'][i]), mask=g['masks'][i])
        assert np.array_equal(to_u8(lib), g['out_u8'][i].astype(np.int32)), (i, int(g['mask_modes'][i]))
        m = g['masks'][i] != 0
        assert (g['out_u8'][i][:, m] == 255).all()                      # masked pixels are white in the reference's output


@pytest.mark.skipif(not ref_import.available(), reason='/root/reference not present')
def test_random_mask_draws_equal_reference():
    """degradation.random_mask_draw against FFHQDegradationDataset.random_mask itself, same seeds: applying the mirrored
    shapes to the same image gives the reference's image (all three kinds occur in 40 draws)."""
    _, DS = ref_import.load_reference_degradations()
    rng = np.random.default_rng(3)
    H, W = 64, 160
    seen = set()
    for seed in range(40):
        img = rng.random((H, W, 3)).astype(np.float32)
        random.seed(seed)
        np.random.seed(seed)
        ref = DS.random_mask(img.copy())
        mode, mask = D.random_mask_draw(H, W, py_random=random.Random(seed), np_random=np.random.RandomState(seed))
        got = dfo.apply_random_mask(img, mode, mask)
        assert np.array_equal(got, np.asarray(ref, dtype=np.float32)), (seed, mode)
        seen.add((mode, bool(mask.any())))
    assert {m for m, _ in seen} == {1, 2}


def test_pil_bicubic_restatement_is_bit_exact_against_pillow():
    """oracle.pil_bicubic_resize (Pillow's 8-bit ImagingResample restated: double-precision coefficient tables, 22-bit fixed
    point, horizontal pass first) against PIL itself, down- and up-scaling, odd sizes included; and the whole 'bicubic' kind
    against the reference's degradations.bicubic."""
    from PIL import Image
    rng = np.random.default_rng(2)
    for (h, w) in [(128, 384), (64, 192), (50, 70), (214, 214), (37, 91)]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        small = dfo.pil_bicubic_resize(img, h // 4, w // 4)
        assert np.array_equal(small, np.asarray(Image.fromarray(img).resize((w // 4, h // 4), Image.BICUBIC)))
        back = dfo.pil_bicubic_resize(small, h, w)
        assert np.array_equal(back, np.asarray(Image.fromarray(small).resize((w, h), Image.BICUBIC)))
    if ref_import.available():
        deg, _ = ref_import.load_reference_degradations()
        img = rng.random((64, 192, 3)).astype(np.float32)
        ref = deg.bicubic(img)
        got = np.array(dfo.pil_bicubic_roundtrip(np.array(img * 255.0, dtype=np.uint8)), dtype=np.float32) / 255.0
        assert np.array_equal(got, ref)
