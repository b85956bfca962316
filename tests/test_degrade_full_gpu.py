"""GPU parity of b200ir_degrade_full (the whole LQ synthesis of FFHQDegradationDataset.__getitem__ in one launch, called
through the C ABI) against
 * tests/golden/degrade_full.npz: outputs of the REFERENCE's own functions run in the build container;
 * oracle/degrade_full_oracle.py on seeded cases that reach every stage and the edge sizes.
Bars: integer stages (pyblur truncation, JPEG) bit-exact -- the low-resolution image after JPEG is an 8-bit image and
must be IDENTICAL to the oracle's for 'pyblur' / no-blur crops; for cv2.filter2D kinds the float blur is a direct fp32
sum (OpenCV uses a DFT for >= 11x11 kernels), so single 8-bit codes may flip before JPEG: bounded as in the CPU test of
the oracle's own direct sum against cv2.  cv2.resize(INTER_LINEAR) runs through Intel IPP in the build container; its
arithmetic is restated (oracle.degrade_full_oracle.resize_linear, pinned bit-exactly on the CPU) and the kernel evaluates
exactly that, so against the oracle's explicit-sum blur the final tensor is required to be IDENTICAL, not close."""
import os

import numpy as np
import pytest
import torch

from oracle import degrade_full_oracle as dfo

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'degrade_full.npz')


def to_u8(x):
    return np.rint((np.asarray(x) * 0.5 + 0.5) * 255).astype(np.int32)


def run_gpu(gt, modes, kernels, sizes, noise, quality, jitter, gray, bgr2rgb=True, bsigma=None, cj=None):
    from image_restoration_b200 import degradation as D
    out, lr = D.degrade_full_batch(torch.from_numpy(gt).cuda(), modes, kernels, sizes, noise=noise, quality=quality,
                                   jitter=jitter, gray=gray, bilateral_sigma=bsigma, color_jitter_pt=cj, bgr2rgb=bgr2rgb,
                                   return_lr=True)
    torch.cuda.synchronize()
    return out.cpu().numpy(), lr.cpu().numpy()


def golden_batch(fname='degrade_full.npz'):
    g = np.load(os.path.join(os.path.dirname(GOLD), fname))
    n, kmax = len(g['seeds']), g['taps'].shape[1]
    kernels = []
    for i in range(n):
        k = int(g['ksize'][i])
        o = (kmax - k) // 2
        kern = g['taps'][i, o:o + k, o:o + k]
        kernels.append(kern if int(g['f64'][i]) else kern.astype(np.float32))
    sizes = [(int(w), int(h)) for w, h in zip(g['lr_w'], g['lr_h'])]
    cj = [[(int(g['cj_op'][i][k]), float(g['cj_f'][i][k])) for k in range(int(g['cj_n'][i]))] for i in range(n)]
    return g, kernels, sizes, cj


@pytest.mark.parametrize('fname', ['degrade_full.npz', 'degrade_full_floatgt.npz', 'degrade_full_bicubic.npz'])
def test_golden_reference_outputs(fname):
    """degrade_full_floatgt.npz: GT images off the 8-bit grid (the dataset's cv2.resize), passed as float32."""
    g, kernels, sizes, cj = golden_batch(fname)
    out, lr = run_gpu(g['gt'], [int(m) for m in g['modes']], kernels, sizes, g['noise'], [int(q) for q in g['quality']],
                      g['jitter'], [int(x) for x in g['gray']], bsigma=[float(x) for x in g['bsigma']], cj=cj)
    exact = 0
    for i in range(len(kernels)):
        diff = np.abs(to_u8(out[i]) - g['out_u8'][i].astype(np.int32))
        if any(op == 1 for op, _ in cj[i]):   # contrast: torch's fp32 mean of the gray image vs the kernel's fp64 sum
            assert diff.max() <= 1 and (diff > 0).mean() < 1e-3, (i, (diff > 0).mean(), diff.max())
        elif int(g['modes'][i]) in (1, 3, 5):  # pyblur / median / bicubic crops: every stage restated bit-exactly -> the reference's output
            assert diff.max() == 0, (i, str(g['kinds'][i]), (diff > 0).mean(), diff.max())
        else:                             # filter2D crops: OpenCV's DFT blur vs the direct sum; bilateral: fp32 sums inside
                                          # OpenCV's SIMD code (see the CPU tests)
            assert (diff > 0).mean() < 0.02 and diff.max() <= 6, (i, str(g['kinds'][i]), (diff > 0).mean(), diff.max())
        exact += int(diff.max() == 0)
    assert exact >= len(kernels) // 2, exact


@pytest.mark.parametrize('fname', ['degrade_full.npz', 'degrade_full_floatgt.npz', 'degrade_full_bicubic.npz'])
def test_golden_against_oracle_stage_by_stage(fname):
    g, kernels, sizes, cj = golden_batch(fname)
    out, lr = run_gpu(g['gt'], [int(m) for m in g['modes']], kernels, sizes, g['noise'], [int(q) for q in g['quality']],
                      g['jitter'], [int(x) for x in g['gray']], bsigma=[float(x) for x in g['bsigma']], cj=cj)
    for i in range(len(kernels)):
        lw, lh = sizes[i]
        ref, ref_lr = dfo.degrade_full(g['gt'][i], int(g['modes'][i]), kernels[i], sizes[i], g['noise'][i, :lh, :lw],
                                       int(g['quality'][i]), g['jitter'][i], int(g['gray'][i]), exact_blur=True,
                                       bilateral_sigma=float(g['bsigma'][i]), cj=cj[i])
        # the oracle's explicit-sum blur is the arithmetic the kernel runs: LR image (8-bit after JPEG) identical
        assert np.array_equal(lr[i, :lh, :lw], ref_lr), (i, str(g['kinds'][i]), np.abs(lr[i, :lh, :lw] - ref_lr).max() * 255)
        d = np.abs(to_u8(out[i]) - to_u8(ref))
        if any(op == 1 for op, _ in cj[i]):
            assert d.max() <= 1 and (d > 0).mean() < 1e-3, (i, d.max(), (d > 0).mean())
        else:
            assert np.array_equal(out[i], ref), (i, str(g['kinds'][i]), d.max())


@pytest.mark.parametrize('H,W', [(128, 384), (64, 256), (48, 80)])
def test_seeded_cases_every_stage(H, W):
    from image_restoration_b200 import degradation as D
    rng = np.random.RandomState(H * 7 + W)
    kernels = [D.BoxKernel(7), D.BoxKernel(21), D.DiskKernel(9), D.DiskKernel(21), D.LineKernel(7, 45, 'full'),
               D.LineKernel(21, 99, 'left'), D.LineKernel(15, 30, 'right'), np.asarray(D.psfDictionary[3], dtype=np.float32),
               np.asarray(D.psfDictionary[77], dtype=np.float32), None,
               D.bivariate_Gaussian(21, 2.0, 2.0, 0, True), D.bivariate_Gaussian(21, 4.0, 0.8, 0.7, False),
               D.motion_kernel(21, True), D.motion_kernel(9, False), D.average_kernel(21), D.average_kernel(5),
               np.zeros((21, 21), np.float32), np.zeros((7, 7), np.float32), D.bilateral_space_kernel(21, 150),
               D.bilateral_space_kernel(21, 250), D.bilateral_space_kernel(9, 201)]
    modes = [1] * 9 + [0] + [2] * 6 + [3, 3, 4, 4, 4]
    bsigma = [0.0] * 18 + [150.0, 250.0, 201.0]
    B = len(kernels)
    gt = rng.randint(0, 256, (B, H, W, 3)).astype(np.uint8)
    gt[1, :, : W // 2] = 255                      # flat white: box-blur sums land on integers (truncation ties)
    gt[2] = (np.indices((H, W)).sum(0) % 256)[..., None].astype(np.uint8)
    gt[3, : H // 2] = 37
    sizes, quality, jitter, gray = [], [], [], []
    for b in range(B):
        scale = rng.uniform(4, 12)
        sizes.append((max(int(W // scale), 2), max(int(H // scale), 2)))
        quality.append([0, 30, 50, 75, 95, 100, 1][b % 7])
        jitter.append(rng.uniform(-20 / 255., 20 / 255., 3).astype(np.float32) if b % 3 == 0 else np.zeros(3, np.float32))
        gray.append(1 if b % 4 == 1 else 0)
    sizes[0], sizes[5] = (W // 4, H // 4), (W // 12, max(H // 12, 2))      # extremes of the scale range
    lwm, lhm = max(s[0] for s in sizes), max(s[1] for s in sizes)
    noise = np.zeros((B, lhm, lwm, 3), np.float32)
    for b, (lw, lh) in enumerate(sizes):
        noise[b, :lh, :lw] = np.float32(rng.randn(lh, lw, 3)) * rng.uniform(0, 20) / 255.
    cj = [[] for _ in range(B)]
    cj[2] = [(3, 0.07), (0, 1.3), (2, 0.0)]                       # hue, brightness, full desaturation
    cj[4] = [(2, 1.5), (3, -0.1), (0, 0.5)]
    cj[6] = [(0, 1.5), (1, 0.5), (3, 0.1), (2, 0.7)]              # with contrast (mean of the gray image)
    cj[9] = [(3, 0.033)]
    cj[11] = [(1, 1.5), (2, 1.2), (3, -0.05), (0, 0.9)]
    out, lr = run_gpu(gt, modes, kernels, sizes, noise, quality, jitter, gray, bsigma=bsigma, cj=cj)
    for b in range(B):
        lw, lh = sizes[b]
        ref, ref_lr = dfo.degrade_full(gt[b], modes[b], kernels[b], sizes[b], noise[b, :lh, :lw], quality[b], jitter[b],
                                       gray[b], exact_blur=True, bilateral_sigma=bsigma[b], cj=cj[b])
        assert np.array_equal(lr[b, :lh, :lw], ref_lr), (b, np.abs(lr[b, :lh, :lw] - ref_lr).max() * 255)
        d = np.abs(to_u8(out[b]) - to_u8(ref))
        if any(op == 1 for op, _ in cj[b]):
            assert d.max() <= 1 and (d > 0).mean() < 1e-3, (b, d.max(), (d > 0).mean())
        else:
            assert np.array_equal(out[b], ref), (b, d.max(), (d > 0).mean())


def test_no_noise_no_jpeg_no_blur_and_channel_order():
    rng = np.random.RandomState(5)
    gt = rng.randint(0, 256, (2, 32, 64, 3)).astype(np.uint8)
    sizes = [(16, 8), (9, 5)]
    for bgr2rgb in (True, False):
        out, lr = run_gpu(gt, [0, 0], [None, None], sizes, None, None, None, None, bgr2rgb=bgr2rgb)
        for b in range(2):
            ref, _ = dfo.degrade_full(gt[b], 0, None, sizes[b], None, 0, None, 0, bgr2rgb=bgr2rgb)
            assert np.array_equal(out[b], ref), (b, np.abs(to_u8(out[b]) - to_u8(ref)).max())


def test_bad_arguments_raise():
    from image_restoration_b200 import degradation as D
    gt = torch.zeros(1, 32, 64, 3, dtype=torch.uint8)
    with pytest.raises(ValueError):
        D.degrade_full_batch(gt, [0], [None], [(8, 4)])                       # CPU tensor: no CPU path
    with pytest.raises(ValueError):
        D.degrade_full_batch(gt.cuda(), [0], [None], [(1, 1)])                # LR image below 2x2


def test_synthesize_pairs_feeds_the_network():
    """The data path of BASELINE config 5: GT crops on the device -> (lq, gt) pair -> forward of the restoration net."""
    import random
    from image_restoration_b200 import GFPGANv1OCR, degradation as D
    opt = dict(blur_kernel_size=21, kernel_list=['iso', 'aniso', 'motion', 'average', 'median', 'bilateral', 'pyblur'],
               kernel_prob=[0.08, 0.08, 0.08, 0.08, 0.08, 0.08, 0.28], blur_sigma=[0.1, 10], downsample_range=[4.0, 12.0],
               noise_range=[0, 20], jpeg_range=[30, 100], color_jitter_prob=0.3, color_jitter_shift=20,
               color_jitter_pt_prob=0.3, gray_prob=0.01)
    rng = np.random.RandomState(1)
    for dtype in (np.uint8, np.float32):
        gt = rng.randint(0, 256, (8, 128, 384, 3)).astype(np.uint8)
        gt_in = gt if dtype == np.uint8 else (gt.astype(np.float32) / 255.)
        pair, prm = D.synthesize_pairs(torch.from_numpy(gt_in).cuda(), opt, py_random=random.Random(3),
                                       np_random=np.random.RandomState(3), torch_generator=torch.Generator().manual_seed(3))
        ref_gt = (gt[..., ::-1].transpose(0, 3, 1, 2).astype(np.float32) / np.float32(255.) - np.float32(0.5)) / np.float32(0.5)
        assert np.array_equal(pair['gt'].cpu().numpy(), ref_gt)
        lq = pair['lq']
        assert lq.shape == (8, 3, 128, 384) and torch.isfinite(lq).all() and lq.min() >= -1 and lq.max() <= 1
        for b in range(8):                      # same draws, oracle on the host
            lw, lh = prm['sizes'][b]
            ref, _ = dfo.degrade_full(gt_in[b], prm['modes'][b], prm['kernels'][b], (lw, lh), prm['noise'][b, :lh, :lw],
                                      prm['quality'][b], prm['jitter'][b], prm['gray'][b], exact_blur=True,
                                      bilateral_sigma=prm['bilateral_sigma'][b], cj=prm['color_jitter_pt'][b])
            d = np.abs(to_u8(lq[b].cpu().numpy()) - to_u8(ref))
            assert d.max() <= 1 and (d > 0).mean() < 1e-3, (b, prm['desc'][b], d.max())
    kw = dict(input_width=384, input_height=128, num_style_feat=256, channel_multiplier=0.5, num_mlp=4,
              input_is_latent=True, different_w=True, narrow=1, sft_half=True)
    torch.manual_seed(0)
    net = GFPGANv1OCR(decoder_load_path=None, fix_decoder=True, **kw).eval().cuda()
    out, _ = net(lq, return_rgb=False, randomize_noise=False)
    assert out.shape == (8, 3, 128, 384) and torch.isfinite(out).all()


def test_random_mask_golden_reference_outputs():
    """`random_mask: true` (ffhq_degradation_dataset.py:153-187, :299-303): the kernel applies the shapes the host mirror drew;
    against the reference's own outputs (tests/golden/degrade_full_mask.npz) and, bit for bit, against the oracle."""
    from image_restoration_b200 import degradation as D
    g, kernels, sizes, cj = golden_batch('degrade_full_mask.npz')
    n = len(kernels)
    out = D.degrade_full_batch(torch.from_numpy(g['gt']).cuda(), [int(m) for m in g['modes']], kernels, sizes, noise=g['noise'],
                               quality=[int(q) for q in g['quality']], jitter=g['jitter'], gray=[int(x) for x in g['gray']],
                               bilateral_sigma=[float(x) for x in g['bsigma']], color_jitter_pt=cj,
                               mask_modes=[int(m) for m in g['mask_modes']], masks=g['masks'])
    torch.cuda.synchronize()
    out = out.cpu().numpy()
    exact = 0
    for i in range(n):
        lw, lh = sizes[i]
        diff = np.abs(to_u8(out[i]) - g['out_u8'][i].astype(np.int32))
        m = g['masks'][i] != 0
        assert (to_u8(out[i])[:, m] == 255).all(), i
        assert (diff > 0).mean() < 0.02 and diff.max() <= 6, (i, str(g['kinds'][i]), (diff > 0).mean(), diff.max())
        exact += int(diff.max() == 0)
        ref, _ = dfo.degrade_full(g['gt'][i], int(g['modes'][i]), kernels[i], sizes[i], g['noise'][i, :lh, :lw],
                                  int(g['quality'][i]), g['jitter'][i], int(g['gray'][i]), exact_blur=True,
                                  bilateral_sigma=float(g['bsigma'][i]), cj=cj[i], mask_mode=int(g['mask_modes'][i]),
                                  mask=g['masks'][i])
        if not any(op == 1 for op, _ in cj[i]):
            assert np.array_equal(out[i], ref), (i, int(g['mask_modes'][i]))
    assert exact >= n // 3, exact


def test_random_mask_through_synthesize_pairs():
    """sample_params draws masks when the dataset options say `random_mask: true`; rows with mask_mode 0 are untouched."""
    import random as pyrandom
    from image_restoration_b200 import degradation as D
    rng = np.random.RandomState(5)
    gt = torch.from_numpy(rng.randint(0, 256, (6, 64, 192, 3)).astype(np.uint8)).cuda()
    opt = dict(blur_kernel_size=21, kernel_list=['iso', 'pyblur'], kernel_prob=[0.5, 0.5], blur_sigma=[0.1, 10],
               downsample_range=[4.0, 12.0], noise_range=[0, 20], jpeg_range=[30, 100], random_mask=True)
    pair, prm = D.synthesize_pairs(gt, opt, py_random=pyrandom.Random(1), np_random=np.random.RandomState(1))
    assert prm['masks'].shape == (6, 64, 192) and len(prm['mask_modes']) == 6
    lq = pair['lq'].cpu().numpy()
    for b in range(6):
        m = prm['masks'][b] != 0
        assert m.any() and (lq[b][:, m] == 1.0).all()
    prm2 = dict(prm, mask_modes=[0] * 6)
    plain = D.degrade_full_batch(gt, **{k: v for k, v in prm2.items() if k != 'masks'}).cpu().numpy()
    for b in range(6):
        if prm['mask_modes'][b] == 1:
            keep = prm['masks'][b] == 0
            assert np.array_equal(lq[b][:, keep], plain[b][:, keep]), b


@pytest.mark.parametrize('H,W', [(128, 384), (64, 192), (40, 72)])
def test_bicubic_kind_uint8_gt(H, W):
    """'bicubic' crops mixed with other kinds on a uint8 GT batch: identical to the oracle (Pillow's integer arithmetic), the
    caller's GT tensor is left untouched (the round trip goes through the scratch buffer)."""
    from image_restoration_b200 import degradation as D
    rng = np.random.RandomState(H + W)
    B = 5
    gt = rng.randint(0, 256, (B, H, W, 3)).astype(np.uint8)
    modes = [5, 0, 5, 2, 5]
    kernels = [np.zeros((21, 21), np.float32), np.zeros((21, 21), np.float32), np.zeros((21, 21), np.float32),
               D.bivariate_Gaussian(21, 2.0, 2.0, 0, isotropic=True).astype(np.float64), np.zeros((21, 21), np.float32)]
    sizes = [(W // 4, H // 4), (W // 5, H // 5), (W // 7, H // 7), (W // 4, H // 4), (W // 9 + 2, H // 9 + 2)]
    gt_d = torch.from_numpy(gt).cuda()
    out, lr = D.degrade_full_batch(gt_d, modes, kernels, sizes, quality=[0, 0, 60, 0, 0], return_lr=True)
    torch.cuda.synchronize()
    assert torch.equal(gt_d.cpu(), torch.from_numpy(gt))
    out = out.cpu().numpy()
    for b in range(B):
        ref, ref_lr = dfo.degrade_full(gt[b], modes[b], kernels[b], sizes[b], quality=[0, 0, 60, 0, 0][b], exact_blur=True)
        lw, lh = sizes[b]
        assert np.array_equal(lr[b, :lh, :lw].cpu().numpy(), ref_lr), (b, modes[b])
        assert np.array_equal(out[b], ref), (b, modes[b])
