"""GPU parity of the fused degradation kernel (through the C ABI) against the CPU oracle (scipy convolve2d + cv2.resize,
i.e. the reference's own library calls) on seeded synthetic crops.
Bar (north_star: "within 1e-5"): the kernel reproduces scipy's summation tree in scipy's arithmetic type and the IPP
resize arithmetic of cv2 (restated and pinned in tests/test_degrade_full_cpu.py), so everything is required to be
IDENTICAL: the convolution sum, the truncated blurred byte (ties included) and the final LQ tensor."""
import numpy as np
import pytest
import torch

from oracle import pyblur_oracle as po

pytestmark = pytest.mark.gpu


def make_cases(rng, H, W):
    from image_restoration_b200 import degradation as dg
    kernels = [dg.BoxKernel(7), dg.BoxKernel(21), dg.DiskKernel(9), dg.DiskKernel(21), dg.LineKernel(7, 45, 'full'),
               dg.LineKernel(21, 99, 'left'), dg.LineKernel(15, 30, 'right'), np.asarray(dg.psfDictionary[3]),
               np.asarray(dg.psfDictionary[77]), None]
    sizes = []
    for _ in kernels:
        scale = rng.uniform(4, 12)
        sizes.append((int(W // scale), int(H // scale)))
    return kernels, sizes


@pytest.mark.parametrize('H,W', [(128, 384), (64, 256), (256, 256)])
def test_degrade_matches_oracle(H, W):
    from image_restoration_b200 import degradation as dg
    rng = np.random.RandomState(H + W)
    kernels, sizes = make_cases(rng, H, W)
    B = len(kernels)
    gt = rng.randint(0, 256, (B, H, W, 3)).astype(np.uint8)
    gt[1, :, : W // 2] = 255                      # flat white region: the truncation-tie case
    gt[2] = (np.indices((H, W)).sum(0) % 256)[..., None].astype(np.uint8)
    lwm, lhm = max(s[0] for s in sizes), max(s[1] for s in sizes)
    noise = np.zeros((B, lhm, lwm, 3), np.float32)
    for b, (lw, lh) in enumerate(sizes):
        noise[b, :lh, :lw] = np.float32(rng.randn(lh, lw, 3)) * rng.uniform(0, 20) / 255.
    out, blur_u8, blur_f32 = dg.degrade_batch(torch.from_numpy(gt).cuda(), kernels, sizes, torch.from_numpy(noise).cuda(),
                                              bgr2rgb=True, return_blur=True)
    torch.cuda.synchronize()
    out, blur_u8, blur_f32 = out.cpu().numpy(), blur_u8.cpu().numpy(), blur_f32.cpu().numpy()
    for b in range(B):
        lw, lh = sizes[b]
        if kernels[b] is not None:
            ref_f = po.blur_f32(gt[b], kernels[b])              # scipy.signal.convolve2d, float64 or float32
            assert np.array_equal(blur_f32[b], ref_f.astype(np.float32)), (b, np.abs(blur_f32[b] - ref_f).max())
            assert np.array_equal(blur_u8[b], ref_f.astype('uint8')), (b, int((blur_u8[b] != ref_f.astype('uint8')).sum()))
        ref = po.degrade(gt[b], kernels[b], sizes[b], noise[b, :lh, :lw], bgr2rgb=True)
        assert np.array_equal(out[b], ref), (b, np.abs(out[b] - ref).max() * 127.5)


def test_pyblur_style_operators():
    from image_restoration_b200 import degradation as dg
    rng = np.random.RandomState(5)
    img = rng.randint(0, 256, (64, 192, 3)).astype(np.uint8)
    for fn, args, k in ((dg.BoxBlur, (9,), po.box_kernel(9)), (dg.DefocusBlur, (11,), po.disk_kernel(11)),
                        (dg.LinearMotionBlur, (13, 45, 'full'), po.line_kernel(13, 45, 'full')),
                        (dg.PsfBlur, (10,), po.psf_kernel(10))):
        got = fn(img, *args)
        assert got.shape == img.shape and got.dtype == np.uint8
        assert np.array_equal(got, po.blur_u8(img, k))
    out = dg.RandomizedBlur(torch.from_numpy(img)[None].cuda(), np.random.RandomState(0))
    assert out.shape == (1, 64, 192, 3) and out.dtype == torch.uint8


def test_random_pipeline_runs_at_batch_256():
    from image_restoration_b200 import degradation as dg
    rng = np.random.RandomState(9)
    B, H, W = 256, 128, 384
    gt = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device='cuda')
    kernels, sizes, nz = dg.random_degradation_params(B, H, W, rng=rng)
    out = dg.degrade_batch(gt, kernels, sizes, torch.from_numpy(nz).cuda())
    torch.cuda.synchronize()
    assert out.shape == (B, 3, H, W) and torch.isfinite(out).all()
    assert out.min().item() >= -1 and out.max().item() <= 1
    q = (out * 0.5 + 0.5) * 255
    assert (q - q.round()).abs().max().item() < 1e-3          # on the 8-bit grid
