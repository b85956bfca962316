"""Small end-to-end case for compute-sanitizer (memcheck / racecheck), one tool per gpurun call:
  compute-sanitizer --tool memcheck python tools/sanitize_case.py
Covers the hand-rolled mbarrier / TMEM pipelines (conv_igemm / conv_row / conv_wgrad), the TMA streaming FIR kernels, the
fused degradation kernel, and the training-step kernels with their fp32 atomics, at sizes the tools finish in minutes."""
import os
import random
import sys

import numpy as np
import torch

sys.path.insert(0, os.getcwd())
from bench import NET_KW  # noqa: E402
from image_restoration_b200 import GFPGANv1OCR, degradation as dg, ops, train  # noqa: E402
from image_restoration_b200.disc import StyleGAN2Discriminator  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else 'all'
torch.manual_seed(0)
dev = torch.device('cuda')
if what in ('all', 'forward'):
    net = GFPGANv1OCR(**NET_KW).eval().to(dev)
    net.engine().use_graphs = False                      # eager launches: the tools see every kernel
    x = torch.rand(2, 3, 128, 384, device=dev) * 2 - 1
    y, rgbs = net(x, return_rgb=True, randomize_noise=False)
    torch.cuda.synchronize()
    print('forward B=2 128x384:', tuple(y.shape), float(y.abs().mean()))
    del net
if what in ('all', 'wgrad'):
    for (B, H, W, cin, cout) in ((2, 32, 96, 256, 256), (2, 64, 96, 32, 32), (1, 16, 48, 512, 64)):
        xx = torch.randn(B, H, W, cin, device=dev).half()
        dy = torch.randn(B, H, W, cout, device=dev).half()
        dw = ops.conv_wgrad(xx, dy)
        torch.cuda.synchronize()
        print('wgrad', (B, H, W, cin, cout), float(dw.abs().mean()))
if what in ('all', 'degrade'):
    B, H, W = 6, 128, 384
    rng = np.random.RandomState(0)
    gt = torch.from_numpy(rng.randint(0, 256, (B, H, W, 3)).astype(np.uint8)).to(dev)
    opt = dict(blur_kernel_size=21, kernel_list=['iso', 'aniso', 'motion', 'average', 'median', 'bilateral', 'pyblur'],
               kernel_prob=[0.1, 0.1, 0.1, 0.1, 0.2, 0.2, 0.2], blur_sigma=[0.1, 10], downsample_range=[4.0, 12.0],
               noise_range=[0, 20], jpeg_range=[30, 100], color_jitter_prob=0.5, color_jitter_shift=20,
               color_jitter_pt_prob=0.5, gray_prob=0.2)
    prm = dg.sample_params(B, H, W, opt, py_random=random.Random(0), np_random=rng)
    lq = dg.degrade_full_batch(gt, **prm)
    torch.cuda.synchronize()
    print('degrade_full', tuple(lq.shape), float(lq.abs().mean()))
if what in ('all', 'train'):
    kw = dict(NET_KW, input_width=96, input_height=32)
    net = GFPGANv1OCR(**kw).to(dev).train()
    netd = StyleGAN2Discriminator(input_width=96, input_height=32, channel_multiplier=1).to(dev)
    tr = train.GFPGANTrainer(net, netd)
    gt = torch.rand(2, 3, 32, 96, device=dev) * 2 - 1
    lq = (gt + 0.1 * torch.randn_like(gt)).clamp(-1, 1)
    for it in range(2):
        tr.feed_data(lq, gt)
        log = tr.optimize_parameters(it + 1)
    torch.cuda.synchronize()
    print('train step 96x32 B=2:', {k: round(float(v), 4) for k, v in log.items()})
print('sanitize_case: done')
