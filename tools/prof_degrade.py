"""Runs the two degradation kernels a few times on the bench workload (256 crops 128x384, seeded draws) for ncu captures.
Usage: python tools/prof_degrade.py"""
import os
import random
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from image_restoration_b200 import degradation as dg  # noqa: E402

DB, H, W = 256, 128, 384
rng = np.random.RandomState(0)
gt = torch.from_numpy(rng.randint(0, 256, (DB, H, W, 3)).astype(np.uint8)).cuda()
kernels, sizes, nz = dg.random_degradation_params(DB, H, W, rng=rng)
nz_d = torch.from_numpy(nz).cuda()
packed = dg.pack_degradation(kernels, sizes, 'cuda')
opt = dict(blur_kernel_size=21, kernel_list=['iso', 'aniso', 'motion', 'average', 'median', 'bilateral', 'pyblur'],
           kernel_prob=[0.08, 0.08, 0.08, 0.08, 0.08, 0.08, 0.28], blur_sigma=[0.1, 10], downsample_range=[4.0, 12.0],
           noise_range=[0, 20], jpeg_range=[30, 100], color_jitter_prob=0.3, color_jitter_shift=20, color_jitter_pt_prob=0.3,
           gray_prob=0.01)
prm = dg.sample_params(DB, H, W, opt, py_random=random.Random(0), np_random=np.random.RandomState(0))
pk = dg.pack_degrade_full(dev='cuda', **prm)
for _ in range(3):
    dg.degrade_batch(gt, kernels, sizes, nz_d, packed=packed)
    dg.degrade_full_batch(gt, packed=pk)
torch.cuda.synchronize()
print('ok')
