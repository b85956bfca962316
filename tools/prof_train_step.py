"""Per-kernel GPU time of one GFPGANTrainer.optimize_parameters step (torch profiler, CUDA activities): which kernels the
training step spends its time in.  Usage: python tools/prof_train_step.py [batch] > gpurun_out/prof_train.txt"""
import os
import sys

import torch

sys.path.insert(0, os.getcwd())
from bench import H, NET_KW, W  # noqa: E402
from image_restoration_b200 import GFPGANv1OCR, train  # noqa: E402
from image_restoration_b200.disc import StyleGAN2Discriminator  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
torch.manual_seed(0)
net = GFPGANv1OCR(**NET_KW).cuda().train()
netd = StyleGAN2Discriminator(input_width=W, input_height=H, channel_multiplier=1).cuda()
tr = train.GFPGANTrainer(net, netd)
gt = torch.rand(B, 3, H, W, device='cuda') * 2 - 1
lq = (gt + 0.1 * torch.randn_like(gt)).clamp(-1, 1)
for it in range(2):
    tr.feed_data(lq, gt)
    tr.optimize_parameters(it + 1)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    tr.feed_data(lq, gt)
    tr.optimize_parameters(3)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=60, max_name_column_width=90))
