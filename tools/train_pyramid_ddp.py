"""Data-parallel image-pyramid-loss steps (the slice of GFPGANModel.optimize_parameters that only involves the trainable part of
net_g): one process per GPU, per-rank batch shard, backward.unet_forward -> L1 pyramid loss -> backward -> NCCL all-reduce of
the flat gradient buffer (grad_sync.GradAllReducer) -> fused Adam + EMA with the 1/world average folded in (optim.FlatAdam).
Checks that the replicas stay bit-identical and reports the step time (CUDA events, max over ranks).

Usage: torchrun --nproc-per-node N --master-addr 127.0.0.1 tools/train_pyramid_ddp.py [batch_per_gpu] [steps]"""
import os
import sys

import torch
import torch.distributed as dist
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import NET_KW, H, W  # noqa: E402
from image_restoration_b200 import GFPGANv1OCR  # noqa: E402
from image_restoration_b200.backward import unet_forward  # noqa: E402
from image_restoration_b200.grad_sync import GradAllReducer  # noqa: E402
from image_restoration_b200.optim import FlatAdam  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 8
rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group('nccl')
NAMES = ('conv_body_first', 'conv_body_down', 'final_conv', 'final_linear', 'conv_body_up', 'condition_scale', 'condition_shift',
         'toRGB')
torch.manual_seed(0)                                   # identical replicas
net = GFPGANv1OCR(**NET_KW)
sd = {k: v.detach().clone().cuda().requires_grad_() for k, v in net.state_dict().items() if k.split('.')[0] in NAMES}
params = list(sd.values())
opt = FlatAdam(params, lr=2e-3, betas=(0.0, 0.99), ema_params=[p.detach().clone() for p in params])
red = GradAllReducer(params) if world > 1 else None
torch.manual_seed(100 + rank)                          # different shard per rank
gt = F.interpolate(torch.rand(B, 3, 8, 24, device='cuda') * 2 - 1, size=(H, W), mode='bilinear', align_corners=False)
lq = (gt + 0.1 * torch.randn_like(gt)).clamp(-1, 1)
pyr = [gt]
for _ in range(4):
    pyr.insert(0, F.interpolate(pyr[0], scale_factor=0.5, mode='bilinear', align_corners=False))
pyr = [p.permute(0, 2, 3, 1).contiguous() for p in pyr]


def step():
    opt.zero_grad()
    _, _, rgbs = unet_forward(sd, lq, num_style_feat=NET_KW['num_style_feat'], return_rgb=True)
    loss = sum((r[..., :3].float() - g).abs().mean() for r, g in zip(rgbs, pyr))
    loss.backward()
    if red is not None:
        red.sync(average=False)
        opt.step(flat_grad=red.flat, grad_scale=1.0 / world, ema_decay=0.999)
    else:
        opt.step(ema_decay=0.999)
    return loss.detach()


losses = [step() for _ in range(3)]                    # warm-up (also part of the optimisation)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
losses += [step() for _ in range(STEPS)]
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / STEPS], device='cuda')
chk = torch.stack([opt.flat.double().sum(), opt.flat.double().abs().sum()])
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    all_chk = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(all_chk, chk)
    same = all(torch.equal(c, all_chk[0]) for c in all_chk)
    lt = torch.stack(losses)
    dist.all_reduce(lt, op=dist.ReduceOp.SUM)
    losses = list(lt / world)
else:
    same = True
if rank == 0:
    print(f'pyramid-loss training, {world} GPU(s) x {B} crops: {ms.item():.2f} ms/step = {world * B / ms.item() * 1e3:.0f} crops/s; '
          f'replicas identical: {same}; mean loss {losses[0].item():.4f} -> {losses[-1].item():.4f}')
    assert same and losses[-1] < losses[0]
if world > 1:
    dist.destroy_process_group()
