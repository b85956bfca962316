"""Training loop on the B200 kernels: the `train_pipeline` of basicsr/train.py for the plate configs
(training_config/train_gfpgan_v4_*_license_*.yml; --fix-decoder freezes the StyleGAN2 decoder, the YAMLs train it), one
process per GPU.

    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 tools/train_plates.py --iters 200 --batch 256 --out /tmp/exp

Per iteration and GPU: GT crops (synthetic plates here: there is no dataset offline) -> on-device pair synthesis
(degradation.synthesize_pairs: the whole FFHQDegradationDataset.__getitem__ chain in one launch) -> GFPGANTrainer.feed_data /
optimize_parameters (net_g: l_g_pix + image pyramid + l_g_gan, EMA; net_d: logistic loss) with the NCCL all-reduce of the flat
gradient buffers.  Every --val-freq iterations the EMA network restores a fixed LQ batch through the inference engine (PSNR
against its GT, calculate_psnr's 10 log10(255^2 / mse) on the uint8 images), and checkpoints are written in the reference's
format: net_g_<iter>.pth = {'params': ..., 'params_ema': ...}, net_d_<iter>.pth = {'params': ...} (base_model.py:193-220)."""
import argparse
import os
import random
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.getcwd())
from bench import H, NET_KW, W  # noqa: E402
from image_restoration_b200 import GFPGANv1OCR, degradation as dg, train  # noqa: E402
from image_restoration_b200.disc import StyleGAN2Discriminator  # noqa: E402

OPT = dict(blur_kernel_size=21, kernel_list=['iso', 'aniso', 'motion', 'average', 'median', 'bilateral', 'pyblur'],
           kernel_prob=[0.08, 0.08, 0.08, 0.08, 0.08, 0.08, 0.28], blur_sigma=[0.1, 10], downsample_range=[4.0, 12.0],
           noise_range=[0, 20], jpeg_range=[30, 100], color_jitter_prob=0.3, color_jitter_shift=20, color_jitter_pt_prob=0.3,
           gray_prob=0.01)


def synthetic_plates(B, rng, dev):
    """Plate-like GT crops: a light background, a dark border and a row of dark glyph blocks (uint8 BGR [B,H,W,3])."""
    img = np.full((B, H, W, 3), 255, np.uint8)
    for b in range(B):
        bg = rng.randint(170, 256, 3)
        img[b] = bg
        img[b, :6], img[b, -6:], img[b, :, :6], img[b, :, -6:] = 20, 20, 20, 20
        x = 20 + rng.randint(0, 10)
        while x < W - 40:
            w, h0 = rng.randint(18, 34), rng.randint(18, 30)
            img[b, h0:H - h0, x:x + w] = rng.randint(0, 60, 3)
            if rng.rand() < 0.5:
                img[b, h0 + 12:H - h0 - 12, x + 6:x + w - 6] = bg
            x += w + rng.randint(8, 16)
    return torch.from_numpy(img).to(dev)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--iters', type=int, default=50)
    ap.add_argument('--batch', type=int, default=64, help='crops per GPU')
    ap.add_argument('--val-freq', type=int, default=25)
    ap.add_argument('--out', default='')
    ap.add_argument('--fix-decoder', action='store_true', help='freeze the StyleGAN2 decoder (the training YAMLs train it)')
    args = ap.parse_args()
    world, rank, local = int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('RANK', '0')), int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    torch.manual_seed(0)                                   # identical replicas
    kw = dict(NET_KW, fix_decoder=args.fix_decoder)
    net_g = GFPGANv1OCR(**kw).to(dev).train()
    net_g_ema = GFPGANv1OCR(**kw).to(dev).eval()
    net_g_ema.load_state_dict(net_g.state_dict())
    net_d = StyleGAN2Discriminator(input_width=W, input_height=H, channel_multiplier=1).to(dev)
    trainer = train.GFPGANTrainer(net_g, net_d, net_g_ema=net_g_ema)
    rng = np.random.RandomState(1234 + rank)               # each rank its own shard of the data
    pyr = random.Random(rank)
    val_gt = synthetic_plates(8, np.random.RandomState(7), dev)
    val_pair, _ = dg.synthesize_pairs(val_gt, OPT, py_random=random.Random(7), np_random=np.random.RandomState(7))
    t0 = time.time()
    for it in range(1, args.iters + 1):
        gt_u8 = synthetic_plates(args.batch, rng, dev)
        pair, _ = dg.synthesize_pairs(gt_u8, OPT, py_random=pyr, np_random=rng)
        trainer.feed_data(pair['lq'], pair['gt'])
        log = trainer.optimize_parameters(it)
        if rank == 0 and (it % 5 == 0 or it == 1):
            rec = float(log['l_g_pix']) + sum(float(v) for k, v in log.items() if k.startswith('l_p_'))
            print(f'iter {it:5d}  l_g_pix+pyramid {rec:.4f}  l_g_gan {float(log["l_g_gan"]):.4f}  l_d {float(log["l_d"]):.4f}  '
                  f'{it * args.batch * world / (time.time() - t0):.0f} crops/s', flush=True)
        if it % args.val_freq == 0 or it == args.iters:
            with torch.no_grad():
                out = net_g_ema(val_pair['lq'], return_rgb=False, randomize_noise=False)[0]
            a = ((out.clamp(-1, 1) + 1) * 127.5).round()
            b = ((val_pair['gt'] + 1) * 127.5).round()
            lq = ((val_pair['lq'] + 1) * 127.5).round()
            psnr = 10 * torch.log10(255.0 ** 2 / ((a - b) ** 2).mean()).item()
            psnr_lq = 10 * torch.log10(255.0 ** 2 / ((lq - b) ** 2).mean()).item()
            if rank == 0:
                print(f'iter {it:5d}  validation (EMA weights, inference engine): PSNR {psnr:.2f} dB (LQ input: {psnr_lq:.2f} dB)', flush=True)
                if args.out:
                    os.makedirs(args.out, exist_ok=True)
                    torch.save({'params': net_g.state_dict(), 'params_ema': net_g_ema.state_dict()}, os.path.join(args.out, f'net_g_{it}.pth'))
                    torch.save({'params': net_d.state_dict()}, os.path.join(args.out, f'net_d_{it}.pth'))
    if world > 1:
        chk = torch.stack([trainer.opt_g.flat.double().sum(), trainer.opt_d.flat.double().sum()])
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        if rank == 0:
            print('replicas bit-identical after training:', bool(torch.equal(lo, hi)), flush=True)
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
