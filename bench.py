#!/usr/bin/env python
"""bench.py — plate crops/s of the GFPGANv1OCR forward pass (BASELINE.json metric) on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch 64] [--impl b200|reference]

N = 1 (BASELINE configs[1]): a "step" is one forward pass over one batch of `--batch` synthetic 3x128x384 plate crops
(stock random-init weights, seed 0; randomize_noise=False).  N > 1 (BASELINE configs[2]): a "step" is one pass over
`--total` (4096) crops sharded data-parallel across the N GPUs -- contiguous per-rank shards (sharding.shard_bounds),
each rank runs its shard in micro-batches of `--batch` through its own engine, no data-path collective; total work is
fixed as N grows ("strong" scaling).
`value` = crops/s with the crops resident in HBM; `e2e` = the same through the host-buffer front end
(host_io.HostPipeline: pinned HOST input, HOST output, H2D + D2H inside the timed region).  Time = max over ranks of
the CUDA-event time; an untimed pre-roll of >= 0.5 s runs after the barrier, immediately before the first event, so
every N is timed at the sustained (power-capped) clock.
`--impl reference` times the CPU fp32 oracle port of the reference forward (oracle/) on the host cores.
The JSON line is the LAST line printed (NCCL_DEBUG output, if enabled by the caller, comes before it).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'plate_crops_per_sec'
UNIT = 'crops/s'
GFLOP_PER_CROP = 90.03          # algorithmic 2*MACs of the reference op list @3x128x384
GEMM_GFLOP_PER_CROP = 89.59     # convs 89.51 + linears 0.08 (everything conv_igemm_kernel executes)
H, W = 128, 384
NET_KW = dict(input_width=W, input_height=H, num_style_feat=256, channel_multiplier=0.5, decoder_load_path=None,
              fix_decoder=True, num_mlp=4, input_is_latent=True, different_w=True, narrow=1, sft_half=True)


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get('bf16_tflops_sustained', 1398.7), d.get('hbm_gbs', 6547.8), 'measured'
    return 1590.0, 6650.0, 'fallback'


class ClockSampler(threading.Thread):
    """Samples SM clock / power / throttle reasons of one GPU through NVML while the timed region runs
    (nvidia-smi as a fallback)."""

    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []          # (sm_mhz, sm_max_mhz, power_w, hw_slowdown, hw_thermal, sw_thermal, sw_power_cap)
        self.stop_flag = threading.Event()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            phys = int(vis.split(',')[index]) if vis and vis.split(',')[index].isdigit() else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
        pw = n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
        r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle) if hasattr(n, 'nvmlDeviceGetCurrentClocksEventReasons') \
            else n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        return (float(sm), float(mx), pw, bool(r & 0x8), bool(r & 0x40), bool(r & 0x20), bool(r & 0x4))

    def _sample_smi(self):
        out = subprocess.run(['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.Q}',
                              '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
        p = [s.strip() for s in out.strip().split(',')]
        act = [s.lower().startswith('active') for s in p[3:7]]
        return (float(p[0]), float(p[1]), float(p[2]), *act)

    def run(self):
        while not self.stop_flag.is_set():
            try:
                self.samples.append(self._sample_nvml() if self.nvml else self._sample_smi())
            except Exception:
                pass
            self.stop_flag.wait(0.02 if self.nvml else 0.1)

    def summary(self):
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unsampled']}
        sm = sorted(s[0] for s in self.samples)
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(s[3 + i] for s in self.samples)]
        return {'sm_mhz': sm[len(sm) // 2], 'sm_min_mhz': sm[0], 'sm_max_mhz': self.samples[0][1],
                'power_w_max': max(s[2] for s in self.samples), 'samples': len(self.samples),
                'source': 'nvml' if self.nvml else 'nvidia-smi', 'reasons': reasons}


def cpu_oracle_rate(seconds_budget=20.0, threads=None, crops=64, micro=8):
    """Times the CPU fp32 oracle port (reference algorithm) on the host cores over a bounded sample of the workload: up to
    one batch of 64 crops in micro-batches of 8 (all host threads), stopped early when the time budget is spent."""
    from oracle.gfpgan_ocr_oracle import OcrNetConfig, gfpgan_ocr_forward
    from image_restoration_b200 import GFPGANv1OCR
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    net = GFPGANv1OCR(**NET_KW).eval()
    sd = net.state_dict()
    cfg = OcrNetConfig(**{k: v for k, v in NET_KW.items() if k not in ('decoder_load_path', 'fix_decoder')})
    gfpgan_ocr_forward(sd, cfg, torch.rand(1, 3, H, W) * 2 - 1, False)          # warm-up (thread pool, allocator)
    x = torch.rand(micro, 3, H, W) * 2 - 1
    t0 = time.time()
    done = 0
    while done < crops and (done == 0 or time.time() - t0 < seconds_budget):
        gfpgan_ocr_forward(sd, cfg, x, False)
        done += micro
    dt = time.time() - t0
    return done / dt, threads, (f'{done} crops 3x{H}x{W} (micro-batches of {micro}) in {dt:.1f} s, fp32 torch-CPU oracle port of '
                                f'the reference forward, after 1 warm-up crop')


def conv_gflop(sd, h, w, first_key):
    """Algorithmic GFLOP (2 x MACs) per crop of a ConvLayer / ResBlock stack described by its state_dict: every 4-D weight is
    a conv over the resolution its block runs at; `*.conv2.*` / `*.skip.*` of a ResBlock produce the halved resolution."""
    total = 0.0
    res = {}
    for k, v in sd.items():
        if v.dim() != 4:
            continue
        blk = k.split('.')[1] if k.startswith('conv_body.') else None
        co, ci, kh, kw = v.shape
        if blk is not None and blk != '0':
            lvl = int(blk)
            hh, ww = h >> (lvl - 1), w >> (lvl - 1)
            if '.conv2.' in k or '.skip.' in k:
                hh, ww = hh // 2, ww // 2
        elif k.startswith('final_conv'):
            n = len({x.split('.')[1] for x in sd if x.startswith('conv_body.')}) - 1
            hh, ww = h >> n, w >> n
        else:
            hh, ww = h, w
        total += 2.0 * co * ci * kh * kw * hh * ww
    for k, v in sd.items():
        if v.dim() == 2:
            total += 2.0 * v.shape[0] * v.shape[1]
    return total / 1e9


def sr_arch_records(dev, timed, tf_peak):
    """options/test/*.yml networks at the shapes of those configs: x4 super-resolution of a batch of
    3x128x128 low-resolution patches.  images/s and the conv kernels' GEMM rate (FLOPs of the launched GEMMs; the 3-channel
    input / output layers are padded to 16 channels)."""
    from image_restoration_b200 import sr_archs
    from image_restoration_b200.ops import ConvOp
    cfgs = [('MSRResNet', dict(num_in_ch=3, num_out_ch=3, num_feat=64, num_block=16, upscale=4), 'test_MSRResNet_x4.yml'),
            ('EDSR', dict(num_in_ch=3, num_out_ch=3, num_feat=64, num_block=16, upscale=4, res_scale=1, img_range=255.,
                          rgb_mean=[0.4488, 0.4371, 0.4040]), 'test_EDSR_Mx4.yml'),
            ('RCAN', dict(num_in_ch=3, num_out_ch=3, num_feat=64, num_group=10, num_block=20, squeeze_factor=16, upscale=4,
                          res_scale=1, img_range=255., rgb_mean=[0.4488, 0.4371, 0.4040]), 'test_RCAN.yml'),
            ('RRDBNet', dict(num_in_ch=3, num_out_ch=3, num_feat=64, num_block=23, num_grow_ch=32), 'test_ESRGAN_x4.yml')]
    Bs, hw = 16, 128
    out = {}
    for name, kw, yml in cfgs:
        torch.manual_seed(0)
        net = getattr(sr_archs, name)(**kw).eval().to(dev)
        x = torch.rand(Bs, 3, hw, hw, device=dev)
        net(x)
        plan = next(iter(net.engine().plans.values()))
        gflop = sum(2.0 * st.desc.m_b * st.desc.m_h * st.desc.m_w * st.desc.cin * st.desc.num_taps * st.desc.cout
                    for st in plan.steps if isinstance(st, ConvOp)) / 1e9
        n_conv = sum(isinstance(st, ConvOp) for st in plan.steps)
        ms = timed(lambda: net(x), 10, 3, preroll=0.1) / 10
        out[name] = {'images_per_s': Bs / (ms / 1e3), 'ms_per_batch': ms, 'batch': Bs, 'input': f'3x{hw}x{hw} -> x4',
                     'config': f'options/test/*/{yml}', 'conv_launches': n_conv, 'gemm_gflop_per_image': gflop / Bs,
                     'gemm_tflops_whole_net': gflop / ms, 'frac_of_tensor_peak': gflop / ms / tf_peak}
        del net, plan
        torch.cuda.empty_cache()
    return out


def training_step_record(dev, world, B, steps, timed, fix_decoder=False, brief=False):
    """BASELINE configs[4]: on-device pair synthesis (fused degradation kernel) + GFPGANModel.optimize_parameters (net_g
    update on l_g_pix + image pyramid + l_g_gan, EMA, net_d update) at `B` crops per GPU, NCCL all-reduce of the flat
    gradient buffers when world > 1.  Returns the record for the JSON line (rank-local timings; the caller's `timed` does the
    barrier / max over ranks)."""
    import random as _random

    import numpy as np
    from image_restoration_b200 import GFPGANv1OCR, degradation as dg, ops, train
    from image_restoration_b200.disc import StyleGAN2Discriminator
    torch.manual_seed(0)
    kw = dict(NET_KW, fix_decoder=fix_decoder)       # the shipped training YAMLs: fix_decoder false (decoder trained as well)
    net = GFPGANv1OCR(**kw).to(dev).train()
    ema = GFPGANv1OCR(**kw).to(dev).eval()
    ema.load_state_dict(net.state_dict())
    netd = StyleGAN2Discriminator(input_width=W, input_height=H, channel_multiplier=1).to(dev)
    # perceptual_opt of the training YAMLs (VGG19, conv1_2 .. conv5_4 before ReLU, perceptual 1 / style 50).  The ImageNet
    # checkpoint is not available offline: seeded random weights of the same architecture (same arithmetic, same cost)
    from image_restoration_b200 import perceptual
    gv = torch.Generator().manual_seed(0)
    vgg_sd, cin, f_vgg, hh, ww = {}, 3, 0.0, H, W
    for idx, name in enumerate(perceptual.VGG19_NAMES[:35]):
        if name.startswith('conv'):
            cout = {'1': 64, '2': 128, '3': 256, '4': 512, '5': 512}[name[4]]
            vgg_sd[f'features.{idx}.weight'] = torch.randn(cout, cin, 3, 3, generator=gv) * (2.0 / (cout * 9)) ** 0.5
            vgg_sd[f'features.{idx}.bias'] = torch.randn(cout, generator=gv) * 0.05
            f_vgg += 2.0 * cout * cin * 9 * hh * ww / 1e9
            cin = cout
        elif name.startswith('pool'):
            hh, ww = hh // 2, ww // 2
    layer_weights = {'conv1_2': 0.1, 'conv2_2': 0.1, 'conv3_4': 1.0, 'conv4_4': 1.0, 'conv5_4': 1.0}
    vgg = perceptual.VGG19Features(vgg_sd, list(layer_weights), dev, use_input_norm=True, range_norm=True)
    tr = train.GFPGANTrainer(net, netd, net_g_ema=ema, perceptual=dict(vgg=vgg, layer_weights=layer_weights,
                                                                        perceptual_weight=1.0, style_weight=50.0))
    # GT crops and degradation parameters are drawn up front (the reference's DataLoader workers do this on the host, beside
    # the step); the synthesis itself (blur / resize / noise / JPEG / jitter -> lq, img2tensor -> gt) runs inside the step
    rng = np.random.RandomState(int(os.environ.get('RANK', '0')))
    gt_u8 = torch.from_numpy(rng.randint(0, 256, (B, H, W, 3)).astype(np.uint8)).to(dev)
    opt = dict(blur_kernel_size=21, kernel_list=['iso', 'aniso', 'motion', 'average', 'median', 'bilateral', 'pyblur'],
               kernel_prob=[0.08, 0.08, 0.08, 0.08, 0.08, 0.08, 0.28], blur_sigma=[0.1, 10], downsample_range=[4.0, 12.0],
               noise_range=[0, 20], jpeg_range=[30, 100], color_jitter_prob=0.3, color_jitter_shift=20,
               color_jitter_pt_prob=0.3, gray_prob=0.01)
    prm = dg.sample_params(B, H, W, opt, py_random=_random.Random(0), np_random=rng)
    pk = dg.pack_degrade_full(dev=dev, **prm)
    gt_t = torch.empty(B, 3, H, W, device=dev)
    it = [0]
    logs = []

    def step():
        lq = dg.degrade_full_batch(gt_u8, packed=pk)
        ops.u8_to_input(gt_u8, gt_t, swap_rb=True)
        tr.feed_data(lq, gt_t)
        it[0] += 1
        logs.append(tr.optimize_parameters(it[0]))
    lib = __import__('image_restoration_b200')._lib.lib()
    torch.cuda.reset_peak_memory_stats(dev)
    step()                                                # first call: packs the frozen decoder, sizes the allocator
    torch.cuda.synchronize()
    n0 = lib.b200ir_launch_count()
    ms = timed(step, steps, 1, preroll=0.0) / steps
    launches = (lib.b200ir_launch_count() - n0) // (steps + 1)
    if brief:
        mem_gb = torch.cuda.max_memory_allocated(dev) / 2 ** 30
        del net, ema, netd, vgg, tr
        torch.cuda.empty_cache()
        return {'ms': ms, 'launches_per_step': int(launches), 'max_memory_gib': mem_gb}
    tr.profile = True
    step()
    phases = tr.phase_ms()
    # an iteration on which the R1 penalty is due (every net_d_reg_every = 16 iterations, gfpgan_model.py:683-689)
    it[0] = 15
    step()
    phases_r1 = tr.phase_ms()
    tr.profile = False
    mem_gb = torch.cuda.max_memory_allocated(dev) / 2 ** 30
    f_d = conv_gflop({k: v for k, v in netd.state_dict().items()}, H, W, 'conv_body.0')
    f_dec = 34.9           # modulated convs of the StyleGAN2 decoder (24.7 plain + 10.2 up-sampling)
    f_unet = GEMM_GFLOP_PER_CROP - f_dec
    # net_g forward + U-Net dgrad and wgrad + decoder dgrad; net_d: forward + dgrad on the output (G step), dgrad + wgrad of
    # that same graph and forward + dgrad + wgrad on the real batch (D step)
    # + the VGG19 of the perceptual loss: forward on output and on gt, input gradient for the output
    # (the reference evaluates net_d on the generator's output twice with identical weights; the trainer does it once and
    # walks that graph twice: 7 net_d passes are executed and counted, not 8)
    # fix_decoder false: + the decoder's weight gradients (one more pass over its modulated convs)
    f_train = GFLOP_PER_CROP + 2 * f_unet + (1 if fix_decoder else 2) * f_dec + 7 * f_d + 3 * f_vgg
    last = {k: float(v) for k, v in logs[-1].items()}
    first = {k: float(v) for k, v in logs[0].items()}
    del net, ema, netd, vgg
    torch.cuda.empty_cache()
    return {'ms': ms, 'batch_per_gpu': B, 'fix_decoder': fix_decoder, 'launches_per_step': int(launches), 'phase_ms': phases,
            'r1_iteration': {'ms_per_step': sum(phases_r1.values()), 'd_r1_ms': phases_r1.get('d_r1'),
                             'every': tr.net_d_reg_every, 'amortised_ms_per_step': phases_r1.get('d_r1', 0.0) / tr.net_d_reg_every},
            'max_memory_gib': mem_gb, 'gflop_per_crop': f_train, 'disc_forward_gflop_per_crop': f_d,
            'vgg19_forward_gflop_per_crop': f_vgg, 'losses_first_step': first, 'losses_last_step': last}


def workload_name(world, B, total, n_local):
    """config.workload of both arms: BASELINE configs[1] on one GPU, configs[2] (sharded) on several."""
    if world == 1:
        return (f'BASELINE configs[1]: GFPGANv1OCR forward (return_rgb=False, randomize_noise=False), batch {B} '
                f'synthetic plate crops 3x{H}x{W} on 1 GPU, stock random-init weights seed 0')
    return (f'BASELINE configs[2]: {total} synthetic plate crops 3x{H}x{W} per step sharded data-parallel '
            f'across {world} GPUs ({n_local} per GPU, contiguous shards, micro-batches of {B} through '
            f'sharding.run_shard; e2e: host-resident crops through host_io.HostPipeline), GFPGANv1OCR forward '
            f'(return_rgb=False, randomize_noise=False), no collective on the path')


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    threads = os.cpu_count()
    from oracle.gfpgan_ocr_oracle import OcrNetConfig, gfpgan_ocr_forward
    from image_restoration_b200 import GFPGANv1OCR
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    net = GFPGANv1OCR(**NET_KW).eval()
    sd = net.state_dict()
    cfg = OcrNetConfig(**{k: v for k, v in NET_KW.items() if k not in ('decoder_load_path', 'fix_decoder')})
    sample_b = 1
    x = torch.rand(sample_b, 3, H, W) * 2 - 1
    for _ in range(min(args.warmup, 2)):
        gfpgan_ocr_forward(sd, cfg, x, False)
    steps = min(args.steps, 8)
    t0 = time.time()
    for _ in range(steps):
        gfpgan_ocr_forward(sd, cfg, x, False)
    dt = time.time() - t0
    value = sample_b * steps / dt
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': steps,
            'warmup': min(args.warmup, 2), 'ms_per_step': dt / steps * 1e3, 'higher_is_better': True,
            'scaling': 'weak' if args.gpus == 1 else 'strong', 'vs_baseline': None, 'dtype': 'fp32', 'data': 'synthetic',
            'config': {'workload': workload_name(args.gpus, args.batch or (64 if args.gpus == 1 else 128), args.total,
                                                 -(-args.total // args.gpus)),
                       'sample': f'{sample_b} crop of that workload per step, on the host cores of rank 0 (the reference '
                                 f'forward treats crops independently; measured crops/s at B = 1 / 4 / 8: within 10 %)',
                       'timing': 'host wall clock'},
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                             'sample': f'{steps} steps of B={sample_b} crop, fp32 oracle port of the reference forward'},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--batch', type=int, default=0, help='micro-batch per launch plan (default: 64 = BASELINE configs[1] at N = 1; '
                    '128 for the sharded configs[2] at N > 1, where the batch per GPU is 512 .. 2048: +2 %% over 64)')
    ap.add_argument('--total', type=int, default=4096, help='N > 1: crops per step over all GPUs (BASELINE configs[2])')
    ap.add_argument('--impl', default='b200')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--train-batch', type=int, default=256, help='crops per GPU of the training-step record (0 = skip)')
    ap.add_argument('--train-steps', type=int, default=3)
    ap.add_argument('--preroll', type=float, default=0.5, help='seconds of untimed steps right before the timed region')
    ap.add_argument('--no-extras', action='store_true', help='skip the secondary records (degradation, tiling, training ...)')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)

    import torch.distributed as dist
    from image_restoration_b200 import GFPGANv1OCR, _lib
    from image_restoration_b200.sharding import micro_batches, run_shard, shard_bounds
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    warmup = max(args.warmup, 3)
    B = args.batch or (64 if world == 1 else 128)
    # crops this rank processes per step: one micro-batch at N = 1, its shard of --total at N > 1
    if world > 1:
        lo, hi = shard_bounds(args.total, rank, world)
        n_local = hi - lo
    else:
        n_local = B

    torch.manual_seed(0)
    net = GFPGANv1OCR(**NET_KW).eval().to(dev)
    eng = net.engine()
    gen = torch.Generator().manual_seed(1000 + rank)
    x_host = torch.empty(n_local, 3, H, W).pin_memory()
    for s, e in micro_batches(0, n_local, 256):
        x_host[s:e] = torch.rand(e - s, 3, H, W, generator=gen) * 2 - 1
    y_host = torch.empty(n_local, 3, H, W).pin_memory()
    x_dev = x_host.to(dev)
    y_dev = torch.empty_like(x_dev)
    plan = eng.plan(B)
    lib = _lib.lib()

    # one eager pass: warms the kernels and counts this library's launches per micro-batch
    plan.x_in.copy_(x_dev[:B])
    for j, buf in enumerate(plan.noise):
        buf.copy_(eng.packed.stored_noise[j].expand_as(buf))
    n0 = lib.b200ir_launch_count()
    plan.launch(False)
    torch.cuda.synchronize()
    launches_per_mb = lib.b200ir_launch_count() - n0
    mbs_per_step = len(micro_batches(0, n_local, B))

    def fwd(xb):
        return net(xb, return_rgb=False, randomize_noise=False)[0]

    def step_resident():
        if world == 1:
            return fwd(x_dev)
        return run_shard(fwd, x_dev, B, out=y_dev)        # this rank's shard, micro-batches of B, results kept in HBM

    from image_restoration_b200.host_io import HostPipeline
    pipe = HostPipeline(net, depth=2, return_rgb=False, randomize_noise=False)

    def step_e2e():
        # host (pinned) in, host (pinned) out; the copies of neighbouring micro-batches overlap the kernels
        for s, e in micro_batches(0, n_local, B):
            pipe.submit(x_host[s:e], y_host[s:e])

    def timed(fn, k, w, preroll=0.5):
        for _ in range(w):
            fn()
        pipe.drain()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        # untimed pre-roll AFTER the barrier, immediately before the first event: the idle of the barrier lets the board
        # leave its power cap and the first ~100 ms would run at boost clocks (round-1 SCALE numbers were inflated by it)
        t_pre = time.time()
        while preroll > 0 and time.time() - t_pre < preroll:
            for _ in range(4):
                fn()
            pipe.drain()
            torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        pipe.join()          # the timed stream waits for every outstanding device->host copy (no-op when none)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
            dist.barrier()
        return ms

    def graphed(fn):
        """fn's launches captured in a CUDA graph; returns its replay.  The forward step itself is a graph replay: timing a
        group of its kernels through per-launch Python / ctypes calls would charge the small pyramid levels (10-20 us kernels)
        with host launch time the step never pays."""
        fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        torch.cuda.synchronize()
        return g.replay

    # clocks take a few hundred ms to settle after the idle -> busy transition: untimed pre-roll before the W warm-ups
    t_pre = time.time()
    while time.time() - t_pre < 1.5:
        for _ in range(10):
            fwd(x_dev[:B])
        torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    ms_total = timed(step_resident, args.steps, warmup, preroll=args.preroll)
    sampler.stop_flag.set()
    sampler.join(timeout=2)
    sampler_e2e = ClockSampler(local)        # the two legs run seconds apart: each gets its own clock record
    sampler_e2e.start()
    ms_e2e = timed(step_e2e, args.steps, warmup, preroll=args.preroll)
    sampler_e2e.stop_flag.set()
    sampler_e2e.join(timeout=2)
    pipe.drain()

    # parity of the benchmarked configuration itself: 4 crops of the batch against the CPU oracle (the checker, not the
    # thing measured); the e2e leg's host output must equal the resident leg's
    parity = None
    if rank == 0 and not args.no_cpu_baseline:
        from oracle.gfpgan_ocr_oracle import OcrNetConfig, gfpgan_ocr_forward, psnr01, to01
        idx = [0, B // 3, (2 * B) // 3, B - 1]
        got = fwd(x_dev[:B])[idx].float().cpu()
        cfg = OcrNetConfig(**{k: v for k, v in NET_KW.items() if k not in ('decoder_load_path', 'fix_decoder')})
        sd_cpu = {k: v.detach().cpu() for k, v in net.state_dict().items()}
        torch.set_num_threads(os.cpu_count())
        ref, _ = gfpgan_ocr_forward(sd_cpu, cfg, x_host[idx], False)
        a, b_ = to01(got), to01(ref)
        parity = {'crops_checked': idx, 'batch': B, 'max_abs_01': (a - b_).abs().max().item(), 'psnr_db': psnr01(a, b_),
                  'bar': 'max-abs <= 2e-2 on [0,1], PSNR >= 45 dB vs the fp32 CPU oracle',
                  'e2e_equals_resident': bool(torch.equal(y_host[:B][idx], got))}
        parity['ok'] = bool(parity['max_abs_01'] <= 2e-2 and parity['psnr_db'] >= 45.0 and parity['e2e_equals_resident'])

    # what the callers call (gfpgan_model.py:803 / api.py:104 defaults): fresh noise per call, the U-Net's toRGB heads
    # returned, uint8 images in and out
    variants = None
    if world == 1 and not args.no_extras:
        variants = {}
        for name, fn in (('randomize_noise_true', lambda: net(x_dev[:B], return_rgb=False, randomize_noise=True)),
                         ('return_rgb_true', lambda: net(x_dev[:B], return_rgb=True, randomize_noise=False)),
                         ('defaults_rgb_and_noise', lambda: net(x_dev[:B]))):
            ms = timed(fn, args.steps, 3, preroll=0.2) / args.steps
            variants[name] = {'crops_per_s': B / (ms / 1e3), 'ms_per_step': ms}
        img_u8 = torch.randint(0, 256, (B, H, W, 3), device=dev, dtype=torch.uint8)
        ms = timed(lambda: net.restore_uint8(img_u8, randomize_noise=False), args.steps, 3, preroll=0.2) / args.steps
        variants['restore_uint8'] = {'crops_per_s': B / (ms / 1e3), 'ms_per_step': ms,
                                     'config': 'uint8 HWC BGR in -> uint8 HWC BGR out (api.py:96-105 on the device)'}

    # dominant kernel: conv_igemm_kernel.  Time only its launches of one step (same buffers as the real step),
    # back to back on the current stream, with CUDA events.
    from image_restoration_b200.ops import ConvOp
    conv_ops = [s for s in plan.steps if isinstance(s, ConvOp)]

    def conv_only():
        for op in conv_ops:
            op()
    ms_conv = timed(graphed(conv_only), args.steps, warmup, preroll=0.3)
    n_conv_ops = len(conv_ops)
    conv_alg_bytes = 0          # algorithmic bytes of the conv launches: input once + output once + weights, fp16
    for op in conv_ops:
        d = op.desc
        m = d.m_b * d.m_h * d.m_w
        conv_alg_bytes += m * d.cin * 2 + (0 if d.no_store else m * d.cout * (4 if d.out_fp32 else 2)) \
            + d.cout * d.num_taps * d.cin * 2

    # memory-bound kernels: every launch of each kind, back to back, against its algorithmic bytes
    from image_restoration_b200.engine import PwOp
    pw_groups = {}
    for st in plan.steps:
        if isinstance(st, PwOp):
            pw_groups.setdefault(st.name, []).append(st)
    pw_report = []
    if world == 1:
        for name, group in pw_groups.items():
            def run_group(g=group):
                for op in g:
                    op()
            ms = timed(graphed(run_group), args.steps, 3, preroll=0.05) / args.steps
            nbytes = sum(op.nbytes for op in group)
            big = max(group, key=lambda op: op.nbytes)        # the largest launch of the kind, alone
            ms_big = timed(graphed(big), args.steps, 3, preroll=0.05) / args.steps
            pw_report.append({'kernel': name, 'launches_per_step': len(group), 'algorithmic_mb_per_step': nbytes / 1e6,
                              'ms_per_step': ms, 'achieved_gbs': nbytes / ms / 1e6,
                              'largest_launch': {'algorithmic_mb': big.nbytes / 1e6, 'ms': ms_big,
                                                 'achieved_gbs': big.nbytes / ms_big / 1e6}})

    # fused degradation kernel (north_star (c)): crops/s of pyblur blur + down-resize + noise + up-resize + quantise,
    # beside the reference's own CPU library calls (oracle/pyblur_oracle.py) on a few crops
    degr = None
    if world == 1 and not args.no_cpu_baseline and not args.no_extras:
        import numpy as np
        from image_restoration_b200 import degradation as dg
        from oracle import pyblur_oracle as po
        rng = np.random.RandomState(0)
        DB = 256
        gt = rng.randint(0, 256, (DB, H, W, 3)).astype(np.uint8)
        kernels, sizes, nz = dg.random_degradation_params(DB, H, W, rng=rng)
        gt_d, nz_d = torch.from_numpy(gt).to(dev), torch.from_numpy(nz).to(dev)
        packed = dg.pack_degradation(kernels, sizes, dev)
        ms = timed(lambda: dg.degrade_batch(gt_d, kernels, sizes, nz_d, packed=packed), 10, 3, preroll=0.1) / 10
        t0 = time.time()
        n_cpu = 8
        for b in range(n_cpu):
            lw, lh = sizes[b]
            po.degrade(gt[b], kernels[b], sizes[b], nz[b, :lh, :lw])
        cpu_ms = (time.time() - t0) / n_cpu * 1e3
        alg = DB * (H * W * 3 + H * W * 3 * 4)
        degr = {'kernel': 'degrade_kernel', 'crops_per_s': DB / (ms / 1e3), 'ms_per_batch': ms, 'batch': DB,
                'params': 'pyblur kernel type/size, scale U[4,12], sigma U[0,20] drawn per crop (np seed 0), explicit noise',
                'algorithmic_mb_per_batch': alg / 1e6, 'achieved_gbs': alg / ms / 1e6,
                'cpu_reference': {'ms_per_crop': cpu_ms, 'crops_per_s': 1e3 / cpu_ms, 'cores': 1, 'kind': 'port',
                                  'sample': f'{n_cpu} crops through scipy.signal.convolve2d + cv2.resize (the '
                                            'reference\'s own library calls, oracle/pyblur_oracle.py)'}}

        # the whole LQ synthesis of __getitem__ (blur kinds of the training YAML, JPEG, jitter, gray) in one launch
        from oracle import degrade_full_oracle as dfo
        opt = dict(blur_kernel_size=21, kernel_list=['iso', 'aniso', 'motion', 'average', 'median', 'bilateral', 'pyblur'],
                   kernel_prob=[0.08, 0.08, 0.08, 0.08, 0.08, 0.08, 0.28], blur_sigma=[0.1, 10], downsample_range=[4.0, 12.0],
                   noise_range=[0, 20], jpeg_range=[30, 100], color_jitter_prob=0.3, color_jitter_shift=20, color_jitter_pt_prob=0.3,
           gray_prob=0.01)
        import random as _random
        prm = dg.sample_params(DB, H, W, opt, py_random=_random.Random(0), np_random=np.random.RandomState(0))
        pk = dg.pack_degrade_full(dev=dev, **prm)
        ms_full = timed(lambda: dg.degrade_full_batch(gt_d, packed=pk), 10, 3, preroll=0.1) / 10
        t0 = time.time()
        for b in range(n_cpu):
            lw, lh = prm['sizes'][b]
            dfo.degrade_full(gt[b], prm['modes'][b], prm['kernels'][b], (lw, lh), prm['noise'][b, :lh, :lw],
                             prm['quality'][b], prm['jitter'][b], prm['gray'][b], exact_blur=False, lib_jpeg=True,
                             bilateral_sigma=prm['bilateral_sigma'][b], cj=prm['color_jitter_pt'][b])
        cpu_full_ms = (time.time() - t0) / n_cpu * 1e3
        # BASELINE config 4: tiled full-frame inference, 3x1080x1920 -> 45 overlapping 256x256 tiles in one batch
        from image_restoration_b200.tiling import TiledRestorer
        kw256 = dict(NET_KW, input_width=256, input_height=256)
        torch.manual_seed(0)
        net256 = GFPGANv1OCR(**kw256).eval().to(dev)
        tiler = TiledRestorer(net256, overlap=32, micro_batch=64)
        frame = torch.rand(3, 1080, 1920, device=dev) * 2 - 1
        with torch.no_grad():
            ms_frame = timed(lambda: tiler(frame, randomize_noise=False), 10, 3, preroll=0.1) / 10
        tiled = {'frames_per_s': 1e3 / ms_frame, 'ms_per_frame': ms_frame, 'tiles_per_frame': 45,
                 'tiles_per_s': 45e3 / ms_frame, 'algorithmic_tflops': 45 * 34.63e9 / ms_frame / 1e9,
                 'config': '3x1080x1920 frame, 256x256 tiles, overlap 32 (5x9 tiles, one batch), gather + forward + '
                           'ramp blend, frame resident in HBM'}
        del net256, tiler
        # network_d of the training YAMLs (forward only: first piece of the training-step row)
        from image_restoration_b200.disc import StyleGAN2Discriminator
        torch.manual_seed(0)
        netd = StyleGAN2Discriminator(input_width=W, input_height=H, channel_multiplier=1).eval().to(dev)
        xd = torch.rand(B, 3, H, W, device=dev) * 2 - 1
        ms_d = timed(lambda: netd(xd), 10, 3, preroll=0.1) / 10
        disc = {'crops_per_s': B / (ms_d / 1e3), 'ms_per_batch': ms_d, 'batch': B,
                'config': f'StyleGAN2Discriminator(input_width={W}, input_height={H}, channel_multiplier=1) forward, eager '
                          'launches'}
        del netd
        # trainable part of net_g (U-Net encoder -> style code, decoder -> SFT conditions) forward + backward: the part of
        # the training-step row that exists; the frozen decoder's input gradients and the losses do not yet
        from image_restoration_b200.backward import unet_forward
        names = ('conv_body_first', 'conv_body_down', 'final_conv', 'final_linear', 'conv_body_up', 'condition_scale',
                 'condition_shift')
        sd_t = {k: v.detach().clone().requires_grad_() for k, v in net.state_dict().items() if k.split('.')[0] in names}
        xt = torch.rand(B, 3, H, W, device=dev) * 2 - 1
        cots = []

        def unet_step():
            for v in sd_t.values():
                v.grad = None
            sc, cd = unet_forward(sd_t, xt, different_w=True, num_style_feat=NET_KW['num_style_feat'])
            if not cots:
                cots.extend(torch.randn_like(t) for t in [sc] + cd)
            torch.autograd.backward([sc] + cd, cots)
        ms_t = timed(unet_step, 4, 2, preroll=0.1) / 4
        unet_train = {'crops_per_s': B / (ms_t / 1e3), 'ms_per_batch': ms_t, 'batch': B,
                      'parameters': sum(v.numel() for v in sd_t.values()),
                      'config': 'U-Net encoder + decoder + SFT heads + final_linear of GFPGANv1OCR (everything optimizer_g updates '
                                'with fix_decoder=True), forward + backward through backward.unet_forward, eager launches '
                                '(host-bound at this batch; 3.4 k crops/s at B=256)'}
        del sd_t, xt, cots
        degr['full_chain'] = {'kernel': 'degrade_full_kernel', 'crops_per_s': DB / (ms_full / 1e3), 'ms_per_batch': ms_full,
                              'batch': DB, 'achieved_gbs': alg / ms_full / 1e6,
                              'stages': 'blur (iso/aniso/motion/average/median/bilateral/pyblur, kernel_list and kernel_prob of '
                                        'training_config/train_gfpgan_v4_square_license_mix_pyblur.yml) + resize + noise + '
                                        'JPEG + resize + colour jitter + gray + color_jitter_pt (same probabilities as that YAML)',
                              'cpu_reference': {'ms_per_crop': cpu_full_ms, 'crops_per_s': 1e3 / cpu_full_ms, 'cores': 1,
                                                'kind': 'port', 'sample': f'{n_cpu} crops through cv2.filter2D / scipy '
                                                'convolve2d + cv2.resize + cv2.imencode/imdecode (the reference\'s calls)'}}

    sr_records = None
    if world == 1 and not args.no_extras:
        sr_records = sr_arch_records(dev, timed, peaks()[0])

    # BASELINE configs[4]: the training step, batch 256 per GPU (all ranks take part: NCCL all-reduce of the gradients)
    training = None
    if not args.no_extras and args.train_batch > 0:
        del plan, conv_ops, pw_groups, conv_only
        pipe.slots.clear()
        eng.plans.clear()
        torch.cuda.empty_cache()
        training = training_step_record(dev, world, args.train_batch, args.train_steps, timed, fix_decoder=False)
        frozen = training_step_record(dev, world, args.train_batch, args.train_steps, timed, fix_decoder=True, brief=True)
        ms_frozen = frozen.pop('ms')
        training['fix_decoder_true'] = dict(frozen, ms_per_step=ms_frozen, crops_per_s=args.train_batch * world / (ms_frozen / 1e3),
                                            note='same step with the StyleGAN2 decoder frozen (input gradients only)')

    if rank == 0:
        tf_peak, hbm_peak, src = peaks()
        crops_per_step = B if world == 1 else args.total          # whole job
        crops = crops_per_step * args.steps
        value = crops / (ms_total / 1e3)
        e2e_value = crops / (ms_e2e / 1e3)
        conv_ms_step = ms_conv / args.steps
        conv_tflops = GEMM_GFLOP_PER_CROP * 1e9 * B / (conv_ms_step / 1e3) / 1e12
        workload = workload_name(world, B, args.total, n_local)
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': warmup,
            'ms_per_step': ms_total / args.steps, 'higher_is_better': True, 'scaling': 'weak' if world == 1 else 'strong',
            'vs_baseline': None, 'dtype': 'fp16', 'data': 'synthetic',
            'config': {'workload': workload,
                       'operands': 'fp16 x fp16 -> fp32 accumulate (tcgen05 kind::f16); bf16 operands fail the '
                                   '2e-2/45 dB parity bar (SURVEY App. D)',
                       'micro_batch': B, 'crops_per_step': crops_per_step, 'crops_per_gpu_per_step': n_local,
                       'parallelism': f'dp{world} (independent shards, no collective)',
                       'l2': 'working set per forward >> 126 MB L2 (activations ~GBs at batch 64); no explicit flush',
                       'executor': 'CUDA graph replay of the launch plan',
                       'kernel_group_timing': 'roofline / memory_bound_kernels: the launches of one step of that kind captured in '
                                              'a CUDA graph and replayed back to back on one stream (CUDA events around the replays)',
                       'timing': 'CUDA events, max over ranks; >= 0.5 s untimed pre-roll after the barrier, immediately '
                                 'before the first event'},
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': crops_per_step * 3 * H * W * 4,
                    'd2h_bytes_per_step': crops_per_step * 3 * H * W * 4, 'ms_per_step': ms_e2e / args.steps,
                    'bytes': 'whole job (all ranks), fp32 NCHW in and out',
                    'clocks': sampler_e2e.summary(),
                    'note': 'same API call per micro-batch as the resident leg plus the H2D / D2H copies on their own streams '
                            '(hidden behind the kernels); the legs are timed seconds apart and the power-capped SM clock '
                            'drifts by 1-2 % in between, so e2e may read slightly above value'},
            'gpu_launches': int(launches_per_mb * mbs_per_step * args.steps * world),
            'launches_per_step': int(launches_per_mb * mbs_per_step * world),
            'launches_per_micro_batch': int(launches_per_mb),
            'roofline': {'bound': 'tensor', 'kernel': 'conv_igemm_kernel', 'achieved': conv_tflops, 'peak': tf_peak,
                         'unit': 'TFLOP/s', 'frac': conv_tflops / tf_peak, 'traffic': None,
                         'peak_source': f'{src} bf16_tflops_sustained', 'launches_per_micro_batch': n_conv_ops,
                         'avg_launch_ms': conv_ms_step / n_conv_ops,
                         'algorithmic_gflop_per_launch': GEMM_GFLOP_PER_CROP * B / n_conv_ops,
                         'conv_share_of_step': conv_ms_step * mbs_per_step / (ms_total / args.steps),
                         'whole_net_frac': value / world * GFLOP_PER_CROP * 1e9 / 1e12 / tf_peak},
            'clocks': sampler.summary(),
        }
        if parity is not None:
            line['parity'] = parity
        if sr_records is not None:
            line['sr_archs'] = sr_records
        if training is not None:
            tb = training['batch_per_gpu'] * world
            cps = tb / (training['ms'] / 1e3)
            training.update({
                'crops_per_s': cps, 'ms_per_step': training.pop('ms'), 'global_batch': tb, 'n_gpus': world,
                'algorithmic_tflops': cps * training['gflop_per_crop'] / 1e3,
                'frac_of_tensor_peak': cps / world * training['gflop_per_crop'] / 1e3 / tf_peak,
                'config': 'BASELINE configs[4]: per step and GPU: fused degradation kernel (training-YAML kernel mix incl. JPEG) '
                          '-> lq, img2tensor -> gt; net_g forward (return_rgb) + l_g_pix 0.1 + image pyramid 1.0 + perceptual 1.0 / '
                          'style 50 (VGG19 conv1_2..conv5_4, seeded random weights: the ImageNet checkpoint is not available '
                          'offline) + l_g_gan 0.1 (wgan_softplus) -> backward (U-Net wgrad/dgrad, StyleGAN2 decoder dgrad + wgrad: fix_decoder false as in training_config/*.yml, '
                          'net_d dgrad, VGG dgrad) -> NCCL all-reduce -> fused Adam + EMA; net_d forward on fake and real -> '
                          'logistic loss -> backward -> all-reduce -> fused Adam; R1 penalty every 16th iteration (timed '
                          'separately: r1_iteration).  Not in the step: facial-component / identity terms (off for plates)',
                'timing': 'CUDA events, max over ranks; phase_ms from one extra profiled step on rank 0'})
            line['training_step'] = training
        if variants is not None:
            line['caller_variants'] = variants
        traffic_file = os.path.join(ROOT, 'profiles', 'conv_traffic.json')
        if os.path.exists(traffic_file):
            with open(traffic_file) as f:
                tr = json.load(f)
            line['roofline']['traffic'] = tr.get('dram_bytes_per_launch')
            line['roofline']['traffic_source'] = tr.get('source')
            line['roofline']['algorithmic_bytes_per_launch'] = conv_alg_bytes / n_conv_ops
        for r in pw_report:
            r['frac_of_hbm_peak'] = r['achieved_gbs'] / hbm_peak
            r['largest_launch']['frac_of_hbm_peak'] = r['largest_launch']['achieved_gbs'] / hbm_peak
        pw_report.sort(key=lambda r: -r['ms_per_step'])
        line['memory_bound_kernels'] = pw_report
        if degr is not None:
            degr['frac_of_hbm_peak'] = degr['achieved_gbs'] / hbm_peak
            line['degradation'] = degr
            line['tiled_full_frame'] = tiled
            line['discriminator_forward'] = disc
            line['unet_forward_backward'] = unet_train
        if pw_report:
            top = pw_report[0]
            line['roofline_hbm'] = {'bound': 'hbm', 'kernel': top['kernel'], 'achieved': top['achieved_gbs'],
                                    'peak': hbm_peak, 'unit': 'GB/s', 'frac': top['achieved_gbs'] / hbm_peak,
                                    'traffic': None, 'peak_source': f'{src} hbm_gbs',
                                    'note': 'all launches of the kind back to back (five pyramid levels, the small ones '
                                            'are latency-bound); largest_launch in memory_bound_kernels is one level'}
        if not args.no_cpu_baseline and world == 1:
            v, cores, sample = cpu_oracle_rate()
            line['cpu_baseline'] = {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample}
        else:
            line['cpu_baseline'] = None
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
        sys.stdout.flush()
        sys.stderr.flush()
        if rank != 0:
            os._exit(0)      # skip the exit handlers: with NCCL_DEBUG=INFO they print after rank 0's JSON line otherwise
        time.sleep(1.0)      # the other ranks are gone by now
    if rank == 0:
        bad = parity is not None and not parity['ok']
        if bad:
            sys.stderr.write('bench.py: the benchmarked batch fails the parity bar against the CPU oracle: ' + json.dumps(parity) + '\n')
        sys.stderr.flush()
        print(json.dumps(line), flush=True)          # the last line of stdout
        if world > 1:
            os._exit(1 if bad else 0)
        if bad:
            sys.exit(1)


if __name__ == '__main__':
    main()
