"""Host logic of the training step on CPU, through the C-ABI simulator (tests/cabi_sim.py): the launch order and buffer wiring
of train.FrozenDecoderFunction / train_forward / GFPGANTrainer are Python above the ABI; here every entry point is restated
with torch CPU ops on the same pointers, and the resulting losses and gradients are compared with torch.autograd over the fp32
oracle (oracle/gfpgan_ocr_oracle.py, oracle/disc_oracle.py) — the reference's own optimize_parameters arithmetic
(gfpgan_model.py:494-691: l_g_pix + image pyramid + l_g_gan, then l_d).  The kernels themselves are checked on the GPU
(tests/test_train_ops_gpu.py, tests/test_train_full_gpu.py)."""
import math

import pytest
import torch
import torch.nn.functional as F

from tests import cabi_sim
from tests.helpers import KW

W, H = 48, 16


def _nets(seed=0):
    from image_restoration_b200 import GFPGANv1OCR
    from image_restoration_b200.disc import StyleGAN2Discriminator
    torch.manual_seed(seed)
    net = GFPGANv1OCR(input_width=W, input_height=H, decoder_load_path=None, fix_decoder=True, **KW)
    netd = StyleGAN2Discriminator(input_width=W, input_height=H, channel_multiplier=1)
    # non-trivial noise gains / biases in the frozen decoder, as a trained checkpoint has (the stock init zeroes them)
    with torch.no_grad():
        for n, p in net.stylegan_decoder.named_parameters():
            if n.endswith('.weight') and p.numel() == 1:
                p.fill_(0.3)
            if n.endswith('activate.bias') or n.endswith('to_rgb1.bias') or '.to_rgbs.' in n and n.endswith('.bias'):
                p.normal_(0, 0.1)
    return net, netd


def _data(B, seed=1):
    g = torch.Generator().manual_seed(seed)
    gt = F.interpolate(torch.rand(B, 3, 4, 12, generator=g) * 2 - 1, size=(H, W), mode='bilinear', align_corners=False)
    lq = (gt + 0.1 * torch.randn(B, 3, H, W, generator=g)).clamp(-1, 1)
    return lq, gt


def _cmp(name, ga, gb, cos_min, rel_max):
    cos = F.cosine_similarity(ga.double().flatten(), gb.double().flatten(), dim=0).item()
    rel = ((ga - gb).double().pow(2).mean().sqrt() / gb.double().pow(2).mean().sqrt().clamp_min(1e-30)).item()
    assert cos >= cos_min and rel <= rel_max, (name, cos, rel)
    return cos, rel


def _oracle_g_losses(net, netd, lq, gt, noises, weights):
    from oracle.disc_oracle import discriminator_forward
    from oracle.gfpgan_ocr_oracle import OcrNetConfig, gfpgan_ocr_forward
    cfg = OcrNetConfig(input_width=W, input_height=H, **KW)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    train = {k for k, p in net.named_parameters() if p.requires_grad}
    for k in train:
        sd[k].requires_grad_()
    out, rgbs = gfpgan_ocr_forward.__wrapped__(sd, cfg, lq, True, noises=noises)
    from image_restoration_b200.train import construct_img_pyramid
    pyr = construct_img_pyramid(gt, len(rgbs))
    sdd = {k: v.detach().clone() for k, v in netd.state_dict().items()}
    l_pix = weights[0] * (out - gt).abs().mean()
    l_pyr = sum(weights[1] * (r - t).abs().mean() for r, t in zip(rgbs, pyr))
    l_gan = weights[2] * F.softplus(-discriminator_forward(sdd, out)).mean()
    return out, (l_pix, l_pyr, l_gan), sd, train


def test_generator_losses_and_gradients_match_oracle_autograd():
    """l_g_pix + pyramid + l_g_gan through train_forward -> FrozenDecoderFunction -> disc (input gradient) vs the oracle."""
    from image_restoration_b200 import train
    net, netd = _nets()
    B = 2
    lq, gt = _data(B)
    g = torch.Generator().manual_seed(5)
    noises = [torch.randn(B, 1, n.shape[2], n.shape[3], generator=g) for n in
              [net.state_dict()[f'stylegan_decoder.noises.noise{j}'] for j in range(2 * (net.log_size - 2) + 1)]]
    weights = (0.1, 1.0, 0.1)
    out_ref, (lp, ly, lg), sd_ref, trainable = _oracle_g_losses(net, netd, lq, gt, noises, weights)
    (lp + ly + lg).backward()

    S = 4096.0 * B
    with cabi_sim.installed() as sim:
        for p in netd.parameters():
            p.requires_grad_(False)
        output, rgbs = train.train_forward(net, lq, return_rgb=True, noise=noises)
        assert output.shape == (B, 3, H, W) and output.dtype == torch.float32 and output.requires_grad
        assert [tuple(r.shape) for r in rgbs] == [(B, 3, 8 * 2 ** i, 24 * 2 ** i) for i in range(len(rgbs))]
        l_pix = train.l1_loss(output, gt, weights[0], S)
        pyr = train.construct_img_pyramid(gt, len(rgbs))
        l_pyr = sum(train.l1_loss(r, t, weights[1], S) for r, t in zip(rgbs, pyr))
        pred = train.disc_forward_image(dict(netd.named_parameters()), output)
        l_gan = train.gan_softplus_loss(pred, True, weights[2], S)
        total = l_pix + l_pyr + l_gan
        total.backward(gradient=torch.full_like(total, S))
        assert sim.launches > 100
    # forward values
    err = (output.detach() - out_ref.detach()).abs().max().item()
    assert err <= 2e-2 * max(1.0, out_ref.abs().max().item()), err
    for a, b in ((l_pix, lp), (l_pyr, ly), (l_gan, lg)):
        assert abs(a.item() - b.item()) <= 3e-3 * abs(b.item()) + 1e-5, (a.item(), b.item())
    # gradients of every trainable parameter (fp16 activations in the simulated kernels: a few leaky-ReLU / L1 sign flips)
    worst = (1.0, 0.0)
    for k, p in net.named_parameters():
        if not p.requires_grad:
            assert p.grad is None, k
            continue
        gb = sd_ref[k].grad
        assert (p.grad is not None) == (gb is not None), k
        if gb is None:
            continue
        cos, rel = _cmp(k, p.grad / S, gb, 0.99, 0.15)
        worst = (min(worst[0], cos), max(worst[1], rel))
    print(f'generator gradients vs oracle autograd: worst cos {worst[0]:.5f}, worst rel rms {worst[1]:.3e}')


def test_decoder_function_gradients_strict():
    """FrozenDecoderFunction alone with fp32-exact inputs for the oracle (the same fp16-rounded style code and conditions on
    both sides): d(style_code) and d(conditions) against autograd through oracle.stylegan_decoder."""
    from image_restoration_b200 import train
    from oracle.gfpgan_ocr_oracle import OcrNetConfig, stylegan_decoder
    net, _ = _nets(seed=3)
    B = 2
    L = net.log_size - 2
    cfg = OcrNetConfig(input_width=W, input_height=H, **KW)
    g = torch.Generator().manual_seed(7)
    code = (0.5 * torch.randn(B, cfg.num_latent, KW['num_style_feat'], generator=g)).half()
    conds, ch = [], None
    sdn = net.state_dict()
    for lvl in range(L):
        h, w = 8 * 2 ** lvl, 24 * 2 ** lvl
        c = sdn[f'condition_scale.{lvl}.2.weight'].shape[0]
        conds += [(1 + 0.3 * torch.randn(B, h, w, c, generator=g)).half(), (0.3 * torch.randn(B, h, w, c, generator=g)).half()]
    noises = [torch.randn(B, 1, sdn[f'stylegan_decoder.noises.noise{j}'].shape[2], sdn[f'stylegan_decoder.noises.noise{j}'].shape[3],
                          generator=g) for j in range(2 * L + 1)]
    cot = torch.randn(B, 3, H, W, generator=g)
    # oracle
    code_r = code.float().requires_grad_()
    conds_r = [c.float().permute(0, 3, 1, 2).contiguous().requires_grad_() for c in conds]
    sd = {k: v.detach().float() for k, v in sdn.items()}
    img_r = stylegan_decoder(sd, cfg, code_r, conds_r, noises)
    (img_r * cot).sum().backward()
    with cabi_sim.installed():
        st = train.decoder_state(net)
        code_a = code.clone().requires_grad_()
        conds_a = [c.clone().requires_grad_() for c in conds]
        img = train.decoder_apply(net, st, noises, code_a, conds_a)
        img.backward(cot)
    assert (img.detach() - img_r.detach()).abs().max().item() <= 1e-2 * img_r.abs().max().item()
    cos, rel = _cmp('d_style_code', code_a.grad.float(), code_r.grad, 0.999, 0.05)
    print(f'd(style_code): cos {cos:.6f} rel {rel:.3e}')
    for i, (ca, cr) in enumerate(zip(conds_a, conds_r)):
        cos, rel = _cmp(f'd_cond{i}', ca.grad.float().permute(0, 3, 1, 2), cr.grad, 0.999, 0.05)
        print(f'd(cond {i}): cos {cos:.6f} rel {rel:.3e}')


def _trainable_decoder_case(dev):
    """Inputs of the fix_decoder=False checks (also used by tests/test_train_full_gpu.py): net with a trainable decoder whose
    noise gains / biases are non-trivial, fp16-rounded style code and conditions, noise planes and a cotangent."""
    from image_restoration_b200 import GFPGANv1OCR
    torch.manual_seed(4)
    net = GFPGANv1OCR(input_width=W, input_height=H, decoder_load_path=None, fix_decoder=False, **KW)
    with torch.no_grad():
        for n, p in net.stylegan_decoder.named_parameters():
            if n.endswith('.weight') and p.numel() == 1:
                p.fill_(0.3)
            if n.endswith('activate.bias') or n.endswith('to_rgb1.bias') or '.to_rgbs.' in n and n.endswith('.bias'):
                p.normal_(0, 0.1)
    net = net.to(dev)
    L = net.log_size - 2
    g = torch.Generator().manual_seed(9)
    B = 3
    sdn = net.state_dict()
    code = (0.5 * torch.randn(B, 2 * L + 2, KW['num_style_feat'], generator=g)).half().to(dev)
    conds = []
    for lvl in range(L):
        h, w = 8 * 2 ** lvl, 24 * 2 ** lvl
        c = sdn[f'condition_scale.{lvl}.2.weight'].shape[0]
        conds += [(1 + 0.3 * torch.randn(B, h, w, c, generator=g)).half().to(dev), (0.3 * torch.randn(B, h, w, c, generator=g)).half().to(dev)]
    noises = [torch.randn(B, 1, *sdn[f'stylegan_decoder.noises.noise{j}'].shape[2:], generator=g).to(dev) for j in range(2 * L + 1)]
    cot = torch.randn(B, 3, H, W, generator=g).to(dev)
    return net, code, conds, noises, cot


def _oracle_decoder_param_grads(net, code, conds, noises, cot):
    from oracle.gfpgan_ocr_oracle import OcrNetConfig, stylegan_decoder
    cfg = OcrNetConfig(input_width=W, input_height=H, **KW)
    sd = {k: v.detach().float().clone() for k, v in net.state_dict().items()}
    names = [f'stylegan_decoder.{n}' for n, _ in net.stylegan_decoder.named_parameters()]
    for k in names:
        sd[k].requires_grad_()
    code_r = code.float().requires_grad_()
    conds_r = [c.float().permute(0, 3, 1, 2).contiguous().requires_grad_() for c in conds]
    img_r = stylegan_decoder(sd, cfg, code_r, conds_r, noises)
    (img_r * cot).sum().backward()
    return img_r.detach(), code_r.grad, {k: sd[k].grad for k in names}


def test_trainable_decoder_parameter_gradients():
    """fix_decoder=False (every shipped training YAML, training_config/*.yml `fix_decoder: false`): DecoderFunction also returns
    the gradient of every decoder parameter on the path — modulated conv weights (direct term + their path through the
    demodulation), modulation linears, noise gains, activation biases, ToRGB weights / biases, the constant input — against
    autograd through oracle.stylegan_decoder; the unused style MLP gets none, as in the reference."""
    from image_restoration_b200 import train
    net, code, conds, noises, cot = _trainable_decoder_case('cpu')
    img_r, dcode_r, pgrads = _oracle_decoder_param_grads(net, code, conds, noises, cot)
    with cabi_sim.installed():
        st = train.decoder_state(net)
        assert st.trainable
        code_a = code.clone().requires_grad_()
        conds_a = [c.clone().requires_grad_() for c in conds]
        img = train.decoder_apply(net, st, noises, code_a, conds_a)
        img.backward(cot)
    assert (img.detach() - img_r).abs().max().item() <= 1e-2 * img_r.abs().max().item()
    _cmp('d_style_code', code_a.grad.float(), dcode_r, 0.999, 0.05)
    seen = 0
    for n, p in net.stylegan_decoder.named_parameters():
        gr = pgrads[f'stylegan_decoder.{n}']
        if gr is None:
            assert p.grad is None and n.startswith('style_mlp'), n
            continue
        assert p.grad is not None and p.grad.shape == p.shape and p.grad.dtype == p.dtype, n
        # the noise gains are scalars: sum over every element of dz * noise with noise ~ N(0, 1) independent of dz, a sum that
        # cancels to ~1 / sqrt(n) of its terms, so the ~1e-3 of flipped leaky-ReLU branches shows up amplified
        cos, rel = _cmp(n, p.grad, gr, 0.999, 0.15 if p.numel() == 1 else 0.05)
        seen += 1
    assert seen == 2 + 5 * (2 * (net.log_size - 2) + 1) + 4 * (net.log_size - 1) - 1, seen


def test_trainer_steps_match_a_torch_reference_loop():
    """GFPGANTrainer.optimize_parameters (G then D, FlatAdam + EMA) for two iterations against the same loop written with the
    fp32 oracle, torch.autograd and torch.optim.Adam (gfpgan_model.py:494-691 without perceptual / R1)."""
    from image_restoration_b200 import GFPGANv1OCR, train
    from oracle.disc_oracle import discriminator_forward
    net, netd = _nets(seed=2)
    B = 2
    lq, gt = _data(B, seed=4)
    # reference copies
    sd_g = {k: v.detach().clone() for k, v in net.state_dict().items()}
    sd_d = {k: v.detach().clone().requires_grad_() for k, v in netd.state_dict().items()}
    g_names = [k for k, p in net.named_parameters() if p.requires_grad]
    for k in g_names:
        sd_g[k].requires_grad_()
    opt_g = torch.optim.Adam([sd_g[k] for k in g_names], lr=2e-3, betas=(0.0, 0.99))
    opt_d = torch.optim.Adam(list(sd_d.values()), lr=2e-3, betas=(0.0, 0.99))
    from oracle.gfpgan_ocr_oracle import OcrNetConfig, gfpgan_ocr_forward
    cfg = OcrNetConfig(input_width=W, input_height=H, **KW)
    ref_logs = []
    for it in range(2):
        opt_g.zero_grad()
        out, rgbs = gfpgan_ocr_forward.__wrapped__(sd_g, cfg, lq, True)          # stored noise buffers (deterministic)
        pyr = train.construct_img_pyramid(gt, len(rgbs))
        sd_d_frozen = {k: v.detach() for k, v in sd_d.items()}
        l_g = 0.1 * (out - gt).abs().mean() + sum((r - t).abs().mean() for r, t in zip(rgbs, pyr)) \
            + 0.1 * F.softplus(-discriminator_forward(sd_d_frozen, out)).mean()
        l_g.backward()
        opt_g.step()
        opt_d.zero_grad()
        l_d = F.softplus(-discriminator_forward(sd_d, gt)).mean() + F.softplus(discriminator_forward(sd_d, out.detach())).mean()
        l_d.backward()
        opt_d.step()
        ref_logs.append((l_g.item(), l_d.item()))
    with cabi_sim.installed():
        torch.manual_seed(2)
        ema = GFPGANv1OCR(input_width=W, input_height=H, decoder_load_path=None, fix_decoder=True, **KW)
        ema.load_state_dict(net.state_dict())
        tr = train.GFPGANTrainer(net, netd, net_g_ema=ema)
        orig_forward = train.train_forward
        train.train_forward = lambda n, x, **kw: orig_forward(n, x, randomize_noise=False, **{k: v for k, v in kw.items() if k != 'randomize_noise'})
        try:
            logs = []
            for it in range(2):
                tr.feed_data(lq, gt)
                log = tr.optimize_parameters(it + 1)
                l_g = log['l_g_pix'] + sum(v for k, v in log.items() if k.startswith('l_p_')) + log['l_g_gan']
                logs.append((l_g.item(), log['l_d'].item()))
        finally:
            train.train_forward = orig_forward
    # iteration 1 sees identical weights; Adam's first step is sign descent (m / sqrt(v) = +-1), so gradients that the fp16
    # activations leave near zero land 2 * lr apart and iteration 2 differs by a few per cent
    for tol, (a_g, a_d), (b_g, b_d) in zip((2e-3, 4e-2), logs, ref_logs):
        assert abs(a_g - b_g) <= tol * abs(b_g) and abs(a_d - b_d) <= tol * abs(b_d), (logs, ref_logs)
    # parameters after two Adam steps (lr 2e-3: each step moves every weight by ~lr whatever the gradient's scale, so a
    # wrong gradient SIGN anywhere shows up as a 2 * lr difference)
    moved = []
    for k, p in net.named_parameters():
        if p.requires_grad:
            d = (p.detach() - sd_g[k].detach()).abs()
            moved.append((d > 3e-3).float().mean().item())
    print('fraction of net_g weights further than 1.5 * lr from the torch loop after two steps:', sum(moved) / len(moved))
    assert sum(moved) / len(moved) < 0.05, sum(moved) / len(moved)
    for k, p in netd.named_parameters():
        d = (p.detach() - sd_d[k].detach()).abs()
        assert (d > 3e-3).float().mean().item() < 0.10, (k, (d > 3e-3).float().mean().item())
    # EMA of the trainable parameters moved towards the new weights
    dec = 0.5 ** (32 / (10 * 1000))
    k0 = g_names[0]
    assert not torch.equal(dict(ema.named_parameters())[k0], dict(net.named_parameters())[k0])


def test_trainer_updates_a_trainable_decoder():
    """GFPGANTrainer with fix_decoder=False (training_config/*.yml): the decoder's parameters are part of optimizer_g, their
    packed copies are rebuilt after every step (FlatAdam bumps the parameter versions), the unused style MLP stays put."""
    from image_restoration_b200 import GFPGANv1OCR, train
    from image_restoration_b200.disc import StyleGAN2Discriminator
    torch.manual_seed(6)
    net = GFPGANv1OCR(input_width=W, input_height=H, decoder_load_path=None, fix_decoder=False, **KW)
    netd = StyleGAN2Discriminator(input_width=W, input_height=H, channel_multiplier=1)
    lq, gt = _data(2, seed=8)
    before = {k: v.detach().clone() for k, v in net.named_parameters()}
    with cabi_sim.installed():
        tr = train.GFPGANTrainer(net, netd)
        assert len(tr.g_params) == len(list(net.parameters()))
        losses = []
        for it in range(2):
            tr.feed_data(lq, gt)
            st0 = train.decoder_state(net)
            log = tr.optimize_parameters(it + 1)
            assert train.decoder_state(net) is not st0                 # repacked from the updated weights
            losses.append(float(log['l_g_pix']))
    assert all(math.isfinite(v) for v in losses)
    for k, p in net.named_parameters():
        moved = (p.detach() - before[k]).abs().max().item()
        if 'style_mlp' in k:
            assert moved == 0.0, k
        else:
            assert 0.0 < moved <= 2.5 * 2 * 2e-3, (k, moved)           # two Adam steps of at most ~lr each


def _ddp_worker(rank, world, port, q):
    import os
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(2)
    from image_restoration_b200 import train
    net, netd = _nets(seed=2)                                  # identical replicas
    lq, gt = _data(2, seed=10 + rank)                          # each rank its own shard of the global batch
    with cabi_sim.installed():
        tr = train.GFPGANTrainer(net, netd)
        orig = train.train_forward
        train.train_forward = lambda n, x, **kw: orig(n, x, randomize_noise=False, **{k: v for k, v in kw.items() if k != 'randomize_noise'})
        tr.feed_data(lq, gt)
        log = tr.optimize_parameters(1)
        # the exchanged gradient is the SUM over ranks in GradAllReducer.flat (the 1 / world average is inside the Adam step)
        g_sum = tr.sync_g.flat.clone()
    q.put((rank, tr.opt_g.flat.double().sum().item(), tr.opt_g.flat.double().abs().sum().item(),
           tr.opt_d.flat.double().sum().item(), g_sum.double().abs().sum().item(), {k: float(v) for k, v in log.items()}))
    dist.destroy_process_group()


def test_trainer_two_rank_gloo_replicas_stay_identical():
    """BASELINE configs[4] is data-parallel: one process per GPU, per-rank shards, one all-reduce of the flat gradient buffer
    per network (base_model.py:62-76).  World-2 gloo run of the trainer on the simulator: after the step both replicas hold
    bit-identical generator and discriminator weights although they saw different crops."""
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, g0, ga0, d0, gs0, log0), (_, g1, ga1, d1, gs1, log1) = res
    assert g0 == g1 and ga0 == ga1 and d0 == d1, res
    assert gs0 == gs1 and gs0 > 0
    assert log0['l_g_pix'] != log1['l_g_pix']                  # different shards
    assert all(math.isfinite(v) for v in list(log0.values()) + list(log1.values()))


def test_r1_penalty_gradients_match_torch_double_backward():
    """r1.r1_penalty_backward (primal backward signals x tangent pass + the Hessian term of the minibatch standard deviation)
    against the reference's own formulation: autograd.grad(real_pred.sum(), gt, create_graph=True) -> penalty -> backward
    (losses.py:492-506, gfpgan_model.py:683-689), on the fp32 oracle discriminator."""
    from image_restoration_b200 import r1
    from oracle.disc_oracle import discriminator_forward
    _, netd = _nets(seed=1)
    B = 4
    _, gt = _data(B, seed=3)
    weight = 10 / 2 * 16                                      # r1_reg_weight / 2 * net_d_reg_every
    sd = {k: v.detach().clone().requires_grad_() for k, v in netd.state_dict().items()}
    x = gt.clone().requires_grad_()
    pred = discriminator_forward(sd, x)
    grad_real = torch.autograd.grad(pred.sum(), x, create_graph=True)[0]
    l_d_r1 = weight * grad_real.pow(2).view(B, -1).sum(1).mean() + 0 * pred[0]
    l_d_r1.sum().backward()
    params = dict(netd.named_parameters())
    S = 512.0
    with cabi_sim.installed():
        for p in params.values():
            p.grad = torch.ones_like(p)                       # the penalty's gradient is ADDED to what l_d.backward() left there
        val = r1.r1_penalty_backward(params, gt, weight, grad_out_scale=S)
    assert abs(val.item() - l_d_r1.sum().item()) <= 3e-3 * l_d_r1.sum().item(), (val.item(), l_d_r1.sum().item())
    for k, p in params.items():
        ga, gb = (p.grad - 1) / S, sd[k].grad
        if gb.abs().max().item() == 0:                        # biases above the statistic layer: no dependence at all
            assert ga.abs().max().item() <= 1e-7, k
            continue
        _cmp(k, ga, gb, 0.999, 0.05)
