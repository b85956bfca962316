"""A real slice of GFPGANModel.optimize_parameters (basicsr/models/gfpgan_model.py:494-796) on the B200 kernels: the image
pyramid loss (lines 502-511, 531-536: L1 between the U-Net's toRGB heads and the bilinear GT pyramid), which involves only the
trainable part of net_g, followed by the optimizer_g step (Adam betas (0, 0.99), lr 2e-3) and the EMA update.
Forward, backward and the optimiser run in libb200ir.so (backward.unet_forward, optim.FlatAdam); torch does the 3-channel
L1 itself and the GT pyramid (data preparation).  Second half: network_d (StyleGAN2Discriminator) forward + backward through
backward.disc_forward against the oracle's autograd, strict tests of its two new pieces, and net_d update steps."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

NAMES = ('conv_body_first', 'conv_body_down', 'final_conv', 'final_linear', 'conv_body_up', 'condition_scale',
         'condition_shift', 'toRGB')


def _pyramid(gt, levels):
    """construct_img_pyramid (gfpgan_model.py:326-332)."""
    out = [gt]
    for _ in range(levels - 1):
        out.insert(0, F.interpolate(out[0], scale_factor=0.5, mode='bilinear', align_corners=False))
    return out


def _pyramid_loss(out_rgbs, pyramid_gt):
    loss = 0.0
    for rgb, gt in zip(out_rgbs, pyramid_gt):
        loss = loss + (rgb[..., :3].float() - gt.permute(0, 2, 3, 1)).abs().mean()       # L1Loss(reduction='mean'), weight 1
    return loss


def _setup(B, seed=0):
    from image_restoration_b200 import GFPGANv1OCR
    from tests.helpers import KW
    torch.manual_seed(seed)
    W, H = 384, 128
    net = GFPGANv1OCR(input_width=W, input_height=H, decoder_load_path=None, fix_decoder=True, **KW)
    sd = {k: v.detach().clone().cuda().requires_grad_() for k, v in net.state_dict().items() if k.split('.')[0] in NAMES}
    # smooth synthetic "plates": low-frequency images in [-1, 1]; the LQ input is the GT plus noise
    gt = F.interpolate(torch.rand(B, 3, 8, 24, device='cuda') * 2 - 1, size=(H, W), mode='bilinear', align_corners=False)
    lq = (gt + 0.1 * torch.randn_like(gt)).clamp(-1, 1)
    return sd, lq, gt, KW


def test_pyramid_loss_gradients_match_oracle_autograd():
    from image_restoration_b200.backward import unet_forward
    from oracle.gfpgan_ocr_oracle import OcrNetConfig, gfpgan_ocr_forward
    from image_restoration_b200 import GFPGANv1OCR
    sd, lq, gt, KW = _setup(2)
    _, _, rgbs = unet_forward(sd, lq, num_style_feat=KW['num_style_feat'], return_rgb=True)
    pyr = _pyramid(gt, len(rgbs))
    loss = _pyramid_loss(rgbs, pyr)
    loss.backward()
    # oracle: the full reference forward in fp32 with autograd; out_rgbs are its toRGB heads
    torch.manual_seed(0)
    full = GFPGANv1OCR(input_width=384, input_height=128, decoder_load_path=None, fix_decoder=True, **KW).state_dict()
    sd_ref = {k: v.detach().clone().cuda() for k, v in full.items()}
    for k in sd:
        sd_ref[k] = sd[k].detach().clone().requires_grad_()
    cfg = OcrNetConfig(input_width=384, input_height=128, **KW)
    _, ref_rgbs = gfpgan_ocr_forward.__wrapped__(sd_ref, cfg, lq, True)
    loss_ref = sum((r - g).abs().mean() for r, g in zip(ref_rgbs, pyr))
    loss_ref.backward()
    torch.cuda.synchronize()
    print(f'pyramid loss {loss.item():.6f} vs oracle {loss_ref.item():.6f}')
    assert abs(loss.item() - loss_ref.item()) <= 2e-3 * abs(loss_ref.item())
    used = [k for k in sd if sd_ref[k].grad is not None]
    assert any(k.startswith('toRGB') for k in used) and any(k.startswith('conv_body_first') for k in used)
    for k in used:
        ga, gb = sd[k].grad, sd_ref[k].grad
        assert ga is not None, k
        cos = F.cosine_similarity(ga.double().flatten(), gb.double().flatten(), dim=0).item()
        rel = ((ga - gb).double().pow(2).mean().sqrt() / gb.double().pow(2).mean().sqrt().clamp_min(1e-30)).item()
        # sign(rgb - gt) flips where the fp16 head lands on the other side of the target, on top of the leaky-ReLU flips
        assert cos >= 0.99 and rel <= 0.15, (k, cos, rel)


def test_pyramid_loss_training_steps_reduce_the_loss():
    from image_restoration_b200.backward import unet_forward
    from image_restoration_b200.optim import FlatAdam
    sd, lq, gt, KW = _setup(4)
    params = list(sd.values())
    ema = [p.detach().clone() for p in params]
    opt = FlatAdam(params, lr=2e-3, betas=(0.0, 0.99), ema_params=ema)
    losses = []
    for it in range(12):
        opt.zero_grad()
        _, _, rgbs = unet_forward(sd, lq, num_style_feat=KW['num_style_feat'], return_rgb=True)
        loss = _pyramid_loss(rgbs, _pyramid(gt, len(rgbs)))
        loss.backward()
        opt.step(ema_decay=0.5 ** (32 / (10 * 1000)))
        losses.append(loss.item())
    torch.cuda.synchronize()
    print('pyramid loss per step:', ' '.join(f'{v:.4f}' for v in losses))
    assert all(math.isfinite(v) for v in losses)
    assert losses[-1] < 0.7 * losses[0], losses
    assert all(torch.isfinite(e).all() for e in ema)


def test_discriminator_forward_backward_against_oracle():
    """network_d (StyleGAN2Discriminator) forward + backward through backward.disc_forward with the logistic loss of the D step
    (gfpgan_model.py: l_d = softplus(-real) + softplus(fake)), against torch.autograd over the fp32 oracle."""
    from image_restoration_b200.backward import disc_forward
    from image_restoration_b200.disc import StyleGAN2Discriminator
    from oracle.disc_oracle import discriminator_forward
    torch.manual_seed(0)
    W, H, B = 384, 128, 4
    netd = StyleGAN2Discriminator(input_width=W, input_height=H, channel_multiplier=1)
    sd_a = {k: v.detach().clone().cuda().requires_grad_() for k, v in netd.state_dict().items()}
    sd_b = {k: v.detach().clone().requires_grad_() for k, v in sd_a.items()}
    x = torch.rand(B, 3, H, W, device='cuda') * 2 - 1
    sign = torch.tensor([1.0, 1.0, -1.0, -1.0], device='cuda').view(B, 1)          # two "real", two "fake" samples
    # fp16 activation gradients need a loss scale: d(score) ~ 0.1 shrinks by 1/sqrt(fan_in) per linear layer and would reach
    # the fp16 subnormals (< 6e-5) at final_conv; the weight gradients come back in fp32 and are unscaled exactly
    S = 4096.0
    score = disc_forward(sd_a, x)
    (F.softplus(-sign * score.float()).mean() * S).backward()
    for v in sd_a.values():
        v.grad /= S
    ref = discriminator_forward(sd_b, x)
    F.softplus(-sign * ref).mean().backward()
    torch.cuda.synchronize()
    err = (score.float() - ref).abs().max().item()
    print(f'disc scores {score.flatten().tolist()} vs oracle {ref.flatten().tolist()} (max err {err:.2e})')
    assert err <= 2e-3 + 2e-2 * ref.abs().max().item()
    stats = {}
    for k in sd_a:
        ga, gb = sd_a[k].grad, sd_b[k].grad
        assert ga is not None and gb is not None, k
        cos = F.cosine_similarity(ga.double().flatten(), gb.double().flatten(), dim=0).item()
        rel = ((ga - gb).double().pow(2).mean().sqrt() / gb.double().pow(2).mean().sqrt().clamp_min(1e-30)).item()
        print(f'disc grad {k}: rel rms {rel:.3e} cos {cos:.6f}')
        stats[k] = (cos, rel)
    # Twelve leaky ReLUs sit in series between the score and the first layer.  After eleven fp16 convs the pre-activations of
    # final_conv differ from the fp32 oracle's by a few 1e-3 of their scale, so ~3e-3 of them take the other branch (each
    # changes that element's gradient five-fold): ~8 % relative RMS from final_conv upwards, measured.  Every component is
    # pinned strictly with the branches held equal (test_backward_gpu.py, test_minibatch_stddev_backward,
    # test_equal_linear_activation_backward below); the two linears behind the last activation show the kernels' own accuracy.
    assert stats['final_linear.1.weight'][1] <= 2e-2 and stats['final_linear.0.weight'][1] <= 2e-2, stats
    for k, (cos, rel) in stats.items():
        assert cos >= 0.99 and rel <= 0.15, (k, cos, rel)


@pytest.mark.parametrize('B,h,w,C,group', [(4, 4, 12, 512, 4), (8, 4, 12, 64, 4), (2, 3, 5, 32, 2)])
def test_minibatch_stddev_backward(B, h, w, C, group):
    from image_restoration_b200.backward import MinibatchStddevFunction
    from oracle.disc_oracle import minibatch_stddev
    torch.manual_seed(B + C)
    x = torch.randn(B, C, h, w, device='cuda').half()
    xg = x.permute(0, 2, 3, 1).contiguous().requires_grad_()
    cat = MinibatchStddevFunction.apply(xg, group)
    cot = torch.randn(B, h, w, C + 1, device='cuda').half()
    (cat[..., :C + 1].float() * cot.float()).sum().backward()
    x_ref = x.float().requires_grad_()
    ref = minibatch_stddev(x_ref, group)
    (ref * cot.float().permute(0, 3, 1, 2)).sum().backward()
    torch.cuda.synchronize()
    assert (cat[..., :C + 1].float().permute(0, 3, 1, 2) - ref).abs().max().item() <= 2e-3
    assert (cat[..., C + 1:] == 0).all()
    err = (xg.grad.float().permute(0, 3, 1, 2) - x_ref.grad).abs().max().item()
    assert err <= 2e-3 * x_ref.grad.abs().max().item(), err


def test_equal_linear_activation_backward():
    """EqualLinear(activation='fused_lrelu') (final_linear.0 of the discriminator) with the branch taken from the kernel's output."""
    from image_restoration_b200.backward import equal_linear
    torch.manual_seed(3)
    B, cin, cout = 4, 24576, 512
    weight = torch.randn(cout, cin, device='cuda', requires_grad=True)
    bias = (0.1 * torch.randn(cout, device='cuda')).requires_grad_()
    x = torch.randn(B, cin, device='cuda').half()
    dy = torch.randn(B, cout, device='cuda').half()
    xg = x.clone().requires_grad_()
    y = equal_linear(xg, weight, bias, 1.0, True)
    y.backward(dy)
    w_ref, b_ref, x_ref = weight.detach().clone().requires_grad_(), bias.detach().clone().requires_grad_(), x.float().requires_grad_()
    z = F.linear(x_ref, w_ref / math.sqrt(cin), b_ref)
    y_ref = z * torch.where(y.detach().float() > 0, math.sqrt(2.0), 0.2 * math.sqrt(2.0))
    y_ref.backward(dy.float())
    torch.cuda.synchronize()
    for name, g, r in (('y', y.detach().float(), y_ref.detach()), ('dx', xg.grad.float(), x_ref.grad), ('dweight', weight.grad, w_ref.grad),
                       ('dbias', bias.grad, b_ref.grad)):
        rel = ((g - r).double().pow(2).mean().sqrt() / r.double().pow(2).mean().sqrt()).item()
        print(f'equal linear + fused_lrelu {name}: rel rms {rel:.3e}')
        assert rel <= 1e-3, (name, rel)


def test_discriminator_steps_reduce_the_logistic_loss():
    """The net_d update of optimize_parameters (gfpgan_model.py: real_d_pred / fake_d_pred through net_d, l_d = gan loss of both,
    optimizer_d.step) on the B200 kernels, without the R1 penalty (needs a double backward): sharp vs blurred synthetic plates.
    A static loss scale keeps the fp16 activation gradients in range; FlatAdam's grad_scale removes it in the fused step."""
    from image_restoration_b200.backward import disc_forward
    from image_restoration_b200.disc import StyleGAN2Discriminator
    from image_restoration_b200.optim import FlatAdam
    torch.manual_seed(0)
    W, H, B = 384, 128, 4
    netd = StyleGAN2Discriminator(input_width=W, input_height=H, channel_multiplier=1)
    sd = {k: v.detach().clone().cuda().requires_grad_() for k, v in netd.state_dict().items()}
    opt = FlatAdam(list(sd.values()), lr=2e-3, betas=(0.0, 0.99))
    real = (torch.rand(B, 3, H, W, device='cuda') * 2 - 1)
    fake = F.avg_pool2d(real, 9, stride=1, padding=4)                     # the "restored" images: blurred copies
    S = 256.0
    losses = []
    for it in range(10):
        opt.zero_grad()
        l_real = F.softplus(-disc_forward(sd, real).float()).mean()
        l_fake = F.softplus(disc_forward(sd, fake).float()).mean()
        ((l_real + l_fake) * S).backward()
        opt.step(grad_scale=1.0 / S)
        losses.append((l_real + l_fake).item())
    torch.cuda.synchronize()
    print('l_d per step:', ' '.join(f'{v:.4f}' for v in losses))
    assert all(math.isfinite(v) for v in losses)
    assert abs(losses[0] - 2 * math.log(2)) < 0.1 and losses[-1] < 0.5 * losses[0], losses
