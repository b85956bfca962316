"""A real slice of GFPGANModel.optimize_parameters (basicsr/models/gfpgan_model.py:494-796) on the B200 kernels: the image
pyramid loss (lines 502-511, 531-536: L1 between the U-Net's toRGB heads and the bilinear GT pyramid), which involves only the
trainable part of net_g, followed by the optimizer_g step (Adam betas (0, 0.99), lr 2e-3) and the EMA update.
Forward, backward and the optimiser run in libb200ir.so (backward.unet_forward, optim.FlatAdam); torch does the 3-channel
L1 itself and the GT pyramid (data preparation)."""
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
