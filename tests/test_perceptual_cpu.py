"""Perceptual + style loss (losses.py:250-356 over vgg_arch.py:56-160): the oracle restatement against torchvision's vgg19
module (the reference's third-party feature extractor) on the same weights, and the product path (perceptual.py) on the
C-ABI simulator against the oracle's autograd."""
import pytest
import torch
import torch.nn.functional as F

from oracle import perceptual_oracle as po
from tests import cabi_sim

LAYER_WEIGHTS = {'conv1_2': 0.1, 'conv2_2': 0.1, 'conv3_4': 1.0, 'conv4_4': 1.0, 'conv5_4': 1.0}      # training YAMLs


def test_oracle_vgg_matches_torchvision_module():
    tv = pytest.importorskip('torchvision')
    sd = po.random_vgg19_state_dict(0)
    net = tv.models.vgg19(weights=None)
    net.load_state_dict(sd, strict=False)
    net.eval()
    x = torch.rand(2, 3, 32, 48) * 2 - 1
    xn = ((x + 1) / 2 - torch.tensor(po.MEAN).view(1, 3, 1, 1)) / torch.tensor(po.STD).view(1, 3, 1, 1)
    ref = {}
    h = xn
    for idx, layer in enumerate(net.features[:35]):
        h = layer(h) if not isinstance(layer, torch.nn.ReLU) else F.relu(h)
        if po.VGG19_NAMES[idx] in LAYER_WEIGHTS:
            ref[po.VGG19_NAMES[idx]] = h.clone()
    got = po.vgg_features(sd, x, list(LAYER_WEIGHTS))
    assert set(got) == set(ref)
    for k in ref:
        assert torch.allclose(got[k], ref[k], rtol=1e-5, atol=1e-6), k


def test_perceptual_and_style_loss_on_simulator_match_oracle_autograd():
    from image_restoration_b200 import perceptual
    sd = po.random_vgg19_state_dict(1)
    g = torch.Generator().manual_seed(3)
    B, H, W = 2, 32, 96
    gt = F.interpolate(torch.rand(B, 3, 4, 12, generator=g) * 2 - 1, size=(H, W), mode='bilinear', align_corners=False)
    x = (gt + 0.2 * torch.randn(B, 3, H, W, generator=g)).clamp(-1, 1)
    xr = x.clone().requires_grad_()
    lp, ls = po.perceptual_loss(sd, xr, gt, LAYER_WEIGHTS, 1.0, 50.0, True, True)
    (lp + ls).backward()
    S = 4096.0 * B
    with cabi_sim.installed():
        vgg = perceptual.VGG19Features(sd, list(LAYER_WEIGHTS), torch.device('cpu'), use_input_norm=True, range_norm=True)
        xa = x.clone().requires_grad_()
        total, p, s = perceptual.perceptual_loss(xa, gt, vgg, LAYER_WEIGHTS, 1.0, 50.0, S)
        total.backward(gradient=torch.full_like(total, S))
    print(f'percep {p.item():.5f} / {lp.item():.5f}   style {s.item():.5f} / {ls.item():.5f}')
    assert abs(p.item() - lp.item()) <= 5e-3 * lp.item() and abs(s.item() - ls.item()) <= 1e-2 * ls.item()
    ga, gb = xa.grad / S, xr.grad
    cos = F.cosine_similarity(ga.flatten().double(), gb.flatten().double(), dim=0).item()
    rel = ((ga - gb).double().pow(2).mean().sqrt() / gb.double().pow(2).mean().sqrt()).item()
    print(f'd/dx: cos {cos:.5f} rel rms {rel:.3e}')
    assert cos >= 0.995 and rel <= 0.1
