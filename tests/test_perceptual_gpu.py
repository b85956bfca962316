"""Perceptual + style loss on the B200 kernels (perceptual.py) at the plate geometry against the fp32 oracle
(oracle/perceptual_oracle.py = losses.py:250-356 over a functional torchvision-vgg19) and its autograd; seeded random VGG19
weights (the ImageNet checkpoint is not available offline)."""
import pytest
import torch
import torch.nn.functional as F

from oracle import perceptual_oracle as po

pytestmark = pytest.mark.gpu
LAYER_WEIGHTS = {'conv1_2': 0.1, 'conv2_2': 0.1, 'conv3_4': 1.0, 'conv4_4': 1.0, 'conv5_4': 1.0}


@pytest.mark.parametrize('H,W,B,style', [(128, 384, 2, 50.0), (32, 96, 3, 50.0), (128, 384, 2, 0.0)])
def test_perceptual_loss_matches_oracle(H, W, B, style):
    from image_restoration_b200 import perceptual
    sd = {k: v.cuda() for k, v in po.random_vgg19_state_dict(1).items()}
    g = torch.Generator().manual_seed(3)
    gt = F.interpolate(torch.rand(B, 3, H // 8, W // 8, generator=g) * 2 - 1, size=(H, W), mode='bilinear', align_corners=False).cuda()
    x = (gt + 0.2 * torch.randn(B, 3, H, W, generator=g).cuda()).clamp(-1, 1)
    xr = x.clone().requires_grad_()
    lp, ls = po.perceptual_loss(sd, xr, gt, LAYER_WEIGHTS, 1.0, style, True, True)
    (lp + (ls if ls is not None else 0)).backward()
    S = 4096.0 * B
    vgg = perceptual.VGG19Features(sd, list(LAYER_WEIGHTS), torch.device('cuda'), use_input_norm=True, range_norm=True)
    xa = x.clone().requires_grad_()
    total, p, s = perceptual.perceptual_loss(xa, gt, vgg, LAYER_WEIGHTS, 1.0, style, S)
    total.backward(gradient=torch.full_like(total, S))
    torch.cuda.synchronize()
    print(f'percep {p.item():.5f} / {lp.item():.5f}   style {s.item():.6f} / {(ls.item() if ls is not None else 0):.6f}')
    assert abs(p.item() - lp.item()) <= 5e-3 * lp.item()
    if style > 0:
        assert abs(s.item() - ls.item()) <= 2e-2 * ls.item()
    ga, gb = xa.grad / S, xr.grad
    cos = F.cosine_similarity(ga.flatten().double(), gb.flatten().double(), dim=0).item()
    rel = ((ga - gb).double().pow(2).mean().sqrt() / gb.double().pow(2).mean().sqrt()).item()
    print(f'd/dx: cos {cos:.5f} rel rms {rel:.3e}')
    # the L1 over fp16 features flips sign(x - t) where the two features round to (nearly) the same value
    assert cos >= 0.995 and rel <= 0.1


def test_trainer_with_perceptual_loss_runs():
    from image_restoration_b200 import perceptual, train
    from tests.test_train_full_gpu import _data, _nets
    net, netd, _ = _nets(seed=0)
    net.train()
    sd = po.random_vgg19_state_dict(0)
    vgg = perceptual.VGG19Features(sd, list(LAYER_WEIGHTS), torch.device('cuda'), use_input_norm=True, range_norm=True)
    tr = train.GFPGANTrainer(net, netd, perceptual=dict(vgg=vgg, layer_weights=LAYER_WEIGHTS, perceptual_weight=1.0, style_weight=50.0))
    lq, gt = _data(4, seed=2)
    first = None
    for it in range(1, 5):
        tr.feed_data(lq, gt)
        log = tr.optimize_parameters(it)
        first = first or {k: v.item() for k, v in log.items()}
    torch.cuda.synchronize()
    last = {k: v.item() for k, v in log.items()}
    print('first', first)
    print('last ', last)
    assert all(torch.isfinite(torch.tensor(v)) for v in last.values())
    assert 'l_g_percep' in last and 'l_g_style' in last and last['l_g_percep'] > 0
