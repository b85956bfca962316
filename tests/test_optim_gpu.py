"""GPU parity of the fused Adam + EMA step (b200ir_adam_step through image_restoration_b200.optim.FlatAdam) against
torch.optim.Adam and the EMA rule of BaseModel.model_ema on the same seeded gradients, with the reference's settings
(gfpgan_model.py:217-248: lr 2e-3, betas (0, 0.99)).  fp32 arithmetic in a different association order: rel 2e-6."""
import copy

import pytest
import torch
from torch import nn

pytestmark = pytest.mark.gpu


def make_net():
    torch.manual_seed(0)
    return nn.Sequential(nn.Conv2d(3, 8, 3), nn.Conv2d(8, 5, 1), nn.Linear(7, 3)).cuda()     # 443 parameters: ragged tail


@pytest.mark.parametrize('betas,wd', [((0.0, 0.99), 0.0), ((0.9, 0.999), 0.01)])
def test_adam_and_ema_match_torch(betas, wd):
    from image_restoration_b200.optim import FlatAdam
    ref_net, ref_ema = make_net(), make_net()
    net, ema = copy.deepcopy(ref_net), copy.deepcopy(ref_ema)
    ref_opt = torch.optim.Adam(ref_net.parameters(), lr=2e-3, betas=betas, weight_decay=wd)
    opt = FlatAdam(net.parameters(), lr=2e-3, betas=betas, weight_decay=wd, ema_params=ema.parameters())
    decay = 0.5 ** (32 / (10 * 1000))
    g = torch.Generator(device='cuda').manual_seed(1)
    for step in range(5):
        for p, q in zip(ref_net.parameters(), net.parameters()):
            gr = torch.randn(p.shape, device='cuda', generator=g) * (10.0 ** (step - 2))
            p.grad, q.grad = gr.clone(), gr.clone()
        ref_opt.step()
        for e, p in zip(ref_ema.parameters(), ref_net.parameters()):
            e.data.mul_(decay).add_(p.data, alpha=1 - decay)
        opt.step(ema_decay=decay)
        for (n, p), q, e, f in zip(ref_net.named_parameters(), net.parameters(), ref_ema.parameters(), ema.parameters()):
            assert torch.allclose(q, p, rtol=2e-6, atol=1e-8), (step, n, (q - p).abs().max().item())
            assert torch.allclose(f, e, rtol=2e-6, atol=1e-8), (step, n)
    # the module still works on the flat storage
    y = net[0](torch.zeros(1, 3, 5, 5, device='cuda'))
    assert torch.isfinite(y).all() and net[0].weight.data_ptr() >= opt.flat.data_ptr()


def test_flat_gradient_buffer_with_world_average():
    from image_restoration_b200.optim import FlatAdam
    ref_net, net = make_net(), make_net()
    ref_opt = torch.optim.Adam(ref_net.parameters(), lr=2e-3, betas=(0.0, 0.99))
    opt = FlatAdam(net.parameters())
    flat = torch.randn(opt.numel, device='cuda')            # "sum over 8 ranks", averaged inside the step
    for p, off in zip(ref_net.parameters(), opt.offsets):   # every parameter starts on a 256-byte boundary (optim.flat_layout)
        assert off % 64 == 0
        p.grad = (flat[off:off + p.numel()] / 8).view_as(p).clone()
    ref_opt.step()
    opt.step(flat_grad=flat, grad_scale=1.0 / 8)
    for p, q in zip(ref_net.parameters(), net.parameters()):
        assert torch.allclose(q, p, rtol=2e-6, atol=1e-8)
    with pytest.raises(ValueError):
        opt.step(ema_decay=0.99)                            # no EMA copy registered


def test_non_finite_gradient_elements_are_dropped():
    """An overflowed fp16 activation gradient (inf / nan in a few elements) must not turn the parameters and the Adam
    moments into NaN: those elements are treated as zero gradient, every other element steps exactly as torch.optim.Adam."""
    import torch
    from image_restoration_b200.optim import FlatAdam
    ref_net, net = make_net(), make_net()
    ref_opt = torch.optim.Adam(ref_net.parameters(), lr=2e-3, betas=(0.0, 0.99))
    opt = FlatAdam(net.parameters())
    g = torch.Generator(device='cuda').manual_seed(0)
    bad = {}
    for i, (p, q) in enumerate(zip(ref_net.parameters(), net.parameters())):
        gr = torch.randn(p.shape, device='cuda', generator=g)
        p.grad = gr.clone()
        gb = gr.clone()
        flat = gb.view(-1)
        flat[0] = float('inf')
        if flat.numel() > 3:
            flat[3] = float('nan')
        p.grad.view(-1)[0] = 0.0
        if flat.numel() > 3:
            p.grad.view(-1)[3] = 0.0
        q.grad = gb
    ref_opt.step()
    opt.step()
    for p, q in zip(ref_net.parameters(), net.parameters()):
        assert torch.isfinite(q).all()
        assert torch.allclose(q, p, rtol=2e-6, atol=1e-8)
    assert torch.isfinite(opt.exp_avg).all() and torch.isfinite(opt.exp_avg_sq).all()
