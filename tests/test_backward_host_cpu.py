"""Host-side logic of the backward path that needs no GPU: the weight packs that turn input gradients into forward convs
(ops.conv_dgrad_weight, ops.convt_merged_weight with swapped channel roles) and the tap / phase bookkeeping of the stride-2
weight gradient, checked against torch's own conv gradients on the CPU."""
import math

import pytest

import torch
import torch.nn.functional as F

from image_restoration_b200 import ops


def _unpack(wp, cin):
    """packed [cout, 9*cin] (tap-major) -> conv weight [cout, cin, 3, 3]."""
    cout = wp.shape[0]
    return wp.view(cout, 3, 3, cin).permute(0, 3, 1, 2).contiguous()


def test_conv_dgrad_weight_is_the_adjoint_conv():
    torch.manual_seed(0)
    cin, cout, B, H, W = 5, 7, 2, 6, 9
    w = torch.randn(cout, cin, 3, 3, dtype=torch.float64)
    wp = w.permute(0, 2, 3, 1).reshape(cout, 9 * cin).contiguous()
    wt = ops.conv_dgrad_weight(wp, cin)                                   # [cin, 9*cout]: a conv from cout to cin channels
    assert wt.shape == (cin, 9 * cout)
    dz = torch.randn(B, cout, H, W, dtype=torch.float64)
    ref = torch.nn.grad.conv2d_input((B, cin, H, W), w, dz, padding=1)
    got = F.conv2d(dz, _unpack(wt, cout), padding=1)
    assert torch.allclose(got, ref, atol=1e-12)


def test_convt_merged_weight_with_swapped_roles_is_the_stride2_dgrad():
    """ResBlockFunction.backward: dgrad of F.conv2d(p, W, stride=2) = conv_transpose2d(dz, W, stride=2); the merged GEMM's
    weight is built from W.permute(1, 0, 2, 3).  Emulates the merged GEMM (four phases x four taps) with torch on the CPU."""
    torch.manual_seed(1)
    cin, cout, B, h, w = 4, 6, 2, 3, 5                                    # dz is h x w, p is (2h+1) x (2w+1)
    W2 = torch.randn(cout, cin, 3, 3).half().double()                    # on the fp16 grid: the pack is exact
    dz = torch.randn(B, cout, h, w, dtype=torch.float64)
    ref = F.conv_transpose2d(dz, W2, stride=2)                            # [B, cin, 2h+1, 2w+1]
    big = ops.convt_merged_weight(W2.permute(1, 0, 2, 3).float(), 1.0).double().view(4, cin, 4, cout)
    out = torch.zeros(B, cin, 2 * h + 2, 2 * w + 2, dtype=torch.float64)
    dzp = F.pad(dz, (1, 1, 1, 1))                                         # zero outside: taps (i - ty, j - tx)
    for py in range(2):
        for px in range(2):
            for ty in range(2):
                for tx in range(2):
                    wk = big[py * 2 + px, :, ty * 2 + tx, :]              # [cin, cout]
                    src = dzp[:, :, 1 - ty:1 - ty + h + 1, 1 - tx:1 - tx + w + 1]          # dz[i - ty, j - tx], i <= h, j <= w
                    out[:, :, py::2, px::2] += torch.einsum('oc,bcij->boij', wk, src)
    assert torch.allclose(out[:, :, :2 * h + 1, :2 * w + 1], ref, atol=1e-10)


def test_stride2_wgrad_phase_and_tap_bookkeeping():
    """conv3x3_s2_wgrad: kernel element (kh, kw) of the stride-2 conv is tap (kh // 2 + 1, kw // 2 + 1) of the stride-1 weight
    gradient over pixel phase (kh % 2, kw % 2) of p; every element is produced exactly once."""
    torch.manual_seed(2)
    cin, cout, B, H, W = 3, 4, 2, 6, 8
    p = torch.randn(B, cin, H + 1, W + 1, dtype=torch.float64)
    dy = torch.randn(B, cout, H // 2, W // 2, dtype=torch.float64)
    ref = torch.nn.grad.conv2d_weight(p, (cout, cin, 3, 3), dy, stride=2)
    got = torch.full_like(ref, float('nan'))
    for ry in range(2):
        for rx in range(2):
            phase = F.pad(p[:, :, ry::2, rx::2], (0, 1, 0, 1))            # view extent (H/2 + 1) x (W/2 + 1), zero beyond
            full = torch.nn.grad.conv2d_weight(F.pad(phase, (1, 1, 1, 1)), (cout, cin, 3, 3),
                                               F.pad(dy, (0, phase.shape[3] - dy.shape[3], 0, phase.shape[2] - dy.shape[2])))
            for sy in ((0, 1) if ry == 0 else (0,)):
                for sx in ((0, 1) if rx == 0 else (0,)):
                    assert torch.isnan(got[:, :, 2 * sy + ry, 2 * sx + rx]).all()
                    got[:, :, 2 * sy + ry, 2 * sx + rx] = full[:, :, sy + 1, sx + 1]
    assert not torch.isnan(got).any()
    assert torch.allclose(got, ref, atol=1e-10)


def test_pack_equal_conv_scale():
    from image_restoration_b200.backward import pack_equal_conv, pack_equal_conv3x3
    w = torch.randn(8, 16, 3, 3)
    wp, s = pack_equal_conv3x3(w)
    assert math.isclose(s, 1 / math.sqrt(16 * 9)) and wp.shape == (8, 144) and wp.dtype == torch.float16
    wp1, s1 = pack_equal_conv(torch.randn(8, 16, 1, 1))
    assert math.isclose(s1, 0.25) and wp1.shape == (8, 16)
    assert torch.allclose(_unpack(wp.float(), 16), (w * s).half().float())


def test_training_functions_have_no_cpu_path():
    """The autograd wrappers refuse CPU tensors instead of falling back to torch (there is no CPU path in the product)."""
    import pytest
    from image_restoration_b200 import backward
    x = torch.zeros(1, 8, 8, 16, dtype=torch.float16)
    w = torch.zeros(16, 16, 3, 3, requires_grad=True)
    b = torch.zeros(16, requires_grad=True)
    with pytest.raises(RuntimeError):
        backward.conv_layer3x3(x, w, b)
    with pytest.raises(RuntimeError):
        backward.res_block(x, w, b, w, b, torch.zeros(16, 16, 1, 1))
    with pytest.raises(RuntimeError):
        backward.res_up_block(x, w, b, w, b, torch.zeros(16, 16, 1, 1))
    with pytest.raises(RuntimeError):
        backward.equal_linear(torch.zeros(2, 16, dtype=torch.float16), torch.zeros(16, 16), torch.zeros(16))
    with pytest.raises(RuntimeError):
        backward.FirstConvFunction.apply(torch.zeros(1, 3, 8, 8), torch.zeros(16, 3, 1, 1), b)
    with pytest.raises(RuntimeError):
        backward.ToRGBHeadFunction.apply(x, torch.zeros(3, 16, 1, 1), torch.zeros(3))
    with pytest.raises(RuntimeError):
        backward.MinibatchStddevFunction.apply(x, 1)
    with pytest.raises(RuntimeError):
        backward.AddFunction.apply(x, x)


@pytest.mark.parametrize('cin,cout,w,f', [(32, 32, 24, 4), (32, 64, 10, 2), (64, 64, 12, 2)])
def test_wgrad_pixel_fold_algebra(cin, cout, w, f, monkeypatch):
    """ops.conv_wgrad on pixel-folded views (low-channel layers) + b200ir_wgrad_unfold == the direct weight gradient, through the
    C-ABI simulator: checks the view arithmetic and the tap-block bookkeeping above the ABI."""
    from image_restoration_b200 import ops
    from tests import cabi_sim
    g = torch.Generator().manual_seed(cin + cout + w)
    B, h = 2, 5
    x = torch.randn(B, h, w, cin, generator=g).half()
    dy = torch.randn(B, h, w, cout, generator=g).half()
    with cabi_sim.installed():
        monkeypatch.setattr(ops, '_WGRAD_FOLD', 1)
        assert ops.wgrad_fold(cin, cout, w) == f
        folded = ops.conv_wgrad(x, dy)
        monkeypatch.setattr(ops, '_WGRAD_FOLD', 0)
        direct = ops.conv_wgrad(x, dy)
    assert folded.shape == direct.shape == (cout, 9, cin)
    assert (folded - direct).abs().max().item() <= 1e-4 * direct.abs().max().item()
