"""CPU check of the folded ConvUpLayer algebra (ops.upfold_weights): conv3x3(bilinear_up2(t)) == per-phase 3x3 convs over
the replicate-padded low-resolution tensor minus the border-ring surplus (four 1-D convs, corner terms added back)."""
import torch
import torch.nn.functional as F

from image_restoration_b200 import ops


def test_folded_conv_up_layer_equals_reference_composition():
    torch.manual_seed(0)
    B, Cin, Cout, h, w = 2, 16, 16, 5, 6
    t = torch.randn(B, Cin, h, w)
    W = torch.randn(Cout, Cin, 3, 3)
    fw = ops.upfold_weights(W, 1.0)
    W16 = W.half().float()
    ref = F.conv2d(F.interpolate(t, scale_factor=2, mode='bilinear', align_corners=False), W, padding=1)
    tp = F.pad(t, (1, 1, 1, 1), mode='replicate')
    main = fw['main'].float().view(2, 2, Cout, 3, 3, Cin).permute(0, 1, 2, 5, 3, 4)      # [py][px][co][ci][dy][dx]
    out = torch.zeros_like(ref)
    for py in range(2):
        for px in range(2):
            out[:, :, py::2, px::2] = F.conv2d(tp, main[py, px])
    tol = 2e-3 * ref.abs().max().item()                                                   # fp16 rounding of the folded weights
    assert (out - ref)[:, :, 1:-1, 1:-1].abs().max().item() <= tol                        # interior: exact up to rounding
    assert (out - ref).abs().max().item() > 10 * tol                                      # the ring differs before correction

    def rowconv(rowpad, wgt):      # (B, Cin, n+2), [2*Cout][3*Cin] -> (B, 2n, Cout)
        k = wgt.float().view(2, Cout, 3, Cin).permute(0, 1, 3, 2).reshape(2 * Cout, Cin, 3)
        o = F.conv1d(rowpad, k)
        n = o.shape[2]
        return o.view(B, 2, Cout, n).permute(0, 3, 1, 2).reshape(B, 2 * n, Cout)
    top, bot = rowconv(tp[:, :, 1, :], fw['top']), rowconv(tp[:, :, h, :], fw['bot'])
    left, right = rowconv(tp[:, :, :, 1], fw['left']), rowconv(tp[:, :, :, w], fw['right'])
    wc = fw['corners']
    assert torch.allclose(wc[1], W[:, :, 0, 2]) and wc.dtype == torch.float32
    top[:, 0] -= torch.einsum('oc,bc->bo', wc[0], t[:, :, 0, 0])
    top[:, -1] -= torch.einsum('oc,bc->bo', wc[1], t[:, :, 0, -1])
    bot[:, 0] -= torch.einsum('oc,bc->bo', wc[2], t[:, :, -1, 0])
    bot[:, -1] -= torch.einsum('oc,bc->bo', wc[3], t[:, :, -1, -1])
    corr = torch.zeros_like(ref)
    corr[:, :, 0, :] += top.permute(0, 2, 1)
    corr[:, :, -1, :] += bot.permute(0, 2, 1)
    corr[:, :, :, 0] += left.permute(0, 2, 1)
    corr[:, :, :, -1] += right.permute(0, 2, 1)
    assert (out - corr - ref).abs().max().item() <= tol
    del W16
