"""GPU parity of the conv weight-gradient kernel (b200ir_conv_wgrad, tcgen05 GEMM over pixels with MN-major operands)
against torch's fp32 conv2d weight gradient computed from the same fp16-rounded operands.  Products of fp16 values are
exact in fp32 and the accumulation is fp32, so only the summation order differs: rel 1e-3 of the largest entry
(observed ~1e-6)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('B,H,W,cin,cout', [(2, 8, 32, 64, 128), (3, 16, 48, 128, 256), (1, 7, 45, 64, 128),
                                              (4, 32, 96, 256, 256), (2, 4, 12, 512, 512)])
def test_wgrad_matches_torch(B, H, W, cin, cout):
    from image_restoration_b200 import ops
    torch.manual_seed(B * 100 + H)
    x = torch.randn(B, cin, H, W, device='cuda').half()
    dy = torch.randn(B, cout, H, W, device='cuda').half()
    ref = torch.nn.grad.conv2d_weight(x.float(), (cout, cin, 3, 3), dy.float(), padding=1)     # [cout, cin, 3, 3]
    ref = ref.permute(0, 2, 3, 1).reshape(cout, 9, cin)
    got = ops.conv_wgrad(x.permute(0, 2, 3, 1).contiguous(), dy.permute(0, 2, 3, 1).contiguous())
    torch.cuda.synchronize()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    print(f'wgrad B{B} {H}x{W} {cin}->{cout}: max err {err:.3e} of {scale:.3e}')
    assert err <= 1e-3 * scale, (err, scale)


def test_wgrad_rejects_unsupported_channels():
    from image_restoration_b200 import ops
    x = torch.zeros(1, 8, 8, 32, device='cuda', dtype=torch.float16)
    dy = torch.zeros(1, 8, 8, 128, device='cuda', dtype=torch.float16)
    with pytest.raises(RuntimeError):
        ops.conv_wgrad(x, dy)
