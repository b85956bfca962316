"""TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline) -- never imported by the product.

CPU fp32 restatement of StyleGAN2Discriminator.forward (Car_Plate-Restoration/basicsr/archs/stylegan2_arch.py:735-805;
ConvLayer :658-703, ResBlock :706-732, EqualLinear :134-175), functional over a state_dict, built from the operator
restatements of oracle/gfpgan_ocr_oracle.py (the blocks are the ones of the restoration network's encoder).

Pinning: tests/test_disc_cpu.py compares it with the unmodified reference class imported from /root/reference (same
seeded weights and inputs) and with fixtures generated from the reference (tests/golden/disc_*.npz,
tests/golden/make_golden_disc.py).  The reference ships no test for this class.
"""
import math

import torch
import torch.nn.functional as F

from .gfpgan_ocr_oracle import conv_layer, fused_lrelu, res_block


def minibatch_stddev(out, stddev_group=4, stddev_feat=1):
    """stylegan2_arch.py:791-800: one extra channel holding the group standard deviation averaged over (c, h, w)."""
    b, c, h, w = out.shape
    group = min(b, stddev_group)
    s = out.view(group, -1, stddev_feat, c // stddev_feat, h, w)
    s = torch.sqrt(s.var(0, unbiased=False) + 1e-8)
    s = s.mean([2, 3, 4], keepdims=True).squeeze(2)
    s = s.repeat(group, 1, h, w)
    return torch.cat([out, s], 1)


def discriminator_forward(sd, x, stddev_group=4):
    """sd: state_dict of StyleGAN2Discriminator (fp32 CPU tensors); x (B,3,H,W) -> (B,1)."""
    sd = {k: v.float() for k, v in sd.items()}
    n_blocks = len({k.split('.')[1] for k in sd if k.startswith('conv_body.')}) - 1
    out = conv_layer(sd, 'conv_body.0', x.float(), 1, False, True, True)
    for i in range(1, n_blocks + 1):
        out = res_block(sd, f'conv_body.{i}', out)
    out = minibatch_stddev(out, stddev_group)
    out = conv_layer(sd, 'final_conv', out, 3, False, True, True)
    out = out.reshape(out.shape[0], -1)
    w0, w1 = sd['final_linear.0.weight'], sd['final_linear.1.weight']
    out = fused_lrelu(F.linear(out, w0 * (1.0 / math.sqrt(w0.shape[1]))), sd['final_linear.0.bias'])   # activation='fused_lrelu'
    return F.linear(out, w1 * (1.0 / math.sqrt(w1.shape[1])), sd['final_linear.1.bias'])
