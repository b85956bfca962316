"""GPU parity of the discriminator forward (through the C ABI) against the committed reference outputs and the oracle.
Operands are fp16 with fp32 accumulation; the scores of a random-init network are O(0.05), so the bar is relative to the
spread of the features: |score - ref| <= 2e-3 + 2e-2 * max|ref| (observed: a few 1e-4)."""
import ast
import glob
import os

import numpy as np
import pytest
import torch

from oracle import disc_oracle

pytestmark = pytest.mark.gpu
GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), 'golden', 'disc_*.npz')))


@pytest.mark.parametrize('path', GOLD)
def test_scores_match_reference_goldens(path):
    from image_restoration_b200.disc import StyleGAN2Discriminator
    g = np.load(path)
    kw, seed, B = ast.literal_eval(str(g['kw'])), int(g['seed']), int(g['B'])
    torch.manual_seed(seed)
    net = StyleGAN2Discriminator(**kw).eval()
    x = torch.rand(B, 3, kw['input_height'], kw['input_width']) * 2 - 1
    y = net.cuda()(x.cuda())
    torch.cuda.synchronize()
    assert y.shape == (B, 1) and y.dtype == x.dtype
    err = np.abs(y.cpu().numpy() - g['score']).max()
    print(os.path.basename(path), 'max err', err, 'max |ref|', np.abs(g['score']).max())
    assert err <= 2e-3 + 2e-2 * np.abs(g['score']).max(), err


def test_perturbed_weights_batches_and_stddev_groups():
    """Non-zero biases, batch sizes 1 / 2 / 8 / 12 (group = min(B, 4)) against the oracle."""
    from image_restoration_b200.disc import StyleGAN2Discriminator
    torch.manual_seed(11)
    net = StyleGAN2Discriminator(input_width=96, input_height=32, channel_multiplier=1).eval()
    with torch.no_grad():
        for n, p in net.named_parameters():
            if n.endswith('bias'):
                p.normal_(0, 0.2)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    net = net.cuda()
    for B in (1, 2, 8, 12):
        x = torch.rand(B, 3, 32, 96) * 2 - 1
        want = disc_oracle.discriminator_forward(sd, x, 4)
        got = net(x.cuda()).cpu()
        err = (got - want).abs().max().item()
        assert err <= 2e-3 + 2e-2 * want.abs().max().item(), (B, err)
    with pytest.raises(ValueError):
        net(torch.zeros(6, 3, 32, 96).cuda())        # 6 is not divisible by the stddev group 4 (the reference fails too)
