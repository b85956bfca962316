"""CPU tests of the discriminator drop-in (network_d of the training YAMLs): parameter names / shapes / seeded
initialisation identical to the reference class, the oracle restatement against the reference forward and against the
committed reference outputs."""
import ast
import glob
import importlib
import os

import numpy as np
import pytest
import torch

from oracle import disc_oracle, ref_import

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), 'golden', 'disc_*.npz')))


def build(kw, seed):
    from image_restoration_b200.disc import StyleGAN2Discriminator
    torch.manual_seed(seed)
    return StyleGAN2Discriminator(**kw).eval()


def test_goldens_present():
    assert len(GOLD) == 3


@pytest.mark.parametrize('path', GOLD)
def test_oracle_reproduces_reference_scores_from_seeded_init(path):
    g = np.load(path)
    kw, seed, B = ast.literal_eval(str(g['kw'])), int(g['seed']), int(g['B'])
    net = build(kw, seed)                       # same RNG stream as the reference constructor
    assert sum(p.numel() for p in net.parameters()) == int(g['n_params'])
    x = torch.rand(B, 3, kw['input_height'], kw['input_width']) * 2 - 1
    y = disc_oracle.discriminator_forward(net.state_dict(), x, net.stddev_group)
    assert y.shape == (B, 1)
    assert np.abs(y.numpy() - g['score']).max() <= 2e-6 * max(1.0, np.abs(g['score']).max())


def test_registered_and_no_cpu_path():
    from image_restoration_b200.registry import ARCH_REGISTRY
    assert 'StyleGAN2Discriminator_B200' in ARCH_REGISTRY
    net = build(dict(input_width=64, input_height=64, channel_multiplier=1), 0)
    with pytest.raises(RuntimeError):
        net(torch.zeros(4, 3, 64, 64))
    with pytest.raises(ValueError):
        build(dict(input_width=64, input_height=64, resample_kernel=(1, 2, 1)), 0)


@pytest.mark.skipif(not ref_import.available(), reason='/root/reference not present')
@pytest.mark.parametrize('kw', [dict(input_width=384, input_height=128, channel_multiplier=1),
                                dict(input_width=64, input_height=64, channel_multiplier=2, narrow=0.5, stddev_group=2)])
def test_state_dict_and_forward_against_reference_class(kw):
    ref_import.load_reference_arch()
    m = importlib.import_module('basicsr.archs.stylegan2_arch')
    torch.manual_seed(5)
    ref = m.StyleGAN2Discriminator(**kw).eval()
    mine = build(kw, 5)
    rs, ms = ref.state_dict(), mine.state_dict()
    assert list(rs.keys()) == list(ms.keys())
    for k in rs:
        assert rs[k].shape == ms[k].shape and torch.equal(rs[k], ms[k]), k       # identical seeded initialisation
    mine.load_state_dict(rs, strict=True)
    x = torch.rand(4, 3, kw['input_height'], kw['input_width']) * 2 - 1
    with torch.no_grad():
        want = ref(x)
    got = disc_oracle.discriminator_forward(rs, x, kw.get('stddev_group', 4))
    assert (got - want).abs().max().item() <= 2e-6 * max(1.0, want.abs().max().item())
