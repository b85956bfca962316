"""CPU tests of the SR-network row: the oracle restatement against the reference's golden outputs and
the reference modules themselves; the B200 classes' state_dict contract (names, shapes, order, seeded init) against
the fixtures' checksums and the reference classes."""
import glob
import json
import os

import numpy as np
import pytest
import torch

from oracle import ref_import
from oracle import sr_oracle
from tests.helpers import GOLDEN_DIR, state_checksum

FWD = {'MSRResNet': lambda sd, x, kw: sr_oracle.msrresnet_forward(sd, x, kw.get('upscale', 4)),
       'EDSR': lambda sd, x, kw: sr_oracle.edsr_forward(sd, x, kw.get('res_scale', 1)),
       'RCAN': lambda sd, x, kw: sr_oracle.rcan_forward(sd, x, kw.get('res_scale', 1)),
       'RRDBNet': lambda sd, x, kw: sr_oracle.rrdbnet_forward(sd, x, kw.get('scale', 4))}


def sr_golden_files():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, 'sr_*.npz')))


def load_sr_golden(path):
    from image_restoration_b200 import sr_archs
    fx = dict(np.load(path, allow_pickle=False))
    kw = json.loads(str(fx['kwargs']))
    torch.manual_seed(int(fx['seed']))
    net = getattr(sr_archs, str(fx['arch']))(**kw).eval()
    ok = state_checksum(net.state_dict()) == str(fx['checksum'])
    return fx, kw, net, ok


@pytest.mark.parametrize('path', sr_golden_files(), ids=lambda p: os.path.basename(p))
def test_seeded_init_and_oracle_match_reference_golden(path):
    fx, kw, net, ok = load_sr_golden(path)
    assert ok, 'same seed must give the reference\'s random-init weights (same RNG draw order and init functions)'
    y = FWD[str(fx['arch'])](net.state_dict(), torch.from_numpy(fx['x']), kw)
    ref = torch.from_numpy(fx['y'])
    assert y.shape == ref.shape
    assert (y - ref).abs().max().item() <= 1e-5 * max(1.0, ref.abs().max().item())


@pytest.mark.skipif(not ref_import.available(), reason='/root/reference not present')
@pytest.mark.parametrize('arch,mod,kw', [
    ('MSRResNet', 'srresnet_arch', dict(num_feat=32, num_block=2, upscale=2)),
    ('EDSR', 'edsr_arch', dict(num_in_ch=3, num_out_ch=3, num_feat=32, num_block=2, upscale=3, res_scale=0.5)),
    ('RCAN', 'rcan_arch', dict(num_in_ch=3, num_out_ch=3, num_feat=32, num_group=2, num_block=1, squeeze_factor=8, upscale=2)),
    ('RRDBNet', 'rrdbnet_arch', dict(num_in_ch=3, num_out_ch=3, scale=1, num_feat=32, num_block=1, num_grow_ch=16)),
])
def test_state_dict_contract_and_oracle_against_reference_modules(arch, mod, kw):
    import importlib
    from image_restoration_b200 import sr_archs
    ref_import.load_reference_arch()
    Ref = getattr(importlib.import_module('basicsr.archs.' + mod), arch)
    torch.manual_seed(21)
    ref = Ref(**kw).eval()
    torch.manual_seed(21)
    ours = getattr(sr_archs, arch)(**kw).eval()
    sd_r, sd_o = ref.state_dict(), ours.state_dict()
    assert list(sd_r.keys()) == list(sd_o.keys())
    for k in sd_r:
        assert sd_r[k].shape == sd_o[k].shape and torch.equal(sd_r[k], sd_o[k]), k
    ours.load_state_dict(sd_r, strict=True)
    g = torch.Generator().manual_seed(22)
    for k, v in sd_r.items():          # non-zero biases so every term is exercised
        if k.endswith('bias'):
            v.add_(torch.randn(v.shape, generator=g) * 0.05)
    ref.load_state_dict(sd_r)
    x = torch.rand(2, 3, 12, 20)
    with torch.no_grad():
        y = ref(x)
    y2 = FWD[arch](sd_r, x, kw)
    assert (y - y2).abs().max().item() <= 1e-5 * max(1.0, y.abs().max().item())


def test_sr_archs_registered_and_refuse_cpu():
    from image_restoration_b200 import sr_archs
    from image_restoration_b200.registry import ARCH_REGISTRY
    for name in ('MSRResNet_B200', 'EDSR_B200', 'RCAN_B200', 'RRDBNet_B200'):
        assert name in ARCH_REGISTRY
    net = ARCH_REGISTRY.get('MSRResNet_B200')(num_feat=16, num_block=1, upscale=2)
    with pytest.raises(RuntimeError):
        net(torch.rand(1, 3, 8, 8))
