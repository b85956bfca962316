"""CPU tests (no GPU): the oracle restatement against (a) the golden vectors generated from the unmodified
reference and (b) the reference itself when /root/reference is present (build container only)."""
import os

import pytest
import torch

from oracle import ref_import
from oracle.gfpgan_ocr_oracle import OcrNetConfig, gfpgan_ocr_forward
from tests.helpers import KW, golden_files, load_golden


@pytest.mark.parametrize('path', golden_files(), ids=lambda p: os.path.basename(p))
def test_oracle_matches_golden(path):
    if '384x128' in path and os.environ.get('B200IR_FAST_TESTS'):
        pytest.skip('fast mode')
    fx, net = load_golden(path)
    if net is None:
        pytest.skip('seeded weights differ from the fixture (different torch RNG stream)')
    cfg = OcrNetConfig(input_width=int(fx['W']), input_height=int(fx['H']), **KW)
    x = torch.from_numpy(fx['x'])
    y, rgbs = gfpgan_ocr_forward(net.state_dict(), cfg, x, True)
    ref = torch.from_numpy(fx['image'])
    scale = ref.abs().max().item()
    assert (y - ref).abs().max().item() <= 1e-4 * scale          # fp32 reassociation only
    for i, r in enumerate(rgbs):
        rr = torch.from_numpy(fx[f'rgb{i}'])
        assert (r - rr).abs().max().item() <= 1e-4 * (rr.abs().max().item() + 1e-6)


@pytest.mark.skipif(not ref_import.available(), reason='/root/reference not present')
@pytest.mark.parametrize('W,H,over', [(48, 16, {}), (32, 32, dict(sft_half=False))])
def test_oracle_matches_reference_modules(W, H, over):
    Ref, _ = ref_import.load_reference_arch()
    kw = dict(KW, **over)
    torch.manual_seed(11)
    net = Ref(input_width=W, input_height=H, decoder_load_path=None, fix_decoder=True, **kw).eval()
    sd = net.state_dict()
    g = torch.Generator().manual_seed(12)
    for k, v in sd.items():   # non-trivial biases / noise gains so every term of the restatement is exercised
        if k.endswith('bias') or (k.endswith('.weight') and v.numel() == 1):
            v.add_(torch.randn(v.shape, generator=g) * 0.3)
    net.load_state_dict(sd)
    x = torch.rand(2, 3, H, W) * 2 - 1
    with torch.no_grad():
        y, rgbs = net(x, return_rgb=True, randomize_noise=False)
    y2, rgbs2 = gfpgan_ocr_forward(sd, OcrNetConfig(input_width=W, input_height=H, **kw), x, True)
    assert (y - y2).abs().max().item() <= 1e-4 * y.abs().max().item()
    for a, b in zip(rgbs, rgbs2):
        assert (a - b).abs().max().item() <= 1e-4 * (a.abs().max().item() + 1e-6)
