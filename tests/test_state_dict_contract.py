"""Drop-in boundary: parameter / buffer names, shapes, order, init values, registry behaviour.
CPU only: constructing the module and moving state dicts needs no GPU; forward() must refuse to run on CPU."""
import copy

import pytest
import torch

from oracle import ref_import
from tests.helpers import KW


def make(W=384, H=128, **over):
    from image_restoration_b200 import GFPGANv1OCR
    return GFPGANv1OCR(input_width=W, input_height=H, decoder_load_path=None, fix_decoder=True, **dict(KW, **over))


def test_key_count_and_param_totals():
    net = make()
    sd = net.state_dict()
    assert len(sd) == 205                                              # SURVEY App. B [probed]
    assert sum(p.numel() for p in net.parameters()) == 73498700
    assert sum(b.numel() for b in net.buffers()) == 130992
    assert sd['final_linear.weight'].shape == (3072, 12288)
    assert sd['stylegan_decoder.constant_input.weight'].shape == (1, 512, 4, 12)
    assert sd['stylegan_decoder.noises.noise10'].shape == (1, 1, 128, 384)
    assert sd['conv_body_down.0.conv2.1.weight'].shape == (64, 32, 3, 3)
    assert sd['stylegan_decoder.to_rgbs.4.bias'].shape == (1, 3, 1, 1)
    assert len(make(256, 256).state_dict()) == 241


def test_init_values_that_matter_for_parity():
    sd = make(48, 16).state_dict()
    assert torch.all(sd['condition_scale.0.2.bias'] == 1) and torch.all(sd['condition_shift.0.2.bias'] == 0)
    assert torch.all(sd['stylegan_decoder.style_conv1.modulated_conv.modulation.bias'] == 1)
    assert sd['stylegan_decoder.style_conv1.weight'].item() == 0
    assert torch.all(sd['stylegan_decoder.to_rgb1.bias'] == 0)
    # EqualLinear with lr_mul 0.01 stores weights divided by lr_mul (stylegan2_ocr_arch.py:156)
    assert sd['stylegan_decoder.style_mlp.1.weight'].std().item() > 50


def test_fix_decoder_and_modes():
    net = make(48, 16)
    assert all(not p.requires_grad for p in net.stylegan_decoder.parameters())
    assert all(p.requires_grad for n, p in net.named_parameters() if not n.startswith('stylegan_decoder'))
    net.train().eval()
    net2 = copy.deepcopy(net)
    assert net2.state_dict().keys() == net.state_dict().keys()
    missing = net2.load_state_dict(net.state_dict(), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys


def test_forward_refuses_cpu():
    net = make(48, 16).eval()
    with pytest.raises(RuntimeError, match='CUDA'):
        net(torch.zeros(1, 3, 16, 48))


def test_registry_api():
    from image_restoration_b200 import ARCH_REGISTRY, GFPGANv1OCR, Registry, build_network, register_into
    assert 'GFPGANv1OCR_B200' in ARCH_REGISTRY
    assert issubclass(ARCH_REGISTRY.get('GFPGANv1OCR_B200'), GFPGANv1OCR)
    net = build_network(dict(type='GFPGANv1OCR_B200', input_width=48, input_height=16, **KW))
    assert isinstance(net, GFPGANv1OCR)
    with pytest.raises(KeyError):
        ARCH_REGISTRY.get('nope')
    r = Registry('arch')

    @r.register()
    class GFPGANv1OCR_dummy:  # noqa: N801
        pass
    with pytest.raises(AssertionError):          # duplicate names assert (registry.py:38-41)
        r.register(GFPGANv1OCR_dummy)
    r._obj_map['GFPGANv1OCR'] = object
    assert register_into(r, override=True) is GFPGANv1OCR and r.get('GFPGANv1OCR') is GFPGANv1OCR
    assert 'GFPGANv1OCR_B200' in [k for k, _ in r] or register_into(r) is not None


@pytest.mark.skipif(not ref_import.available(), reason='/root/reference not present')
@pytest.mark.parametrize('W,H,over', [(384, 128, {}), (256, 256, {}), (256, 64, {}),
                                      (64, 32, dict(sft_half=False, different_w=False, input_is_latent=False,
                                                    num_mlp=8, num_style_feat=512, channel_multiplier=1))])
def test_same_keys_shapes_order_and_seeded_init_as_reference(W, H, over):
    Ref, _ = ref_import.load_reference_arch()
    kw = dict(KW, **over)
    torch.manual_seed(3)
    ref = Ref(input_width=W, input_height=H, decoder_load_path=None, fix_decoder=True, **kw)
    torch.manual_seed(3)
    ours = make(W, H, **over)
    a, b = ref.state_dict(), ours.state_dict()
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert a[k].shape == b[k].shape, k
        assert torch.equal(a[k], b[k]), k           # same RNG draw order => identical random init
    assert [n for n, _ in ref.named_parameters()] == [n for n, _ in ours.named_parameters()]
    assert [p.requires_grad for p in ref.parameters()] == [p.requires_grad for p in ours.parameters()]
    # both directions of a strict load
    ours.load_state_dict(a, strict=True)
    ref.load_state_dict(b, strict=True)
