"""Drop-in GFPGANv1OCR for B200.

Mirrors the reference class `GFPGANv1OCR` (Car_Plate-Restoration/basicsr/archs/gfpganv1_ocr_arch.py:228-393):
same constructor keywords, same parameter / buffer names and shapes (so `load_state_dict(ckpt['params_ema'])`
works unchanged), same `forward(x, return_latents, save_feat_path, load_feat_path, return_rgb,
randomize_noise) -> (image, out_rgbs)`.  The modules below only *hold* parameters; all arithmetic of the forward
pass runs in libb200ir.so through `engine.OcrEngine` (hand-written sm_100a kernels, no torch fallback).

Parameter initialisation follows the reference's distributions and draws from torch's RNG in the same order as the
reference constructors, so `torch.manual_seed(s); GFPGANv1OCR(...)` yields the same random-init weights as the
reference does for seed s (checked in tests/test_state_dict_contract.py).
"""
import math
import weakref

import torch
from torch import nn

from .registry import ARCH_REGISTRY, USING_BASICSR_REGISTRY

_ENGINES = weakref.WeakKeyDictionary()  # module -> OcrEngine (kept out of __dict__ so deepcopy / pickling still work)


def _randn_param(*shape, div=None):
    t = torch.randn(*shape)
    if div is not None:
        t.div_(div)
    return nn.Parameter(t)


class _Bias(nn.Module):
    """Holds `.bias` (FusedLeakyReLU's only parameter, fused_act.py:81-91)."""

    def __init__(self, channels):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(channels))


class _Conv(nn.Module):
    """Holds `.weight` (+ optional `.bias`) of an EqualConv2d (stylegan2_ocr_arch.py:609-655)."""

    def __init__(self, cin, cout, k, bias=False, bias_init_val=0.0):
        super().__init__()
        self.weight = _randn_param(cout, cin, k, k)
        if bias:
            self.bias = nn.Parameter(torch.zeros(cout).fill_(bias_init_val))
        else:
            self.register_parameter('bias', None)


class _Linear(nn.Module):
    """EqualLinear parameters (stylegan2_ocr_arch.py:134-163)."""

    def __init__(self, cin, cout, bias_init_val=0.0, lr_mul=1.0):
        super().__init__()
        self.lr_mul = lr_mul
        self.weight = _randn_param(cout, cin, div=lr_mul)
        self.bias = nn.Parameter(torch.zeros(cout).fill_(bias_init_val))


class _Empty(nn.Module):
    """Parameter-free slot (UpFirDnSmooth / ScaledLeakyReLU / NormStyleCode positions inside a Sequential)."""


def _conv_layer(cin, cout, k, downsample=False, bias=True, activate=True):
    """Slots of ConvLayer(nn.Sequential) (stylegan2_ocr_arch.py:658-705): [smooth] conv [FusedLeakyReLU]."""
    mods = []
    if downsample:
        mods.append(_Empty())
    mods.append(_Conv(cin, cout, k, bias=bias and not activate))
    if activate:
        mods.append(_Bias(cout) if bias else _Empty())
    return nn.Sequential(*mods)


class _ResBlock(nn.Module):
    """stylegan2_ocr_arch.py:708-734."""

    def __init__(self, cin, cout):
        super().__init__()
        self.conv1 = _conv_layer(cin, cin, 3)
        self.conv2 = _conv_layer(cin, cout, 3, downsample=True)
        self.skip = _conv_layer(cin, cout, 1, downsample=True, bias=False, activate=False)


class _ConvUp(nn.Module):
    """ConvUpLayer parameters (gfpganv1_ocr_arch.py:139-186)."""

    def __init__(self, cin, cout, k, bias=True, activate=True):
        super().__init__()
        self.weight = _randn_param(cout, cin, k, k)
        self.register_parameter('bias', None)
        if activate and bias:
            self.activation = _Bias(cout)


class _ResUpBlock(nn.Module):
    """gfpganv1_ocr_arch.py:205-218."""

    def __init__(self, cin, cout):
        super().__init__()
        self.conv1 = _conv_layer(cin, cin, 3)
        self.conv2 = _ConvUp(cin, cout, 3)
        self.skip = _ConvUp(cin, cout, 1, bias=False, activate=False)


class _ModConv(nn.Module):
    """ModulatedConv2d parameters (stylegan2_ocr_arch.py:200-237): modulation is created before weight."""

    def __init__(self, cin, cout, k, num_style_feat):
        super().__init__()
        self.modulation = _Linear(num_style_feat, cin, bias_init_val=1.0)
        self.weight = _randn_param(1, cout, cin, k, k)


class _StyleConv(nn.Module):
    """stylegan2_ocr_arch.py:303-321."""

    def __init__(self, cin, cout, num_style_feat):
        super().__init__()
        self.modulated_conv = _ModConv(cin, cout, 3, num_style_feat)
        self.weight = nn.Parameter(torch.zeros(1))
        self.activate = _Bias(cout)


class _ToRGB(nn.Module):
    """stylegan2_ocr_arch.py:346-355."""

    def __init__(self, cin, num_style_feat):
        super().__init__()
        self.modulated_conv = _ModConv(cin, 3, 1, num_style_feat)
        self.bias = nn.Parameter(torch.zeros(1, 3, 1, 1))


class _StyleGANDecoder(nn.Module):
    """StyleGAN2OCRGeneratorSFT parameters (stylegan2_ocr_arch.py:408-497)."""

    def __init__(self, input_width, input_height, num_style_feat, num_mlp, channel_multiplier, lr_mlp, narrow):
        super().__init__()
        mlp = [_Empty()]
        for _ in range(num_mlp):
            mlp.append(_Linear(num_style_feat, num_style_feat, lr_mul=lr_mlp))
        self.style_mlp = nn.Sequential(*mlp)
        ch = _channels(narrow, channel_multiplier)
        ratio = int(input_width / input_height)
        self.constant_input = nn.Module()
        self.constant_input.weight = _randn_param(1, ch[4], 4, 4 * ratio)
        self.style_conv1 = _StyleConv(ch[4], ch[4], num_style_feat)
        self.to_rgb1 = _ToRGB(ch[4], num_style_feat)
        log_size = int(math.log(min(input_width, input_height), 2))
        num_layers = (log_size - 2) * 2 + 1
        self.style_convs = nn.ModuleList()
        self.to_rgbs = nn.ModuleList()
        self.noises = nn.Module()
        for layer_idx in range(num_layers):
            rh = 2 ** ((layer_idx + 5) // 2)
            self.noises.register_buffer(f'noise{layer_idx}', torch.randn(1, 1, rh, rh * ratio))
        cin = ch[4]
        for i in range(3, log_size + 1):
            cout = ch[2 ** i]
            self.style_convs.append(_StyleConv(cin, cout, num_style_feat))
            self.style_convs.append(_StyleConv(cout, cout, num_style_feat))
            self.to_rgbs.append(_ToRGB(cout, num_style_feat))
            cin = cout


def _channels(narrow, channel_multiplier):
    return {4: int(512 * narrow), 8: int(512 * narrow), 16: int(512 * narrow), 32: int(512 * narrow),
            64: int(256 * channel_multiplier * narrow), 128: int(128 * channel_multiplier * narrow),
            256: int(64 * channel_multiplier * narrow), 512: int(32 * channel_multiplier * narrow),
            1024: int(16 * channel_multiplier * narrow)}


class GFPGANv1OCR(nn.Module):
    """U-Net + StyleGAN2 decoder with SFT; B200-native forward.  See module docstring."""

    def __init__(self, input_width=768, input_height=32, num_style_feat=512, channel_multiplier=1,
                 resample_kernel=(1, 3, 3, 1), decoder_load_path=None, fix_decoder=True, num_mlp=8, lr_mlp=0.01,
                 input_is_latent=False, different_w=False, narrow=1, sft_half=False):
        super().__init__()
        if tuple(resample_kernel) != (1, 3, 3, 1):
            raise ValueError('only resample_kernel=(1,3,3,1) is implemented (the value every reference config uses)')
        self.input_width, self.input_height = input_width, input_height
        self.input_is_latent = input_is_latent
        self.different_w = different_w
        self.num_style_feat = num_style_feat
        self.sft_half = sft_half
        self.num_mlp = num_mlp
        self.channel_multiplier = channel_multiplier
        self.narrow = narrow
        out_size = min(input_width, input_height)
        self.log_size = int(math.log(out_size, 2))
        ch = _channels(narrow * 0.5, channel_multiplier)
        first = 2 ** self.log_size

        self.conv_body_first = _conv_layer(3, ch[first], 1)
        cin = ch[first]
        self.conv_body_down = nn.ModuleList()
        for i in range(self.log_size, 2, -1):
            cout = ch[2 ** (i - 1)]
            self.conv_body_down.append(_ResBlock(cin, cout))
            cin = cout
        self.final_conv = _conv_layer(cin, ch[4], 3)
        cin = ch[4]
        self.conv_body_up = nn.ModuleList()
        for i in range(3, self.log_size + 1):
            cout = ch[2 ** i]
            self.conv_body_up.append(_ResUpBlock(cin, cout))
            cin = cout
        self.toRGB = nn.ModuleList()
        for i in range(3, self.log_size + 1):
            self.toRGB.append(_Conv(ch[2 ** i], 3, 1, bias=True))
        linear_out = (self.log_size * 2 - 2) * num_style_feat if different_w else num_style_feat
        self.final_linear = _Linear(ch[4] * 4 * 4 * int(input_width / input_height), linear_out)
        self.stylegan_decoder = _StyleGANDecoder(input_width, input_height, num_style_feat, num_mlp,
                                                 channel_multiplier, lr_mlp, narrow)
        if decoder_load_path:
            self.stylegan_decoder.load_state_dict(
                torch.load(decoder_load_path, map_location=lambda storage, loc: storage)['params_ema'])
        if fix_decoder:
            for _, p in self.stylegan_decoder.named_parameters():
                p.requires_grad = False
        self.condition_scale = nn.ModuleList()
        self.condition_shift = nn.ModuleList()
        for i in range(3, self.log_size + 1):
            c = ch[2 ** i]
            c_sft = c if sft_half else c * 2
            self.condition_scale.append(nn.Sequential(_Conv(c, c, 3, bias=True), _Empty(),
                                                      _Conv(c, c_sft, 3, bias=True, bias_init_val=1.0)))
            self.condition_shift.append(nn.Sequential(_Conv(c, c, 3, bias=True), _Empty(),
                                                      _Conv(c, c_sft, 3, bias=True, bias_init_val=0.0)))

    # ------------------------------------------------------------------ engine plumbing
    def engine(self):
        from .engine import OcrEngine
        eng = _ENGINES.get(self)
        if eng is None or eng.stale():
            eng = OcrEngine(self)
            _ENGINES[self] = eng
        return eng

    def invalidate_engine(self):
        """Drop packed weights / plans (call after changing parameters in place without load_state_dict)."""
        _ENGINES.pop(self, None)

    def _apply(self, fn, *a, **kw):
        _ENGINES.pop(self, None)  # device / dtype moved: packed weights are stale
        return super()._apply(fn, *a, **kw)

    def load_state_dict(self, *a, **kw):
        _ENGINES.pop(self, None)
        return super().load_state_dict(*a, **kw)

    def forward(self, x, return_latents=False, save_feat_path=None, load_feat_path=None, return_rgb=True,
                randomize_noise=True):
        """Same contract as gfpganv1_ocr_arch.py:341-393.  `x`: (B,3,H,W) float CUDA tensor."""
        if not x.is_cuda:
            raise RuntimeError('image_restoration_b200.GFPGANv1OCR runs on a CUDA B200 only; there is no CPU path '
                               '(move the module and the input to cuda)')
        if self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            # training call (GFPGANModel.optimize_parameters, gfpgan_model.py:508): the differentiable path -- U-Net through
            # the autograd Functions of backward.py, StyleGAN2 decoder through train.DecoderFunction.  The inference
            # engine below runs under no_grad and would hand back tensors without history.
            if save_feat_path is not None or load_feat_path is not None:
                raise NotImplementedError('save_feat_path / load_feat_path are inference options (call .eval() first)')
            from .train import train_forward
            return train_forward(self, x, return_rgb=return_rgb, randomize_noise=randomize_noise)
        image, out_rgbs = self.engine().forward(x, return_rgb=return_rgb, randomize_noise=randomize_noise,
                                                save_feat_path=save_feat_path, load_feat_path=load_feat_path)
        return image, out_rgbs


    def restore_uint8(self, img, bgr=True, randomize_noise=True):
        """The whole `restoration()` body of api.py:92-117 for a batch already at the network size: uint8 HWC images
        (B,H,W,3) on the device -> restored uint8 HWC images.  img2tensor + normalize, forward(return_rgb=False) and
        tensor2img(min_max=(-1,1)) all run on the device."""
        if not img.is_cuda:
            raise RuntimeError('restore_uint8 takes a CUDA uint8 tensor (use host_io.HostPipeline for host images)')
        return self.engine().forward(img, return_rgb=False, randomize_noise=randomize_noise, uint8_io=True, bgr=bgr)[0]


def register_into(registry=None, name='GFPGANv1OCR_B200', override=False):
    """Registers the B200 class in a basicsr-style Registry (basicsr/utils/registry.py:4-82).

    `override=False`: adds it under `name` (select it with `network_g.type: GFPGANv1OCR_B200` in the YAML).
    `override=True`: replaces the entry `GFPGANv1OCR` so unmodified configs pick the B200 class; Registry asserts on
    duplicate names (registry.py:38-41), so the map entry is replaced directly."""
    registry = registry if registry is not None else ARCH_REGISTRY
    if override:
        registry._obj_map['GFPGANv1OCR'] = GFPGANv1OCR
        return GFPGANv1OCR
    cls = type(name, (GFPGANv1OCR,), {'__doc__': GFPGANv1OCR.__doc__})
    if name not in registry:
        registry.register(cls)
    return registry.get(name)


register_into(ARCH_REGISTRY, 'GFPGANv1OCR_B200')
if not USING_BASICSR_REGISTRY and 'GFPGANv1OCR' not in ARCH_REGISTRY:
    ARCH_REGISTRY.register(GFPGANv1OCR)  # stand-alone use: the reference name resolves to the B200 class
