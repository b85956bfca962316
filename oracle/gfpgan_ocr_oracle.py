"""ORACLE (test infrastructure, NOT product code).

CPU fp32 restatement of the reference's GFPGANv1OCR forward pass, written as
pure functions over a ``state_dict`` so it can travel to the GPU box where
``/root/reference`` does not exist.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import this.

Parity status: the reference ships no tests or golden vectors,
so this restatement is pinned against the reference's own modules executed in
the build container (``tests/test_oracle_vs_reference.py``, skipped where
``/root/reference`` is absent) and against fixtures generated from that run
(``tests/golden/``, made by ``tests/golden/make_golden.py``).

Every function cites the reference file:line it follows.  Shorthand:
  G = Car_Plate-Restoration/basicsr/archs/gfpganv1_ocr_arch.py
  S = Car_Plate-Restoration/basicsr/archs/stylegan2_ocr_arch.py
  U = Car_Plate-Restoration/basicsr/ops/upfirdn2d/upfirdn2d.py
  A = Car_Plate-Restoration/basicsr/ops/fused_act/src/fused_bias_act_kernel.cu
"""
import math

import torch
import torch.nn.functional as F

SQRT2 = math.sqrt(2.0)


# --------------------------------------------------------------------------- config
class OcrNetConfig:
    """Shape bookkeeping of GFPGANv1OCR.__init__ (G:232-339) and
    StyleGAN2OCRGenerator.__init__ (S:408-497)."""

    def __init__(self, input_width=768, input_height=32, num_style_feat=512, channel_multiplier=1,
                 num_mlp=8, input_is_latent=False, different_w=False, narrow=1, sft_half=False):
        self.input_width, self.input_height = input_width, input_height
        self.num_style_feat = num_style_feat
        self.num_mlp = num_mlp
        self.input_is_latent = input_is_latent
        self.different_w = different_w
        self.sft_half = sft_half
        out_size = min(input_width, input_height)
        self.log_size = int(math.log(out_size, 2))           # G:267
        self.ratio = int(input_width / input_height)         # S:466, G:301
        un = narrow * 0.5                                    # G:253

        def table(n):
            return {4: int(512 * n), 8: int(512 * n), 16: int(512 * n), 32: int(512 * n),
                    64: int(256 * channel_multiplier * n), 128: int(128 * channel_multiplier * n),
                    256: int(64 * channel_multiplier * n), 512: int(32 * channel_multiplier * n),
                    1024: int(16 * channel_multiplier * n)}
        self.unet_ch = table(un)                             # G:254-264
        self.dec_ch = table(narrow)                          # S:432-442
        self.num_levels = self.log_size - 2
        self.num_layers = self.num_levels * 2 + 1            # S:457
        self.num_latent = self.log_size * 2 - 2              # S:458


# --------------------------------------------------------------------------- primitives
def fused_lrelu(x, bias=None, slope=0.2, scale=SQRT2):
    """A:27-48 (act=3, grad=0): y = lrelu(x + b_c) * scale; fused_act.py:81-95."""
    if bias is not None:
        x = x + bias.view(1, -1, *([1] * (x.dim() - 2)))
    return F.leaky_relu(x, slope) * scale


def fir_kernel():
    """S:26-40 make_resample_kernel((1,3,3,1))."""
    k = torch.tensor([1.0, 3.0, 3.0, 1.0])
    k = k[None, :] * k[:, None]
    return k / k.sum()


def upfirdn(x, kernel, up=1, down=1, pad=(0, 0)):
    """U:162-192 upfirdn2d_native: zero-stuff, zero-pad, correlate with the flipped
    kernel, decimate.  (Only non-negative pads occur on this path.)"""
    b, c, h, w = x.shape
    if up > 1:
        z = x.new_zeros(b, c, h, up, w, up)
        z[:, :, :, 0, :, 0] = x
        x = z.view(b, c, h * up, w * up)
    x = F.pad(x, [pad[0], pad[1], pad[0], pad[1]])
    kf = torch.flip(kernel, [0, 1]).to(x)[None, None]
    y = F.conv2d(x.reshape(b * c, 1, x.shape[2], x.shape[3]), kf)
    y = y.view(b, c, y.shape[2], y.shape[3])
    return y[:, :, ::down, ::down]


def equal_conv(x, w, bias=None, stride=1, padding=0):
    """S:639-648 EqualConv2d.forward."""
    scale = 1.0 / math.sqrt(w.shape[1] * w.shape[2] * w.shape[3])
    return F.conv2d(x, w * scale, bias=bias, stride=stride, padding=padding)


def equal_linear(x, w, bias, lr_mul=1.0):
    """S:165-175 EqualLinear.forward (activation=None)."""
    scale = (1.0 / math.sqrt(w.shape[1])) * lr_mul
    return F.linear(x, w * scale, bias=None if bias is None else bias * lr_mul)


def conv_layer(sd, pre, x, k, downsample, bias, activate):
    """S:658-705 ConvLayer(nn.Sequential): [smooth] + EqualConv2d + [FusedLeakyReLU | ScaledLeakyReLU].
    Index of the conv inside the Sequential is 1 when a smooth layer precedes it."""
    ci = 1 if downsample else 0
    if downsample:
        p = (4 - 2) + (k - 1)                                # S:116-121
        x = upfirdn(x, fir_kernel(), pad=((p + 1) // 2, p // 2))
        stride, padding = 2, 0
    else:
        stride, padding = 1, k // 2
    conv_bias = sd.get(f'{pre}.{ci}.bias') if (bias and not activate) else None
    y = equal_conv(x, sd[f'{pre}.{ci}.weight'], conv_bias, stride, padding)
    if activate:
        y = fused_lrelu(y, sd[f'{pre}.{ci + 1}.bias'] if bias else None)
    return y


def res_block(sd, pre, x):
    """S:708-734 ResBlock.forward."""
    out = conv_layer(sd, f'{pre}.conv1', x, 3, False, True, True)
    out = conv_layer(sd, f'{pre}.conv2', out, 3, True, True, True)
    skip = conv_layer(sd, f'{pre}.skip', x, 1, True, False, False)
    return (out + skip) / SQRT2


def conv_up_layer(x, w, act_bias, padding, activate):
    """G:188-202 ConvUpLayer.forward: bilinear x2 then conv (then FusedLeakyReLU)."""
    x = F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=False)
    y = equal_conv(x, w, None, 1, padding)
    return fused_lrelu(y, act_bias) if activate else y


def res_up_block(sd, pre, x):
    """G:205-225 ResUpBlock.forward."""
    out = conv_layer(sd, f'{pre}.conv1', x, 3, False, True, True)
    out = conv_up_layer(out, sd[f'{pre}.conv2.weight'], sd[f'{pre}.conv2.activation.bias'], 1, True)
    skip = conv_up_layer(x, sd[f'{pre}.skip.weight'], None, 0, False)
    return (out + skip) / SQRT2


def sft_head(sd, pre, x):
    """G:322-339: EqualConv2d 3x3 -> ScaledLeakyReLU -> EqualConv2d 3x3."""
    y = equal_conv(x, sd[f'{pre}.0.weight'], sd[f'{pre}.0.bias'], 1, 1)
    y = fused_lrelu(y, None)                                 # S:604-606 ScaledLeakyReLU
    return equal_conv(y, sd[f'{pre}.2.weight'], sd[f'{pre}.2.bias'], 1, 1)


def modulated_conv(sd, pre, x, style, demodulate, upsample):
    """S:239-279 ModulatedConv2d.forward, restated per sample without the groups=b trick."""
    w = sd[f'{pre}.weight'][0]                               # (cout, cin, k, k)
    cout, cin, k, _ = w.shape
    s = equal_linear(style, sd[f'{pre}.modulation.weight'], sd[f'{pre}.modulation.bias'])  # (b, cin)
    scale = 1.0 / math.sqrt(cin * k * k)
    outs = []
    for bi in range(x.shape[0]):
        wb = scale * w * s[bi].view(1, cin, 1, 1)
        if demodulate:
            wb = wb * torch.rsqrt(wb.pow(2).sum([1, 2, 3]) + 1e-8).view(cout, 1, 1, 1)
        xb = x[bi:bi + 1]
        if upsample:
            yb = F.conv_transpose2d(xb, wb.transpose(0, 1), stride=2, padding=0)
            p = (4 - 2) - (k - 1)                            # S:112-115
            yb = upfirdn(yb, fir_kernel() * 4, pad=((p + 1) // 2 + 1, p // 2 + 1))
        else:
            yb = F.conv2d(xb, wb, padding=k // 2)
        outs.append(yb)
    return torch.cat(outs, 0)


def style_conv(sd, pre, x, style, noise, upsample):
    """S:323-333 StyleConv.forward (noise must be given: the oracle is deterministic)."""
    out = modulated_conv(sd, f'{pre}.modulated_conv', x, style, True, upsample)
    out = out + sd[f'{pre}.weight'] * noise
    return fused_lrelu(out, sd[f'{pre}.activate.bias'])


def to_rgb(sd, pre, x, style, skip, upsample):
    """S:357-374 ToRGB.forward; skip upsample = UpFirDnUpsample S:43-69 (pad (2,1), kernel*4)."""
    out = modulated_conv(sd, f'{pre}.modulated_conv', x, style, False, False) + sd[f'{pre}.bias']
    if skip is not None:
        if upsample:
            skip = upfirdn(skip, fir_kernel() * 4, up=2, pad=(2, 1))
        out = out + skip
    return out


def style_mlp(sd, cfg, z):
    """S:12-23 NormStyleCode + S:424-430 style MLP (EqualLinear lr_mul 0.01 + fused lrelu)."""
    z = z * torch.rsqrt(torch.mean(z ** 2, dim=1, keepdim=True) + 1e-8)
    for i in range(1, cfg.num_mlp + 1):
        w = sd[f'stylegan_decoder.style_mlp.{i}.weight']
        z = F.linear(z, w * ((1.0 / math.sqrt(w.shape[1])) * 0.01))
        z = fused_lrelu(z, sd[f'stylegan_decoder.style_mlp.{i}.bias'] * 0.01)
    return z


# --------------------------------------------------------------------------- whole net
def stylegan_decoder(sd, cfg, style_code, conditions, noises):
    """G:50-136 StyleGAN2OCRGeneratorSFT.forward with one style entry (as G:387-391 calls it)."""
    D = 'stylegan_decoder'
    styles = style_code if cfg.input_is_latent else style_mlp(sd, cfg, style_code)
    latent = styles if styles.ndim == 3 else styles.unsqueeze(1).repeat(1, cfg.num_latent, 1)   # G:92-99
    b = latent.shape[0]
    out = sd[f'{D}.constant_input.weight'].repeat(b, 1, 1, 1)                                # S:389-391
    out = style_conv(sd, f'{D}.style_conv1', out, latent[:, 0], noises[0], False)
    skip = to_rgb(sd, f'{D}.to_rgb1', out, latent[:, 1], None, False)
    i = 1
    for lvl in range(cfg.num_levels):                                                          # G:112-129
        out = style_conv(sd, f'{D}.style_convs.{2 * lvl}', out, latent[:, i], noises[2 * lvl + 1], True)
        if i < len(conditions):
            if cfg.sft_half:
                half = out.shape[1] // 2
                out = torch.cat([out[:, :half], out[:, half:] * conditions[i - 1] + conditions[i]], 1)
            else:
                out = out * conditions[i - 1] + conditions[i]
        out = style_conv(sd, f'{D}.style_convs.{2 * lvl + 1}', out, latent[:, i + 1], noises[2 * lvl + 2], False)
        skip = to_rgb(sd, f'{D}.to_rgbs.{lvl}', out, latent[:, i + 2], skip, True)
        i += 2
    return skip


def stored_noises(sd, cfg):
    """randomize_noise=False path, G:82-83."""
    return [sd[f'stylegan_decoder.noises.noise{i}'] for i in range(cfg.num_layers)]


@torch.no_grad()
def gfpgan_ocr_forward(sd, cfg, x, return_rgb=True, noises=None, taps=None):
    """G:341-393 GFPGANv1OCR.forward.  `noises`: list of (b|1,1,h,w) tensors, default = stored buffers.
    `taps` (optional dict) receives named intermediates for per-stage parity tests."""
    sd = {k: v.float() for k, v in sd.items()}
    x = x.float()
    if noises is None:
        noises = stored_noises(sd, cfg)
    L = cfg.num_levels
    feat = conv_layer(sd, 'conv_body_first', x, 1, False, True, True)                         # G:353
    if taps is not None:
        taps['first'] = feat
    skips = []
    for i in range(L):                                                                         # G:354-356
        feat = res_block(sd, f'conv_body_down.{i}', feat)
        skips.insert(0, feat)
        if taps is not None:
            taps[f'down{i}'] = feat
    feat = conv_layer(sd, 'final_conv', feat, 3, False, True, True)                           # G:358
    style_code = equal_linear(feat.reshape(feat.shape[0], -1), sd['final_linear.weight'], sd['final_linear.bias'])
    if cfg.different_w:
        style_code = style_code.view(style_code.shape[0], -1, cfg.num_style_feat)             # G:361-363
    if taps is not None:
        taps['final_conv'] = feat
        taps['style_code'] = style_code
    conditions, out_rgbs = [], []
    for i in range(L):                                                                         # G:366-378
        feat = feat + skips[i]
        feat = res_up_block(sd, f'conv_body_up.{i}', feat)
        conditions.append(sft_head(sd, f'condition_scale.{i}', feat))
        conditions.append(sft_head(sd, f'condition_shift.{i}', feat))
        if return_rgb:
            out_rgbs.append(equal_conv(feat, sd[f'toRGB.{i}.weight'], sd[f'toRGB.{i}.bias'], 1, 0))
        if taps is not None:
            taps[f'up{i}'] = feat
            taps[f'scale{i}'] = conditions[-2]
            taps[f'shift{i}'] = conditions[-1]
    image = stylegan_decoder(sd, cfg, style_code, conditions, noises)                         # G:387-391
    return image, out_rgbs


# --------------------------------------------------------------------------- comparison protocol
def to01(y):
    """tensor2img(min_max=(-1,1)) range mapping, basicsr/utils/img_util.py:66-67."""
    return (y.float().clamp(-1, 1) + 1) / 2


def psnr01(a, b):
    """psnr_ssim.py:47 on a [0,1] scale: 10*log10(1/mse)."""
    mse = torch.mean((a.double() - b.double()) ** 2).item()
    return float('inf') if mse == 0 else 10.0 * math.log10(1.0 / mse)
