"""ORACLE (test infrastructure, NOT product code): functional torch-fp32 restatement of the plain-conv SR networks the
reference's options/*.yml select.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import this.

Follows, line by line:
  MSRResNet.forward   Car_Plate-Restoration/basicsr/archs/srresnet_arch.py:55-68
  EDSR.forward        .../edsr_arch.py:59-72
  RCAN.forward        .../rcan_arch.py:122-135 (ChannelAttention :8-24, RCAB :27-45, ResidualGroup :48-66)
  RRDBNet.forward     .../rrdbnet_arch.py:105-123 (ResidualDenseBlock :9-39, RRDB :42-63), pixel_unshuffle arch_util.py:185-201
  ResidualBlockNoBN   .../arch_util.py:66-93 ; Upsample :96-109

Pinned: against the unmodified reference modules imported in the build container (tests/test_sr_cpu.py, max rel diff
1e-5) and against fixtures generated from them (tests/golden/sr_*.npz, script tests/golden/make_golden_sr.py).  The
reference ships no tests or golden vectors of its own for these networks.
"""
import math

import torch
import torch.nn.functional as F


def _conv(sd, key, x):
    return F.conv2d(x, sd[key + '.weight'], sd.get(key + '.bias'), 1, 1)


def _res_block(sd, p, x, res_scale):                    # arch_util.py:90-93
    out = _conv(sd, p + '.conv2', F.relu(_conv(sd, p + '.conv1', x)))
    return x + out * res_scale


def _count(sd, prefix):
    return len({k[len(prefix):].split('.')[0] for k in sd if k.startswith(prefix)})


def msrresnet_forward(sd, x, upscale=4):
    feat = F.leaky_relu(_conv(sd, 'conv_first', x), 0.1)
    out = feat
    for i in range(_count(sd, 'body.')):
        out = _res_block(sd, f'body.{i}', out, 1)
    if upscale == 4:
        out = F.leaky_relu(F.pixel_shuffle(_conv(sd, 'upconv1', out), 2), 0.1)
        out = F.leaky_relu(F.pixel_shuffle(_conv(sd, 'upconv2', out), 2), 0.1)
    else:
        out = F.leaky_relu(F.pixel_shuffle(_conv(sd, 'upconv1', out), upscale), 0.1)
    out = _conv(sd, 'conv_last', F.leaky_relu(_conv(sd, 'conv_hr', out), 0.1))
    base = F.interpolate(x, scale_factor=upscale, mode='bilinear', align_corners=False)
    return out + base


def _upsample(sd, x):                                   # arch_util.py:96-109
    for i in sorted({int(k.split('.')[1]) for k in sd if k.startswith('upsample.')}):
        w = sd[f'upsample.{i}.weight']
        r = int(round(math.sqrt(w.shape[0] // w.shape[1])))
        x = F.pixel_shuffle(_conv(sd, f'upsample.{i}', x), r)
    return x


def edsr_forward(sd, x, res_scale=1, img_range=255., rgb_mean=(0.4488, 0.4371, 0.4040)):
    mean = torch.tensor(rgb_mean, dtype=x.dtype).view(1, 3, 1, 1)
    x = (x - mean) * img_range
    x = _conv(sd, 'conv_first', x)
    out = x
    for i in range(_count(sd, 'body.')):
        out = _res_block(sd, f'body.{i}', out, res_scale)
    res = _conv(sd, 'conv_after_body', out) + x
    x = _conv(sd, 'conv_last', _upsample(sd, res))
    return x / img_range + mean


def rcan_forward(sd, x, res_scale=1, img_range=255., rgb_mean=(0.4488, 0.4371, 0.4040)):
    mean = torch.tensor(rgb_mean, dtype=x.dtype).view(1, 3, 1, 1)
    x = (x - mean) * img_range
    x = _conv(sd, 'conv_first', x)
    out = x
    for g in range(_count(sd, 'body.')):
        g_in = out
        for b in range(_count(sd, f'body.{g}.residual_group.')):
            p = f'body.{g}.residual_group.{b}.rcab'
            t = _conv(sd, p + '.2', F.relu(_conv(sd, p + '.0', out)))
            y = t.mean((2, 3), keepdim=True)                                      # nn.AdaptiveAvgPool2d(1)
            y = F.relu(F.conv2d(y, sd[p + '.3.attention.1.weight'], sd[p + '.3.attention.1.bias']))
            y = torch.sigmoid(F.conv2d(y, sd[p + '.3.attention.3.weight'], sd[p + '.3.attention.3.bias']))
            out = t * y * res_scale + out                                          # rcan_arch.py:43-45
        out = _conv(sd, f'body.{g}.conv', out) + g_in
    res = _conv(sd, 'conv_after_body', out) + x
    x = _conv(sd, 'conv_last', _upsample(sd, res))
    return x / img_range + mean


def _rdb(sd, p, x):                                     # rrdbnet_arch.py:31-39
    lr = lambda t: F.leaky_relu(t, 0.2)                 # noqa: E731
    x1 = lr(_conv(sd, p + '.conv1', x))
    x2 = lr(_conv(sd, p + '.conv2', torch.cat((x, x1), 1)))
    x3 = lr(_conv(sd, p + '.conv3', torch.cat((x, x1, x2), 1)))
    x4 = lr(_conv(sd, p + '.conv4', torch.cat((x, x1, x2, x3), 1)))
    x5 = _conv(sd, p + '.conv5', torch.cat((x, x1, x2, x3, x4), 1))
    return x5 * 0.2 + x


def _pixel_unshuffle(x, scale):                         # arch_util.py:185-201
    b, c, hh, hw = x.shape
    h, w = hh // scale, hw // scale
    return x.view(b, c, h, scale, w, scale).permute(0, 1, 3, 5, 2, 4).reshape(b, c * scale * scale, h, w)


def rrdbnet_forward(sd, x, scale=4):                    # rrdbnet_arch.py:105-123
    feat = _pixel_unshuffle(x, 2) if scale == 2 else (_pixel_unshuffle(x, 4) if scale == 1 else x)
    feat = _conv(sd, 'conv_first', feat)
    out = feat
    for i in range(_count(sd, 'body.')):
        y = out
        for j in (1, 2, 3):
            y = _rdb(sd, f'body.{i}.rdb{j}', y)
        out = y * 0.2 + out                             # RRDB.forward :59-63
    feat = feat + _conv(sd, 'conv_body', out)
    feat = F.leaky_relu(_conv(sd, 'conv_up1', F.interpolate(feat, scale_factor=2, mode='nearest')), 0.2)
    feat = F.leaky_relu(_conv(sd, 'conv_up2', F.interpolate(feat, scale_factor=2, mode='nearest')), 0.2)
    return _conv(sd, 'conv_last', F.leaky_relu(_conv(sd, 'conv_hr', feat), 0.2))


def psnr01(a, b):
    mse = ((a.clamp(0, 1) - b.clamp(0, 1)).double() ** 2).mean().item()
    return float('inf') if mse == 0 else 10 * math.log10(1.0 / mse)
