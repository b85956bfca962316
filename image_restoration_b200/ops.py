"""Host-side operator layer over the C ABI: builds b200ir_conv_desc records and launches the kernels on torch's
current CUDA stream.  torch is used for device memory and streams only; all arithmetic happens in libb200ir.so.

Naming follows the reference operators these replace (EqualConv2d / ModulatedConv2d / ConvUpLayer /
upfirdn2d / fused_leaky_relu — see include/b200ir.h for file:line).
"""
import ctypes as C
import functools
import math
import os

import torch

from . import _lib
from ._lib import ConvDesc, View, check

SQRT2 = math.sqrt(2.0)
INV_SQRT2 = 1.0 / SQRT2


_raw_stream = getattr(torch._C, '_cuda_getCurrentRawStream', None)


def _stream():
    """torch's current CUDA stream of the current device as a cudaStream_t (the raw-handle accessor when this torch has it:
    torch.cuda.current_stream() builds a Python Stream object per call, ~16 us, and every launch asks)."""
    if _raw_stream is not None:
        return C.c_void_p(_raw_stream(torch.cuda.current_device()))
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _req(t, dtype, name):
    _lib.require_cuda(t, f'ops ({name})')
    if not (t.dtype == dtype and t.is_contiguous()):
        raise ValueError(f'{name}: expected contiguous CUDA {dtype}, got {t.dtype} on {t.device}')


@functools.lru_cache(maxsize=4096)
def pick_tile(m_w, m_h, m_b, min_w=1, max_b=128):
    """Choose (tile_w, tile_h, tile_b), powers of two with product 128, minimising the number of tiles (optionally
    with a minimum row-segment length, for DRAM-friendly TMA boxes, and a cap on images per tile)."""
    best = None
    for lw in range(8):
        for lh in range(8 - lw):
            tw, th, tb = 1 << lw, 1 << lh, 1 << (7 - lw - lh)
            if tw < min(min_w, m_w) or tb > max_b:
                continue
            tiles = -(-m_w // tw) * -(-m_h // th) * -(-m_b // tb)
            key = (tiles, -tw, -th)
            if best is None or key < best[0]:
                best = (key, (tw, th, tb))
    return best[1]


def pick_block_n(cout):
    for n in (256, 128, 64, 32, 16):
        if cout % n == 0:
            return n
    raise ValueError(f'cout={cout} must be a multiple of 16')


def nhwc_view(t, ch=None):
    """b200ir_view of a contiguous NHWC tensor [B,H,W,C]."""
    b, h, w, c = t.shape
    return View(t.data_ptr(), ch or c, w, h, b, c, w * c, h * w * c)


class ConvOp:
    """One prepared b200ir_conv_igemm launch (descriptor built once, launched many times)."""

    def __init__(self, views, weight, cin, cout, taps, m_whb, out, out_strides, *, out_fp32=False, out_c_off=0,
                 out_mul_off=(1, 0, 1, 0), bias=None, demod=None, noise=None, noise_gain=None, noise_strides=(0, 0),
                 act=False, res=None, res_mode=0, res_strides=(0, 0, 0), res_wh=(0, 0), res_scale=1.0, block_n=None,
                 tile=None, max_ctas=0, row_mode=0, out_scale=None, rgb_w=None, rgb_part=None, rgb_hw=(0, 0),
                 no_store=False, act_slope=None, res_mul=0.0, ps_r=0, ps_c=0, demod_c=0, tap_mask=None, corr=None,
                 w_per_image=False):
        d = ConvDesc()
        assert 1 <= len(views) <= _lib.MAX_VIEWS and 1 <= len(taps) <= _lib.MAX_TAPS
        for i, v in enumerate(views):
            d.a[i] = v
        d.num_views = len(views)
        _req(weight, torch.float16, 'weight')
        if w_per_image:      # one weight matrix per image: [m_b, cout, taps * cin]; a tile must not span images
            assert weight.shape == (m_whb[2], cout, len(taps) * cin), (weight.shape, m_whb, cout, len(taps), cin)
            d.w_per_image = 1
            tile = tile or pick_tile(m_whb[0], m_whb[1], m_whb[2], max_b=1)
            assert tile[2] == 1 and not row_mode
        else:
            assert weight.shape == (cout, len(taps) * cin), (weight.shape, cout, len(taps), cin)
        d.weight = weight.data_ptr()
        d.cin, d.cout, d.num_taps = cin, cout, len(taps)
        for i, (v, dx, dy) in enumerate(taps):
            d.tap_view[i], d.tap_dx[i], d.tap_dy[i] = v, dx, dy
        d.m_w, d.m_h, d.m_b = m_whb
        d.tile_w, d.tile_h, d.tile_b = tile or pick_tile(*m_whb)
        d.block_n = block_n or pick_block_n(cout)
        d.out = out.data_ptr() if out is not None else None
        d.out_fp32 = 1 if out_fp32 else 0
        d.out_stride_x, d.out_stride_y, d.out_stride_b = out_strides
        d.out_c_off = out_c_off
        d.out_x_mul, d.out_x_off, d.out_y_mul, d.out_y_off = out_mul_off
        for name, t in (('bias', bias), ('demod', demod), ('noise', noise), ('noise_gain', noise_gain)):
            if t is not None:
                _req(t, torch.float32, name)
                setattr(d, name, t.data_ptr())
        d.noise_stride_b, d.noise_stride_y = noise_strides
        # act=True: FusedLeakyReLU (slope 0.2, gain sqrt 2); act_slope given: max(v, act_slope * v) without gain
        d.act = 2 if act_slope is not None else (1 if act else 0)
        d.act_slope = float(act_slope) if act_slope is not None else 0.0
        d.res_mul = res_mul
        d.ps_r = ps_r
        d.ps_c, d.demod_c = ps_c, demod_c
        if corr is not None:        # (top, bottom, left, right) fp32 correction buffers of the folded ConvUpLayer
            for name, t in zip(('corr_top', 'corr_bot', 'corr_left', 'corr_right'), corr):
                _req(t, torch.float32, name)
                setattr(d, name, t.data_ptr())
            self._keep_corr = corr
        if tap_mask is not None:
            assert len(tap_mask) <= 8
            d.use_tap_mask = 1
            for i, m in enumerate(tap_mask):
                d.tap_mask[i] = m
        d.res_mode = res_mode
        if res is not None:
            _req(res, torch.float16, 'res')
            d.res = res.data_ptr()
        d.res_stride_x, d.res_stride_y, d.res_stride_b = res_strides
        d.res_w, d.res_h = res_wh
        d.res_scale = res_scale
        d.max_ctas = max_ctas
        d.row_mode = row_mode
        for name, t in (('out_scale', out_scale), ('rgb_w', rgb_w), ('rgb_part', rgb_part)):
            if t is not None:
                _req(t, torch.float32, name)
                setattr(d, name, t.data_ptr())
        if rgb_w is not None:
            assert rgb_part is not None and rgb_part.shape[0] == cout // d.block_n, (rgb_part.shape, cout, d.block_n)
        d.rgb_h, d.rgb_w_px = rgb_hw
        d.no_store = 1 if no_store else 0
        self.desc = d
        # keep every tensor alive as long as the op exists
        self._keep = (weight, out, bias, demod, noise, noise_gain, res, out_scale, rgb_w, rgb_part)
        self._plan = None

    def attach_rgb(self, rgb_w, hw, no_store=False):
        """Fuse ToRGB into this conv's epilogue: rgb_w fp32 [B][3][cout]; returns the partial-plane buffer
        [cout/block_n][B][3][h][w] that b200ir_rgb_combine reduces."""
        d = self.desc
        _req(rgb_w, torch.float32, 'rgb_w')
        assert rgb_w.shape == (d.m_b, 3, d.cout), rgb_w.shape
        part = torch.empty(d.cout // d.block_n, d.m_b, 3, hw[0], hw[1], device=rgb_w.device, dtype=torch.float32)
        d.rgb_w, d.rgb_part = rgb_w.data_ptr(), part.data_ptr()
        d.rgb_h, d.rgb_w_px = hw
        d.no_store = 1 if no_store else 0
        self._keep = self._keep + (rgb_w, part)
        self._plan = None
        return part

    def __call__(self):
        """Launch through a b200ir_conv_plan: validation, tensor-map encoding and tile sizing happen once per distinct
        descriptor (pointers included), not once per launch."""
        if self._plan is None:
            self._plan = _conv_plan(self.desc)
        check(_lib.lib().b200ir_conv_plan_launch(self._plan, _stream()), 'b200ir_conv_plan_launch')


class _PlanCache:
    """Descriptor bytes -> b200ir_conv_plan handle (LRU).  The training-step Functions build a fresh ConvOp per call, but
    the caching allocator hands the same buffers back step after step, so the same descriptors recur."""

    def __init__(self, capacity=8192):
        self.capacity = capacity
        self.plans = {}

    def get(self, desc):
        key = bytes(desc)
        plan = self.plans.pop(key, None)
        if plan is None:
            handle = C.c_void_p()
            check(_lib.lib().b200ir_conv_plan_create(C.byref(desc), C.byref(handle)), 'b200ir_conv_plan_create')
            plan = handle
            while len(self.plans) >= self.capacity:
                _lib.lib().b200ir_conv_plan_destroy(self.plans.pop(next(iter(self.plans))))
        self.plans[key] = plan              # most recently used last
        return plan


_PLANS = _PlanCache()


def _conv_plan(desc):
    return _PLANS.get(desc)


def taps_3x3():
    return [(0, kw - 1, kh - 1) for kh in range(3) for kw in range(3)]


NUM_SMS = 148


def row_mode_ok(b, h, w, cin, cout):
    """Mirror of the eligibility test of the row-sliding conv variant in b200ir_conv_igemm: low channel counts at
    high resolution (L2-bound with generic tiles), weights resident in shared memory, enough work items."""
    if cin > 128 or cout not in (16, 32, 64) or w < 128:
        return False
    block_k = 64 if cin % 64 == 0 else (32 if cin % 32 == 0 else 16)
    kc = cin // block_k
    w_bytes = 9 * kc * cout * block_k * 2
    slots = (232448 - 1024 - (512 + 512 * 4 + 3 * 2560 * 4) - w_bytes) // (136 * block_k * 2)
    if slots < 2 * kc:
        return False
    return b * -(-w // 128) * -(-h // 8) >= 2 * NUM_SMS


def conv_same(x, weight, out, ksize, **kw):
    """Stride-1 'same' conv (k = 1 or 3) of NHWC x [B,H,W,Cin] into NHWC out [B,H,W,Ctot] (channel offset via
    out_c_off).  EqualConv2d / plain ModulatedConv2d / ConvUpLayer conv."""
    b, h, w, cin = x.shape
    cout = weight.shape[-2]            # [cout, K], or [B, cout, K] with w_per_image
    taps = taps_3x3() if ksize == 3 else [(0, 0, 0)]
    oc = out.shape[3] if out is not None else cout
    if ksize == 3 and 'tile' not in kw and 'row_mode' not in kw and not kw.get('w_per_image') and \
            row_mode_ok(b, h, w, cin, cout):
        kw.update(tile=(128, 1, 1), row_mode=1, block_n=cout)
    return ConvOp([nhwc_view(x)], weight, cin, cout, taps, (w, h, b), out, (oc, w * oc, h * w * oc), **kw)


def conv3x3_s2(p, h, w, weight, out, **kw):
    """3x3 stride-2 conv over the FIR-smoothed buffer p [B,H+2,W+2,C] (valid (H+1)x(W+1)) -> out [B,H/2,W/2,Cout].
    EqualConv2d(stride=2, padding=0) after UpFirDnSmooth, stylegan2_ocr_arch.py:685-697."""
    b, hp, wp, c = p.shape
    assert hp == h + 2 and wp == w + 2 and h % 2 == 0 and w % 2 == 0
    cout = weight.shape[0]
    if c == 32 and _S2_FOLD:       # 64-byte pixel rows: the pixel-folded form is faster (118 -> 87 us at B=64, 128x384, 32->64;
        return conv3x3_s2_folded(p, h, w, s2_fold_weight(weight, c), out, **kw)   # 64 channels and up: slower, tools/time_conv_s2.py)
    views = []
    for py in range(2):
        for px in range(2):
            views.append(View(p.data_ptr() + 2 * (py * wp + px) * c, c, wp // 2, hp // 2, b, 2 * c, 2 * wp * c,
                              hp * wp * c))
    taps = [((kh % 2) * 2 + (kw % 2), kw // 2, kh // 2) for kh in range(3) for kw in range(3)]
    oh, ow = h // 2, w // 2
    return ConvOp(views, weight, c, cout, taps, (ow, oh, b), out, (cout, ow * cout, oh * ow * cout), **kw)


_S2_FOLD = os.environ.get('B200IR_S2_FOLD', '1') != '0'


def s2_fold_weight(weight, cin):
    """Packed weights [cout, 9*cin] (tap-major) of the stride-2 conv -> [cout, 6 * 2*cin] for conv3x3_s2_folded: taps
    (kh, kw' = 0): pixel pair (kw 0 | kw 1), (kh, kw' = 1): (kw 2 | zeros)."""
    cout = weight.shape[0]
    w = weight.view(cout, 3, 3, cin)
    out = torch.zeros(cout, 3, 2, 2, cin, device=weight.device, dtype=weight.dtype)
    out[:, :, 0, 0], out[:, :, 0, 1], out[:, :, 1, 0] = w[:, :, 0], w[:, :, 1], w[:, :, 2]
    return out.reshape(cout, 12 * cin).contiguous()


def conv3x3_s2_folded(p, h, w, weight_folded, out, **kw):
    """conv3x3_s2 on a pixel-folded view of p: horizontally adjacent pixel pairs are read as 2*C channels of one pixel
    ([B,H+2,(W+2)/2,2C], a free view), which turns the stride-2 sampling along x into a dense stride-1 conv with two taps per
    kernel row (the third kw shares a pixel pair with a zero block); rows keep their stride-2 phase views.  Dense TMA boxes with
    128-byte rows for C = 32 instead of element-strided ones, at 4/3 of the MMA work."""
    b, hp, wp, c = p.shape
    assert hp == h + 2 and wp == w + 2 and h % 2 == 0 and w % 2 == 0
    cout = weight_folded.shape[0]
    assert weight_folded.shape[1] == 12 * c
    views = [View(p.data_ptr() + 2 * py * wp * c, 2 * c, wp // 2, hp // 2, b, 2 * c, 2 * wp * c, hp * wp * c) for py in range(2)]
    taps = [(kh % 2, kx, kh // 2) for kh in range(3) for kx in range(2)]
    oh, ow = h // 2, w // 2
    return ConvOp(views, weight_folded, 2 * c, cout, taps, (ow, oh, b), out, (cout, ow * cout, oh * ow * cout), **kw)


CONVT_PHASES = [(py, px) for py in range(2) for px in range(2)]


def convt_phase_taps(py, px):
    """Taps (kh, kw) of a 3x3 stride-2 transposed conv that land on output rows 2i+py / cols 2j+px."""
    khs = (0, 2) if py == 0 else (1,)
    kws = (0, 2) if px == 0 else (1,)
    return [(kh, kw) for kh in khs for kw in kws]


def convt_s2_phase(x, weight_phase, py, px, raw, **kw):
    """One output phase of conv_transpose2d(stride 2, padding 0, 3x3) (stylegan2_ocr_arch.py:265): x [B,h,w,Cin] ->
    raw[b, 2i+py, 2j+px, :] for the (2h+1)x(2w+1) valid region of raw [B,RH,RW,Cout]."""
    b, h, w, cin = x.shape
    _, rh, rw, cout = raw.shape
    taps = [(0, -(kwi // 2), -(khi // 2)) for (khi, kwi) in convt_phase_taps(py, px)]
    m_h = h + 1 if py == 0 else h
    m_w = w + 1 if px == 0 else w
    return ConvOp([nhwc_view(x)], weight_phase, cin, cout, taps, (m_w, m_h, b), raw,
                  (cout, rw * cout, rh * rw * cout), out_mul_off=(2, px, 2, py), **kw)


def convt_merged_weight(w, scale):
    """(cout, cin, 3, 3) fp32 -> fp16 [4*cout][4*cin] for convt_s2_merged: row block = output phase (py, px), column
    block = input tap (ty, tx) (input offset (-ty, -tx)); phase (py, px) uses tap (ty, tx) through kernel element
    (py + 2 ty, px + 2 tx) when that index is <= 2, zero otherwise."""
    cout, cin = w.shape[:2]
    big = w.new_zeros(4, cout, 4, cin)
    for py in range(2):
        for px in range(2):
            for ty in range(2):
                for tx in range(2):
                    kh, kw = py + 2 * ty, px + 2 * tx
                    if kh <= 2 and kw <= 2:
                        big[py * 2 + px, :, ty * 2 + tx, :] = w[:, :, kh, kw] * scale
    return big.reshape(4 * cout, 4 * cin).contiguous().to(torch.float16)


def convt_s2_merged(x, w_big, raw, demod):
    """conv_transpose2d(stride 2, padding 0, 3x3) (stylegan2_ocr_arch.py:265) as ONE implicit GEMM: the four output
    phases are column blocks of N = 4*cout (N = 256 MMAs even for cout = 64), the four input taps (i - ty, j - tx) are
    K blocks, every N-tile skips the taps its phases do not use, and the epilogue stores column block (py, px) to
    raw[b, 2i + py, 2j + px, :].  M runs over (h+1) x (w+1); the surplus row / column of the odd phases lands in the
    padding row / column of `raw` that nothing reads."""
    b, h, w, cin = x.shape
    _, rh, rw, cout = raw.shape
    assert rh >= 2 * h + 2 and rw >= 2 * w + 2 and w_big.shape == (4 * cout, 4 * cin)
    taps = [(0, -tx, -ty) for ty in range(2) for tx in range(2)]
    bn = min(256, 4 * cout)
    masks = []
    for j in range(4 * cout // bn):
        m = 0
        for ph in range(j * bn // cout, ((j + 1) * bn - 1) // cout + 1):
            py, px = ph // 2, ph % 2
            for ty in range(2):
                for tx in range(2):
                    if py + 2 * ty <= 2 and px + 2 * tx <= 2:
                        m |= 1 << (ty * 2 + tx)
        masks.append(m)
    tile = pick_tile(w + 1, h + 1, b, min_w=8, max_b=max(1, 2560 // bn))
    return ConvOp([nhwc_view(x)], w_big, cin, 4 * cout, taps, (w + 1, h + 1, b), raw, (cout, rw * cout, rh * rw * cout),
                  demod=demod, block_n=bn, tile=tile, ps_r=2, ps_c=cout, demod_c=cout, tap_mask=masks)


# ------------------------------------------------------------------------------------------ folded ConvUpLayer
_BILIN = ((.75, .25, 0.), (.25, .75, 0.), (0., .75, .25), (0., .25, .75))   # A[r + 1][dy + 1]: weight of t[i + dy] in up(2i + r)


def upfold_weights(w, scale):
    """ConvUpLayer (gfpganv1_ocr_arch.py:188-202) folded: conv3x3(bilinear_up2(t)) == for every output phase (py, px) a 3x3
    conv over the replicate-padded low-resolution t with  Wp = sum_{kh,kw} W[kh,kw] A[py+kh][dy] A[px+kw][dx].
    w (cout, cin, 3, 3) fp32 ->  dict(main fp16 [4*cout][9*cin] (phase-major rows, taps (dy, dx) row-major),
    top / bot fp16 [2*cout][3*cin], left / right fp16 [2*cout][3*cin] (1-D surplus convs of the border ring),
    corners fp32 [4][cout][cin])."""
    dev = w.device
    A = torch.tensor(_BILIN, dtype=torch.float32)
    w = w.detach().float().cpu() * scale      # one-time host-side packing (no device GEMMs on behalf of weight prep)
    cout, cin = w.shape[:2]
    Ay = torch.stack([A[py:py + 3] for py in range(2)])                 # [py][kh][dy]
    wp = torch.einsum('oikl,pkd,qle->pqodei', w, Ay, Ay)                # [py][px][co][dy][dx][ci]
    row = lambda kh: torch.einsum('oil,qle->qoei', w[:, :, kh, :], Ay)  # noqa: E731  [px][co][dx][ci]
    col = lambda kw: torch.einsum('oik,pkd->podi', w[:, :, :, kw], Ay)  # noqa: E731  [py][co][dy][ci]
    f16 = lambda t, rows: t.reshape(rows, -1).contiguous().to(torch.float16).to(dev)  # noqa: E731
    return dict(main=f16(wp, 4 * cout), top=f16(row(0), 2 * cout), bot=f16(row(2), 2 * cout), left=f16(col(0), 2 * cout),
                right=f16(col(2), 2 * cout),
                corners=torch.stack([w[:, :, 0, 0], w[:, :, 0, 2], w[:, :, 2, 0], w[:, :, 2, 2]]).contiguous().to(dev))


class UpFoldConv:
    """Prepared launches of one folded ConvUpLayer + ResUpBlock merge: tp [B, h+2, w+2, cin] is the replicate-padded
    low-resolution input (interior written by the producer), out [B, 2h, 2w, cout]."""

    def __init__(self, tp, fw, bias, out, res, res_scale):
        b, hp, wp_, cin = tp.shape
        h, w = hp - 2, wp_ - 2
        cout = out.shape[3]
        dev = tp.device
        self.tp, self.h, self.w, self.cin, self.cout, self.fw = tp, h, w, cin, cout, fw
        e32 = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)  # noqa: E731
        self.top, self.bot = e32(b, 2 * w, cout), e32(b, 2 * w, cout)
        self.left, self.right = e32(b, 2 * h, cout), e32(b, 2 * h, cout)
        sb, sy = hp * wp_ * cin, wp_ * cin
        base = tp.data_ptr()
        taps3 = [(0, d, 0) for d in range(3)]

        def border(ptr, n, stride_n, wgt, dst):
            v = View(ptr, cin, n + 2, 1, b, stride_n, sb, sb)
            return ConvOp([v], wgt, cin, 2 * cout, taps3, (n, 1, b), dst, (2 * cout, n * 2 * cout, n * 2 * cout),
                          out_fp32=True)
        self.border_ops = [border(base + 2 * (1 * sy), w, cin, fw['top'], self.top),             # row t[0]
                           border(base + 2 * (h * sy), w, cin, fw['bot'], self.bot),             # row t[h-1]
                           border(base + 2 * (1 * cin), h, sy, fw['left'], self.left),           # column t[:, 0]
                           border(base + 2 * (w * cin), h, sy, fw['right'], self.right)]         # column t[:, w-1]
        taps9 = [(0, dx, dy) for dy in range(3) for dx in range(3)]
        view = View(base, cin, wp_, hp, b, cin, sy, sb)
        cr = res.shape[3]
        self.main = ConvOp([view], fw['main'], cin, 4 * cout, taps9, (w, h, b), out,
                           (cout, 2 * w * cout, 4 * h * w * cout), bias=bias, act=True, res=res, res_mode=2,
                           res_strides=(cr, w * cr, h * w * cr), res_wh=(w, h), res_scale=res_scale,
                           block_n=min(256, 4 * cout), ps_r=2, ps_c=cout,
                           corr=(self.top, self.bot, self.left, self.right))

    def pad(self):
        check(_lib.lib().b200ir_replicate_border(_ptr(self.tp), self.tp.shape[0], self.h, self.w, self.cin, _stream()),
              'replicate_border')

    def corners(self):
        check(_lib.lib().b200ir_upfold_corners(_ptr(self.tp), _ptr(self.fw['corners']), _ptr(self.top), _ptr(self.bot),
                                               self.tp.shape[0], self.h, self.w, self.cin, self.cout, _stream()),
              'upfold_corners')


def linear_as_conv(x2d, weight, out, **kw):
    """EqualLinear as a 1x1 conv over a [B,1,1,K] view; out fp32 [B,N]."""
    b, k = x2d.shape
    n = weight.shape[0]
    v = View(x2d.data_ptr(), k, 1, 1, b, k, k, k)
    return ConvOp([v], weight, k, n, [(0, 0, 0)], (1, 1, b), out, (n, n, n), out_fp32=True, tile=(1, 1, 128), **kw)


# ------------------------------------------------------------------------------------------ memory-bound stages
def first_conv(x, w, bias, out):
    b, _, h, wd = x.shape
    check(_lib.lib().b200ir_first_conv(_ptr(x), _ptr(w), _ptr(bias), _ptr(out), b, h, wd, w.shape[0], _stream()),
          'first_conv')


def u8_to_input(img, x, swap_rb=True):
    """uint8 HWC batch [B,H,W,3] -> fp32 NCHW [-1,1] (img2tensor + normalize of api.py:96-101)."""
    b, h, w, _ = img.shape
    check(_lib.lib().b200ir_u8_to_input(_ptr(img), _ptr(x), b, h, w, 1 if swap_rb else 0, _stream()), 'u8_to_input')


def image_to_u8(x, img, swap_rb=True):
    """fp32 NCHW network output -> uint8 HWC (tensor2img with min_max=(-1,1), api.py:105)."""
    b, _, h, w = x.shape
    check(_lib.lib().b200ir_image_to_u8(_ptr(x), _ptr(img), b, h, w, 1 if swap_rb else 0, _stream()), 'image_to_u8')


def fir_pad22(x, out):
    b, h, w, c = x.shape
    check(_lib.lib().b200ir_fir_pad22(_ptr(x), _ptr(out), b, h, w, c, out.shape[1], out.shape[2], _stream()),
          'fir_pad22')


def fir_down2(x, out):
    b, h, w, c = x.shape
    check(_lib.lib().b200ir_fir_down2(_ptr(x), _ptr(out), b, h, w, c, _stream()), 'fir_down2')


def bilinear_up2(x, out):
    b, h, w, c = x.shape
    check(_lib.lib().b200ir_bilinear_up2(_ptr(x), _ptr(out), b, h, w, c, _stream()), 'bilinear_up2')


def add(a, b, out):
    check(_lib.lib().b200ir_add(_ptr(a), _ptr(b), _ptr(out), a.numel(), _stream()), 'add')


def upfir_act(raw, out, noise, noise_stride_b, noise_gain, bias, scale, shift, c_sft, s_next):
    b, h2, w2, c = out.shape
    check(_lib.lib().b200ir_upfir_act(_ptr(raw), _ptr(out), b, h2, w2, c, raw.shape[1], raw.shape[2], _ptr(noise),
                                      noise_stride_b, _ptr(noise_gain), _ptr(bias), _ptr(scale), _ptr(shift), c_sft,
                                      _ptr(s_next), _stream()), 'upfir_act')


def to_rgb(x, wrgb, s, bias, skip, rgb, s_next=None, xs_out=None):
    b, h, w, c = x.shape
    check(_lib.lib().b200ir_to_rgb(_ptr(x), b, h, w, c, _ptr(wrgb), _ptr(s), _ptr(bias), _ptr(skip), _ptr(rgb),
                                   _ptr(s_next), _ptr(xs_out), _stream()), 'to_rgb')


def rgb_combine(part, bias, skip, rgb):
    """rgb = bias + sum of the partial planes the conv epilogues wrote (+ up-sampled skip): second half of ToRGB."""
    b, _, h, w = rgb.shape
    check(_lib.lib().b200ir_rgb_combine(_ptr(part), part.shape[0], _ptr(bias), _ptr(skip), _ptr(rgb), b, h, w,
                                        _stream()), 'rgb_combine')


def rgb_wmod(w, s, wm):
    """wm[b][o][c] = w[o][c] * s[b][c]: per-image ToRGB weights (ModulatedConv2d without demodulation)."""
    b, _, c = wm.shape
    check(_lib.lib().b200ir_rgb_wmod(_ptr(w), _ptr(s), _ptr(wm), b, c, _stream()), 'rgb_wmod')


def modulate_const(cst, s, out):
    b, h, w, c = out.shape
    check(_lib.lib().b200ir_modulate_const(_ptr(cst), _ptr(s), _ptr(out), b, h * w, c, _stream()), 'modulate_const')


def mod_linear(latent, lat_idx, w, bias, wscale, s):
    b, L, f = latent.shape
    check(_lib.lib().b200ir_mod_linear(_ptr(latent), L, f, lat_idx, _ptr(w), _ptr(bias), wscale, _ptr(s), b,
                                       w.shape[0], _stream()), 'mod_linear')


def demod(s, wsq, scale2, d):
    b, cin = s.shape
    check(_lib.lib().b200ir_demod(_ptr(s), _ptr(wsq), scale2, _ptr(d), b, cin, wsq.shape[0], _stream()), 'demod')


def _device_records(records, dev):
    """ctypes structure array -> device byte tensor (the *_multi kernels read their layer tables from HBM)."""
    arr = (type(records[0]) * len(records))(*records)
    return torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev)


class ModLinearMulti:
    """All modulation linears of one forward (ModulatedConv2d.modulation, stylegan2_ocr_arch.py:229-234,249) in one
    launch.  layers: list of (weight [cin,F], bias [cin], lat_idx, s_out [B,cin])."""

    def __init__(self, latent, layers, wscale):
        self.latent, self.wscale = latent, wscale
        self.keep = layers
        recs = [_lib.ModLayer(w.data_ptr(), b.data_ptr(), s.data_ptr(), li, w.shape[0]) for (w, b, li, s) in layers]
        self.table = _device_records(recs, latent.device)
        self.n = len(recs)
        self.max_cin = max(w.shape[0] for (w, _, _, _) in layers)

    def __call__(self):
        b, L, f = self.latent.shape
        check(_lib.lib().b200ir_mod_linear_multi(_ptr(self.latent), L, f, _ptr(self.table), self.n, self.max_cin,
                                                 self.wscale, b, _stream()), 'mod_linear_multi')


class DemodMulti:
    """All demodulation tables of one forward (stylegan2_ocr_arch.py:253-257) in one launch.
    layers: list of (s [B,cin], wsq [cout,cin], scale2, d_out [B,cout])."""

    def __init__(self, layers):
        self.keep = layers
        recs = [_lib.DemodLayer(s.data_ptr(), wsq.data_ptr(), d.data_ptr(), sc2, wsq.shape[1], wsq.shape[0])
                for (s, wsq, sc2, d) in layers]
        self.table = _device_records(recs, layers[0][0].device)
        self.n = len(recs)
        self.max_cout = max(wsq.shape[0] for (_, wsq, _, _) in layers)
        assert all(wsq.shape[1] <= 512 for (_, wsq, _, _) in layers), 'demod_multi keeps a wsq row in registers: cin <= 512'
        self.b = layers[0][0].shape[0]

    def __call__(self):
        check(_lib.lib().b200ir_demod_multi(_ptr(self.table), self.n, self.max_cout, self.b, _stream()), 'demod_multi')


# ------------------------------------------------------------------------------------------ SR-network stages
def nchw_to_nhwc_pad(x, out, sub=None, mul=1.0, unshuffle=1):
    b, c = x.shape[:2]
    check(_lib.lib().b200ir_nchw_to_nhwc_pad(_ptr(x), _ptr(out), b, c, out.shape[1], out.shape[2], out.shape[3], _ptr(sub),
                                             mul, unshuffle, _stream()), 'nchw_to_nhwc_pad')


def nearest_up2(x, out):
    b, h, w, c = x.shape
    check(_lib.lib().b200ir_nearest_up2(_ptr(x), _ptr(out), b, h, w, c, _stream()), 'nearest_up2')


def sr_output(y, out, mul=1.0, add=None, base=None, scale=1):
    b, c, h, w = out.shape
    check(_lib.lib().b200ir_sr_output(_ptr(y), _ptr(out), b, c, h, w, y.shape[3], mul, _ptr(add), _ptr(base), scale,
                                      _stream()), 'sr_output')


def channel_mean(x, mean):
    b, h, w, c = x.shape
    check(_lib.lib().b200ir_channel_mean(_ptr(x), _ptr(mean), b, h * w, c, _stream()), 'channel_mean')


def ca_mlp(mean, w1, b1, w2, b2, att):
    b, c = mean.shape
    check(_lib.lib().b200ir_ca_mlp(_ptr(mean), _ptr(w1), _ptr(b1), _ptr(w2), _ptr(b2), _ptr(att), b, c, w1.shape[0],
                                   _stream()), 'ca_mlp')


def ca_scale_add(x, att, identity, out, res_scale=1.0, id_stride=None, out_stride=None):
    """out = x * att * res_scale + identity; identity / out may be the leading channels of wider NHWC buffers (pass the
    tensors and their pixel strides)."""
    b, h, w, c = x.shape
    check(_lib.lib().b200ir_ca_scale_add(_ptr(x), _ptr(att), _ptr(identity), _ptr(out), res_scale, b, h * w, c,
                                         id_stride or c, out_stride or c, _stream()), 'ca_scale_add')


def style_mlp(z, w, bias, out, lr_mul):
    b, f = z.shape
    check(_lib.lib().b200ir_style_mlp(_ptr(z), _ptr(w), _ptr(bias), _ptr(out), b, f, w.shape[0], lr_mul, _stream()),
          'style_mlp')


def nhwc_to_nchw_f32(x, out):
    b, h, w, c = x.shape
    check(_lib.lib().b200ir_nhwc_to_nchw_f32(_ptr(x), _ptr(out), b, h * w, c, _stream()), 'nhwc_to_nchw_f32')


def conv_wgrad(x, dy, dw=None):
    """Weight gradient of the 3x3 stride-1 'same' conv: x NHWC fp16 [B,H,W,Cin], dy NHWC fp16 [B,H,W,Cout] ->
    dw fp32 [Cout, 9, Cin] (tap-major like the packed forward weights).  Backward of F.conv2d(x, W, padding=1) w.r.t. W."""
    b, h, w, cin = x.shape
    cout = dy.shape[3]
    assert tuple(dy.shape[:3]) == (b, h, w) and x.dtype == torch.float16 and dy.dtype == torch.float16
    if dw is None:
        dw = torch.empty(cout, 9, cin, device=x.device, dtype=torch.float32)
    f = wgrad_fold(cin, cout, w)
    if f > 1 and x.is_contiguous() and dy.is_contiguous():
        # low-channel layers: f adjacent pixels read as f * C channels (a free view of NHWC), GEMM on the folded shapes, then
        # the tap blocks are collected (b200ir_wgrad_unfold)
        G = torch.empty(f * cout, 9, f * cin, device=x.device, dtype=torch.float32)
        check(_lib.lib().b200ir_conv_wgrad(_ptr(x), _ptr(dy), _ptr(G), b, h, w // f, f * cin, f * cout, _stream()), 'conv_wgrad')
        check(_lib.lib().b200ir_wgrad_unfold(_ptr(G), _ptr(dw), f, cin, cout, _stream()), 'wgrad_unfold')
        return dw
    check(_lib.lib().b200ir_conv_wgrad(_ptr(x), _ptr(dy), _ptr(dw), b, h, w, cin, cout, _stream()), 'conv_wgrad')
    return dw


_WGRAD_FOLD = int(os.environ.get('B200IR_WGRAD_FOLD', '1'))        # 0: never fold (A/B switch of tools/time_wgrad_lowc.py)


def wgrad_fold(cin, cout, w):
    """Pixel-fold factor of conv_wgrad.  The weight-gradient GEMM pays for a 128 (cout) x 64 (cin) tile whatever the channel
    counts, so 32 -> 32 runs as 128 -> 128 over a quarter of the pixels (1 / f of the folded result is kept): measured at
    B = 256, 128x384: 32->32 1882 -> 698 us, 32->64 1994 -> 1036 us, 64->32 2075 -> 1158 us, 64->64 1893 -> 1321 us."""
    if _WGRAD_FOLD == 0:
        return 1
    if cin == 32 and cout == 32 and w % 4 == 0:
        return 4
    if (cin, cout) in ((32, 64), (64, 32), (64, 64)) and w % 2 == 0:
        return 2
    return 1


def conv_dgrad_weight(weight, cin):
    """Packed forward weights [Cout, 9*Cin] (tap-major, fp16) -> the packed weights of the input-gradient conv
    [Cin, 9*Cout]: dx = conv_same(dz, Wt) with Wt[ci][kh][kw][co] = W[co][2-kh][2-kw][ci] (the 3x3 stride-1 'same' conv
    is its own adjoint up to this flip / transpose, so dgrad runs on b200ir_conv_igemm)."""
    cout = weight.shape[0]
    assert weight.shape == (cout, 9 * cin)
    return weight.view(cout, 3, 3, cin).flip(1, 2).permute(3, 1, 2, 0).reshape(cin, 9 * cout).contiguous()


def conv_dgrad(dz, weight_t, dx, **kw):
    """Input gradient of F.conv2d(x, W, padding=1): dz NHWC fp16 [B,H,W,Cout], weight_t from conv_dgrad_weight ->
    dx NHWC fp16 [B,H,W,Cin].  Returns the prepared ConvOp (call it to launch)."""
    return conv_same(dz, weight_t, dx, 3, **kw)


def lrelu_bias_bwd(dy, y, dz=None, dbias=None, slope=0.2, scale=2 ** 0.5, want_bias=True):
    """Backward of FusedLeakyReLU + bias gradient (b200ir_lrelu_bias_bwd): dy, y NHWC fp16 [..., C] -> (dz, dbias).
    y=None: no activation, only the bias gradient sum_p dy (pass scale=1.0); returns (None, dbias)."""
    assert dy.dtype == torch.float16 and (y is None or (dy.shape == y.shape and y.dtype == torch.float16))
    c = dy.shape[-1]
    if dz is None and y is not None:
        dz = torch.empty_like(dy)
    if dbias is None and want_bias:
        dbias = torch.empty(c, device=dy.device, dtype=torch.float32)
    check(_lib.lib().b200ir_lrelu_bias_bwd(_ptr(dy), _ptr(y), _ptr(dz), _ptr(dbias) if dbias is not None else None,
                                           dy.numel() // c, c, slope, scale, _stream()), 'lrelu_bias_bwd')
    return dz, dbias


def conv_wgrad_view(view, dy, dw=None, tap_mask=0x1FF):
    """b200ir_conv_wgrad_view: weight gradient over a strided view of x and a subset of the nine taps (see
    include/b200ir.h).  dy NHWC fp16 [B,H,W,Cout] -> dw fp32 [Cout, 9, Cin] (taps outside the mask are zero)."""
    b, h, w, cout = dy.shape
    assert dy.dtype == torch.float16 and dy.is_contiguous()
    if dw is None:
        dw = torch.empty(cout, 9, view.c, device=dy.device, dtype=torch.float32)
    check(_lib.lib().b200ir_conv_wgrad_view(C.byref(view), _ptr(dy), _ptr(dw), b, h, w, cout, tap_mask, _stream()),
          'conv_wgrad_view')
    return dw


def conv1x1_wgrad(x, dy):
    """Weight gradient of a 1x1 conv (EqualConv2d k = 1): x NHWC fp16 [B,H,W,Cin], dy [B,H,W,Cout] -> fp32 [Cout, Cin]."""
    assert x.dtype == torch.float16 and x.is_contiguous() and tuple(x.shape[:3]) == tuple(dy.shape[:3])
    return conv_wgrad_view(nhwc_view(x), dy, tap_mask=1 << 4)[:, 4, :]


def conv3x3_s2_wgrad(p, h, w, dy):
    """Weight gradient of conv3x3_s2 (the stride-2 conv of ResBlock.conv2 over the FIR-smoothed buffer p [B,H+2,W+2,C]):
    dy [B,H/2,W/2,Cout] -> fp32 [Cout, 9, C].  One masked launch per pixel phase of p; kernel element (2 sy + ry, 2 sx + rx)
    is tap (sy + 1, sx + 1) of the launch over phase (ry, rx)."""
    b, hp, wp, c = p.shape
    assert hp == h + 2 and wp == w + 2 and h % 2 == 0 and w % 2 == 0 and p.dtype == torch.float16 and p.is_contiguous()
    cout = dy.shape[3]
    assert tuple(dy.shape[:3]) == (b, h // 2, w // 2)
    dw = torch.empty(cout, 3, 3, c, device=p.device, dtype=torch.float32)
    for ry in range(2):
        for rx in range(2):
            view = View(p.data_ptr() + 2 * (ry * wp + rx) * c, c, wp // 2, hp // 2, b, 2 * c, 2 * wp * c, hp * wp * c)
            sys_, sxs = ((0, 1) if ry == 0 else (0,)), ((0, 1) if rx == 0 else (0,))
            mask = sum(1 << ((sy + 1) * 3 + sx + 1) for sy in sys_ for sx in sxs)
            part = conv_wgrad_view(view, dy, tap_mask=mask).view(cout, 3, 3, c)
            for sy in sys_:
                for sx in sxs:
                    dw[:, 2 * sy + ry, 2 * sx + rx] = part[:, sy + 1, sx + 1]
    return dw.view(cout, 9, c)


def fir_pad11(raw, out):
    """Adjoint of fir_pad22: raw [B,RH,RW,C] (valid (H+1)x(W+1) at the origin) -> out [B,H,W,C]."""
    b, h, w, c = out.shape
    check(_lib.lib().b200ir_fir_pad11(_ptr(raw), _ptr(out), b, h, w, c, raw.shape[1], raw.shape[2], _stream()), 'fir_pad11')


def fir_down2_adjoint(d, out, add=None):
    """Adjoint of fir_down2: d [B,h,w,C] -> out [B,2h,2w,C] (+ add)."""
    b, h, w, c = d.shape
    assert tuple(out.shape) == (b, 2 * h, 2 * w, c) and (add is None or add.shape == out.shape)
    check(_lib.lib().b200ir_fir_down2_adjoint(_ptr(d), _ptr(add), _ptr(out), b, h, w, c, _stream()), 'fir_down2_adjoint')


def bilinear_up2_adjoint(d, out, scale=1.0):
    """Adjoint of bilinear_up2: d [B,2h,2w,C] -> out [B,h,w,C] (times scale)."""
    b, h, w, c = out.shape
    assert tuple(d.shape) == (b, 2 * h, 2 * w, c)
    check(_lib.lib().b200ir_bilinear_up2_adjoint(_ptr(d), _ptr(out), b, h, w, c, scale, _stream()), 'bilinear_up2_adjoint')


def first_conv_wgrad(x, dz):
    """Weight gradient of first_conv: x fp32 NCHW [B,3,H,W], dz NHWC fp16 [B,H,W,Cout] -> fp32 [Cout, 3]."""
    b, _, h, w = x.shape
    cout = dz.shape[3]
    _req(x, torch.float32, 'x')
    dw = torch.empty(cout, 3, device=x.device, dtype=torch.float32)
    check(_lib.lib().b200ir_first_conv_wgrad(_ptr(x), _ptr(dz), _ptr(dw), b, h, w, cout, _stream()), 'first_conv_wgrad')
    return dw


# ------------------------------------------------------------------------------------------ training step (train_ops.cu)
def _zeros32(dev, *shape):
    return torch.zeros(*shape, device=dev, dtype=torch.float32)


def sft_mod(a, scale, shift, s_next, out):
    """SFT + next modulation, training forward (b200ir_sft_mod): a, out NHWC fp16 [B,h,w,C]; scale / shift [B,h,w,c_sft] or None."""
    b, h, w, c = a.shape
    check(_lib.lib().b200ir_sft_mod(_ptr(a), _ptr(scale), _ptr(shift), scale.shape[3] if scale is not None else 0,
                                    _ptr(s_next), _ptr(out), b, h * w, c, _stream()), 'sft_mod')


def sft_mod_bwd(g, a, scale, shift, s_next, da, accumulate, dscale, dshift, ds, a_broadcast=False):
    """b200ir_sft_mod_bwd: g NHWC fp16 [B,h,w,C]; a [B,h,w,C] (or [h,w,C] shared by all images when a_broadcast)."""
    b, h, w, c = g.shape
    check(_lib.lib().b200ir_sft_mod_bwd(_ptr(g), _ptr(a), 0 if a_broadcast else h * w * c, _ptr(scale), _ptr(shift),
                                        scale.shape[3] if scale is not None else 0, _ptr(s_next), _ptr(da),
                                        1 if accumulate else 0, _ptr(dscale), _ptr(dshift), _ptr(ds), b, h * w, c, _stream()),
          'sft_mod_bwd')


def style_act_bwd(da, a, noise, noise_gain, bias, oscale, mul, out, dd):
    b, h, w, c = a.shape
    check(_lib.lib().b200ir_style_act_bwd(_ptr(da), _ptr(a), _ptr(noise), h * w, _ptr(noise_gain), _ptr(bias), _ptr(oscale),
                                          float(mul), _ptr(out), _ptr(dd), b, h * w, c, _stream()), 'style_act_bwd')


def to_rgb_bwd(drgb, a, w, s, da, accumulate, ds):
    b, h, wd, c = a.shape
    _req(drgb, torch.float32, 'drgb')
    check(_lib.lib().b200ir_to_rgb_bwd(_ptr(drgb), _ptr(a), _ptr(w), _ptr(s), _ptr(da), 1 if accumulate else 0, _ptr(ds), b,
                                       h * wd, c, _stream()), 'to_rgb_bwd')


def rgb_up_adjoint(d, out):
    b, c, h, w = out.shape
    _req(d, torch.float32, 'd')
    check(_lib.lib().b200ir_rgb_up_adjoint(_ptr(d), _ptr(out), b * c, h, w, _stream()), 'rgb_up_adjoint')


def demod_bwd(ds, s, dd, d, wsq, scale2):
    b, cin = ds.shape
    check(_lib.lib().b200ir_demod_bwd(_ptr(ds), _ptr(s), _ptr(dd), _ptr(d), _ptr(wsq), float(scale2), b, cin, wsq.shape[0],
                                      _stream()), 'demod_bwd')


def mod_linear_bwd(ds, w, wscale, dlat, lat_idx):
    b, L, f = dlat.shape
    check(_lib.lib().b200ir_mod_linear_bwd(_ptr(ds), _ptr(w), float(wscale), _ptr(dlat), L, f, lat_idx, b, ds.shape[1],
                                           _stream()), 'mod_linear_bwd')


def first_conv_dgrad(dz, w, dx, accumulate=False):
    b, h, wd, cout = dz.shape
    check(_lib.lib().b200ir_first_conv_dgrad(_ptr(dz), _ptr(w), _ptr(dx), 1 if accumulate else 0, b, h, wd, cout, _stream()),
          'first_conv_dgrad')


def head_to_nchw(head, rgb):
    b, h, w, cpad = head.shape
    check(_lib.lib().b200ir_head_to_nchw(_ptr(head), _ptr(rgb), b, h * w, cpad, _stream()), 'head_to_nchw')


def nchw_to_head(drgb, dhead):
    b, h, w, cpad = dhead.shape
    _req(drgb, torch.float32, 'drgb')
    check(_lib.lib().b200ir_nchw_to_head(_ptr(drgb), _ptr(dhead), b, h * w, cpad, _stream()), 'nchw_to_head')


def l1_loss(x, t, weight, grad_scale, loss, grad):
    """loss[0] += weight * mean|x - t|; grad = grad_scale * weight / n * sign(x - t)  (fp32, contiguous, same shape)."""
    _req(x, torch.float32, 'x')
    _req(t, torch.float32, 't')
    assert x.shape == t.shape
    check(_lib.lib().b200ir_l1_loss(_ptr(x), _ptr(t), x.numel(), float(weight), float(grad_scale), _ptr(loss), _ptr(grad),
                                    _stream()), 'l1_loss')


def softplus_loss(pred, sign, weight, grad_scale, loss, dpred):
    """GANLoss('wgan_softplus') on fp16 scores [B, 1] (contiguous): sign = -1 for target_is_real."""
    n = pred.shape[0]
    stride = pred.stride(0)
    assert dpred is None or dpred.stride(0) == stride, 'pred and dpred must share the element stride'
    check(_lib.lib().b200ir_softplus_loss(_ptr(pred), n, stride, float(sign), float(weight), float(grad_scale), _ptr(loss),
                                          _ptr(dpred), _stream()), 'softplus_loss')


def pack_weights(w, scale, mode=0, cin_pad=0):
    """b200ir_pack_weights: fp32 [cout, cin, kh, kw] -> fp16 [cout, taps*cin_pad] (mode 0) / [cin, taps*cout] (mode 1)."""
    cout, cin, kh, kw = w.shape
    _req(w, torch.float32, 'w')
    if mode == 0:
        out = torch.empty(cout, kh * kw * (cin_pad or cin), device=w.device, dtype=torch.float16)
    else:
        out = torch.empty(cin, kh * kw * cout, device=w.device, dtype=torch.float16)
    check(_lib.lib().b200ir_pack_weights(_ptr(w), _ptr(out), cout, cin, kh, kw, float(scale), mode, cin_pad, _stream()),
          'pack_weights')
    return out


def sum_squares(x, scale, out):
    _req(x, torch.float32, 'x')
    check(_lib.lib().b200ir_sum_squares(_ptr(x), x.numel(), float(scale), _ptr(out), _stream()), 'sum_squares')


def minibatch_stddev(x, group):
    """b200ir_minibatch_stddev: x NHWC fp16 [B,h,w,C] -> [B,h,w,c_pad] = concat(x, group statistic, zero padding)."""
    b, h, w, c = x.shape
    c_pad = (c + 1 + 15) // 16 * 16
    out = torch.zeros(b, h, w, c_pad, device=x.device, dtype=torch.float16)
    s_buf = torch.empty(b // group, device=x.device, dtype=torch.float32)
    check(_lib.lib().b200ir_minibatch_stddev(_ptr(x), _ptr(s_buf), _ptr(out), b, h * w, c, c_pad, group, _stream()),
          'minibatch_stddev')
    return out


def minibatch_stddev_bwd(x, dcat, ds, dx, group):
    b, h, w, c = x.shape
    check(_lib.lib().b200ir_minibatch_stddev_bwd(_ptr(x), _ptr(dcat), _ptr(ds), _ptr(dx), b, h * w, c, dcat.shape[3], group,
                                                 _stream()), 'minibatch_stddev_bwd')


def minibatch_stddev_jvp(x, t, group):
    """Tangent of minibatch_stddev's output along t (the tangent of x): [B,h,w,c_pad]."""
    b, h, w, c = x.shape
    c_pad = (c + 1 + 15) // 16 * 16
    tcat = torch.empty(b, h, w, c_pad, device=x.device, dtype=torch.float16)
    ts = torch.empty(b // group, device=x.device, dtype=torch.float32)
    check(_lib.lib().b200ir_minibatch_stddev_jvp(_ptr(x), _ptr(t), _ptr(ts), _ptr(tcat), b, h * w, c, c_pad, group, _stream()),
          'minibatch_stddev_jvp')
    return tcat


def minibatch_stddev_hvp(x, t, a, group):
    """Hessian of (a . statistic)(x) times t: NHWC fp16 [B,h,w,C]; a fp32 [B / group]."""
    b, h, w, c = x.shape
    q = torch.empty_like(x)
    _req(a, torch.float32, 'a')
    check(_lib.lib().b200ir_minibatch_stddev_hvp(_ptr(x), _ptr(t), _ptr(a), _ptr(q), b, h * w, c, group, _stream()),
          'minibatch_stddev_hvp')
    return q


def maxpool2_relu(z, out):
    b, h, w, c = z.shape
    check(_lib.lib().b200ir_maxpool2_relu(_ptr(z), _ptr(out), b, h, w, c, _stream()), 'maxpool2_relu')


def maxpool2_relu_bwd(z, dpool, add, dz):
    b, h, w, c = z.shape
    check(_lib.lib().b200ir_maxpool2_relu_bwd(_ptr(z), _ptr(dpool), _ptr(add), _ptr(dz), b, h, w, c, _stream()),
          'maxpool2_relu_bwd')


def l1_loss_f16(x, t, weight, grad_scale, loss, grad):
    _req(x, torch.float16, 'x')
    _req(t, torch.float16, 't')
    assert x.shape == t.shape
    check(_lib.lib().b200ir_l1_loss_f16(_ptr(x), _ptr(t), x.numel(), float(weight), float(grad_scale), _ptr(loss), _ptr(grad),
                                        _stream()), 'l1_loss_f16')


def gram_batched(x, dy=None, out=None):
    """out[b] = sum over pixels of dy[b,p,:] (x) x[b,p,:] (fp32 [B, cout, cin]); dy = x: the un-normalised Gram matrices of
    PerceptualLoss._gram_mat for the whole batch in one launch (b200ir_gram_batched)."""
    dy = x if dy is None else dy
    b, h, w, cin = x.shape
    cout = dy.shape[3]
    _req(x, torch.float16, 'x')
    _req(dy, torch.float16, 'dy')
    if out is None:
        out = torch.empty(b, cout, cin, device=x.device, dtype=torch.float32)
    check(_lib.lib().b200ir_gram_batched(_ptr(x), _ptr(dy), _ptr(out), b, h, w, cin, cout, _stream()), 'gram_batched')
    return out


def style_act_bwd_params(da, a, noise, noise_gain, bias, oscale, mul, out, dd, db, dn):
    """style_act_bwd + the per-image reductions for StyleConv.activate.bias (db) and the noise gain (dn): fp32 [B, C], zeroed by the caller."""
    b, h, w, c = a.shape
    check(_lib.lib().b200ir_style_act_bwd_params(_ptr(da), _ptr(a), _ptr(noise), h * w, _ptr(noise_gain), _ptr(bias), _ptr(oscale),
                                                 float(mul), _ptr(out), _ptr(dd), _ptr(db), _ptr(dn), b, h * w, c, _stream()),
          'style_act_bwd_params')


def to_rgb_bwd_params(drgb, a, w, s, da, accumulate, ds, R):
    """to_rgb_bwd + R[b, o, c] += sum_p drgb[b, o, p] a[b, p, c] (fp32 [B, 3, C], zeroed by the caller)."""
    b, h, wd, c = a.shape
    _req(drgb, torch.float32, 'drgb')
    check(_lib.lib().b200ir_to_rgb_bwd_params(_ptr(drgb), _ptr(a), _ptr(w), _ptr(s), _ptr(da), 1 if accumulate else 0, _ptr(ds),
                                              _ptr(R), b, h * wd, c, _stream()), 'to_rgb_bwd_params')


def plane_sums(x, out):
    """out[c] += sum over batch and pixels of fp32 NCHW x[b, c, :, :]."""
    b, c, h, w = x.shape
    _req(x, torch.float32, 'x')
    check(_lib.lib().b200ir_plane_sums(_ptr(x), _ptr(out), b, c, h * w, _stream()), 'plane_sums')


def table_colsum(tab, out=None, mul=None, scale=1.0):
    """out[j] = scale * sum_b tab[b, j] * (mul[b, j % m] if mul is given): tab fp32 or fp16 [B, ...], mul fp32 [B, m] -> fp32 [n]."""
    b = tab.shape[0]
    n = tab.numel() // b
    assert tab.is_contiguous() and tab.dtype in (torch.float32, torch.float16)
    if out is None:
        out = torch.empty(n, device=tab.device, dtype=torch.float32)
    m = mul.shape[1] if mul is not None else 0
    check(_lib.lib().b200ir_table_colsum(_ptr(tab), 1 if tab.dtype == torch.float16 else 0, _ptr(mul), m, float(scale), _ptr(out),
                                         b, n, _stream()), 'table_colsum')
    return out


def mod_linear_wgrad(ds, latent, wscale, lat_idx):
    """Weight gradient of a modulation EqualLinear: ds fp32 [B, cin], latent fp32 [B, L, F] -> fp32 [cin, F]."""
    b, L, f = latent.shape
    dw = torch.empty(ds.shape[1], f, device=ds.device, dtype=torch.float32)
    check(_lib.lib().b200ir_mod_linear_wgrad(_ptr(ds), _ptr(latent), float(wscale), _ptr(dw), L, f, lat_idx, b, ds.shape[1],
                                             _stream()), 'mod_linear_wgrad')
    return dw


def modconv_wgrad(G, transposed, w_raw, s, dd, d, scale):
    """ModulatedConv2d weight gradient in the reference layout [cout, cin, kh, kw] (see include/b200ir.h)."""
    cout, cin, kh, kw = w_raw.shape
    _req(w_raw, torch.float32, 'w_raw')
    _req(G, torch.float32, 'G')
    dw = torch.empty_like(w_raw)
    check(_lib.lib().b200ir_modconv_wgrad(_ptr(G), 1 if transposed else 0, _ptr(w_raw), _ptr(s), _ptr(dd), _ptr(d), float(scale),
                                          _ptr(dw), s.shape[0], cin, cout, kh * kw, _stream()), 'modconv_wgrad')
    return dw
