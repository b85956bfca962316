"""TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline) -- never imported by the product.

CPU restatement of the JPEG round trip the reference applies to every low-quality crop:
    degradations.add_jpg_compression (Car_Plate-Restoration/basicsr/data/degradations.py:876-892)
        img = clip(img, 0, 1); cv2.imencode('.jpg', img * 255., [IMWRITE_JPEG_QUALITY, int(quality)]); cv2.imdecode(..., 1) / 255.
The arithmetic lives in a third-party dependency that is not under /root/reference: OpenCV (pinned opencv-python==4.6.0.66,
requirements.txt:24; this container: 4.13.0) which bundles libjpeg-turbo (here 3.1.2).  Entropy coding is lossless, so
encode + decode is: 8-bit conversion (cv::saturate_cast: round half to even) -> RGB->YCbCr (jccolor.c, 16-bit fixed
point) -> edge replication to whole 16x16 MCUs -> 2x2 chroma down-sampling with the alternating 1,2 bias (jcsample.c
h2v2_downsample) -> forward DCT (jfdctint.c, "islow") -> quantisation with the quality-scaled Annex K tables
(jcparam.c jpeg_quality_scaling / jpeg_add_quant_table, jcdctmgr.c) -> de-quantisation + inverse DCT (jidctint.c) ->
"fancy" triangle chroma up-sampling (jdsample.c h2v2_fancy_upsample) -> YCbCr->RGB (jdcolor.c).  All integer.

Pinning: tests/test_jpeg_cpu.py compares this restatement BIT-EXACTLY with cv2.imencode / cv2.imdecode run in this
container over random and smooth images, odd sizes and qualities 1..100, and tests/golden/jpeg_roundtrip.npz holds
cv2's outputs for the GPU box (made by tests/golden/make_golden_jpeg.py).
"""
import numpy as np

STD_LUMA = np.array([
    16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56,
    14, 17, 22, 29, 51, 87, 80, 62, 18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92,
    49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99], dtype=np.int64).reshape(8, 8)
STD_CHROMA = np.array([
    17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99,
    47, 66, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
    99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99], dtype=np.int64).reshape(8, 8)

F_0_298631336, F_0_390180644, F_0_541196100, F_0_765366865 = 2446, 3196, 4433, 6270
F_0_899976223, F_1_175875602, F_1_501321110, F_1_847759065 = 7373, 9633, 12299, 15137
F_1_961570560, F_2_053119869, F_2_562915447, F_3_072711026 = 16069, 16819, 20995, 25172
CONST_BITS, PASS1_BITS = 13, 2


def quant_tables(quality):
    """jcparam.c: jpeg_quality_scaling + jpeg_add_quant_table(force_baseline=TRUE)."""
    q = int(quality)
    q = 1 if q <= 0 else (100 if q > 100 else q)
    scale = 5000 // q if q < 50 else 200 - q * 2
    out = []
    for base in (STD_LUMA, STD_CHROMA):
        t = (base * scale + 50) // 100
        out.append(np.clip(t, 1, 255))
    return out


def _fix(x):
    return int(x * 65536 + 0.5)


def rgb_to_ycc(rgb):
    """jccolor.c rgb_ycc_convert: uint8 [H,W,3] (R,G,B) -> int64 planes Y, Cb, Cr."""
    r, g, b = (rgb[..., i].astype(np.int64) for i in range(3))
    half, off = 1 << 15, 128 << 16
    y = (_fix(0.29900) * r + _fix(0.58700) * g + _fix(0.11400) * b + half) >> 16
    cb = (-_fix(0.16874) * r - _fix(0.33126) * g + _fix(0.50000) * b + off + half - 1) >> 16
    cr = (_fix(0.50000) * r - _fix(0.41869) * g - _fix(0.08131) * b + off + half - 1) >> 16
    return y, cb, cr


def _pad_edge(p, mh, mw):
    h, w = p.shape
    H, W = -(-h // mh) * mh, -(-w // mw) * mw
    return np.pad(p, ((0, H - h), (0, W - w)), mode='edge')


def h2v2_downsample(p):
    """jcsample.c h2v2_downsample: (sum of the 2x2 block + bias) >> 2, bias 1, 2, 1, 2 ... along each row."""
    s = p[0::2, 0::2] + p[0::2, 1::2] + p[1::2, 0::2] + p[1::2, 1::2]
    bias = np.where(np.arange(s.shape[1]) % 2 == 0, 1, 2)[None, :]
    return (s + bias) >> 2


def _descale(x, n):
    return (x + (1 << (n - 1))) >> n


def _fdct_1d(d, first):
    """One pass of jfdctint.c jpeg_fdct_islow over the last axis (8 samples)."""
    d0, d1, d2, d3, d4, d5, d6, d7 = (d[..., i] for i in range(8))
    tmp0, tmp7, tmp1, tmp6 = d0 + d7, d0 - d7, d1 + d6, d1 - d6
    tmp2, tmp5, tmp3, tmp4 = d2 + d5, d2 - d5, d3 + d4, d3 - d4
    tmp10, tmp13, tmp11, tmp12 = tmp0 + tmp3, tmp0 - tmp3, tmp1 + tmp2, tmp1 - tmp2
    o = [None] * 8
    if first:
        o[0] = (tmp10 + tmp11) << PASS1_BITS
        o[4] = (tmp10 - tmp11) << PASS1_BITS
        sh = CONST_BITS - PASS1_BITS
    else:
        o[0] = _descale(tmp10 + tmp11, PASS1_BITS)
        o[4] = _descale(tmp10 - tmp11, PASS1_BITS)
        sh = CONST_BITS + PASS1_BITS
    z1 = (tmp12 + tmp13) * F_0_541196100
    o[2] = _descale(z1 + tmp13 * F_0_765366865, sh)
    o[6] = _descale(z1 + tmp12 * (-F_1_847759065), sh)
    z1, z2, z3, z4 = tmp4 + tmp7, tmp5 + tmp6, tmp4 + tmp6, tmp5 + tmp7
    z5 = (z3 + z4) * F_1_175875602
    tmp4, tmp5 = tmp4 * F_0_298631336, tmp5 * F_2_053119869
    tmp6, tmp7 = tmp6 * F_3_072711026, tmp7 * F_1_501321110
    z1, z2 = z1 * (-F_0_899976223), z2 * (-F_2_562915447)
    z3, z4 = z3 * (-F_1_961570560) + z5, z4 * (-F_0_390180644) + z5
    o[7] = _descale(tmp4 + z1 + z3, sh)
    o[5] = _descale(tmp5 + z2 + z4, sh)
    o[3] = _descale(tmp6 + z2 + z3, sh)
    o[1] = _descale(tmp7 + z1 + z4, sh)
    return np.stack(o, axis=-1)


def fdct_islow(blocks):
    """blocks int64 [..., 8(row), 8(col)] of samples - 128 -> coefficients scaled by 8."""
    t = _fdct_1d(blocks, True)                                    # pass 1: rows
    t = _fdct_1d(np.swapaxes(t, -1, -2), False)                   # pass 2: columns
    return np.swapaxes(t, -1, -2)


def quantize(coef, qtbl):
    """jcdctmgr.c quantize: divisor = qtbl << 3 (islow output is scaled by 8), round half away from zero."""
    q = qtbl << 3
    a = np.abs(coef)
    r = (a + (q >> 1)) // q
    return np.where(coef < 0, -r, r)


def _idct_1d(c, first):
    """One pass of jidctint.c jpeg_idct_islow over the last axis."""
    i0, i1, i2, i3, i4, i5, i6, i7 = (c[..., i] for i in range(8))
    z2, z3 = i2, i6
    z1 = (z2 + z3) * F_0_541196100
    tmp2 = z1 + z3 * (-F_1_847759065)
    tmp3 = z1 + z2 * F_0_765366865
    tmp0 = (i0 + i4) << CONST_BITS
    tmp1 = (i0 - i4) << CONST_BITS
    tmp10, tmp13, tmp11, tmp12 = tmp0 + tmp3, tmp0 - tmp3, tmp1 + tmp2, tmp1 - tmp2
    tmp0, tmp1, tmp2, tmp3 = i7, i5, i3, i1
    z1, z2, z3, z4 = tmp0 + tmp3, tmp1 + tmp2, tmp0 + tmp2, tmp1 + tmp3
    z5 = (z3 + z4) * F_1_175875602
    tmp0, tmp1 = tmp0 * F_0_298631336, tmp1 * F_2_053119869
    tmp2, tmp3 = tmp2 * F_3_072711026, tmp3 * F_1_501321110
    z1, z2 = z1 * (-F_0_899976223), z2 * (-F_2_562915447)
    z3, z4 = z3 * (-F_1_961570560) + z5, z4 * (-F_0_390180644) + z5
    tmp0, tmp1, tmp2, tmp3 = tmp0 + z1 + z3, tmp1 + z2 + z4, tmp2 + z2 + z3, tmp3 + z1 + z4
    sh = CONST_BITS - PASS1_BITS if first else CONST_BITS + PASS1_BITS + 3
    o = [_descale(tmp10 + tmp3, sh), _descale(tmp11 + tmp2, sh), _descale(tmp12 + tmp1, sh), _descale(tmp13 + tmp0, sh),
         _descale(tmp13 - tmp0, sh), _descale(tmp12 - tmp1, sh), _descale(tmp11 - tmp2, sh), _descale(tmp10 - tmp3, sh)]
    return np.stack(o, axis=-1)


def idct_islow(coef):
    """De-quantised coefficients [..., 8(row), 8(col)] -> samples 0..255 (range_limit of value + 128)."""
    t = _idct_1d(np.swapaxes(coef, -1, -2), True)                 # pass 1: columns
    t = _idct_1d(np.swapaxes(t, -1, -2), False)                   # pass 2: rows
    return np.clip(t + 128, 0, 255)


def _blocks(p):
    h, w = p.shape
    return p.reshape(h // 8, 8, w // 8, 8).transpose(0, 2, 1, 3)


def _unblocks(b):
    n, m = b.shape[:2]
    return b.transpose(0, 2, 1, 3).reshape(n * 8, m * 8)


def codec_plane(p, qtbl):
    c = quantize(fdct_islow(_blocks(p) - 128), qtbl)
    return _unblocks(idct_islow(c * qtbl))


def h2v2_fancy_upsample(c, ch, cw):
    """jdsample.c h2v2_fancy_upsample on the real ch x cw chroma samples (edges replicated: jdmainct.c context rows,
    first / last column special cases): 3/4 nearer + 1/4 farther per axis, rounding constants 8 (even output columns)
    and 7 (odd output columns)."""
    c = c[:ch, :cw]
    up = np.vstack([c[:1], c[:-1]])          # row above (row 0 duplicated)
    dn = np.vstack([c[1:], c[-1:]])          # row below (last row duplicated)
    out = np.empty((2 * ch, 2 * cw), dtype=np.int64)
    for v, far in ((0, up), (1, dn)):
        s = 3 * c + far                      # column sums
        left = np.hstack([s[:, :1], s[:, :-1]])
        right = np.hstack([s[:, 1:], s[:, -1:]])
        out[v::2, 0::2] = (3 * s + left + 8) >> 4
        out[v::2, 1::2] = (3 * s + right + 7) >> 4
    return out


def ycc_to_rgb(y, cb, cr):
    """jdcolor.c ycc_rgb_convert."""
    half = 1 << 15
    xb, xr = cb - 128, cr - 128
    r = y + ((_fix(1.40200) * xr + half) >> 16)
    g = y + ((-_fix(0.34414) * xb + half - _fix(0.71414) * xr) >> 16)
    b = y + ((_fix(1.77200) * xb + half) >> 16)
    return np.clip(np.stack([r, g, b], axis=-1), 0, 255).astype(np.uint8)


def jpeg_roundtrip_u8(bgr, quality):
    """uint8 BGR [H,W,3] -> uint8 BGR [H,W,3]: cv2.imdecode(cv2.imencode('.jpg', bgr, [IMWRITE_JPEG_QUALITY, q])[1], 1)."""
    h, w = bgr.shape[:2]
    ql, qc = quant_tables(quality)
    y, cb, cr = rgb_to_ycc(bgr[..., ::-1])
    # jcprepct.c pre_process_data: every colour-converted row is widened to whole MCUs by repeating its last sample
    # (expand_right_edge), the rows only to a whole row GROUP (2 rows, expand_bottom_edge on the colour buffer); after
    # down-sampling each component is then padded to a whole MCU height by repeating its last DOWN-SAMPLED row -- for
    # an even height the chroma padding is the average of the last two image rows, not the last row.
    y, cb, cr = (_pad_edge(p, 2, 16) for p in (y, cb, cr))
    yq = codec_plane(_pad_edge(y, 16, 16), ql)
    ch, cw = -(-h // 2), -(-w // 2)
    planes = []
    for p in (cb, cr):
        d = codec_plane(_pad_edge(h2v2_downsample(p), 8, 8), qc)
        planes.append(h2v2_fancy_upsample(d, ch, cw)[:h, :w])
    rgb = ycc_to_rgb(yq[:h, :w], planes[0], planes[1])
    return rgb[..., ::-1].copy()


def to_u8_saturate(img255):
    """cv::saturate_cast<uchar>(float): cvRound (round half to even), clamped -- what imencode does with a float image."""
    return np.clip(np.rint(img255), 0, 255).astype(np.uint8)


def add_jpg_compression(img, quality):
    """degradations.py:876-892 on a float32 BGR image in [0, 1]."""
    img = np.clip(img, 0, 1)
    u8 = to_u8_saturate(img * np.float32(255.0))
    return jpeg_roundtrip_u8(u8, int(quality)).astype(np.float32) / np.float32(255.0)
