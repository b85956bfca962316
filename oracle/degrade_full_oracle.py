"""TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline) -- never imported by the product.

CPU restatement of the LQ synthesis of FFHQDegradationDataset.__getitem__
(Car_Plate-Restoration/basicsr/data/ffhq_degradation_dataset.py:242-311) for the stages b200ir_degrade_full runs, as a
deterministic function of explicit parameters (the random draws are made by the caller):

    blur    'iso' / 'aniso' / 'generalized_*' / 'plateau_*' / 'motion' / 'average': cv2.filter2D(img, -1, kernel)
            (degradations.py:460-515); 'median': cv2.medianBlur; 'bilateral': cv2.bilateralFilter (:353-361)
            'pyblur': scipy convolve2d on the uint8 image, truncated to uint8 (oracle/pyblur_oracle.py; explicit
                      arithmetic: convolve2d_same_fill)
    down    cv2.resize(img, (lr_w, lr_h), INTER_LINEAR)                         (:255-256)
    noise   clip(img + noise, 0, 1)                                             (degradations.py:660-669)
    JPEG    add_jpg_compression                                                 (degradations.py:876-892)
    up      cv2.resize(img, (W, H), INTER_LINEAR)                               (:272)
    jitter  clip(img + shift, 0, 1)                                             (:90-95)
    gray    cv2.cvtColor(BGR2GRAY), tiled                                       (:283-285)
    jitter2 color_jitter_pt: torchvision adjust_* in the drawn order            (:187-207, :290-296)
    tail    clamp(round(x * 255), 0, 255) / 255; (x - 0.5) / 0.5; BGR -> RGB; CHW   (:288, :307-311)

The heavy lifting is the same third-party calls the reference makes (OpenCV 4.13 here, pinned 4.6.0.66 in
requirements.txt:24).  Two stages also exist as explicit arithmetic because the CUDA kernel has to match them bit for
bit: the JPEG round trip (oracle/jpeg_oracle.py, bit-exact against cv2) and the filter2D sum (`filter2d_direct`: fp32
products and adds in kernel order, BORDER_REFLECT_101 -- OpenCV itself uses a DFT for kernels >= 11x11, so its result
differs from ANY direct sum in the last bits; tests bound that difference).

Pinning: tests/test_degrade_full_cpu.py runs the reference's own functions (imported from /root/reference, seeded)
beside `image_restoration_b200.degradation.sample_params` + this module and requires identical results;
tests/golden/degrade_full.npz holds reference outputs for the GPU box (tests/golden/make_golden_degrade_full.py).
"""
import cv2
import numpy as np

from . import jpeg_oracle, pyblur_oracle


def filter2d_direct(img, kernel):
    """cv2.filter2D(img, -1, kernel) for float32 HxWx3: correlation, anchor at the centre, BORDER_REFLECT_101.  fp32
    product and add per tap, taps in row-major order, zero taps skipped (adding an exact zero changes nothing)."""
    k = np.asarray(kernel, dtype=np.float32)
    r = k.shape[0] // 2
    pad = np.pad(img, ((r, r), (r, r), (0, 0)), mode='reflect')
    h, w = img.shape[:2]
    acc = np.zeros_like(img, dtype=np.float32)
    for i in range(k.shape[0]):
        for j in range(k.shape[1]):
            if k[i, j] != 0:
                acc = acc + k[i, j] * pad[i:i + h, j:j + w]
    return acc


def convolve2d_same_fill(img_u8, kernel, fill=255.0):
    """scipy.signal.convolve2d(float32(img), kernel, mode='same', fillvalue=fill) per channel, as explicit arithmetic in
    the type scipy computes in (result_type of float32 and the kernel's dtype).  scipy 1.18 (this container) walks kernel
    rows in ascending order; within a row it adds blocks of four columns as ((p0 + p1) + p2) + p3 to the running sum and
    the remaining columns one at a time; products and sums are rounded separately.  Found by probing and pinned
    bit-exactly against scipy in tests/test_degrade_full_cpu.py."""
    k = np.asarray(kernel)
    dt = np.result_type(np.float32, k.dtype)
    n = k.shape[0]
    c = (n - 1) // 2
    H, W = img_u8.shape[:2]
    pad = np.pad(img_u8.astype(dt), ((c, c), (c, c), (0, 0)), constant_values=fill)
    k = k.astype(dt)
    acc = np.zeros(img_u8.shape, dt)

    def prod(i, j):
        return k[i, j] * pad[2 * c - i:2 * c - i + H, 2 * c - j:2 * c - j + W]
    for i in range(n):
        j = 0
        while j + 4 <= n:
            if np.any(k[i, j:j + 4] != 0):
                acc = acc + (((prod(i, j) + prod(i, j + 1)) + prod(i, j + 2)) + prod(i, j + 3))
            j += 4
        while j < n:
            if k[i, j] != 0:
                acc = acc + prod(i, j)
            j += 1
    return acc


def _fma32(a, b, c):
    """fp32 fused multiply-add: the float64 product of two fp32 values is exact, so one float64 add + one rounding to
    fp32 reproduces it (up to double-rounding cases of probability ~2^-29)."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def resize_linear(img, dsize):
    """cv2.resize(img, dsize, interpolation=INTER_LINEAR) for float32 HxWxC as this container's OpenCV 4.13 executes it
    (Intel IPP): source coordinate (d + 0.5) * src / dst - 0.5 in float64, fraction rounded to fp32, indices clamped,
    each pass fma(S1 - S0, f, S0), horizontal pass then vertical.  Found by probing; pinned bit-exactly against
    cv2.resize in tests/test_degrade_full_cpu.py.  This is what the CUDA kernels evaluate."""
    W, H = dsize
    h, w = img.shape[:2]

    def coords(dst_n, src_n):
        f = (np.arange(dst_n) + 0.5) * (np.float64(src_n) / dst_n) - 0.5
        s = np.floor(f)
        return (np.clip(s, 0, src_n - 1).astype(int), np.clip(s + 1, 0, src_n - 1).astype(int), (f - s).astype(np.float32))

    def lerp(a, b, f):
        return _fma32(b - a, np.broadcast_to(f, a.shape), a)
    x0, x1, fx = coords(W, w)
    y0, y1, fy = coords(H, h)
    t = lerp(img[:, x0], img[:, x1], fx[None, :, None])
    return lerp(t[y0], t[y1], fy[:, None, None])


def median_u8(img_u8, k):
    """cv2.medianBlur(uint8 HxWx3, k): exact median of the k x k window per channel, BORDER_REPLICATE."""
    r = k // 2
    pad = np.pad(img_u8, ((r, r), (r, r), (0, 0)), mode='edge')
    win = np.lib.stride_tricks.sliding_window_view(pad, (k, k), axis=(0, 1))
    h, w = img_u8.shape[:2]
    return np.sort(win.reshape(h, w, 3, -1), axis=-1)[..., (k * k) // 2].astype(np.uint8)


def bilateral_u8(img_u8, d, sigma):
    """cv2.bilateralFilter(uint8 HxWx3, d, sigma, sigma) (bilateral_filter.dispatch.cpp / .simd.hpp, 8-bit 3-channel):
    BORDER_REFLECT_101; taps with sqrt(i^2 + j^2) <= d // 2 in row-major order; colour table (float)exp(n^2 * gc) over
    n = |db| + |dg| + |dr|, space weights (float)exp(r^2 * gs), both evaluated in double; weight = space * colour in fp32,
    sum = fma(value, weight, sum), result cvRound(sum / wsum).  Pinned against cv2 in tests/test_degrade_full_cpu.py."""
    radius = max(d // 2, 1)
    gc = gs = -0.5 / (float(sigma) * float(sigma))
    n = np.arange(256 * 3, dtype=np.float64)
    cw = np.exp(n * n * gc).astype(np.float32)
    pad = np.pad(img_u8.astype(np.int32), ((radius, radius), (radius, radius), (0, 0)), mode='reflect')
    h, w = img_u8.shape[:2]
    acc = np.zeros((h, w, 3), np.float32)
    wsum = np.zeros((h, w), np.float32)
    c = img_u8.astype(np.int32)
    for i in range(-radius, radius + 1):
        for j in range(-radius, radius + 1):
            rr = np.sqrt(float(i * i + j * j))
            if rr > radius:
                continue
            sw = np.float32(np.exp(rr * rr * gs))
            v = pad[radius + i:radius + i + h, radius + j:radius + j + w]
            wt = (sw * cw[np.abs(v - c).sum(axis=2)]).astype(np.float32)
            wsum = wsum + wt
            acc = _fma32(v.astype(np.float32), np.broadcast_to(wt[..., None], v.shape), acc)
    return np.clip(np.rint(acc / wsum[..., None]), 0, 255).astype(np.uint8)


def gray_bgr(img):
    """cv2.cvtColor(img, COLOR_BGR2GRAY) on float32 as this container's OpenCV evaluates it:
    fma(r, 0.299, fma(b, 0.114, g * 0.587)) (found by comparing all orderings against cv2; bit-exact)."""
    b, g, r = (img[..., i].astype(np.float64) for i in range(3))
    c = [float(np.float32(v)) for v in (0.114, 0.587, 0.299)]
    t = (g * c[1]).astype(np.float32).astype(np.float64)
    t = (b * c[0] + t).astype(np.float32).astype(np.float64)
    return (r * c[2] + t).astype(np.float32)


def color_jitter_pt(img_bgr, steps):
    """FFHQDegradationDataset.color_jitter_pt (ffhq_degradation_dataset.py:187-207) with the drawn (op, factor) list
    made explicit: torchvision's adjust_brightness / contrast / saturation / hue (the reference's own library calls;
    torchvision 0.14.0 pinned in requirements.txt:45, 0.26 here) on the RGB CHW tensor img2tensor builds (:288)."""
    import torch
    import torchvision.transforms.functional as TF
    t = torch.from_numpy(np.ascontiguousarray(img_bgr[..., ::-1].transpose(2, 0, 1))).float()
    fns = (TF.adjust_brightness, TF.adjust_contrast, TF.adjust_saturation, TF.adjust_hue)
    for op, f in steps:
        t = fns[op](t, f)
    return np.ascontiguousarray(t.numpy().transpose(1, 2, 0)[..., ::-1])


def lq_image(gt_u8, mode, kernel, lr_size, noise=None, quality=0, jitter=None, gray=0, exact_blur=True, lib_jpeg=False,
             bilateral_sigma=0.0, cj=None):
    """uint8 BGR [H,W,3] (or the float32 image in [0,1] the dataset holds after its cv2.resize) -> float32 BGR [H,W,3] LQ
    image before the 8-bit rounding, plus the LR image after noise/JPEG."""
    H, W = gt_u8.shape[:2]
    if gt_u8.dtype == np.uint8:
        img = gt_u8.astype(np.float32) / np.float32(255.)
    else:       # random_pyblur / median_blur / bilateral_blur quantise with np.array(img * 255.0, dtype=np.uint8)
        img = gt_u8
        gt_u8 = np.array(img * 255.0, dtype=np.uint8)
    if mode == 2:
        k = np.asarray(kernel, dtype=np.float32)
        img = filter2d_direct(img, k) if exact_blur else cv2.filter2D(img, -1, np.asarray(kernel))
    elif mode == 1:     # the kernel's dtype decides the arithmetic type, as in the reference (float64 box / disk / line)
        blur = convolve2d_same_fill(gt_u8, kernel) if exact_blur else pyblur_oracle.blur_f32(gt_u8, np.asarray(kernel))
        img = blur.astype('uint8').astype(np.float32) / np.float32(255.)
    elif mode == 3:     # median_blur (degradations.py:353-355)
        k = np.asarray(kernel).shape[0]
        blur = median_u8(gt_u8, k) if exact_blur else cv2.medianBlur(gt_u8, k)
        img = blur.astype(np.float32) / np.float32(255.)
    elif mode == 4:     # bilateral_blur (degradations.py:358-361)
        k = np.asarray(kernel).shape[0]
        blur = (bilateral_u8(gt_u8, k, bilateral_sigma) if exact_blur
                else cv2.bilateralFilter(gt_u8, k, bilateral_sigma, bilateral_sigma))
        img = blur.astype(np.float32) / np.float32(255.)
    elif mode == 5:     # bicubic (degradations.py:379-385): Pillow round trip of np.array(img * 255.0, uint8)
        if exact_blur:
            blur = pil_bicubic_roundtrip(gt_u8)
        else:
            from PIL import Image
            pil = Image.fromarray(gt_u8)
            blur = np.asarray(pil.resize((W // 4, H // 4), Image.BICUBIC).resize((W, H), Image.BICUBIC))
        img = np.array(blur, dtype=np.float32) / 255.0
    lr = cv2.resize(img, tuple(lr_size), interpolation=cv2.INTER_LINEAR)
    if noise is not None:
        lr = np.clip(lr + noise, 0, 1)
    if quality:
        if lib_jpeg:
            enc = cv2.imencode('.jpg', np.clip(lr, 0, 1) * 255.0, [int(cv2.IMWRITE_JPEG_QUALITY), int(quality)])[1]
            lr = np.float32(cv2.imdecode(enc, 1)) / 255.0
        else:
            lr = jpeg_oracle.add_jpg_compression(lr, quality)
    up = cv2.resize(lr, (W, H), interpolation=cv2.INTER_LINEAR)
    if jitter is not None and np.any(np.asarray(jitter) != 0):
        up = np.clip(up + np.asarray(jitter, dtype=np.float32), 0, 1)
    if gray:
        up = np.tile(gray_bgr(up)[:, :, None], [1, 1, 3])
    if cj:
        up = color_jitter_pt(up, cj)
    return up, lr


# ---- 'bicubic' kind (degradations.py:379-385): Pillow's 8-bit ImagingResample (src/libImaging/Resample.c) restated.
_PIL_PRECISION_BITS = 32 - 8 - 2


def _pil_bicubic_filter(x):
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def pil_bicubic_tables(in_size, out_size):
    """precompute_coeffs + normalize_coeffs_8bpc for the bicubic filter (support 2): per output index (xmin, n) and the
    fixed-point coefficients (int)(+-0.5 + w / sum(w) * 2^22); all in double, in Pillow's order of operations."""
    import math
    scale = float(np.float32(in_size) - np.float32(0)) / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int64)
    kk = np.zeros((out_size, ksize), np.int64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [_pil_bicubic_filter((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x, v in enumerate(w):
            if ww != 0.0:
                v = v / ww
            kk[xx, x] = int(-0.5 + v * (1 << _PIL_PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << _PIL_PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _pil_resample_axis(img, out_size, axis):
    bounds, kk = pil_bicubic_tables(img.shape[axis], out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.zeros((out_size,) + src.shape[1:], np.int64)
    for xx in range(out_size):
        xmin, n = bounds[xx]
        acc = np.full(src.shape[1:], 1 << (_PIL_PRECISION_BITS - 1), np.int64)
        for x in range(n):
            acc = acc + src[xmin + x] * int(kk[xx, x])
        out[xx] = np.clip(acc >> _PIL_PRECISION_BITS, 0, 255)
    return np.moveaxis(out, 0, axis).astype(np.uint8)


def pil_bicubic_resize(img_u8, h, w):
    """Image.fromarray(img).resize((w, h), BICUBIC) on a uint8 HWC image: horizontal pass first, then vertical."""
    t = _pil_resample_axis(img_u8, w, 1) if w != img_u8.shape[1] else img_u8
    return _pil_resample_axis(t, h, 0) if h != img_u8.shape[0] else t


def pil_bicubic_roundtrip(img_u8):
    """degradations.bicubic on the uint8 image: x 1/4 and back to the original size."""
    h, w = img_u8.shape[:2]
    return pil_bicubic_resize(pil_bicubic_resize(img_u8, h // 4, w // 4), h, w)


def apply_random_mask(up, mask_mode, mask):
    """FFHQDegradationDataset.random_mask applied to the image of __getitem__ (ffhq_degradation_dataset.py:299-303) once its
    shapes are drawn (degradation.random_mask_draw mirrors the draws): tensor2img(out_type=float32) clamps to [0, 1]; the
    regular / half kinds (:95-110, :163-186) write 1.0 into the masked pixels; the irregular kind (:112-151) takes
    np.array(img * 255.0, dtype=np.uint8) of the whole image, draws 255 and returns mask.astype('float32') / 255.0."""
    img = np.clip(up, 0, 1).astype(np.float32)
    if mask_mode == 2:
        m = np.array(img * 255.0, dtype=np.uint8)
        m[np.asarray(mask) != 0] = 255
        return m.astype('float32') / 255.0
    if mask_mode == 1:
        img = img.copy()
        img[np.asarray(mask) != 0] = 1.0
    return img


def lq_tensor(up, bgr2rgb=True):
    """Tail of __getitem__: 8-bit grid, normalise with mean = std = 0.5, HWC -> CHW (numpy float32)."""
    x = np.clip(np.rint(np.clip(up, 0, 1) * np.float32(255.)), 0, 255) / np.float32(255.)
    x = (x - np.float32(0.5)) / np.float32(0.5)
    if bgr2rgb:
        x = x[..., ::-1]
    return np.ascontiguousarray(x.transpose(2, 0, 1)).astype(np.float32)


def degrade_full(gt_u8, mode, kernel, lr_size, noise=None, quality=0, jitter=None, gray=0, bgr2rgb=True, exact_blur=True,
                 lib_jpeg=False, bilateral_sigma=0.0, cj=None, mask_mode=0, mask=None):
    up, lr = lq_image(gt_u8, mode, kernel, lr_size, noise, quality, jitter, gray, exact_blur, lib_jpeg, bilateral_sigma, cj)
    if mask_mode:
        up = apply_random_mask(up, mask_mode, mask)
    return lq_tensor(up, bgr2rgb), lr
