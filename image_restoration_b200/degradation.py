"""Host side of the fused B200 degradation: mirrors the reference's pyblur operators
(Car_Plate-Restoration/pyblur/pyblur/{BoxBlur,DefocusBlur,LinearMotionBlur,PsfBlur,RandomizedBlur}.py) and the
blur -> downsample -> noise -> upsample -> quantise -> normalise part of FFHQDegradationDataset.__getitem__
(basicsr/data/ffhq_degradation_dataset.py:244-272,307-311), but on batches of crops resident on the GPU: the blur
kernels (a few hundred floats) are built on the host exactly as pyblur builds them, everything per-pixel runs in one
launch of libb200ir's degrade kernel.  There is no CPU path for the per-pixel work.

Same names and argument meaning as pyblur: BoxKernel(dim), DiskKernel(dim), LineKernel(dim, angle, linetype),
psfDictionary[id]; BoxBlur / DefocusBlur / LinearMotionBlur / PsfBlur / RandomizedBlur take a uint8 HxWx3 image (numpy,
PIL-compatible array, or a uint8 CUDA tensor [B,H,W,3]) and return the blurred uint8 image(s).

Differences from the reference, on purpose:
  * LineKernel does not mutate the shared line dictionary (the reference's 'left'/'right' calls overwrite the
    dictionary entry, LinearMotionBlur.py:37-43, so its output depends on call history); this implementation always
    behaves like the reference with a fresh dictionary.
"""
import ctypes as C
import math
import os

import numpy as np
import torch

from . import _lib

KERNEL_DIMS = [7, 9, 11, 13, 15, 17, 19, 21]      # boxKernelDims / defocusKernelDims / lineLengths
LINE_TYPES = ['full', 'right', 'left']            # LinearMotionBlur.py:11
_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data', 'psf_kernels.npz')


# ------------------------------------------------------------------------------------------ kernels (host, exact)
def BoxKernel(dim):
    """BoxBlur.py:21-26."""
    k = np.ones((dim, dim), dtype=np.float32)
    return k / np.count_nonzero(k)


def _ellipse_mask(dim):
    """skimage.draw.disk((dim/2, dim/2), dim/2) of scikit-image 0.19.3 (DefocusBlur.py:25-31): points of the
    bounding box [ceil(c-r), floor(c+r)] with ((y-c)/r)^2 + ((x-c)/r)^2 < 1, clipped to the kernel."""
    c = r = dim / 2
    lo, hi = int(math.ceil(c - r)), int(math.floor(c + r))
    ys, xs = np.mgrid[lo:hi + 1, lo:hi + 1].astype(np.float64)
    inside = ((ys - c) / r) ** 2 + ((xs - c) / r) ** 2 < 1
    m = np.zeros((dim, dim), dtype=bool)
    sel = inside & (ys < dim) & (xs < dim) & (ys >= 0) & (xs >= 0)
    m[ys[sel].astype(int), xs[sel].astype(int)] = True
    return m


def DiskKernel(dim):
    """DefocusBlur.py:25-39 (incl. the corner adjustment for dim 3 and 5)."""
    k = np.zeros((dim, dim), dtype=np.float32)
    k[_ellipse_mask(dim)] = 1
    if dim in (3, 5):
        k[0, 0] = k[0, dim - 1] = k[dim - 1, 0] = k[dim - 1, dim - 1] = 0
    return k / np.count_nonzero(k)


def line_anchors(n):
    """LineDictionary.createNxNLines (LineDictionary.py:74-96) as a list indexed by angle step; the hard-coded 7x7
    table differs from the generator in one entry (90 degrees: endpoints swapped, LineDictionary.py:47)."""
    assert n % 2 == 1 and n >= 7
    out = []
    for i in range((n - 1) // 2, n):
        out.append([i, 0, n - 1 - i, n - 1])
    for j in range(1, (n + 1) // 2):
        out.append([n - 1, j, 0, n - 1 - j])
    for j in range((n + 1) // 2, n):
        out.append([0, n - 1 - j, n - 1, j])
    for i in range(1, (n - 1) // 2):
        out.append([i, 0, n - 1 - i, n - 1])
    if n == 7:
        out[6] = [0, 3, 6, 3]
    return out


def _bresenham(r0, c0, r1, c1):
    """skimage.draw.line (integer Bresenham, end point forced)."""
    pts = []
    r, c = r0, c0
    dr, dc = abs(r1 - r0), abs(c1 - c0)
    sc = 1 if c1 > c0 else -1
    sr = 1 if r1 > r0 else -1
    steep = dr > dc
    if steep:
        r, c, dr, dc, sr, sc = c, r, dc, dr, sc, sr
    d = 2 * dr - dc
    for _ in range(dc):
        pts.append((c, r) if steep else (r, c))
        while d >= 0:
            r += sr
            d -= 2 * dc
        c += sc
        d += 2 * dr
    pts.append((r1, c1))
    return pts


def sanitize_angle_index(dim, angle):
    """SanitizeAngleValue (LinearMotionBlur.py:51-56): index of the nearest valid angle."""
    center = dim // 2
    valid = np.linspace(0, 180, center * 4, endpoint=False)
    return int(np.abs(valid - math.fmod(angle, 180.0)).argmin())


def LineKernel(dim, angle, linetype):
    """LinearMotionBlur.py:32-49 with a fresh dictionary."""
    center = dim // 2
    a = list(line_anchors(dim)[sanitize_angle_index(dim, angle)])
    if linetype == 'right':
        a[0] = a[1] = center
    if linetype == 'left':
        a[2] = a[3] = center
    k = np.zeros((dim, dim), dtype=np.float32)
    for r, c in _bresenham(*a):
        k[r, c] = 1
    return k / np.count_nonzero(k)


class _Psf(dict):
    def __missing__(self, key):
        with np.load(_DATA) as z:
            for i in range(100):
                self[i] = z[f'psf{i}']
        return dict.__getitem__(self, key)

    def __len__(self):
        return 100


psfDictionary = _Psf()      # PsfBlur.py:10-11 (same table, converted by tools/convert_psf.py)


def random_angle(dim, rng=np.random):
    """LinearMotionBlur.randomAngle (:63-68)."""
    valid = np.linspace(0, 180, (dim // 2) * 4, endpoint=False)
    return int(valid[rng.randint(0, len(valid))])


def random_blur_kernel(rng=np.random):
    """Draws a kernel the way RandomizedBlur does (RandomizedBlur.py:8-12 and the *_random functions), consuming the
    same np.random calls in the same order.  Returns (kernel, description)."""
    which = rng.randint(0, 4)
    if which == 0:
        dim = KERNEL_DIMS[rng.randint(0, len(KERNEL_DIMS))]
        return BoxKernel(dim), ('box', dim)
    if which == 1:
        dim = KERNEL_DIMS[rng.randint(0, len(KERNEL_DIMS))]
        return DiskKernel(dim), ('disk', dim)
    if which == 2:
        li = rng.randint(0, len(KERNEL_DIMS))
        ti = rng.randint(0, len(LINE_TYPES))
        dim, lt = KERNEL_DIMS[li], LINE_TYPES[ti]
        ang = random_angle(dim, rng)
        return LineKernel(dim, ang, lt), ('line', dim, ang, lt)
    pid = rng.randint(0, 100)
    return np.asarray(psfDictionary[pid], dtype=np.float32), ('psf', pid)


# ------------------------------------------------------------------------------------------ device launch
def _pack_kernels(kernels):
    """List of KxK arrays (or None = no blur) -> (taps [B,kmax,kmax] fp32 centred, ksize [B] int32)."""
    sizes = [0 if k is None else k.shape[0] for k in kernels]
    kmax = max(max(sizes), 1)
    taps = np.zeros((len(kernels), kmax, kmax), dtype=np.float32)
    for b, k in enumerate(kernels):
        if k is None:
            continue
        assert k.shape[0] == k.shape[1] and k.shape[0] % 2 == 1, 'blur kernels must be odd and square'
        o = (kmax - k.shape[0]) // 2
        taps[b, o:o + k.shape[0], o:o + k.shape[0]] = k
    return taps, np.asarray(sizes, dtype=np.int32), kmax


def pack_degradation(kernels, lr_sizes, dev):
    """Device-side parameter block of a batch (blur taps, kernel sizes, low-resolution sizes): build once, reuse."""
    taps, ksize, kmax = _pack_kernels(kernels)
    lw = np.asarray([s[0] for s in lr_sizes], dtype=np.int32)
    lh = np.asarray([s[1] for s in lr_sizes], dtype=np.int32)
    assert lw.min() >= 1 and lh.min() >= 1
    return dict(taps=torch.from_numpy(taps).to(dev), ks=torch.from_numpy(ksize).to(dev), lw=torch.from_numpy(lw).to(dev),
                lh=torch.from_numpy(lh).to(dev), kmax=kmax, lr_wmax=int(lw.max()), lr_hmax=int(lh.max()), n=len(lw))


def degrade_batch(gt_u8, kernels, lr_sizes, noise=None, bgr2rgb=True, return_blur=False, packed=None):
    """Runs the fused degradation on a batch.

    gt_u8    : uint8 CUDA tensor [B,H,W,3] (BGR as cv2 gives it) — what random_pyblur feeds pyblur.
    kernels  : list of B blur kernels (numpy KxK, odd) or None entries (no blur).
    lr_sizes : list of B (lr_w, lr_h) = (int(w // scale), int(h // scale)) (ffhq_degradation_dataset.py:255-256).
    noise    : None or fp32 tensor [B,lr_hmax,lr_wmax,3], already multiplied by sigma/255 (degradations.py:567).
    Returns the LQ batch fp32 [B,3,H,W] in [-1,1] (and, if return_blur, the blurred uint8 and fp32 images)."""
    if not (gt_u8.is_cuda and gt_u8.dtype == torch.uint8 and gt_u8.dim() == 4 and gt_u8.shape[3] == 3):
        raise ValueError('gt_u8 must be a uint8 CUDA tensor [B,H,W,3]; image_restoration_b200 has no CPU path')
    gt_u8 = gt_u8.contiguous()
    B, H, W, _ = gt_u8.shape
    dev = gt_u8.device
    pk = packed if packed is not None else pack_degradation(kernels, lr_sizes, dev)
    assert pk['n'] == B
    kmax, lr_wmax, lr_hmax = pk['kmax'], pk['lr_wmax'], pk['lr_hmax']
    if noise is not None:
        assert noise.is_cuda and noise.dtype == torch.float32 and tuple(noise.shape) == (B, lr_hmax, lr_wmax, 3)
        noise = noise.contiguous()
    t_taps, t_ks, t_lw, t_lh = pk['taps'], pk['ks'], pk['lw'], pk['lh']
    out = torch.empty(B, 3, H, W, device=dev, dtype=torch.float32)
    blur_u8 = torch.empty(B, H, W, 3, device=dev, dtype=torch.uint8) if return_blur else None
    blur_f32 = torch.empty(B, H, W, 3, device=dev, dtype=torch.float32) if return_blur else None
    p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)  # noqa: E731
    with torch.cuda.device(dev):
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(_lib.lib().b200ir_degrade(p(gt_u8), p(t_taps), p(t_ks), kmax, p(t_lw), p(t_lh), p(noise), lr_wmax,
                                             lr_hmax, p(out), p(blur_u8), p(blur_f32), B, H, W, 1 if bgr2rgb else 0,
                                             st), 'b200ir_degrade')
    if return_blur:
        return out, blur_u8, blur_f32
    return out


def _blur_images(img, kernel):
    """Shared body of the pyblur-style operators: img uint8 [H,W,3] (numpy / PIL) or CUDA tensor [B,H,W,3]."""
    single = not torch.is_tensor(img)
    t = torch.from_numpy(np.ascontiguousarray(np.asarray(img, dtype=np.uint8)))[None].cuda() if single else img
    B, H, W, _ = t.shape
    _, blur_u8, _ = degrade_batch(t, [kernel] * B, [(max(W // 4, 1), max(H // 4, 1))] * B, return_blur=True)
    return blur_u8[0].cpu().numpy() if single else blur_u8


def BoxBlur(img, dim):
    return _blur_images(img, BoxKernel(dim))


def DefocusBlur(img, dim):
    return _blur_images(img, DiskKernel(dim))


def LinearMotionBlur(img, dim, angle, linetype):
    return _blur_images(img, LineKernel(dim, angle, linetype))


def PsfBlur(img, psfid):
    return _blur_images(img, np.asarray(psfDictionary[psfid], dtype=np.float32))


def RandomizedBlur(img, rng=np.random):
    return _blur_images(img, random_blur_kernel(rng)[0])


def random_degradation_params(B, H, W, downsample_range=(4, 12), noise_range=(0, 20), rng=np.random):
    """Per-crop random draws of the reference pipeline for the stages fused here (pyblur kernel, scale, sigma, noise),
    in the reference's order per crop.  Noise is drawn on the host with numpy like the reference (degradations.py:567)."""
    kernels, sizes, noises = [], [], []
    for _ in range(B):
        k, _ = random_blur_kernel(rng)
        kernels.append(k)
        scale = rng.uniform(downsample_range[0], downsample_range[1])
        lw, lh = int(W // scale), int(H // scale)
        sizes.append((lw, lh))
        sigma = rng.uniform(noise_range[0], noise_range[1])
        rng.uniform()                       # the gray-noise coin of random_generate_gaussian_noise (degradations.py:652)
        noises.append(np.float32(rng.randn(lh, lw, 3)) * sigma / 255.)
    lr_wmax, lr_hmax = max(s[0] for s in sizes), max(s[1] for s in sizes)
    nz = np.zeros((B, lr_hmax, lr_wmax, 3), dtype=np.float32)
    for b, n in enumerate(noises):
        nz[b, :n.shape[0], :n.shape[1]] = n
    return kernels, sizes, nz
