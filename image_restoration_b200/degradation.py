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
def _pack_kernels(kernels, dtype=np.float32):
    """List of KxK arrays (or None = no blur) -> (taps [B,kmax,kmax] centred, ksize [B] int32)."""
    sizes = [0 if k is None else k.shape[0] for k in kernels]
    kmax = max(max(sizes), 1)
    taps = np.zeros((len(kernels), kmax, kmax), dtype=dtype)
    for b, k in enumerate(kernels):
        if k is None:
            continue
        assert k.shape[0] == k.shape[1] and k.shape[0] % 2 == 1, 'blur kernels must be odd and square'
        o = (kmax - k.shape[0]) // 2
        taps[b, o:o + k.shape[0], o:o + k.shape[0]] = k
    return taps, np.asarray(sizes, dtype=np.int32), kmax


def pack_degradation(kernels, lr_sizes, dev):
    """Device-side parameter block of a batch (blur taps, kernel sizes, low-resolution sizes): build once, reuse."""
    taps, ksize, kmax = _pack_kernels(kernels, np.float64)
    # the dtype of each kernel selects the arithmetic type of the reference's convolve2d (float64 box / disk / line
    # kernels under NumPy 2, float32 psf kernels)
    f64 = np.asarray([1 if (k is not None and np.asarray(k).dtype == np.float64) else 0 for k in kernels], dtype=np.int32)
    lw = np.asarray([s[0] for s in lr_sizes], dtype=np.int32)
    lh = np.asarray([s[1] for s in lr_sizes], dtype=np.int32)
    assert lw.min() >= 1 and lh.min() >= 1
    return dict(taps=torch.from_numpy(taps).to(dev), ks=torch.from_numpy(ksize).to(dev), f64=torch.from_numpy(f64).to(dev),
                lw=torch.from_numpy(lw).to(dev),
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
        _lib.check(_lib.lib().b200ir_degrade(p(gt_u8), p(t_taps), p(t_ks), p(pk['f64']), kmax, p(t_lw), p(t_lh), p(noise), lr_wmax,
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


# ====================================================================================== full LQ synthesis (b200ir_degrade_full)
# Host mirror of the kernel builders and of the order of random draws of random_mixed_kernels (degradations.py:419-523)
# and FFHQDegradationDataset.__getitem__ (ffhq_degradation_dataset.py:242-285).  `random` and `np.random` are the two
# generators the reference draws from; pass seeded stand-ins (random.Random / np.random.RandomState) to reproduce a
# reference run draw for draw.
FILTER2D_KINDS = ('iso', 'aniso', 'generalized_iso', 'generalized_aniso', 'plateau_iso', 'plateau_aniso', 'motion',
                  'average')
UNSUPPORTED_KINDS = ('pyblur_motion', 'random_cover')     # undefined in the reference itself (RandomMotion / RandomCover)


def mesh_axis(kernel_size):
    """Coordinates of mesh_grid (degradations.py:35-50): -k//2 + 1 ... k//2."""
    return np.arange(-kernel_size // 2 + 1., kernel_size // 2 + 1.)


def bivariate_Gaussian(kernel_size, sig_x, sig_y, theta, isotropic=True):
    """degradations.py:87-112 (with sigma_matrix2 :19-32 and pdf2 :53-66): exp(-x^T S^-1 x / 2) on the centred grid,
    S = R diag(sx^2, sy^2) R^T, normalised to sum 1.  Returns float64 [k, k] indexed [y, x]."""
    if isotropic:
        cov = np.array([[sig_x ** 2, 0.], [0., sig_x ** 2]])
    else:
        rot = np.array([[np.cos(theta), -np.sin(theta)], [np.sin(theta), np.cos(theta)]])
        cov = rot @ np.diag([sig_x ** 2, sig_y ** 2]) @ rot.T
    inv = np.linalg.inv(cov)
    ax = mesh_axis(kernel_size)
    xx, yy = np.meshgrid(ax, ax)
    pts = np.stack([xx, yy], axis=-1)                                  # [y, x, (x, y)]
    k = np.exp(-0.5 * np.sum((pts @ inv) * pts, axis=2))
    return k / np.sum(k)


def _quad_form(kernel_size, sig_x, sig_y, theta, isotropic):
    """x^T S^-1 x on the centred grid (shared by the three kernel families, degradations.py:87-176)."""
    if isotropic:
        cov = np.array([[sig_x ** 2, 0.], [0., sig_x ** 2]])
    else:
        rot = np.array([[np.cos(theta), -np.sin(theta)], [np.sin(theta), np.cos(theta)]])
        cov = rot @ np.diag([sig_x ** 2, sig_y ** 2]) @ rot.T
    inv = np.linalg.inv(cov)
    ax = mesh_axis(kernel_size)
    xx, yy = np.meshgrid(ax, ax)
    pts = np.stack([xx, yy], axis=-1)
    return np.sum((pts @ inv) * pts, axis=2)


def bivariate_generalized_Gaussian(kernel_size, sig_x, sig_y, theta, beta, isotropic=True):
    """degradations.py:115-147: exp(-0.5 * (x^T S^-1 x)^beta), normalised."""
    k = np.exp(-0.5 * np.power(_quad_form(kernel_size, sig_x, sig_y, theta, isotropic), beta))
    return k / np.sum(k)


def bivariate_plateau(kernel_size, sig_x, sig_y, theta, beta, isotropic=True):
    """degradations.py:150-176: 1 / ((x^T S^-1 x)^beta + 1), normalised."""
    k = np.reciprocal(np.power(_quad_form(kernel_size, sig_x, sig_y, theta, isotropic), beta) + 1)
    return k / np.sum(k)


def motion_kernel(kernel_size, horizontal):
    """motion_blur (degradations.py:330-341): a centred horizontal or vertical line of 1/k."""
    k = np.zeros((kernel_size, kernel_size))
    c = int((kernel_size - 1) / 2)
    if horizontal:
        k[c, :] = 1
    else:
        k[:, c] = 1
    return k / kernel_size


def average_kernel(kernel_size):
    """average_blur (degradations.py:344-350): float32 ones / k^2."""
    return np.ones((kernel_size, kernel_size), np.float32) / (kernel_size * kernel_size)


def bilateral_space_kernel(d, sigma_space):
    """Space weights of cv2.bilateralFilter (bilateral_filter.dispatch.cpp): radius = d // 2, for every (i, j) with
    sqrt(i^2 + j^2) <= radius the weight (float)exp(r^2 * -0.5 / sigma^2); zero outside the circle.  float32 [2r+1, 2r+1]."""
    radius = max(d // 2, 1)
    gs = -0.5 / (float(sigma_space) * float(sigma_space))
    ax = np.arange(-radius, radius + 1, dtype=np.float64)
    r = np.sqrt(ax[:, None] ** 2 + ax[None, :] ** 2)
    k = np.exp(r * r * gs).astype(np.float32)
    k[r > radius] = 0
    return k


def _pad_to(k, size):
    pad = (size - k.shape[0]) // 2
    return np.pad(k, ((pad, pad), (pad, pad))) if pad > 0 else k


def random_mixed_kernel(kernel_list, kernel_prob, kernel_size=21, sigma_x_range=(0.6, 5), sigma_y_range=(0.6, 5),
                        rotation_range=(-math.pi, math.pi), pad_kernel=False, pad_kernel_size=21, py_random=None,
                        np_random=np.random, betag_range=(0.5, 8), betap_range=(0.5, 8)):
    """The kernel random_mixed_kernels would apply (degradations.py:419-523), drawn with the same calls in the same
    order.  Returns (blur_mode, kernel, description): blur_mode 2 = cv2.filter2D kinds, 1 = 'pyblur', 3 = 'median'
    (kernel: zeros, only its size is used), 4 = 'bilateral' (kernel: the space weights; description carries sigma), 5 = 'bicubic'
    (kernel unused)."""
    import random as _random
    py_random = py_random or _random
    kind = py_random.choices(kernel_list, kernel_prob)[0]
    if kind in ('iso', 'aniso'):
        assert kernel_size % 2 == 1 and sigma_x_range[0] < sigma_x_range[1]
        sx = np_random.uniform(sigma_x_range[0], sigma_x_range[1])
        sy, rot = sx, 0
        if kind == 'aniso':
            sy = np_random.uniform(sigma_y_range[0], sigma_y_range[1])
            rot = np_random.uniform(rotation_range[0], rotation_range[1])
        k = bivariate_Gaussian(kernel_size, sx, sy, rot, isotropic=(kind == 'iso'))
        k = k / np.sum(k)                       # random_bivariate_Gaussian normalises once more (:222)
        desc = (kind, sx, sy, rot)
    elif kind in ('generalized_iso', 'generalized_aniso', 'plateau_iso', 'plateau_aniso'):
        # random_bivariate_generalized_Gaussian / random_bivariate_plateau (degradations.py:226-327): sigma_x, [sigma_y,
        # rotation], the beta coin, beta
        iso = kind.endswith('_iso')
        sx = np_random.uniform(sigma_x_range[0], sigma_x_range[1])
        sy, rot = sx, 0
        if not iso:
            sy = np_random.uniform(sigma_y_range[0], sigma_y_range[1])
            rot = np_random.uniform(rotation_range[0], rotation_range[1])
        br = betag_range if kind.startswith('generalized') else betap_range
        beta = np_random.uniform(br[0], 1) if np_random.uniform() < 0.5 else np_random.uniform(1, br[1])
        build = bivariate_generalized_Gaussian if kind.startswith('generalized') else bivariate_plateau
        k = build(kernel_size, sx, sy, rot, beta, isotropic=iso)
        k = k / np.sum(k)
        desc = (kind, sx, sy, rot, beta)
    elif kind == 'motion':
        horizontal = py_random.random() > 0.5
        k, desc = motion_kernel(kernel_size, horizontal), (kind, horizontal)
    elif kind == 'average':
        k, desc = average_kernel(kernel_size), (kind,)
    elif kind == 'pyblur':
        k, d = random_blur_kernel(np_random)
        return 1, k, ('pyblur',) + tuple(d)
    elif kind == 'median':          # median_blur (degradations.py:353-355): cv2.medianBlur(uint8 image, kernel_size)
        return 3, np.zeros((kernel_size, kernel_size), np.float32), ('median', kernel_size)
    elif kind == 'bilateral':       # bilateral_blur (degradations.py:358-361): sigma = random.randint(150, 250)
        sigma = py_random.randint(150, 250)
        return 4, bilateral_space_kernel(kernel_size, sigma), ('bilateral', kernel_size, sigma)
    elif kind == 'bicubic':         # bicubic (degradations.py:379-385): Pillow x1/4 and back, evaluated on the device
        return 5, np.zeros((kernel_size, kernel_size), np.float32), ('bicubic',)
    elif kind in ('pyblur_motion', 'random_cover'):
        raise NotImplementedError(f"blur kind '{kind}' cannot run in the reference either: degradations.py:369-377 calls "
                                  "RandomMotion / RandomCover, which its pyblur package does not define")
    else:
        raise NotImplementedError(f"blur kind '{kind}' has no B200 implementation (supported: "
                                  f"{FILTER2D_KINDS + ('pyblur', 'median', 'bilateral', 'bicubic')})")
    if pad_kernel:
        k = _pad_to(k, pad_kernel_size)
    return 2, k, desc


def random_mask_draw(H, W, py_random=None, np_random=np.random):
    """The draws of FFHQDegradationDataset.random_mask (ffhq_degradation_dataset.py:95-187) for an H x W image, with the same
    calls in the same order.  Returns (mask_mode, mask uint8 [H, W]): mode 1 = regular rectangles / half masks (masked pixels
    become 1.0), mode 2 = irregular mask (lines, circles, ellipses rasterised by the reference's own cv2 calls; the image is
    also truncated to the 8-bit grid, see include/b200ir.h).  The reference indexes its shapes with size[0] = H for x and
    size[1] = W for y (:121-141); that is kept."""
    import random as _random
    rnd = py_random or _random
    mask = np.zeros((H, W), dtype=np.uint8)
    if rnd.random() > 0.3:
        if rnd.random() > 0.5:                                   # random_regular_mask (:95-110)
            n_mask = rnd.randint(1, 5)
            limx, limy = H - H / (n_mask + 1), W - W / (n_mask + 1)
            for _ in range(n_mask):
                x, y = rnd.randint(0, int(limx)), rnd.randint(0, int(limy))
                range_x = x + rnd.randint(int(H / (n_mask + 7)), int(H - x))
                range_y = y + rnd.randint(int(W / (n_mask + 7)), int(W - y))
                mask[int(x):int(range_x), int(y):int(range_y)] = 1
            return 1, mask
        if H < 64 or W < 64:                                     # random_irregular_mask (:112-151)
            raise Exception('Width and Height of mask must be at least 64!')
        import cv2
        max_width = 20
        for _ in range(rnd.randint(16, 64)):
            model = rnd.random()
            if model < 0.6:
                x1, x2 = rnd.randint(1, H), rnd.randint(1, H)
                y1, y2 = rnd.randint(1, W), rnd.randint(1, W)
                cv2.line(mask, (x1, y1), (x2, y2), 255, rnd.randint(4, max_width))
            elif 0.6 < model < 0.8:
                x1, y1 = rnd.randint(1, H), rnd.randint(1, W)
                cv2.circle(mask, (x1, y1), rnd.randint(4, max_width), 255, -1)
            elif model > 0.8:
                x1, y1 = rnd.randint(1, H), rnd.randint(1, W)
                s1, s2 = rnd.randint(1, H), rnd.randint(1, W)
                a1, a2, a3 = rnd.randint(3, 180), rnd.randint(3, 180), rnd.randint(3, 180)
                cv2.ellipse(mask, (x1, y1), (s1, s2), a1, a2, a3, 255, rnd.randint(4, max_width))
        return 2, (mask == 255).astype(np.uint8)
    half_h, half_w = int(H / 2), int(W / 2)                      # half masks (:163-186)
    if rnd.random() > 0.5:
        start = np_random.uniform(0.0, 7 / 8)
        end = np_random.uniform(start, 1.0)
        if end - start > 0.5:
            end -= 0.5
        if rnd.random() > 0.5:
            mask[int(start * half_h):int(end * half_h), :] = 1
        else:
            mask[:, int(start * half_w):int(end * half_w)] = 1
    else:
        tmp = rnd.random()
        if tmp > 0.75:
            mask[:half_h] = 1
        elif tmp > 0.50:
            mask[half_h:] = 1
        elif tmp > 0.25:
            mask[:, :half_w] = 1
        else:
            mask[:, half_w:] = 1
    return 1, mask


def sample_params(B, H, W, opt, py_random=None, np_random=np.random, torch_generator=None):
    """Per-crop random draws of __getitem__ for the stages b200ir_degrade_full runs, in the reference's order: blur
    kind + kernel, scale, noise sigma (+ the gray-noise coin) + noise field, JPEG quality, colour-jitter coin / shifts,
    gray coin, color_jitter_pt coin + torch.randperm(4) + one torch uniform per adjustment (torch_generator=None draws
    from torch's global generator, as the reference does).  opt carries the dataset options of the training YAML (kernel_list, kernel_prob, blur_kernel_size,
    blur_sigma, downsample_range, noise_range, jpeg_range, color_jitter_prob, color_jitter_shift, gray_prob).
    Returns a dict of host arrays ready for degrade_full_batch."""
    ks = opt['blur_kernel_size']
    modes, kernels, sizes, noises, quality, jitter, gray, desc, bsigma, cj = [], [], [], [], [], [], [], [], [], []
    mask_modes, masks = [], []
    for _ in range(B):
        m, k, d = random_mixed_kernel(opt['kernel_list'], opt['kernel_prob'], ks, opt['blur_sigma'], opt['blur_sigma'],
                                      (-math.pi, math.pi), pad_kernel=True, pad_kernel_size=ks, py_random=py_random,
                                      np_random=np_random)
        modes.append(m)
        bsigma.append(float(d[2]) if m == 4 else 0.0)
        kernels.append(np.asarray(k))       # dtype kept: it selects the arithmetic type of the reference's blur
        desc.append(d)
        scale = np_random.uniform(opt['downsample_range'][0], opt['downsample_range'][1])
        lw, lh = int(W // scale), int(H // scale)
        sizes.append((lw, lh))
        if opt.get('noise_range') is not None:
            sigma = np_random.uniform(opt['noise_range'][0], opt['noise_range'][1])
            np_random.uniform()                 # gray-noise coin of random_generate_gaussian_noise (degradations.py:652)
            noises.append(np.float32(np_random.randn(lh, lw, 3)) * sigma / 255.)
        else:
            noises.append(None)
        if opt.get('jpeg_range') is not None:
            quality.append(int(np_random.uniform(opt['jpeg_range'][0], opt['jpeg_range'][1])))
        else:
            quality.append(0)
        j = np.zeros(3, dtype=np.float32)
        if opt.get('color_jitter_prob') is not None and np_random.uniform() < opt['color_jitter_prob']:
            shift = opt.get('color_jitter_shift', 20) / 255.
            j = np_random.uniform(-shift, shift, 3).astype(np.float32)
        jitter.append(j)
        gray.append(1 if (opt.get('gray_prob') and np_random.uniform() < opt['gray_prob']) else 0)
        steps = []
        if opt.get('color_jitter_pt_prob') is not None and np_random.uniform() < opt['color_jitter_pt_prob']:
            ranges = (opt.get('brightness', (0.5, 1.5)), opt.get('contrast', (0.5, 1.5)), opt.get('saturation', (0, 1.5)),
                      opt.get('hue', (-0.1, 0.1)))
            for fn_id in torch.randperm(4, generator=torch_generator).tolist():
                if ranges[fn_id] is not None:
                    f = torch.tensor(1.0).uniform_(ranges[fn_id][0], ranges[fn_id][1], generator=torch_generator).item()
                    steps.append((fn_id, f))
        cj.append(steps)
        if opt.get('random_mask'):              # :299-303, after color_jitter_pt
            mm, mk = random_mask_draw(H, W, py_random=py_random, np_random=np_random)
            mask_modes.append(mm)
            masks.append(mk)
    lr_wmax, lr_hmax = max(s[0] for s in sizes), max(s[1] for s in sizes)
    nz = None
    if any(n is not None for n in noises):
        nz = np.zeros((B, lr_hmax, lr_wmax, 3), dtype=np.float32)
        for b, n in enumerate(noises):
            if n is not None:
                nz[b, :n.shape[0], :n.shape[1]] = n
    return dict(modes=modes, kernels=kernels, sizes=sizes, noise=nz, quality=quality, jitter=jitter, gray=gray,
                bilateral_sigma=bsigma, color_jitter_pt=cj, desc=desc,
                mask_modes=mask_modes or None, masks=np.stack(masks) if masks else None)


def pack_degrade_full(modes, kernels, sizes, noise=None, quality=None, jitter=None, gray=None, bilateral_sigma=None,
                      color_jitter_pt=None, dev='cuda', mask_modes=None, masks=None, **_unused):
    """Device-side parameter block of a batch for b200ir_degrade_full (taps, per-crop records, noise): build once per
    batch of draws, reuse across launches.  Arguments as returned by sample_params."""
    B = len(modes)
    taps, ksize, kmax = _pack_kernels([k if m != 0 else None for k, m in zip(kernels, modes)], np.float64)
    crops = (_lib.DegradeCrop * B)()
    for b in range(B):
        c = crops[b]
        c.blur_mode, c.ksize = int(modes[b]), int(ksize[b])
        c.blur_f64 = 1 if (c.blur_mode == 1 and np.asarray(kernels[b]).dtype == np.float64) else 0
        c.lr_w, c.lr_h = int(sizes[b][0]), int(sizes[b][1])
        if c.lr_w < 2 or c.lr_h < 2:
            raise ValueError('low-resolution size must be at least 2x2')
        c.jpeg_quality = int(quality[b]) if quality is not None else 0
        c.gray = int(gray[b]) if gray is not None else 0
        c.bilateral_sigma = float(bilateral_sigma[b]) if bilateral_sigma is not None else 0.0
        if c.blur_mode == 4 and not c.bilateral_sigma > 0:
            raise ValueError('bilateral blur needs bilateral_sigma > 0')
        for i in range(3):
            c.jitter[i] = float(jitter[b][i]) if jitter is not None else 0.0
        steps = color_jitter_pt[b] if color_jitter_pt is not None else []
        c.cj_count = len(steps)
        for i, (op, f) in enumerate(steps):     # torchvision's _blend: ratio and (1.0 - ratio) become float32 scalars
            c.cj_order[i], c.cj_factor[i], c.cj_one_minus[i] = int(op), float(f), float(1.0 - float(f))
        c.mask_mode = int(mask_modes[b]) if mask_modes is not None else 0
        if c.mask_mode not in (0, 1, 2):
            raise ValueError('mask_mode must be 0, 1 or 2')
    lr_wmax, lr_hmax = max(s[0] for s in sizes), max(s[1] for s in sizes)
    if noise is not None:
        noise = torch.as_tensor(noise, dtype=torch.float32).to(dev).contiguous()
        assert tuple(noise.shape) == (B, lr_hmax, lr_wmax, 3)
    mask_t = None
    if mask_modes is not None and any(int(m) for m in mask_modes):
        if masks is None:
            raise ValueError('mask_modes given without masks')
        mask_t = torch.as_tensor(np.asarray(masks), dtype=torch.uint8).to(dev).contiguous()
        assert mask_t.dim() == 3 and mask_t.shape[0] == B, 'masks: uint8 [B, H, W]'
    return dict(crops=torch.frombuffer(bytearray(bytes(crops)), dtype=torch.uint8).to(dev),
                taps=torch.from_numpy(taps).to(dev), kmax=kmax, noise=noise, lr_wmax=lr_wmax, lr_hmax=lr_hmax, n=B, mask=mask_t,
                has_bicubic=any(int(m) == 5 for m in modes))


def degrade_full_batch(gt_u8, modes=None, kernels=None, sizes=None, noise=None, quality=None, jitter=None, gray=None,
                       bilateral_sigma=None, color_jitter_pt=None, bgr2rgb=True, return_lr=False, packed=None, mask_modes=None,
                       masks=None, **_unused):
    """One launch of b200ir_degrade_full over a batch (see include/b200ir.h).  gt_u8: uint8 CUDA tensor [B,H,W,3] in
    the reference's channel order (BGR), or a float32 CUDA tensor [B,H,W,3] in [0,1] (the reference's img_gt after
    cv2.resize, not on the 8-bit grid: filter2D kinds then work on the float values as the reference does); the other arguments as returned by sample_params (or packed= the result of
    pack_degrade_full).  Returns the LQ batch fp32 [B,3,H,W] in [-1,1] (and the low-resolution image after noise /
    JPEG if return_lr)."""
    if not (gt_u8.is_cuda and gt_u8.dtype in (torch.uint8, torch.float32) and gt_u8.dim() == 4 and gt_u8.shape[3] == 3):
        raise ValueError('gt_u8 must be a uint8 or float32 CUDA tensor [B,H,W,3]; image_restoration_b200 has no CPU path')
    gt_u8 = gt_u8.contiguous()
    B, H, W, _ = gt_u8.shape
    dev = gt_u8.device
    gt_f32 = None
    if gt_u8.dtype == torch.float32:
        gt_f32, gt_u8 = gt_u8, torch.empty(B, H, W, 3, device=dev, dtype=torch.uint8)   # scratch the launch fills
    pk = packed if packed is not None else pack_degrade_full(modes, kernels, sizes, noise, quality, jitter, gray,
                                                             bilateral_sigma, color_jitter_pt, dev, mask_modes, masks)
    assert pk['n'] == B
    mask_t = pk.get('mask')
    if mask_t is not None and tuple(mask_t.shape) != (B, H, W):
        raise ValueError(f'masks must be uint8 [B, H, W] = {(B, H, W)}, got {tuple(mask_t.shape)}')
    out = torch.empty(B, 3, H, W, device=dev, dtype=torch.float32)
    scratch = torch.empty(B, H, W, 3, device=dev, dtype=torch.uint8) if pk.get('has_bicubic') else None   # 'bicubic' crops
    lr = torch.zeros(B, pk['lr_hmax'], pk['lr_wmax'], 3, device=dev, dtype=torch.float32) if return_lr else None
    p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)  # noqa: E731
    with torch.cuda.device(dev):
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(_lib.lib().b200ir_degrade_full_ex(p(gt_u8), p(gt_f32), p(pk['taps']), pk['kmax'], p(pk['crops']),
                                                     p(pk['noise']), pk['lr_wmax'], pk['lr_hmax'], p(mask_t), p(scratch), p(out),
                                                     p(lr), B, H, W, 1 if bgr2rgb else 0, st), 'b200ir_degrade_full')
    return (out, lr) if return_lr else out


def synthesize_pairs(gt, opt, py_random=None, np_random=np.random, torch_generator=None):
    """The training pair FFHQDegradationDataset.__getitem__ returns (ffhq_degradation_dataset.py:221-331), for a batch of
    GT crops already at the network size and resident on the device: gt uint8 [B,H,W,3] (BGR) or float32 in [0,1].
    Returns {'lq': fp32 [B,3,H,W], 'gt': fp32 [B,3,H,W]} (RGB, normalised with mean = std = 0.5) and the drawn parameters.
    Two launches: b200ir_degrade_full for lq, img2tensor + normalize for gt."""
    B, H, W, _ = gt.shape
    prm = sample_params(B, H, W, opt, py_random=py_random, np_random=np_random, torch_generator=torch_generator)
    lq = degrade_full_batch(gt, **prm)
    gt_t = torch.empty(B, 3, H, W, device=gt.device, dtype=torch.float32)
    fn = _lib.lib().b200ir_u8_to_input if gt.dtype == torch.uint8 else _lib.lib().b200ir_f32_to_input
    with torch.cuda.device(gt.device):
        _lib.check(fn(C.c_void_p(gt.contiguous().data_ptr()), C.c_void_p(gt_t.data_ptr()), B, H, W, 1,
                      C.c_void_p(torch.cuda.current_stream().cuda_stream)), 'gt img2tensor + normalize')
    return {'lq': lq, 'gt': gt_t}, prm
