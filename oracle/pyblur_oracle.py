"""ORACLE (test infrastructure, NOT product code).

CPU restatement of the degradation part of the reference's hot path:
  P = Car_Plate-Restoration/pyblur/pyblur/           (BoxBlur.py, DefocusBlur.py, LinearMotionBlur.py, LineDictionary.py,
                                                      PsfBlur.py, RandomizedBlur.py)
  D = Car_Plate-Restoration/basicsr/data/degradations.py
  F = Car_Plate-Restoration/basicsr/data/ffhq_degradation_dataset.py
The per-pixel arithmetic goes through the same library calls the reference makes (scipy.signal.convolve2d,
cv2.resize, numpy casts), so the oracle is "those calls as executed on this machine" (SURVEY.md §8c).

Third-party algorithm not under /root/reference: skimage.draw.{disk,line} (scikit-image==0.19.3,
CPR/requirements.txt:37) — restated here from the published algorithms.  The reference has no test pinning their
output, so kernel taps are pinned by (a) the reference pyblur run with the same restated primitives in the build
container (tests/test_degradation_cpu.py) and (b) the known answers of SURVEY.md §8c.
Parity note: pyblur truncates the fp32 convolution to uint8; where the fp32 value lies within rounding of an integer
the truncated byte depends on summation order (scipy's own order is build dependent), so parity is asserted on the
fp32 value (1e-5 of 255) and on the byte up to such ties.
"""
import math
import os

import numpy as np

_PSF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'image_restoration_b200', 'data',
                    'psf_kernels.npz')


def box_kernel(dim):
    """P/BoxBlur.py:21-26."""
    k = np.ones((dim, dim), np.float32)
    return k / np.count_nonzero(k)


def sk_disk(center, radius):
    """skimage.draw.disk -> ellipse: bounding box ceil/floor, strict '< 1' test."""
    r0, c0 = center
    ul = np.ceil([r0 - radius, c0 - radius]).astype(int)
    lr = np.floor([r0 + radius, c0 + radius]).astype(int)
    rr, cc = np.ogrid[ul[0]:lr[0] + 1, ul[1]:lr[1] + 1]
    a, b = np.nonzero(((rr - r0) / radius) ** 2 + ((cc - c0) / radius) ** 2 < 1)
    return a + ul[0], b + ul[1]


def disk_kernel(dim):
    """P/DefocusBlur.py:25-39."""
    k = np.zeros((dim, dim), np.float32)
    rr, cc = sk_disk((dim / 2, dim / 2), dim / 2)
    k[rr, cc] = 1
    if dim in (3, 5):
        k[0, 0] = k[0, -1] = k[-1, 0] = k[-1, -1] = 0
    return k / np.count_nonzero(k)


def sk_line(r0, c0, r1, c1):
    """skimage.draw.line: integer Bresenham with the end point forced."""
    steep = 0
    r, c = r0, c0
    dr, dc = abs(r1 - r0), abs(c1 - c0)
    sc = 1 if (c1 - c) > 0 else -1
    sr = 1 if (r1 - r) > 0 else -1
    if dr > dc:
        steep = 1
        c, r, dc, dr, sc, sr = r, c, dr, dc, sr, sc
    d = 2 * dr - dc
    rr = np.zeros(dc + 1, np.intp)
    cc = np.zeros(dc + 1, np.intp)
    for i in range(dc):
        if steep:
            rr[i], cc[i] = c, r
        else:
            rr[i], cc[i] = r, c
        while d >= 0:
            r += sr
            d -= 2 * dc
        c += sc
        d += 2 * dr
    rr[dc], cc[dc] = r1, c1
    return rr, cc


def line_table(n):
    """P/LineDictionary.py: hard-coded 7x7 (:36-51) and 9x9 (:53-72) tables, createNxNLines (:74-96) otherwise.
    Returned as {angle: [r0, c0, r1, c1]}."""
    if n == 7:
        a = [[3, 0, 3, 6], [4, 0, 2, 6], [5, 0, 1, 6], [6, 0, 0, 6], [6, 1, 0, 5], [6, 2, 0, 4], [0, 3, 6, 3],
             [0, 2, 6, 4], [0, 1, 6, 5], [0, 0, 6, 6], [1, 0, 5, 6], [2, 0, 4, 6]]
        return {15.0 * i: v for i, v in enumerate(a)}
    lines = {}
    unit = 180.0 / (2 * n - 2)
    cnt = 0
    for i in range(int((n - 1) / 2), n):
        lines[cnt * unit] = [i, 0, n - 1 - i, n - 1]
        cnt += 1
    for j in range(1, int((n + 1) / 2)):
        lines[cnt * unit] = [n - 1, j, 0, n - 1 - j]
        cnt += 1
    for j in range(int((n + 1) / 2), n):
        lines[cnt * unit] = [0, n - 1 - j, n - 1, j]
        cnt += 1
    for i in range(1, int((n - 1) / 2)):
        lines[cnt * unit] = [i, 0, n - 1 - i, n - 1]
        cnt += 1
    return lines


def line_kernel(dim, angle, linetype):
    """P/LinearMotionBlur.py:32-61 evaluated on a fresh dictionary (the reference mutates its shared one)."""
    center = int(math.floor(dim / 2))
    valid = np.linspace(0, 180, center * 4, endpoint=False)
    ang = valid[(np.abs(valid - math.fmod(angle, 180.0))).argmin()]
    table = line_table(dim)
    key = min(table.keys(), key=lambda k: abs(k - ang))
    a = list(table[key])
    if linetype == 'right':
        a[0] = a[1] = center
    if linetype == 'left':
        a[2] = a[3] = center
    k = np.zeros((dim, dim), np.float32)
    rr, cc = sk_line(*a)
    k[rr, cc] = 1
    return k / np.count_nonzero(k)


def psf_kernel(psfid):
    """P/PsfBlur.py:10-11,16 — the table itself is data (converted from psf.pkl by tools/convert_psf.py)."""
    with np.load(_PSF) as z:
        return np.asarray(z[f'psf{psfid}'], np.float32)


def blur_f32(img_u8, kernel):
    """The convolution of P/*Blur.py before the uint8 cast: per channel convolve2d(mode='same', fillvalue=255)."""
    from scipy.signal import convolve2d
    a = np.array(img_u8, dtype='float32')
    return np.stack([convolve2d(a[:, :, i], kernel, mode='same', fillvalue=255.0) for i in range(3)], axis=2)


def blur_u8(img_u8, kernel):
    return blur_f32(img_u8, kernel).astype('uint8')


def degrade(gt_u8, kernel, lr_size, noise, bgr2rgb=True):
    """F:244-272,307-311 for one crop with explicit random draws.
    gt_u8 uint8 [H,W,3] (= np.array(img*255, dtype=uint8) of D:364); kernel KxK or None; lr_size (lw, lh);
    noise fp32 [lh,lw,3] already scaled by sigma/255 or None.  Returns fp32 [3,H,W] in [-1,1]."""
    import cv2
    h, w = gt_u8.shape[:2]
    img = gt_u8 if kernel is None else blur_u8(gt_u8, kernel)
    img = np.array(img, dtype=np.float32) / 255.0                                       # D:366
    img = cv2.resize(img, lr_size, interpolation=cv2.INTER_LINEAR)                     # F:256
    if noise is not None:
        img = np.clip(img + noise, 0, 1)                                               # D:660-669
    img = cv2.resize(img, (w, h), interpolation=cv2.INTER_LINEAR)                      # F:272
    if bgr2rgb:
        img = img[:, :, ::-1]                                                          # img2tensor(bgr2rgb=True)
    t = np.ascontiguousarray(img.transpose(2, 0, 1)).astype(np.float32)
    t = np.clip(t, 0, 1)                                                               # tensor2img clamp, F:301
    t = np.clip(np.round(t * np.float32(255.0)), 0, 255) / np.float32(255.)            # F:308 (round half to even)
    return ((t - np.float32(0.5)) / np.float32(0.5)).astype(np.float32)                # F:310-311
