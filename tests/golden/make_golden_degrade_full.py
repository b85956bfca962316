"""Generates tests/golden/degrade_full.npz by running the REFERENCE's own degradation functions
(basicsr/data/degradations.py, ffhq_degradation_dataset.py, imported from /root/reference) on seeded synthetic crops.
Run in the build container: python tests/golden/make_golden_degrade_full.py"""
import math
import os
import random
import sys

import cv2
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402
from image_restoration_b200 import degradation as D  # noqa: E402

# kernel_list of training_config/train_gfpgan_v4_square_license_mix_pyblur.yml:26-28 (probabilities evened out so that 20
# crops reach every kind)
OPT = dict(blur_kernel_size=21, kernel_list=['iso', 'aniso', 'motion', 'average', 'median', 'bilateral', 'pyblur'],
           kernel_prob=[0.12, 0.12, 0.12, 0.12, 0.16, 0.16, 0.2], blur_sigma=[0.1, 10], downsample_range=[4.0, 12.0],
           noise_range=[0, 20], jpeg_range=[30, 100], color_jitter_prob=0.3, color_jitter_shift=20, gray_prob=0.25,
           color_jitter_pt_prob=0.4)


def smooth_crop(rng, h, w):
    """A plate-like synthetic crop: smooth background + sharp strokes, uint8 BGR."""
    base = cv2.resize(rng.random((h // 16 + 2, w // 16 + 2, 3)).astype(np.float32), (w, h), interpolation=cv2.INTER_CUBIC)
    img = np.clip(base, 0, 1)
    for _ in range(12):
        x0, y0 = int(rng.integers(0, w - 8)), int(rng.integers(0, h - 8))
        img[y0:y0 + int(rng.integers(4, 40)), x0:x0 + int(rng.integers(2, 12))] = rng.random(3)
    return (img * 255).astype(np.uint8)


def reference_lq(deg, DS, gt, opt):
    """__getitem__ lines 242-311 for one crop, calling the reference's functions (random state: the global
    `random` / `np.random` / torch generator, seeded by the caller).  gt: uint8 image (imfrombytes float32 = u8 / 255) or the
    float32 image in [0, 1] the dataset holds after its cv2.resize (:230)."""
    img_gt = gt.astype(np.float32) / 255. if gt.dtype == np.uint8 else gt
    h, w, _ = img_gt.shape
    img_lq = deg.random_mixed_kernels(img=img_gt, kernel_list=opt['kernel_list'], kernel_prob=opt['kernel_prob'],
                                      kernel_size=opt['blur_kernel_size'], sigma_x_range=opt['blur_sigma'],
                                      sigma_y_range=opt['blur_sigma'], rotation_range=[-math.pi, math.pi],
                                      noise_range=None, pad_kernel=True, pad_kernel_size=opt['blur_kernel_size'])
    scale = np.random.uniform(opt['downsample_range'][0], opt['downsample_range'][1])
    img_lq = cv2.resize(img_lq, (int(w // scale), int(h // scale)), interpolation=cv2.INTER_LINEAR)
    if opt['noise_range'] is not None:
        img_lq = deg.random_add_gaussian_noise(img_lq, opt['noise_range'])
    if opt['jpeg_range'] is not None:
        img_lq = deg.random_add_jpg_compression(img_lq, opt['jpeg_range'])
    img_lq = cv2.resize(img_lq, (w, h), interpolation=cv2.INTER_LINEAR)
    if opt['color_jitter_prob'] is not None and (np.random.uniform() < opt['color_jitter_prob']):
        img_lq = DS.color_jitter(img_lq, opt['color_jitter_shift'] / 255.)
    if opt['gray_prob'] and np.random.uniform() < opt['gray_prob']:
        img_lq = cv2.cvtColor(img_lq, cv2.COLOR_BGR2GRAY)
        img_lq = np.tile(img_lq[:, :, None], [1, 1, 3])
    t = torch.from_numpy(np.ascontiguousarray(img_lq[..., ::-1].transpose(2, 0, 1))).float()   # img2tensor(bgr2rgb)
    if opt.get('color_jitter_pt_prob') is not None and (np.random.uniform() < opt['color_jitter_pt_prob']):
        t = DS.color_jitter_pt(t, opt.get('brightness', (0.5, 1.5)), opt.get('contrast', (0.5, 1.5)),
                               opt.get('saturation', (0, 1.5)), opt.get('hue', (-0.1, 0.1)))
    t = t.clamp(0, 1)                                             # tensor2img(out_type=float32) / img2tensor round trip (:299-303)
    if opt.get('random_mask'):                                    # :299-303: tensor2img -> random_mask -> img2tensor
        img = np.ascontiguousarray(t.numpy().transpose(1, 2, 0)[..., ::-1])
        img = DS.random_mask(img)
        t = torch.from_numpy(np.ascontiguousarray(img[..., ::-1].transpose(2, 0, 1))).float()
    t = torch.clamp((t * 255.0).round(), 0, 255) / 255.
    return ((t - 0.5) / 0.5).numpy()


def main():
    deg, DS = ref_import.load_reference_degradations()
    if len(sys.argv) > 1 and sys.argv[1] == 'mask':     # random_mask: true (regular / irregular / half masks), 14 small crops
        make(deg, DS, 'degrade_full_mask.npz', 64, 192, 14, 5000, np.random.default_rng(11), float_gt=False,
             opt=dict(OPT, random_mask=True))
        return
    if len(sys.argv) > 1 and sys.argv[1] == 'bicubic':  # the 'bicubic' kind (Pillow x1/4 and back), float GT off the 8-bit grid
        make(deg, DS, 'degrade_full_bicubic.npz', 64, 192, 8, 7000, np.random.default_rng(13), float_gt=True,
             opt=dict(OPT, kernel_list=['bicubic', 'iso'], kernel_prob=[0.75, 0.25]))
        return
    make(deg, DS, 'degrade_full.npz', 128, 384, 20, 1000, np.random.default_rng(7), float_gt=False)
    # GT images that are not on the 8-bit grid: the dataset resizes every image to the network size (:230)
    make(deg, DS, 'degrade_full_floatgt.npz', 64, 192, 12, 3000, np.random.default_rng(9), float_gt=True)


def make(deg, DS, fname, H, W, N, seed0, rng, float_gt, opt=None):
    opt = opt or OPT
    gts, outs, seeds = [], [], []
    for i in range(N):
        gt = smooth_crop(rng, H, W)
        if float_gt:
            big = smooth_crop(rng, H * 2 + 7, W * 2 + 5).astype(np.float32) / 255.
            gt = cv2.resize(big, (W, H), interpolation=cv2.INTER_LINEAR)
        seed = seed0 + i
        ref_import.load_reference_pyblur()          # fresh LineDictionary (the reference mutates it)
        random.seed(seed)
        np.random.seed(seed)
        torch.manual_seed(seed)
        outs.append(reference_lq(deg, DS, gt, opt))
        gts.append(gt)
        seeds.append(seed)
    # parameters the host mirror draws from the same seeds (stored so the GPU box needs no reference)
    recs = []
    for gt, seed in zip(gts, seeds):
        pr, nr = random.Random(seed), np.random.RandomState(seed)
        recs.append(D.sample_params(1, H, W, opt, py_random=pr, np_random=nr,
                                    torch_generator=torch.Generator().manual_seed(seed)))
    kmax = 29
    taps = np.zeros((N, kmax, kmax), np.float64)
    for i, r in enumerate(recs):
        k = r['kernels'][0]
        o = (kmax - k.shape[0]) // 2
        taps[i, o:o + k.shape[0], o:o + k.shape[0]] = k
    lw = np.array([r['sizes'][0][0] for r in recs]); lh = np.array([r['sizes'][0][1] for r in recs])
    noise = np.zeros((N, lh.max(), lw.max(), 3), np.float32)
    for i, r in enumerate(recs):
        noise[i, :lh[i], :lw[i]] = r['noise'][0]
    extra = {}
    if opt.get('random_mask'):
        extra = dict(mask_modes=np.array([r['mask_modes'][0] for r in recs]), masks=np.stack([r['masks'][0] for r in recs]))
        print('mask modes', extra['mask_modes'].tolist(), 'masked fraction', [round(float(m.mean()), 3) for m in extra['masks']])
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', fname), gt=np.stack(gts), **extra,
                        out_u8=np.rint((np.stack(outs) * 0.5 + 0.5) * 255).astype(np.uint8), seeds=np.array(seeds),
                        modes=np.array([r['modes'][0] for r in recs]), taps=taps,
                        ksize=np.array([r['kernels'][0].shape[0] for r in recs]),
                        f64=np.array([int(r['kernels'][0].dtype == np.float64) for r in recs]), lr_w=lw, lr_h=lh, noise=noise,
                        quality=np.array([r['quality'][0] for r in recs]),
                        jitter=np.stack([r['jitter'][0] for r in recs]), gray=np.array([r['gray'][0] for r in recs]),
                        bsigma=np.array([r['bilateral_sigma'][0] for r in recs], dtype=np.float32),
                        cj_n=np.array([len(r['color_jitter_pt'][0]) for r in recs]),
                        cj_op=np.array([[s[0] for s in r['color_jitter_pt'][0]] + [0] * (4 - len(r['color_jitter_pt'][0]))
                                        for r in recs]),
                        cj_f=np.array([[s[1] for s in r['color_jitter_pt'][0]] + [0.] * (4 - len(r['color_jitter_pt'][0]))
                                       for r in recs], dtype=np.float64),
                        kinds=np.array([r['desc'][0][0] for r in recs]))
    print('kinds', [r['desc'][0][0] for r in recs])
    print('gray', [r['gray'][0] for r in recs], 'quality', [r['quality'][0] for r in recs])
    print('cj', [r['color_jitter_pt'][0] for r in recs])


if __name__ == '__main__':
    main()
