"""ORACLE SUPPORT (test infrastructure, NOT product code).

Imports the *unmodified* reference modules from /root/reference in the build
container so the restatements in this directory can be pinned against them and
golden fixtures can be generated.  /root/reference does not exist on the GPU
box: callers must check ``available()`` and skip.

Shims, none of which alter reference arithmetic:
  * ``basicsr`` is registered as a bare package so ``basicsr/__init__.py`` (whose
    star-imports need torchvision.transforms.functional_tensor) is skipped;
  * ``fused_act_ext.fused_bias_act`` has no CPU implementation in the reference
    (fused_act.py:19-27); a torch restatement of fused_bias_act_kernel.cu:27-48
    (act=3, grad=0) is injected;
  * ``skimage.draw.{line,disk}`` (scikit-image 0.19.3, not installed, not vendored)
    are restated from their published algorithms (integer Bresenham; ellipse
    with strict ``< 1`` test inside the ceil/floor bounding box).
"""
import copy
import importlib
import os
import sys
import types

import numpy as np

REF_ROOT = '/root/reference/Car_Plate-Restoration'


def available():
    return os.path.isdir(os.path.join(REF_ROOT, 'basicsr', 'archs'))


_arch_cache = {}


def load_reference_arch():
    """Returns (GFPGANv1OCR class, ARCH_REGISTRY) from the reference tree."""
    if 'cls' in _arch_cache:
        return _arch_cache['cls'], _arch_cache['reg']
    import torch.nn.functional as F
    sys.dont_write_bytecode = True
    if 'basicsr' not in sys.modules:
        pkg = types.ModuleType('basicsr')
        pkg.__path__ = [os.path.join(REF_ROOT, 'basicsr')]
        sys.modules['basicsr'] = pkg
    fa = importlib.import_module('basicsr.ops.fused_act.fused_act')

    class _Ext:
        @staticmethod
        def fused_bias_act(x, b, ref, act, grad, alpha, scale):
            assert act == 3 and grad == 0
            if b.numel():
                x = x + b.view(1, -1, *([1] * (x.dim() - 2)))
            return F.leaky_relu(x, alpha) * scale
    fa.fused_act_ext = _Ext
    mod = importlib.import_module('basicsr.archs.gfpganv1_ocr_arch')
    reg = importlib.import_module('basicsr.utils.registry').ARCH_REGISTRY
    _arch_cache['cls'], _arch_cache['reg'] = mod.GFPGANv1OCR, reg
    return mod.GFPGANv1OCR, reg


# ---- scikit-image 0.19.3 primitives restated (skimage/draw/_draw.pyx `_line`, draw.py `disk`/`ellipse`)
def sk_line(r0, c0, r1, c1):
    steep = 0
    r, c = r0, c0
    dr, dc = abs(r1 - r0), abs(c1 - c0)
    sc = 1 if (c1 - c) > 0 else -1
    sr = 1 if (r1 - r) > 0 else -1
    if dr > dc:
        steep = 1
        c, r = r, c
        dc, dr = dr, dc
        sc, sr = sr, sc
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


def sk_disk(center, radius, shape=None):
    r, c = center
    ul = np.ceil([r - radius, c - radius]).astype(int)
    lr = np.floor([r + radius, c + radius]).astype(int)
    rl, cl = np.ogrid[0:float(lr[0] - ul[0] + 1), 0:float(lr[1] - ul[1] + 1)]
    a, b = np.nonzero(((rl - (r - ul[0])) / radius) ** 2 + ((cl - (c - ul[1])) / radius) ** 2 < 1)
    return a + ul[0], b + ul[1]


_pyblur_cache = {}


def load_reference_pyblur():
    """Returns the reference `pyblur` package (fresh LineDictionary state on every call:
    LinearMotionBlur.LineKernel mutates the shared dictionary, LinearMotionBlur.py:37-43)."""
    if 'mod' not in _pyblur_cache:
        sys.dont_write_bytecode = True
        if 'skimage' not in sys.modules:
            sk = types.ModuleType('skimage')
            skd = types.ModuleType('skimage.draw')
            skd.line, skd.disk = sk_line, sk_disk
            sk.draw = skd
            sys.modules['skimage'] = sk
            sys.modules['skimage.draw'] = skd
        inner = os.path.join(REF_ROOT, 'pyblur')
        if inner not in sys.path:
            sys.path.insert(0, inner)
        mod = importlib.import_module('pyblur')
        lmb = importlib.import_module('pyblur.LinearMotionBlur')
        _pyblur_cache['mod'] = mod
        _pyblur_cache['lmb'] = lmb
        _pyblur_cache['lines0'] = copy.deepcopy(lmb.lineDict.lines)
    _pyblur_cache['lmb'].lineDict.lines = copy.deepcopy(_pyblur_cache['lines0'])
    return _pyblur_cache['mod']


def load_reference_degradations():
    """Returns (basicsr.data.degradations, FFHQDegradationDataset) of the reference.  One more stand-in module:
    torchvision.transforms.functional_tensor (removed from torchvision 0.26; degradations.py:11 only takes
    rgb_to_grayscale from it, which still exists in torchvision.transforms.functional)."""
    load_reference_arch()
    load_reference_pyblur()
    if 'torchvision.transforms.functional_tensor' not in sys.modules:
        import torchvision.transforms.functional as TF
        m = types.ModuleType('torchvision.transforms.functional_tensor')
        m.rgb_to_grayscale = TF.rgb_to_grayscale
        sys.modules['torchvision.transforms.functional_tensor'] = m
    deg = importlib.import_module('basicsr.data.degradations')
    ds = importlib.import_module('basicsr.data.ffhq_degradation_dataset')
    return deg, ds.FFHQDegradationDataset
