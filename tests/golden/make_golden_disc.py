"""Generates tests/golden/disc_*.npz: outputs of the REFERENCE StyleGAN2Discriminator (basicsr/archs/stylegan2_arch.py,
imported from /root/reference) on seeded weights and inputs.  Only the seed and the scores are stored: the B200 class
reproduces the reference's seeded initialisation.  Run: python tests/golden/make_golden_disc.py"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402

CASES = [dict(name='disc_384x128_seed0', seed=0, B=4, kw=dict(input_width=384, input_height=128, channel_multiplier=1)),
         dict(name='disc_64x64_seed1', seed=1, B=8, kw=dict(input_width=64, input_height=64, channel_multiplier=2)),
         dict(name='disc_96x32_seed2', seed=2, B=2, kw=dict(input_width=96, input_height=32, channel_multiplier=1, narrow=0.5))]


def main():
    ref_import.load_reference_arch()
    m = importlib.import_module('basicsr.archs.stylegan2_arch')
    for c in CASES:
        torch.manual_seed(c['seed'])
        net = m.StyleGAN2Discriminator(**c['kw']).eval()
        x = torch.rand(c['B'], 3, c['kw']['input_height'], c['kw']['input_width']) * 2 - 1
        with torch.no_grad():
            y = net(x)
        np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', c['name'] + '.npz'), seed=c['seed'], B=c['B'],
                            kw=np.array(repr(c['kw'])), score=y.numpy(),
                            n_params=sum(p.numel() for p in net.parameters()))
        print(c['name'], y.flatten()[:4].tolist())


if __name__ == '__main__':
    main()
