"""Generates tests/golden/sr_*.npz by running the UNMODIFIED reference MSRResNet / EDSR / RCAN (imported from
/root/reference through oracle/ref_import.py) on seeded random-init weights and synthetic low-resolution images in
[0, 1].  Run in the build container:   python tests/golden/make_golden_sr.py
Each fixture stores the input, the reference output, the constructor kwargs and a checksum of the seeded state_dict."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402
from tests.helpers import state_checksum  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = [  # name, arch, kwargs, (B, H, W), seed
    ('sr_msrresnet_x4_seed0', 'MSRResNet', dict(num_in_ch=3, num_out_ch=3, num_feat=64, num_block=4, upscale=4), (2, 24, 40), 0),
    ('sr_msrresnet_x3_seed1', 'MSRResNet', dict(num_in_ch=3, num_out_ch=3, num_feat=32, num_block=2, upscale=3), (1, 16, 24), 1),
    ('sr_edsr_x2_seed2', 'EDSR', dict(num_in_ch=3, num_out_ch=3, num_feat=64, num_block=4, upscale=2, res_scale=0.1), (2, 24, 32), 2),
    ('sr_rrdbnet_x4_seed4', 'RRDBNet', dict(num_in_ch=3, num_out_ch=3, scale=4, num_feat=64, num_block=2, num_grow_ch=32), (2, 16, 24), 4),
    ('sr_rrdbnet_x2_seed5', 'RRDBNet', dict(num_in_ch=3, num_out_ch=3, scale=2, num_feat=32, num_block=1, num_grow_ch=16), (1, 32, 48), 5),
    ('sr_rcan_x4_seed3', 'RCAN', dict(num_in_ch=3, num_out_ch=3, num_feat=64, num_group=2, num_block=2, squeeze_factor=16, upscale=4), (2, 16, 24), 3),
]


def main():
    ref_import.load_reference_arch()                      # registers every reference arch
    import importlib
    mods = {'MSRResNet': 'srresnet_arch', 'EDSR': 'edsr_arch', 'RCAN': 'rcan_arch', 'RRDBNet': 'rrdbnet_arch'}
    for name, arch, kw, (B, H, W), seed in CASES:
        cls = getattr(importlib.import_module('basicsr.archs.' + mods[arch]), arch)
        torch.manual_seed(seed)
        net = cls(**kw).eval()
        x = torch.rand(B, 3, H, W)
        with torch.no_grad():
            y = net(x)
        np.savez_compressed(os.path.join(HERE, name + '.npz'), x=x.numpy(), y=y.numpy(), arch=arch, kwargs=json.dumps(kw),
                            seed=seed, checksum=state_checksum(net.state_dict()))
        print(name, tuple(y.shape), float(y.abs().max()))


if __name__ == '__main__':
    main()
