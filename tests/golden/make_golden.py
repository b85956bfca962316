"""Generates tests/golden/*.npz by running the UNMODIFIED reference GFPGANv1OCR (imported from /root/reference
through oracle/ref_import.py) on seeded random-init weights and synthetic crops.  Run in the build container:

    python tests/golden/make_golden.py

Each fixture stores the input crops, the reference output image (+ out_rgbs) and a checksum of the seeded
state_dict, so a test can rebuild the same weights from the seed and verify it did.
"""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
KW = dict(num_style_feat=256, channel_multiplier=0.5, num_mlp=4, input_is_latent=True, different_w=True, narrow=1,
          sft_half=True)
CASES = [  # name, W, H, batch, seed
    ('gfpgan_ocr_384x128_seed0', 384, 128, 1, 0),
    ('gfpgan_ocr_48x16_seed1', 48, 16, 2, 1),
    ('gfpgan_ocr_64x64_seed2', 64, 64, 2, 2),
]


def state_checksum(sd):
    h = hashlib.sha256()
    for k in sorted(sd.keys()):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def main():
    Ref, _ = ref_import.load_reference_arch()
    for name, W, H, B, seed in CASES:
        torch.manual_seed(seed)
        net = Ref(input_width=W, input_height=H, decoder_load_path=None, fix_decoder=True, **KW).eval()
        x = torch.rand(B, 3, H, W) * 2 - 1
        with torch.no_grad():
            y, rgbs = net(x, return_rgb=True, randomize_noise=False)
        out = dict(x=x.numpy(), image=y.numpy(), W=W, H=H, seed=seed, checksum=state_checksum(net.state_dict()))
        for i, r in enumerate(rgbs):
            out[f'rgb{i}'] = r.numpy()
        np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
        print(name, tuple(y.shape), out['checksum'][:16])


if __name__ == '__main__':
    main()
