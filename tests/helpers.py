import glob
import hashlib
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
KW = dict(num_style_feat=256, channel_multiplier=0.5, num_mlp=4, input_is_latent=True, different_w=True, narrow=1,
          sft_half=True)


def state_checksum(sd):
    h = hashlib.sha256()
    for k in sorted(sd.keys()):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def golden_files():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, 'gfpgan_ocr_*.npz')))


def load_golden(path):
    """Returns (fixture dict, seeded B200 module) or (fixture, None) when the seeded weights cannot be reproduced."""
    from image_restoration_b200 import GFPGANv1OCR
    fx = dict(np.load(path, allow_pickle=False))
    W, H, seed = int(fx['W']), int(fx['H']), int(fx['seed'])
    torch.manual_seed(seed)
    net = GFPGANv1OCR(input_width=W, input_height=H, decoder_load_path=None, fix_decoder=True, **KW).eval()
    if state_checksum(net.state_dict()) != str(fx['checksum']):
        return fx, None
    return fx, net
