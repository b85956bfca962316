"""Converts the reference's PSF table (pyblur/pyblur/psf.pkl: a dict {0..99: float32 KxK array}, DATA not code) into
image_restoration_b200/data/psf_kernels.npz so PsfBlur works without the reference tree.  Run in the build container."""
import os
import pickle
import sys

import numpy as np

SRC = '/root/reference/Car_Plate-Restoration/pyblur/pyblur/psf.pkl'
DST = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'image_restoration_b200', 'data',
                   'psf_kernels.npz')
with open(SRC, 'rb') as f:
    d = pickle.load(f, encoding='latin1')
assert sorted(d.keys()) == list(range(100))
np.savez_compressed(DST, **{f'psf{i}': np.asarray(d[i], dtype=np.float32) for i in range(100)})
print(DST, os.path.getsize(DST), 'bytes', file=sys.stderr)
