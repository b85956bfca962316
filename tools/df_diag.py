import sys, numpy as np, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import test_degrade_full_gpu as T
from oracle import degrade_full_oracle as dfo
g, kernels, sizes, cj = T.golden_batch(sys.argv[1] if len(sys.argv) > 1 else 'degrade_full.npz')
out, lr = T.run_gpu(g['gt'], [int(m) for m in g['modes']], kernels, sizes, g['noise'], [int(q) for q in g['quality']], g['jitter'], [int(x) for x in g['gray']], bsigma=[float(x) for x in g['bsigma']], cj=cj)
for i in range(len(kernels)):
    lw, lh = sizes[i]
    ref, ref_lr = dfo.degrade_full(g['gt'][i], int(g['modes'][i]), kernels[i], sizes[i], g['noise'][i, :lh, :lw], int(g['quality'][i]), g['jitter'][i], int(g['gray'][i]), exact_blur=True, bilateral_sigma=float(g['bsigma'][i]), cj=cj[i])
    dg = np.abs(T.to_u8(out[i]) - g['out_u8'][i].astype(np.int32)); do = np.abs(T.to_u8(out[i]) - T.to_u8(ref))
    dl = np.abs(lr[i,:lh,:lw]-ref_lr)*255
    print(i, g['kinds'][i], [op for op, _ in cj[i]], 'vs golden: frac %.4f max %d | vs oracle: frac %.5f max %d | lr: frac %.5f max %.3f' % ((dg>0).mean(), dg.max(), (do>0).mean(), do.max(), (dl>0).mean(), dl.max()))
