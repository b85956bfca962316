"""Row-sliding conv (forced with row_mode=2 at any batch) vs the generic tiles (row_mode=0) vs a torch reference."""
import math
import sys, os
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from image_restoration_b200 import ops

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
DEV = 'cuda'
for (B, H, W, cin, cout) in [(1, 128, 384, 64, 64), (2, 64, 192, 64, 64), (1, 128, 384, 32, 32), (3, 40, 200, 32, 64),
                             (2, 64, 192, 64, 128), (1, 128, 384, 64, 32), (2, 16, 130, 16, 16)]:
    torch.manual_seed(0)
    x = torch.randn(B, cin, H, W, device=DEV)
    w = torch.randn(cout, cin, 3, 3, device=DEV) / math.sqrt(cin * 9)
    bias = torch.randn(cout, device=DEV) * 0.1
    xh = x.permute(0, 2, 3, 1).contiguous().half()
    wh = w.permute(0, 2, 3, 1).reshape(cout, 9 * cin).contiguous().half()
    wr = wh.float().view(cout, 3, 3, cin).permute(0, 3, 1, 2)
    ref = F.leaky_relu(F.conv2d(xh.float().permute(0, 3, 1, 2), wr, bias, padding=1), 0.2) * math.sqrt(2)
    for mode in (0, 2):
        out = torch.zeros(B, H, W, cout, device=DEV, dtype=torch.float16)
        op = ops.conv_same(xh, wh, out, 3, bias=bias, act=True, tile=(128, 1, 1), row_mode=mode, block_n=cout)
        op()
        torch.cuda.synchronize()
        got = out.float().permute(0, 3, 1, 2)
        err = (got - ref).abs().max().item()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            op()
        e1.record(); torch.cuda.synchronize()
        print(f'B{B} {H}x{W} {cin}->{cout} row_mode={mode}: max err {err:.3e} (scale {ref.abs().max().item():.2f}) '
              f'{e0.elapsed_time(e1) / 5 * 1e3:.1f} us', flush=True)
