"""Where does the end-to-end step lose time against the resident step?  Times K steps of: resident, H2D only, D2H only,
both (HostPipeline), for pipeline depths 2 and 3."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import NET_KW, H, W  # noqa: E402
from image_restoration_b200 import GFPGANv1OCR  # noqa: E402
from image_restoration_b200.host_io import HostPipeline  # noqa: E402

B, K = 64, 30
torch.manual_seed(0)
net = GFPGANv1OCR(**NET_KW).eval().cuda()
x_host = (torch.rand(B, 3, H, W) * 2 - 1).pin_memory()
y_host = torch.empty(B, 3, H, W).pin_memory()
x_dev = x_host.cuda()


def timed(fn, fin=lambda: None):
    for _ in range(3):
        fn()
    fin()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        fn()
    fin()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K, (time.perf_counter() - t0) * 1e3 / K


print('resident            ms/step (events, wall): %.3f %.3f' % timed(lambda: net(x_dev, return_rgb=False, randomize_noise=False)))
copy = torch.cuda.Stream()


def h2d_only():
    with torch.cuda.stream(copy):
        xd = x_host.to('cuda', non_blocking=True)
    torch.cuda.current_stream().wait_stream(copy)
    net(xd, return_rgb=False, randomize_noise=False)


print('naive h2d + compute ms/step: %.3f %.3f' % timed(h2d_only))
for depth in (2, 3, 4):
    pipe = HostPipeline(net, depth=depth)
    print(f'pipeline depth {depth}    ms/step: %.3f %.3f' % timed(lambda: pipe.submit(x_host, y_host), pipe.join))
    pipe.drain()
# raw PCIe copies alone
for name, fn in (('H2D 37.7MB', lambda: x_dev.copy_(x_host, non_blocking=True)), ('D2H 37.7MB', lambda: y_host.copy_(x_dev, non_blocking=True))):
    print(name, 'ms: %.3f %.3f' % timed(fn))
