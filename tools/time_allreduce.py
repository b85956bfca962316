"""Times the training-step gradient exchange (BASELINE config 5; reference: DDP in base_model.py:70-73) under torchrun:
flat-buffer bucketed NCCL all-reduce of net_g (73.5 M fp32 gradients @128x384) and net_d (32.3 M), device-timed, max
over ranks.   torchrun --nproc-per-node N tools/time_allreduce.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from image_restoration_b200.grad_sync import GradAllReducer  # noqa: E402

local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
world, rank = dist.get_world_size(), dist.get_rank()
for name, n in (('net_g 73.5M', 73_498_700), ('net_d 32.3M', 32_300_000)):
    p = torch.nn.Parameter(torch.randn(n, device=dev))
    p.grad = torch.full((n,), float(rank + 1), device=dev)
    for bucket_mb in (32, 128, 512):
        red = GradAllReducer([p], bucket_mb=bucket_mb)
        for _ in range(3):
            red.sync(average=False)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        K = 10
        for _ in range(K):
            red.sync(average=False)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / K], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            gb = n * 4 / 1e9
            print(f'{name}: world {world}, bucket {bucket_mb:4d} MB: {t.item():7.3f} ms per sync (pack + all-reduce + unpack), '
                  f'algorithm bandwidth {gb / (t.item() / 1e3):7.1f} GB/s, bus {gb * 2 * (world - 1) / world / (t.item() / 1e3):7.1f} GB/s')
        p.grad.fill_(float(rank + 1))
dist.destroy_process_group()
