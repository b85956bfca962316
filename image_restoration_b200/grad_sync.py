"""Gradient all-reduce for the optional data-parallel training step (BASELINE config 5).

The reference wraps net_g / net_d in DistributedDataParallel (basicsr/models/base_model.py:62-76,
`find_unused_parameters: true`), i.e. one bucketed NCCL all-reduce(sum) + divide per network per step.  This is the
same exchange without the wrapper: gradients are packed into ONE persistent flat fp32 buffer per network (73.5 M
elements for GFPGANv1OCR @128x384, 32.3 M for the discriminator), all-reduced in a few large buckets on a side stream
(NVSwitch gives full bandwidth to every peer, so buckets are sized for launch latency and overlap, not link count), and
unpacked in place.  Parameters that received no gradient contribute zeros, which is what find_unused_parameters does.

torch.distributed (backend nccl on GPUs, gloo in the CPU tests) is the transport; there is no collective on the
inference path.
"""
import torch
import torch.distributed as dist


class GradAllReducer:
    def __init__(self, params, bucket_mb=64, group=None):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError('no trainable parameters')
        dev = self.params[0].device
        self.group = group
        from .optim import flat_layout
        offsets, self.numel = flat_layout(self.params)        # the layout of optim.FlatAdam: reduce_flat() feeds its step
        self.flat = torch.zeros(self.numel, device=dev, dtype=torch.float32)
        self.views = [self.flat[off:off + p.numel()].view_as(p) for p, off in zip(self.params, offsets)]
        per = max(1, int(bucket_mb * (1 << 20) // 4))
        self.buckets = [(s, min(s + per, self.numel)) for s in range(0, self.numel, per)]
        self.stream = torch.cuda.Stream(device=dev) if dev.type == 'cuda' else None

    @torch.no_grad()
    def reduce_flat(self):
        """Packs the current .grad of every parameter into the flat buffer, all-reduces it (SUM, bucketed, on the side stream)
        and returns the flat buffer WITHOUT averaging or unpacking: optim.FlatAdam.step(flat_grad=..., grad_scale=1 / world)
        consumes it directly, which saves the unpack / re-pack passes of sync()."""
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()
            else:
                v.copy_(p.grad)
        if self.stream is not None:
            self.stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.stream):
                works = [dist.all_reduce(self.flat[s:e], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
                         for s, e in self.buckets]
                for w in works:
                    w.wait()
            torch.cuda.current_stream().wait_stream(self.stream)
        else:
            for s, e in self.buckets:
                dist.all_reduce(self.flat[s:e], op=dist.ReduceOp.SUM, group=self.group)
        return self.flat

    @torch.no_grad()
    def sync(self, average=True):
        """All-reduces the current .grad of every parameter across the group (sum, then / world when `average`) and
        writes the result back into .grad (allocating zeros for parameters that had none, like DDP with
        find_unused_parameters)."""
        world = dist.get_world_size(self.group)
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()
            else:
                v.copy_(p.grad)
        if self.stream is not None:
            self.stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.stream):
                works = [dist.all_reduce(self.flat[s:e], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
                         for s, e in self.buckets]
                for w in works:
                    w.wait()
                if average:
                    self.flat.div_(world)
            torch.cuda.current_stream().wait_stream(self.stream)
        else:
            for s, e in self.buckets:
                dist.all_reduce(self.flat[s:e], op=dist.ReduceOp.SUM, group=self.group)
            if average:
                self.flat.div_(world)
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                p.grad = v.clone()
            else:
                p.grad.copy_(v)
        return self.numel * 4
