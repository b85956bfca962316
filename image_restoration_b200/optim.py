"""Optimiser step of the optional training loop (BASELINE config 5): torch.optim.Adam as the
reference configures it (basicsr/models/gfpgan_model.py:217-248: lr 2e-3, betas (0, 0.99), no weight decay) and the EMA
of BaseModel.model_ema (basicsr/models/base_model.py:50-57), as ONE launch over one flat fp32 buffer per network
(b200ir_adam_step) instead of ~600 small ATen kernels per step (205 parameter tensors x foreach chunks).

The parameters of the module are re-pointed at views of the flat buffer (the layout DDP / apex use), so the module,
its state_dict and the forward engine keep working on the same storage.  Gradients arrive either as the .grad of the
parameters or as the flat buffer of grad_sync.GradAllReducer (same parameter order), whose 1 / world average is folded
into the step (grad_scale).
"""
import ctypes as C

import torch

from . import _lib, ops


FLAT_ALIGN = 64     # elements: every parameter starts on a 256-byte boundary of the flat buffer


def flat_layout(params, align=FLAT_ALIGN):
    """Offsets of the parameters inside a flat buffer and its total length.  Each view starts on a 256-byte boundary: the
    kernels read bias / table rows as 16-byte vectors and TMA wants 16-byte aligned bases, and parameters re-pointed at the
    flat buffer are handed to them as they are.  The padding elements stay zero (zero gradient: Adam leaves them at zero)."""
    offs, off = [], 0
    for p in params:
        offs.append(off)
        off += -(-p.numel() // align) * align
    return offs, off


class FlatAdam:
    def __init__(self, params, lr=2e-3, betas=(0.0, 0.99), eps=1e-8, weight_decay=0.0, ema_params=None):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError('no trainable parameters')
        dev = self.params[0].device
        _lib.require_cuda(self.params[0], 'optim.FlatAdam')
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.offsets, self.numel = flat_layout(self.params)
        self.flat = torch.zeros(self.numel, device=dev, dtype=torch.float32)
        self.grad = torch.zeros(self.numel, device=dev, dtype=torch.float32)
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.ema = None
        ema_params = list(ema_params) if ema_params is not None else None
        self.ema_params = ema_params
        if ema_params is not None:
            assert [tuple(e.shape) for e in ema_params] == [tuple(p.shape) for p in self.params]
            self.ema = torch.zeros_like(self.flat)
        self.grad_views = []
        with torch.no_grad():
            for i, (p, off) in enumerate(zip(self.params, self.offsets)):
                n = p.numel()
                self.flat[off:off + n].copy_(p.detach().reshape(-1).float())
                p.data = self.flat[off:off + n].view_as(p)
                self.grad_views.append(self.grad[off:off + n].view_as(p))
                if self.ema is not None:
                    self.ema[off:off + n].copy_(ema_params[i].detach().reshape(-1).float())
                    ema_params[i].data = self.ema[off:off + n].view_as(p)
        self.step_count = 0

    @torch.no_grad()
    def step(self, flat_grad=None, grad_scale=1.0, ema_decay=None):
        """One Adam step.  flat_grad: a flat fp32 gradient buffer in parameter order (e.g. GradAllReducer.flat, not yet
        averaged: pass grad_scale = 1 / world); None = gather the .grad of the parameters.  ema_decay: also update the
        EMA copy (needs ema_params at construction)."""
        if flat_grad is None:
            for p, v in zip(self.params, self.grad_views):
                if p.grad is None:
                    v.zero_()
                else:
                    v.copy_(p.grad)
            flat_grad = self.grad
        assert flat_grad.numel() == self.numel and flat_grad.dtype == torch.float32 and flat_grad.device == self.flat.device
        if ema_decay is not None and self.ema is None:
            raise ValueError('ema_decay given but no ema_params were registered')
        self.step_count += 1
        ema_ptr = C.c_void_p(self.ema.data_ptr()) if ema_decay is not None else C.c_void_p(0)
        with _lib.device_ctx(self.flat.device):
            st = ops._stream()
            _lib.check(_lib.lib().b200ir_adam_step(
                C.c_void_p(self.flat.data_ptr()), C.c_void_p(flat_grad.data_ptr()), C.c_void_p(self.exp_avg.data_ptr()),
                C.c_void_p(self.exp_avg_sq.data_ptr()), self.numel, float(self.lr), float(self.betas[0]),
                float(self.betas[1]), float(self.eps), float(self.weight_decay), self.step_count, float(grad_scale),
                ema_ptr, float(ema_decay if ema_decay is not None else 0.0), st), 'b200ir_adam_step')
        # the kernel wrote through raw pointers: tell autograd and the forward engines (OcrEngine.stale() compares the
        # parameters' version counters) that the values changed, so saved-tensor checks and packed weights stay honest
        torch.autograd.graph.increment_version(self.params)
        if ema_decay is not None:
            torch.autograd.graph.increment_version(self.ema_params)

    def zero_grad(self):
        for p in self.params:
            p.grad = None
