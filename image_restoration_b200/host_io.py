"""Host-buffer front end of the B200 GFPGANv1OCR: what inference.py / api.py do around the network call
(`img.unsqueeze(0).to('cuda')` ... `tensor2img(out)` = device->host, api.py:96-105, inference.py:68-71), but
pipelined: the host->device copy of batch i+1 and the device->host copy of batch i-1 run on their own CUDA streams
while batch i is in the kernels, so PCIe time disappears behind compute.

    pipe = HostPipeline(net, depth=2)
    t = pipe.submit(x_host_pinned, y_host_pinned)     # asynchronous
    pipe.wait(t)                                      # y_host_pinned now holds net(x)[0]

torch is used for streams, events and pinned/device memory only; every arithmetic op is a libb200ir kernel.
"""
import torch


class _Slot:
    def __init__(self):
        self.x_dev = None
        self.image = None
        self.rgbs = None
        self.h2d_done = torch.cuda.Event()
        self.compute_done = torch.cuda.Event()
        self.d2h_done = torch.cuda.Event()
        self.busy = False


class HostPipeline:
    def __init__(self, net, depth=2, return_rgb=False, randomize_noise=False, bgr=True):
        dev = next(net.parameters()).device
        if dev.type != 'cuda':
            raise RuntimeError('HostPipeline needs the module on a CUDA B200 (no CPU path)')
        self.net, self.dev, self.depth = net, dev, depth
        self.return_rgb, self.randomize_noise, self.bgr = return_rgb, randomize_noise, bgr
        with torch.cuda.device(dev):
            self.h2d = torch.cuda.Stream()
            self.d2h = torch.cuda.Stream()
            self.compute = torch.cuda.current_stream()
            self.slots = [_Slot() for _ in range(depth)]
        self.n = 0

    def submit(self, x_host, y_host, rgbs_host=None):
        """x_host: (B,3,H,W) float32 host tensor in [-1,1], or (B,H,W,3) uint8 images (then y_host is uint8 (B,H,W,3)
        too); pinned for true asynchrony.  Returns a ticket for wait()."""
        if x_host.is_cuda or y_host.is_cuda:
            raise ValueError('HostPipeline.submit takes HOST tensors; call the module directly for device tensors')
        s = self.slots[self.n % self.depth]
        if s.busy:
            s.d2h_done.synchronize()       # the slot's previous job has fully left the device
        with torch.cuda.device(self.dev):
            compute = torch.cuda.current_stream()
            with torch.cuda.stream(self.h2d):
                if s.busy:
                    self.h2d.wait_event(s.compute_done)     # previous input of this slot consumed
                if s.x_dev is None or s.x_dev.shape != x_host.shape or s.x_dev.dtype != x_host.dtype:
                    s.x_dev = torch.empty(x_host.shape, device=self.dev, dtype=x_host.dtype)
                s.x_dev.copy_(x_host, non_blocking=True)
                s.h2d_done.record(self.h2d)
            compute.wait_event(s.h2d_done)
            if x_host.dtype == torch.uint8:     # uint8 HWC images in, uint8 HWC images out (api.py:96-105 on the device)
                s.image = self.net.restore_uint8(s.x_dev, bgr=self.bgr, randomize_noise=self.randomize_noise)
                s.rgbs = []
            else:
                s.image, s.rgbs = self.net(s.x_dev, return_rgb=self.return_rgb, randomize_noise=self.randomize_noise)
            s.compute_done.record(compute)
            with torch.cuda.stream(self.d2h):
                self.d2h.wait_event(s.compute_done)
                s.image.record_stream(self.d2h)
                y_host.copy_(s.image, non_blocking=True)
                if rgbs_host is not None:
                    for dst, src in zip(rgbs_host, s.rgbs):
                        src.record_stream(self.d2h)
                        dst.copy_(src, non_blocking=True)
                s.d2h_done.record(self.d2h)
        s.busy = True
        self.n += 1
        return self.n - 1

    def wait(self, ticket):
        """Blocks the host until the job's outputs are in the host buffers."""
        if ticket < self.n - self.depth:
            return                          # slot already recycled: its job completed before the reuse
        self.slots[ticket % self.depth].d2h_done.synchronize()

    def join(self, stream=None):
        """Makes `stream` (default: the current stream) wait for every outstanding device->host copy, without blocking
        the host: used to close a device-timed region."""
        stream = stream or torch.cuda.current_stream()
        for s in self.slots:
            if s.busy:
                stream.wait_event(s.d2h_done)

    def drain(self):
        for s in self.slots:
            if s.busy:
                s.d2h_done.synchronize()


def restore_host(net, x_host, y_host=None, chunks=1):
    """Synchronous helper: runs net on a host batch and returns the image on the host (pinned staging inside)."""
    pipe = getattr(net, '_host_pipe', None)
    if pipe is None:
        pipe = HostPipeline(net)
        net._host_pipe = pipe
    if y_host is None:       # uint8 HWC images come back as uint8 HWC images (api.py:105), float batches as float32
        y_host = torch.empty(x_host.shape, dtype=torch.uint8 if x_host.dtype == torch.uint8 else torch.float32).pin_memory()
    elif (x_host.dtype == torch.uint8) != (y_host.dtype == torch.uint8):
        raise ValueError('restore_host: uint8 input needs a uint8 output buffer (and float input a float one)')
    b = x_host.shape[0]
    step = -(-b // chunks)
    tickets = [pipe.submit(x_host[i:i + step], y_host[i:i + step]) for i in range(0, b, step)]
    for t in tickets:
        pipe.wait(t)
    return y_host
