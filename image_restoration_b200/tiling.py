"""Tiled full-frame inference (BASELINE config 4): overlapping tiles, batched through the network, blended back.

The reference has no tiling code (api_plate_oto.restoration_car resizes the whole car image to the network size,
api_plate_oto.py:376-401), so this module defines the scheme: tiles of the network's input size on a regular stride
(tile - overlap), the last row / column of tiles aligned to the frame border, linear-ramp blending normalised by the
weight sum.  The gather and the blend are libb200ir kernels; torch only owns the buffers.
"""
import ctypes as C

import torch

from . import _lib
from .ops import _ptr, _stream


def tile_positions(size, tile, overlap):
    """Start offsets of the tiles along one axis: stride tile - overlap, last one flush with the border."""
    if size < tile:
        raise ValueError(f'frame extent {size} is smaller than the tile {tile}')
    if not 0 <= overlap < tile:
        raise ValueError(f'overlap {overlap} must satisfy 0 <= overlap < tile ({tile})')
    stride = tile - overlap
    pos = list(range(0, size - tile, stride)) + [size - tile]
    return pos


def ramp_weights(size, tile, overlap, pos):
    """1-D blend weight of every tile (oracle restatement of tile_ramp in pointwise.cu): [len(pos), tile] float32."""
    w = torch.ones(len(pos), tile)
    i = torch.arange(tile, dtype=torch.float32)
    for k, p0 in enumerate(pos):
        if p0 != 0:
            w[k] = torch.minimum(w[k], (i + 1) / (overlap + 1))
        if p0 + tile != size:
            w[k] = torch.minimum(w[k], (tile - i) / (overlap + 1))
    return w


class TiledRestorer:
    """frame (3,H,W) float32 CUDA tensor in the network's input range -> restored frame (3,H,W)."""

    def __init__(self, net, overlap=32, micro_batch=64):
        if net.input_width != net.input_height:
            raise ValueError('tiling uses square tiles: build the network with input_width == input_height')
        if not 0 <= overlap < net.input_width:
            raise ValueError(f'overlap {overlap} must satisfy 0 <= overlap < tile ({net.input_width})')
        self.net, self.tile, self.overlap, self.micro_batch = net, net.input_width, overlap, micro_batch
        self._grid = {}

    def grid(self, H, W, dev):
        key = (H, W, str(dev))
        if key not in self._grid:
            ty, tx = tile_positions(H, self.tile, self.overlap), tile_positions(W, self.tile, self.overlap)
            self._grid[key] = (ty, tx, torch.tensor(ty, dtype=torch.int32, device=dev),
                               torch.tensor(tx, dtype=torch.int32, device=dev))
        return self._grid[key]

    @torch.no_grad()
    def __call__(self, frame, randomize_noise=True):
        if not (frame.is_cuda and frame.dtype == torch.float32 and frame.dim() == 3 and frame.shape[0] == 3):
            raise ValueError('expected a float32 CUDA frame (3,H,W)')
        frame = frame.contiguous()
        _, H, W = frame.shape
        T = self.tile
        ty, tx, ty_d, tx_d = self.grid(H, W, frame.device)
        n = len(ty) * len(tx)
        lib = _lib.lib()
        tiles = torch.empty(n, 3, T, T, device=frame.device, dtype=torch.float32)
        _lib.check(lib.b200ir_tiles_gather(_ptr(frame), _ptr(tiles), 3, H, W, T, _ptr(ty_d), _ptr(tx_d), len(ty), len(tx),
                                           _stream()), 'tiles_gather')
        outs = [self.net(tiles[s:s + self.micro_batch], return_rgb=False, randomize_noise=randomize_noise)[0]
                for s in range(0, n, self.micro_batch)]
        restored = outs[0] if len(outs) == 1 else torch.cat(outs, 0)
        out = torch.empty_like(frame)
        _lib.check(lib.b200ir_tiles_blend(_ptr(restored.contiguous()), _ptr(out), 3, H, W, T, self.overlap, _ptr(ty_d),
                                          _ptr(tx_d), len(ty), len(tx), _stream()), 'tiles_blend')
        return out


def blend_reference(tiles, H, W, tile, overlap):
    """CPU / torch restatement of the blend (test oracle): tiles [n,3,T,T] in row-major tile order -> frame (3,H,W)."""
    ty, tx = tile_positions(H, tile, overlap), tile_positions(W, tile, overlap)
    wy, wx = ramp_weights(H, tile, overlap, ty), ramp_weights(W, tile, overlap, tx)
    num = torch.zeros(3, H, W, dtype=torch.float64)
    den = torch.zeros(H, W, dtype=torch.float64)
    for a, y0 in enumerate(ty):
        for b, x0 in enumerate(tx):
            w = (wy[a][:, None] * wx[b][None, :]).double()
            num[:, y0:y0 + tile, x0:x0 + tile] += w * tiles[a * len(tx) + b].double()
            den[y0:y0 + tile, x0:x0 + tile] += w
    return (num / den).float()
