"""Any-resolution driver: the tile / fold logic of demo_any_resolution.py (:35-52, :116-139) on the device.

The reference pads the image onto a square zero canvas, unfolds it into 256x256 tiles with stride 128, runs the model
on one tile at a time, re-concatenates the outputs (O(n^2) copies), folds them back with F.fold and divides by the
folded cover count.  Here the tiles of this rank's contiguous range are cut straight from the image by a gather kernel
(no canvas, no unfold), run through the model in batches, and overlap-added into a canvas by a gather-form fold kernel
(no atomics); with several ranks the canvases are summed by ONE reduce to rank 0, which normalises, crops and clamps.
The channel count of the result comes from the model output (the fork's 1-channel head makes the reference script
fail for more than one tile).
"""
import ctypes
import math

import torch
import torch.distributed as dist

from . import _lib
from .shard import tile_range


def canvas_geometry(h, w, kernel=256, stride=128):
    X = int(math.ceil(max(h, w) / float(kernel)) * kernel)
    n = (X - kernel) // stride + 1
    return X, n


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


@torch.no_grad()
def extract_tiles(img, first, count, kernel=256, stride=128):
    img = _lib.require_cuda(img, "img")
    _, C, h, w = img.shape
    tiles = torch.empty(count, C, kernel, kernel, device=img.device, dtype=torch.float32)
    with torch.cuda.device(img.device):
        _lib.check(_lib.load().sunet_tiles_extract(_p(img), C, h, w, kernel, stride, _p(tiles), first, count, _lib.stream_ptr(img.device)))
    return tiles


@torch.no_grad()
def fold_tiles(tiles, first, h, w, acc=None, kernel=256, stride=128):
    tiles = _lib.require_cuda(tiles, "tiles")
    count, C = tiles.shape[0], tiles.shape[1]
    X, _ = canvas_geometry(h, w, kernel, stride)
    if acc is None:
        acc = torch.zeros(C, X, X, device=tiles.device, dtype=torch.float32)
    with torch.cuda.device(tiles.device):
        _lib.check(_lib.load().sunet_tiles_fold(_p(tiles), C, h, w, kernel, stride, first, count, _p(acc), _lib.stream_ptr(tiles.device)))
    return acc


@torch.no_grad()
def finish(acc, h, w, kernel=256, stride=128):
    C = acc.shape[0]
    out = torch.empty(1, C, h, w, device=acc.device, dtype=torch.float32)
    with torch.cuda.device(acc.device):
        _lib.check(_lib.load().sunet_tiles_finish(_p(acc), C, h, w, kernel, stride, _p(out), _lib.stream_ptr(acc.device)))
    return out


@torch.no_grad()
def denoise_any_resolution(model, img, kernel=256, stride=128, tile_batch=64, rank=0, world_size=1, events=None):
    """img (1, C, h, w) fp32 CUDA in [0,1] -> restored (1, C_out, h, w), clamped to [0,1] (valid on rank 0).
    Tiles are sharded contiguously over ranks; each rank needs the full input image.
    ``events`` (optional dict) receives CUDA events recorded on the current stream around the one collective of the path:
    'reduce_begin' / 'reduce_end' (+ 'reduce_bytes'), so a caller can report the time spent in the canvas reduce."""
    img = _lib.require_cuda(img, "img")
    _, _, h, w = img.shape
    X, n = canvas_geometry(h, w, kernel, stride)
    lo, hi = tile_range(n * n, rank, world_size)
    acc = None
    for first in range(lo, hi, tile_batch):
        count = min(tile_batch, hi - first)
        out = model(extract_tiles(img, first, count, kernel, stride))
        acc = fold_tiles(out, first, h, w, acc, kernel, stride)
    if acc is None:  # more ranks than tiles
        c_out = getattr(getattr(model, "swin_unet", model), "out_chans", 1)
        acc = torch.zeros(c_out, X, X, device=img.device, dtype=torch.float32)
    if world_size > 1:
        if events is not None:
            events["reduce_begin"] = torch.cuda.Event(enable_timing=True)
            events["reduce_end"] = torch.cuda.Event(enable_timing=True)
            events["reduce_bytes"] = acc.numel() * 4
            events["reduce_begin"].record()
        dist.reduce(acc, dst=0, op=dist.ReduceOp.SUM)
        if events is not None:
            events["reduce_end"].record()
        if rank != 0:
            return None
    return finish(acc, h, w, kernel, stride)
