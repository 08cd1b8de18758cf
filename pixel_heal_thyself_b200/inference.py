"""Full-frame tiled inference with halo overlap (north_star: "inference on full
frames is tiled per GPU with halo overlap").

The reference has no inference script; the oracle is its ``AFGSANet.eval()`` run
on the whole frame (the network is fully convolutional for H, W multiples of 8).
A tile computed with a halo of >= 48 px on its interior sides and an origin
aligned to the 8-px attention block grid reproduces the full-frame result
exactly (the dependency front grows 2 px in the encoders, to
8*ceil((c+3)/8)+2 in each of the 5 blocks, +3 in the decoder = 45 -> 48); sides
on the true image border need no halo.  Tiles are independent, so N GPUs each
take a strided subset and only the 3-channel cores are gathered.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.distributed as dist

EXACT_HALO = 48
ALIGN = 8


@dataclass(frozen=True)
class Tile:
    y0: int   # core (output) region in frame coordinates
    y1: int
    x0: int
    x1: int
    ty0: int  # haloed (input) region
    ty1: int
    tx0: int
    tx1: int


def _splits(n: int, parts: int) -> list[tuple[int, int]]:
    """Split [0, n) into `parts` contiguous ranges whose boundaries are multiples of ALIGN."""
    units = n // ALIGN
    parts = max(1, min(parts, units))
    edges = [ALIGN * ((units * i) // parts) for i in range(parts)] + [n]
    return [(edges[i], edges[i + 1]) for i in range(parts)]


def _uniform(lo: int, hi: int, size: int, n: int) -> tuple[int, int]:
    """Grow the haloed range [lo, hi) to `size` inside [0, n): extra halo is harmless (still exact), the origin stays a
    multiple of ALIGN because lo, size and n are."""
    lo = max(0, min(lo, n - size))
    return lo, lo + size


def plan_tiles(H: int, W: int, rows: int, cols: int, halo: int = EXACT_HALO, uniform: bool = True) -> list[Tile]:
    """``uniform``: every tile's haloed input region gets the SAME shape (the largest one; border tiles take a wider
    halo on their inner side), so the generator runs on one activation arena instead of one per tile shape."""
    if H % ALIGN or W % ALIGN:
        raise AssertionError("feature map dimensions must be divisible by the block size")
    if halo % ALIGN:
        raise ValueError("halo must be a multiple of the 8-px attention block")
    tiles = []
    for (y0, y1) in _splits(H, rows):
        for (x0, x1) in _splits(W, cols):
            tiles.append(Tile(y0, y1, x0, x1, max(0, y0 - halo), min(H, y1 + halo), max(0, x0 - halo), min(W, x1 + halo)))
    if uniform and tiles:
        th = max(t.ty1 - t.ty0 for t in tiles)
        tw = max(t.tx1 - t.tx0 for t in tiles)
        tiles = [Tile(t.y0, t.y1, t.x0, t.x1, *_uniform(t.ty0, t.ty1, th, H), *_uniform(t.tx0, t.tx1, tw, W)) for t in tiles]
    return tiles


def denoise_frame(net, x: torch.Tensor, aux: torch.Tensor, rows: int = 2, cols: int = 4, halo: int = EXACT_HALO,
                  rank: int = 0, world: int = 1, gather: bool = True) -> torch.Tensor | None:
    """x [1,3,H,W], aux [1,7,H,W] preprocessed NCHW fp32 on this rank's device.  Every rank calls this with the
    same frame; rank r computes tiles r, r+world, ...  Returns the stitched [1,3,H,W] frame on rank 0 (all
    ranks when world == 1), None elsewhere."""
    _, _, H, W = x.shape
    tiles = plan_tiles(H, W, rows, cols, halo)
    out = torch.zeros_like(x)
    with torch.no_grad():
        for i in range(rank, len(tiles), world):
            t = tiles[i]
            o = net(x[:, :, t.ty0:t.ty1, t.tx0:t.tx1].contiguous(), aux[:, :, t.ty0:t.ty1, t.tx0:t.tx1].contiguous())
            out[:, :, t.y0:t.y1, t.x0:t.x1] = o[:, :, t.y0 - t.ty0:t.y1 - t.ty0, t.x0 - t.tx0:t.x1 - t.tx0]
    if world > 1 and gather:
        # cores are disjoint and everything else is zero: a SUM reduce onto rank 0 is the gather
        dist.reduce(out, dst=0, op=dist.ReduceOp.SUM)
        return out if rank == 0 else None
    return out
