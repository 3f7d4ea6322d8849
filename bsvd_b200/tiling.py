"""Spatial-tile mode for frames too large for one GPU's time budget (BASELINE config 5: 4K clips).

The frame is cut into rows x cols tiles, one per GPU (one process per GPU).  Each rank gathers
the input pixels of a `HALO`-pixel (80) ring around its tile from its neighbours (one NCCL all_gather of
the 4-channel input tiles over NVLink — the input is the only tensor that ever crosses GPUs), runs
the unmodified BSVD-64 path on the enlarged tile and keeps the centre.

Why this is exact: the receptive field of one DenBlock reaches 40 px (full-res convs 2+2, stride-2
convs 1+2, half-res 4+6, quarter-res 20, and the two PixelShuffles, whose sub-pixel of a coarse cell
adds up to 2+1 px), 80 px for temp1+temp2.  With HALO = 80 (a multiple of 4, so the /2 and /4
grids of the tile coincide with those of the full frame) every
output pixel of the centre sees exactly the operands it sees in the full-frame run, in the same
per-pixel summation order, and true image borders still get the conv's zero padding because the
enlarged tile is clamped to the frame.  The result is bit-identical to the single-GPU forward
(tests/test_gpu_network.py::test_spatial_tiles_bit_exact).  The cost is redundant compute on the
ring ((th+160)(tw+160)/(th*tw) for an interior tile; 1.26x on average for 4x2 tiles of 2160x3840) instead of ~32 latency-bound
per-layer halo exchanges per frame.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

HALO = 80   # receptive-field radius of the two DenBlocks (2 x 40 px); a multiple of 4


@dataclass(frozen=True)
class Tile:
    y0: int
    y1: int
    x0: int
    x1: int           # centre region owned by the tile (half-open)
    hy0: int
    hy1: int
    hx0: int
    hx1: int          # enlarged region actually processed (clamped to the frame)


def tile_plan(H: int, W: int, rows: int, cols: int, halo: int = HALO):
    """rows x cols tiles whose edges are multiples of 4; row-major order (tile r*cols+c)."""
    if H % 4 or W % 4:
        raise ValueError("H and W must be multiples of 4")
    if halo % 4:
        raise ValueError("halo must be a multiple of 4")

    def cuts(n, k):
        q = n // 4
        edges = [4 * ((q * i) // k) for i in range(k)] + [n]
        if len(set(edges)) != len(edges):
            raise ValueError(f"cannot cut {n} pixels into {k} tiles of at least 4")
        return edges

    ys, xs = cuts(H, rows), cuts(W, cols)
    tiles = []
    for r in range(rows):
        for c in range(cols):
            y0, y1, x0, x1 = ys[r], ys[r + 1], xs[c], xs[c + 1]
            tiles.append(Tile(y0, y1, x0, x1, max(0, y0 - halo), min(H, y1 + halo),
                              max(0, x0 - halo), min(W, x1 + halo)))
    return tiles


def run_tile(forward, x_full: torch.Tensor, t: Tile) -> torch.Tensor:
    """forward: callable [T,C,h,w] -> [T,3,h,w].  Returns the tile's centre [T,3,y1-y0,x1-x0]."""
    y = forward(x_full[:, :, t.hy0:t.hy1, t.hx0:t.hx1].contiguous())
    return y[:, :, t.y0 - t.hy0:t.y1 - t.hy0, t.x0 - t.hx0:t.x1 - t.hx0]


def forward_tiled_local(forward, x_full: torch.Tensor, rows: int, cols: int, halo: int = HALO):
    """All tiles on the calling device, one after the other (checks / single-GPU use)."""
    T, _, H, W = x_full.shape
    out = None
    for t in tile_plan(H, W, rows, cols, halo):
        y = run_tile(forward, x_full, t)
        if out is None:
            out = torch.empty((T, y.shape[1], H, W), dtype=y.dtype, device=y.device)
        out[:, :, t.y0:t.y1, t.x0:t.x1] = y
    return out


def forward_tiled_distributed(forward, x_tile: torch.Tensor, H: int, W: int, rows: int, cols: int,
                              halo: int = HALO, group=None, gather_output: bool = True):
    """One tile per rank.  x_tile: this rank's input tile [T,C,th,tw] (rank = r*cols+c).
    1. all_gather the input tiles (only the input crosses GPUs), 2. crop the enlarged region,
    3. run, 4. (optionally) all_gather the output centres into the full [T,3,H,W] frame."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    tiles = tile_plan(H, W, rows, cols, halo)
    if world != len(tiles):
        raise ValueError(f"{len(tiles)} tiles need {len(tiles)} ranks, got {world}")
    T, C = x_tile.shape[:2]
    mine = tiles[rank]
    assert x_tile.shape[2:] == (mine.y1 - mine.y0, mine.x1 - mine.x0), "tile shape mismatch"
    # tiles differ in size by at most 4 px: pad to the largest so one all_gather does the exchange
    th = max(t.y1 - t.y0 for t in tiles)
    tw = max(t.x1 - t.x0 for t in tiles)
    pad = x_tile.new_zeros((T, C, th, tw))
    pad[:, :, :x_tile.shape[2], :x_tile.shape[3]] = x_tile
    gathered = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(gathered, pad, group=group)
    region = x_tile.new_empty((T, C, mine.hy1 - mine.hy0, mine.hx1 - mine.hx0))
    for t, g in zip(tiles, gathered):          # paste the parts of every tile that fall in my region
        ya, yb = max(t.y0, mine.hy0), min(t.y1, mine.hy1)
        xa, xb = max(t.x0, mine.hx0), min(t.x1, mine.hx1)
        if ya < yb and xa < xb:
            region[:, :, ya - mine.hy0:yb - mine.hy0, xa - mine.hx0:xb - mine.hx0] = \
                g[:, :, ya - t.y0:yb - t.y0, xa - t.x0:xb - t.x0]
    y = forward(region)
    centre = y[:, :, mine.y0 - mine.hy0:mine.y1 - mine.hy0,
               mine.x0 - mine.hx0:mine.x1 - mine.hx0].contiguous()
    if not gather_output:
        return centre
    opad = centre.new_zeros((T, centre.shape[1], th, tw))
    opad[:, :, :centre.shape[2], :centre.shape[3]] = centre
    outs = [torch.empty_like(opad) for _ in range(world)]
    dist.all_gather(outs, opad, group=group)
    full = centre.new_empty((T, centre.shape[1], H, W))
    for t, o in zip(tiles, outs):
        full[:, :, t.y0:t.y1, t.x0:t.x1] = o[:, :, :t.y1 - t.y0, :t.x1 - t.x0]
    return full
