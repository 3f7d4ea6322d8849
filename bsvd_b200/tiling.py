"""Spatial-tile mode for frames too large for one GPU's time budget (BASELINE config 5: 4K clips).

The frame is cut into rows x cols tiles, one per GPU (one process per GPU).  Each rank receives
the input pixels of a `HALO`-pixel (80) ring around its tile from the (up to 8) neighbours that own
them — `TileExchange`: copy-engine puts into the neighbour's peer-mapped buffer over NVLink, only the
strips that are needed, the 4-channel input being the only tensor that ever crosses GPUs — runs the
unmodified BSVD-64 path on the enlarged tile and writes the centre of its output straight into the
owner's full frame.  (`forward_tiled_distributed` is the older variant that all_gathers whole tiles
through NCCL; it is kept for the CPU/gloo test of the paste logic.)

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


def tile_plan(H: int, W: int, rows: int, cols: int, halo: int = HALO, balance: bool = False):
    """rows x cols tiles whose edges are multiples of 4; row-major order (tile r*cols+c).
    balance=True equalises the ENLARGED sizes instead of the owned ones: a tile in the middle carries the
    halo on both sides, a tile at the frame border on one, so border tiles own `halo` more pixels and every
    rank computes the same area (the slowest rank sets the step time)."""
    if H % 4 or W % 4:
        raise ValueError("H and W must be multiples of 4")
    if halo % 4:
        raise ValueError("halo must be a multiple of 4")

    def cuts(n, k):
        q = n // 4
        edges = [4 * ((q * i) // k) for i in range(k)] + [n]
        if balance and k > 2:
            e = (n + 2 * halo * (k - 1)) / k                      # common enlarged size
            inner = max(4, int(round((e - 2 * halo) / 4)) * 4)    # owned by a tile with two halos
            first = (n - inner * (k - 2)) // 2 // 4 * 4           # owned by the two border tiles (the last takes the rest)
            if first >= 4 and n - first - inner * (k - 2) >= 4:
                edges = [0] + [first + inner * i for i in range(k - 1)] + [n]
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


# ---------------------------------------------------------------------------------------------------
# neighbour-only exchange over NVLink (peer-mapped memory, copy engines)
# ---------------------------------------------------------------------------------------------------
def overlap(a: Tile, b: Tile):
    """Part of tile a's OWN pixels that falls inside tile b's enlarged region: (ya, yb, xa, xb) or None."""
    ya, yb = max(a.y0, b.hy0), min(a.y1, b.hy1)
    xa, xb = max(a.x0, b.hx0), min(a.x1, b.hx1)
    return (ya, yb, xa, xb) if ya < yb and xa < xb else None


def exchange_plan(tiles):
    """sends[r] = [(dst, ya, yb, xa, xb)]: the strips rank r owns that rank dst needs (dst == r: its own
    centre).  Only tiles whose regions actually overlap appear: at most 8 neighbours + self."""
    sends = []
    for r, a in enumerate(tiles):
        lst = []
        for d, b in enumerate(tiles):
            o = overlap(a, b)
            if o is not None:
                lst.append((d,) + o)
        sends.append(lst)
    recv_from = [[r for r, lst in enumerate(sends) if any(d == me for d, *_ in lst)] for me in range(len(tiles))]
    return sends, recv_from


class TileExchange:
    """One spatial tile per rank; per step: neighbours' strips in, forward on the enlarged tile, centre
    out to the owner rank's full frame.  All transfers are bsvd_peer_put3d copies on the side stream
    (one per neighbour: [T*C planes][rows][cols] in a single copy-engine transfer), ordered by flags:

      in_ready [k][src]   on dst : src's strip for slot k has landed              (value step+1)
      in_free  [k][dst]   on src : dst's forward has consumed its region slot k   (value step+1)
      out_ready[k][src]   on owner: src's centre is in full-frame slot k
      out_free [k]        on all  : the owner has consumed full-frame slot k
    """

    def __init__(self, T, C, H, W, rows, cols, halo: int = HALO, owner: int = 0, depth: int = 2, group=None,
                 backend: str = "cuda", balance: bool = False):
        from .peer import PeerGroup
        import torch.distributed as dist
        self.T, self.C, self.H, self.W, self.depth, self.owner = T, C, H, W, depth, owner
        self.tiles = tile_plan(H, W, rows, cols, halo, balance=balance)
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        if world != len(self.tiles):
            raise ValueError(f"{len(self.tiles)} tiles need {len(self.tiles)} ranks, got {world}")
        self.sends, self.recv_from = exchange_plan(self.tiles)
        rh = max(t.hy1 - t.hy0 for t in self.tiles)
        rw = max(t.hx1 - t.hx0 for t in self.tiles)
        self.region_bytes = (T * C * rh * rw * 4 + 1023) // 1024 * 1024
        self.full_bytes = (T * 3 * H * W * 4 + 1023) // 1024 * 1024
        self.off_region = lambda k: k * self.region_bytes
        self.off_full = lambda k: depth * self.region_bytes + k * self.full_bytes
        nbytes = depth * (self.region_bytes + self.full_bytes)
        w = world
        self.f_in_ready = lambda k, src: k * w + src
        self.f_in_free = lambda k, dst: depth * w + k * w + dst
        self.f_out_ready = lambda k, src: 2 * depth * w + k * w + src
        self.f_out_free = lambda k: 3 * depth * w + k
        self.pg = PeerGroup(nbytes, 3 * depth * w + depth, group, backend=backend)
        self.rank, self.world = self.pg.rank, world
        self.me = self.tiles[self.rank]
        self.ev_in = self.pg.new_event()
        self.ev_fwd = self.pg.new_event()
        self.ev_own = self.pg.new_event()
        me = self.me
        self.ring_bytes = T * C * 4 * ((me.hy1 - me.hy0) * (me.hx1 - me.hx0) - (me.y1 - me.y0) * (me.x1 - me.x0))
        self.received_bytes = 0

    def step(self, forward, x_tile: torch.Tensor, step: int):
        """x_tile: this rank's [T,C,th,tw] fp32 input tile (device, contiguous).  Enqueues the whole step;
        returns the owner's full-frame view [T,3,H,W] of this step on the owner rank (valid after
        `wait_full(step)`), None elsewhere."""
        pg, k, me, T, C = self.pg, step % self.depth, self.me, self.T, self.C
        th, tw = me.y1 - me.y0, me.x1 - me.x0
        assert tuple(x_tile.shape) == (T, C, th, tw) and x_tile.is_contiguous() and x_tile.dtype == torch.float32
        cur = pg.current_stream()
        self.ev_in.record(cur)
        pg.side.wait_event(self.ev_in)
        for dst, ya, yb, xa, xb in self.sends[self.rank]:
            d = self.tiles[dst]
            dh, dw = d.hy1 - d.hy0, d.hx1 - d.hx0
            if step >= self.depth:
                pg.wait(self.f_in_free(k, dst), step - self.depth + 1, stream=pg.side)
            src_off = ((ya - me.y0) * tw + (xa - me.x0)) * 4
            dst_off = self.off_region(k) + ((ya - d.hy0) * dw + (xa - d.hx0)) * 4
            pg.put3d(dst, dst_off, dw * 4, dh, x_tile, tw * 4, th, (xb - xa) * 4, yb - ya, T * C, stream=pg.side,
                     src_off=src_off)
            pg.signal(dst, self.f_in_ready(k, self.rank), step + 1, stream=pg.side)
        if x_tile.is_cuda:
            x_tile.record_stream(pg.side)
        # ---- forward on the enlarged tile once every contributor's strip has landed
        for src in self.recv_from[self.rank]:
            pg.wait(self.f_in_ready(k, src), step + 1)
            if src != self.rank and step == 0:
                ya, yb, xa, xb = overlap(self.tiles[src], me)
                self.received_bytes += T * C * 4 * (yb - ya) * (xb - xa)
        hh, ww = me.hy1 - me.hy0, me.hx1 - me.hx0
        region = pg.local_tensor(self.off_region(k), (T, C, hh, ww))
        y = forward(region)                                    # [T,3,hh,ww]
        self.ev_fwd.record(cur)
        pg.side.wait_event(self.ev_fwd)
        for src in self.recv_from[self.rank]:                  # my region slot k may be refilled
            pg.signal(src, self.f_in_free(k, self.rank), step + 1, stream=pg.side)
        # ---- centre of the output straight into the owner's full frame
        if step >= self.depth:
            pg.wait(self.f_out_free(k), step - self.depth + 1, stream=pg.side)
        y = y.contiguous()
        src_off = ((me.y0 - me.hy0) * ww + (me.x0 - me.hx0)) * 4
        dst_off = self.off_full(k) + (me.y0 * self.W + me.x0) * 4
        pg.put3d(self.owner, dst_off, self.W * 4, self.H, y, ww * 4, hh, tw * 4, th, T * 3, stream=pg.side,
                 src_off=src_off)
        pg.signal(self.owner, self.f_out_ready(k, self.rank), step + 1, stream=pg.side)
        if y.is_cuda:
            y.record_stream(pg.side)
        if self.rank == self.owner:
            return pg.local_tensor(self.off_full(k), (T, 3, self.H, self.W))
        return None

    def wait_full(self, step: int):
        """Owner: the calling stream waits until every rank's centre of `step` has landed."""
        k = step % self.depth
        for src in range(self.world):
            self.pg.wait(self.f_out_ready(k, src), step + 1)

    def release_full(self, step: int):
        """Owner: reads of the full frame of `step` enqueued on the current stream so far are the last."""
        pg, k = self.pg, step % self.depth
        self.ev_own.record(pg.current_stream())
        pg.side.wait_event(self.ev_own)
        for r in range(self.world):
            pg.signal(r, self.f_out_free(k), step + 1, stream=pg.side)

    def close(self):
        self.pg.close()
