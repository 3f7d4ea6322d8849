"""Spatial-tile mode (bsvd_b200/tiling.py): plan arithmetic, exactness of the halo argument with
the CPU oracle as the network, and the world_size-2 exchange path over gloo."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bsvd_b200 import tiling
from oracle import bsvd_oracle as O


def test_tile_plan_covers_frame_and_is_aligned():
    for (H, W, r, c) in [(2160, 3840, 4, 2), (2160, 3840, 2, 4), (540, 960, 1, 2), (136, 264, 2, 3)]:
        tiles = tiling.tile_plan(H, W, r, c)
        assert len(tiles) == r * c
        cover = torch.zeros(H, W, dtype=torch.int32)
        for t in tiles:
            cover[t.y0:t.y1, t.x0:t.x1] += 1
            assert all(v % 4 == 0 for v in (t.y0, t.y1, t.x0, t.x1, t.hy0, t.hy1, t.hx0, t.hx1))
            assert t.hy0 == max(0, t.y0 - tiling.HALO) and t.hx1 == min(W, t.x1 + tiling.HALO)
        assert int(cover.min()) == 1 and int(cover.max()) == 1
    with pytest.raises(ValueError):
        tiling.tile_plan(18, 32, 1, 1)
    # balanced plan: same coverage / alignment, equal ENLARGED sizes along an axis with more than two tiles
    for (H, W, r, c) in [(2160, 3840, 4, 2), (2160, 3840, 2, 4), (540, 960, 1, 5)]:
        tiles = tiling.tile_plan(H, W, r, c, balance=True)
        cover = torch.zeros(H, W, dtype=torch.int32)
        for t in tiles:
            cover[t.y0:t.y1, t.x0:t.x1] += 1
            assert all(v % 4 == 0 for v in (t.y0, t.y1, t.x0, t.x1, t.hy0, t.hy1, t.hx0, t.hx1))
        assert int(cover.min()) == 1 and int(cover.max()) == 1
        areas = [(t.hy1 - t.hy0) * (t.hx1 - t.hx0) for t in tiles]
        plain = [(t.hy1 - t.hy0) * (t.hx1 - t.hx0) for t in tiling.tile_plan(H, W, r, c)]
        assert max(areas) < max(plain) and max(areas) - min(areas) <= 0.02 * max(areas)


def _oracle_forward():
    layers = O.layers_from_tsn_state(O.make_synthetic_params(0, 0.5))
    return lambda x: O.forward_clip(layers, x)


def test_halo_is_sufficient_with_oracle_network():
    """Receptive-field claim behind HALO=80: tiled == untiled for the fp32 oracle (CPU conv picks
    kernels by shape, so allow fp32 summation-order noise), and a too-small halo is NOT enough."""
    fwd = _oracle_forward()
    x, _ = O.make_synthetic_clip(2, 96, 336, seed=5)
    full = fwd(x)
    tiled = tiling.forward_tiled_local(fwd, x, 1, 2)
    assert float((tiled - full).abs().max()) < 2e-5
    small = tiling.forward_tiled_local(fwd, x, 1, 2, halo=8)
    assert float((small - full).abs().max()) > 1e-4


def _worker(rank, world, port, H, W, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    fwd = _oracle_forward()
    x, _ = O.make_synthetic_clip(1, H, W, seed=6)
    t = tiling.tile_plan(H, W, 1, world)[rank]
    out = tiling.forward_tiled_distributed(fwd, x[:, :, t.y0:t.y1, t.x0:t.x1].contiguous(), H, W,
                                           1, world)
    if rank == 0:
        q.put(out.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_distributed_exchange_world2_gloo():
    H, W = 48, 328
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, H, W, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = torch.from_numpy(q.get(timeout=240))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    x, _ = O.make_synthetic_clip(1, H, W, seed=6)
    full = _oracle_forward()(x)
    assert got.shape == full.shape
    assert float((got - full).abs().max()) < 2e-5


def test_exchange_plan_strips_rebuild_every_enlarged_region():
    """Neighbour-only exchange (TileExchange): the strips each rank sends, pasted at the receiver, are
    exactly the receiver's enlarged region; nobody sends to a tile that does not need it; the bytes a
    tile receives equal its ring."""
    H, W = 176, 352
    x = torch.arange(H * W, dtype=torch.float32).reshape(1, 1, H, W)
    for rows, cols in ((1, 2), (2, 2), (2, 4), (4, 2)):
        tiles = tiling.tile_plan(H, W, rows, cols)
        sends, recv_from = tiling.exchange_plan(tiles)
        for me, t in enumerate(tiles):
            reg = torch.full((1, 1, t.hy1 - t.hy0, t.hx1 - t.hx0), -1.0)
            got = 0
            for src, lst in enumerate(sends):
                for d, ya, yb, xa, xb in lst:
                    if d == me:
                        assert src in recv_from[me]
                        reg[:, :, ya - t.hy0:yb - t.hy0, xa - t.hx0:xb - t.hx0] = x[:, :, ya:yb, xa:xb]
                        if src != me:
                            got += (yb - ya) * (xb - xa)
            assert torch.equal(reg, x[:, :, t.hy0:t.hy1, t.hx0:t.hx1])
            ring = (t.hy1 - t.hy0) * (t.hx1 - t.hx0) - (t.y1 - t.y0) * (t.x1 - t.x0)
            assert got == ring
    # 4K, 2x4: at most 8 neighbours + self, and far tiles exchange nothing
    tiles = tiling.tile_plan(2160, 3840, 2, 4)
    sends, _ = tiling.exchange_plan(tiles)
    assert max(len(s) for s in sends) <= 9 and all(len(s) < len(tiles) for s in sends)


# ---------------------------------------------------------------------------------------------------
# the one-sided protocols of the multi-GPU path (bsvd_b200/peer.py ClipGather, tiling.TileExchange) under
# world_size-2 gloo on the CPU: same host code as on the GPUs, PeerGroup(backend="shm") in place of the
# CUDA-IPC buffers (ring slots, ready / free flags, strip geometry, slot reuse over more steps than slots)
# ---------------------------------------------------------------------------------------------------
def _peer_worker(rank, world, port, mode, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    ok = True
    if mode == "gather":
        from bsvd_b200.peer import ClipGather
        shape = (2, 3, 12, 20)
        g = ClipGather(shape, torch.float32, depth=2, backend="shm")
        for i in range(7):                                  # 7 steps through 2 slots: free flags are exercised
            torch.manual_seed(100 * i + rank)
            y = torch.randn(shape)
            want = [torch.empty(shape) for _ in range(world)]
            dist.all_gather(want, y)
            g.put(y, i)
            got = g.wait(i).clone()
            g.release(i)
            ok = ok and torch.equal(got, torch.stack(want))
        ok = ok and g.pg.read_flag(g.lay["ready"](0, 1 - rank)) == 7      # last even step 6 -> value 7
        g.close()
        q.put((rank, ok))
    else:
        fwd = _oracle_forward()
        T, H, W = 1, 48, 328
        ex = tiling.TileExchange(T, 4, H, W, 1, world, owner=0, backend="shm")
        me = ex.me
        outs = []
        for i in range(3):                                  # 3 steps through 2 slots
            x, _ = O.make_synthetic_clip(T, H, W, seed=6 + i)
            full = ex.step(fwd, x[:, :, me.y0:me.y1, me.x0:me.x1].contiguous(), i)
            if rank == 0:
                ex.wait_full(i)
                outs.append(full.clone().numpy())
                ex.release_full(i)
        info = (ex.received_bytes, ex.ring_bytes)
        ex.close()
        q.put((rank, outs if rank == 0 else None, info))
    dist.barrier()
    dist.destroy_process_group()


def _run_peer_workers(mode):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_peer_worker, args=(r, 2, port, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return sorted(res, key=lambda r: r[0])


def test_clip_gather_protocol_world2_gloo_shared_memory():
    res = _run_peer_workers("gather")
    assert all(r[1] for r in res), res


def test_tile_exchange_protocol_world2_gloo_shared_memory():
    res = _run_peer_workers("tiles")
    outs, (recv, ring) = res[0][1], res[0][2]
    assert recv == ring and res[1][2][0] == res[1][2][1]
    fwd = _oracle_forward()
    for i, got in enumerate(outs):
        x, _ = O.make_synthetic_clip(1, 48, 328, seed=6 + i)
        full = fwd(x)
        assert float((torch.from_numpy(got) - full).abs().max()) < 2e-5
