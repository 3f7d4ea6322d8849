"""torchrun worker for tests/test_gpu_multi.py (one process per GPU).  Prints one JSON line on rank 0.

  gather : ClipGather (copy-engine puts into peer-mapped rings) against NCCL all_gather over 6 steps
  tiles  : TileExchange on a small frame, rows x cols = 1 x world, against the single-GPU forward
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import bsvd_oracle as O  # noqa: E402


def make_net(dev):
    from bsvd_b200.arch import BSVD
    net = BSVD(chns=[64, 128, 256], mid_ch=64, shift_input=False, norm='none', interm_ch=64,
               act='relu6', pretrain_ckpt=None)
    net.load_tsn_state(O.make_synthetic_params(0, 0.5))
    return net.to(dev).eval()


def main():
    mode = sys.argv[1]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    res = {"world": world}
    if mode == "gather":
        from bsvd_b200.peer import ClipGather
        shape = (3, 3, 40, 56)
        g = ClipGather(shape, torch.float32, depth=2)
        consumer = torch.cuda.Stream(device=dev)
        ok, steps = True, 6
        for i in range(steps):
            torch.manual_seed(100 * i + rank)
            y = torch.randn(shape, device=dev)
            ref = torch.empty((world,) + shape, device=dev)
            dist.all_gather_into_tensor(ref, y)
            g.put(y, i)
            with torch.cuda.stream(consumer):
                got = g.wait(i)
                snap = got.clone()
                g.release(i)
            consumer.synchronize()
            ok = ok and bool(torch.equal(snap, ref))
        flags = [g.pg.read_flag(g.lay["ready"](k, r)) for k in range(2) for r in range(world)]
        res.update(ok=ok, steps=steps, ready_flags=flags)
        t = torch.tensor([1.0 if ok else 0.0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        res["ok"] = bool(t.item() == 1.0)
        g.close()
    elif mode == "tiles":
        from bsvd_b200 import tiling
        T, H, W = 3, 176, 352 if world == 2 else 176 * world
        rows, cols = 1, world
        net = make_net(dev)
        x, _ = O.make_synthetic_clip(T, H, W, seed=5)
        ex = tiling.TileExchange(T, 4, H, W, rows, cols, owner=0)
        me = ex.me
        fwd = lambda r: net(r[None])[0]  # noqa: E731
        exact = True
        with torch.no_grad():
            whole = net(x[None].to(dev))[0] if rank == 0 else None
            for i in range(4):          # more steps than ring slots: exercises the free flags
                xi = x if i % 2 == 0 else x.flip(0)
                x_tile = xi[:, :, me.y0:me.y1, me.x0:me.x1].contiguous().to(dev)
                full = ex.step(fwd, x_tile, i)
                if rank == 0:
                    ex.wait_full(i)
                    snap = full.clone()
                    ex.release_full(i)
                    want = whole if i % 2 == 0 else net(xi[None].to(dev))[0]
                    exact = exact and bool(torch.equal(snap, want))
        torch.cuda.synchronize()
        res.update(bit_exact=exact, received_bytes=ex.received_bytes, ring_bytes=ex.ring_bytes,
                   tiles=[(t.y0, t.y1, t.x0, t.x1) for t in ex.tiles])
        ex.close()
    if rank == 0:
        print(json.dumps(res), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
