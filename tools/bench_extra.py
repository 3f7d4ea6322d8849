"""Secondary BASELINE.json configurations (bench.py owns configs[1] / the driver contract):

  python tools/bench_extra.py stream   [--frames 100] [--precision bf16]     # configs[2]
  torchrun ... tools/bench_extra.py tiles --rows 4 --cols 2                   # configs[4] (world = rows*cols)

Each prints one JSON line (rank 0).  CUDA-event timing, max over ranks for multi-GPU."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bsvd_b200.arch import BSVD  # noqa: E402
from bsvd_b200 import tiling  # noqa: E402
from oracle import bsvd_oracle as O  # noqa: E402


def make_net(prec, dev):
    net = BSVD(chns=[64, 128, 256], mid_ch=64, shift_input=False, norm='none', interm_ch=64,
               act='relu6', pretrain_ckpt=None, precision=prec)
    net.load_tsn_state(O.make_synthetic_params(0, 0.5))
    return net.to(dev).eval()


def stream(args):
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    net = make_net(args.precision, dev)
    H, W, F = 540, 960, args.frames
    x, _ = O.make_synthetic_clip(min(F, 20), H, W, seed=1)
    frames = [x[i % x.shape[0]:i % x.shape[0] + 1].to(dev) for i in range(F)]

    def run():
        net.reset()
        outs = 0
        for f in frames:
            if net.feedin_one_element(f) is not None:
                outs += 1
        while outs < F:
            if net.feedin_one_element(None) is not None:
                outs += 1
        net.reset()
        return outs

    with torch.no_grad():
        run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            run()
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    # parity of the streaming schedule on a short prefix (fp32 oracle, CPU)
    xs = x[:3]
    with torch.no_grad():
        net.reset()
        outs = [net.feedin_one_element(xs[i:i + 1].to(dev)) for i in range(3)]
        while sum(o is not None for o in outs) < 3:
            outs.append(net.feedin_one_element(None))
        net.reset()
    y = torch.cat([o for o in outs if o is not None]).float().cpu()
    ref = O.forward_clip(O.layers_from_tsn_state(O.make_synthetic_params(0, 0.5)), xs)
    print(json.dumps({
        "metric": "denoised frames/sec at 540x960 (c=64)", "value": F / (ms / 1e3), "unit": "frames/s",
        "n_gpus": 1, "steps": args.steps, "ms_per_step": ms, "dtype": args.precision,
        "config": {"workload": f"BSVD-64 streaming bidirectional-buffer mode, {F}-frame 540x960 "
                               f"sequence, {args.precision} (BASELINE.json configs[2]); one "
                               "feedin_one_element call per frame + 16 drain calls",
                   "latency_frames": 16},
        "parity": {"max_abs": float((y - ref).abs().max()),
                   "tolerance": 1e-2 if args.precision == "bf16" else 1e-3},
        "workspace_bytes": None}), flush=True)


def c32(args):
    """The blind c32 configuration (options/train/0402_*_blind_c32.yml) and the fused caller entry
    (bsvd_denoise_clip) on a [10,3,540,960] clip: frames/s, parity vs the fp32 oracle on 2 frames."""
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    c = O.C32
    sd = O.make_synthetic_params(0, 0.5, in_ch=3, chns=c["chns"], mid_ch=c["mid_ch"], interm_ch=c["interm_ch"])
    net = BSVD(chns=list(c["chns"]), mid_ch=c["mid_ch"], shift_input=False, norm='none',
               interm_ch=c["interm_ch"], act=c["act"], blind=True, pretrain_ckpt=None,
               precision=args.precision)
    net.load_tsn_state(sd)
    net = net.to(dev).eval()
    H, W, T = 540, 960, args.frames
    x, _ = O.make_synthetic_clip(T, H, W, seed=1)
    x3 = x[:, :3].contiguous().to(dev)
    lines = []
    for name, fn in (("c32 forward", lambda: net(x3[None])[0]),
                     ("c32 denoise_sequence (fused pad/clamp/crop entry)", lambda: net.denoise_sequence(x3, None))):
        with torch.no_grad():
            for _ in range(3):
                y = fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                y = fn()
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        lines.append((name, T / ms * 1e3, ms))
    with torch.no_grad():
        ys = net(x3[None, :2])[0].float().cpu()
    ref = O.forward_clip(O.layers_from_tsn_state(sd), x[:2, :3], act=c["act"])
    for name, fps, ms in lines:
        print(json.dumps({
            "metric": "denoised frames/sec at 540x960 (c=32, blind)", "value": fps, "unit": "frames/s",
            "n_gpus": 1, "steps": args.steps, "ms_per_step": ms, "dtype": args.precision,
            "config": {"workload": name + f", 1 clip [1,{T},3,{H},{W}]; zero-padded onto the 64-channel kernels"},
            "parity": {"max_abs": float((ys - ref).abs().max()), "tolerance": 1e-3}}), flush=True)


def u8(args):
    """Decoded frames end to end: pinned uint8 [10,540,960,3] on the host -> H2D -> bsvd_denoise_clip_u8
    -> D2H of the uint8 result, every step (31 MB over PCIe per clip instead of 145 MB as fp32)."""
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    net = make_net(args.precision, dev)
    T, H, W = args.frames, 540, 960
    x, _ = O.make_synthetic_clip(T, H, W, seed=1)
    frames = (x[:, :3].clamp(0, 1) * 255).round().to(torch.uint8).permute(0, 2, 3, 1).contiguous().pin_memory()
    out_h = torch.empty_like(frames).pin_memory()
    sigma = 20.0 / 255.0

    def step():
        d = frames.to(dev, non_blocking=True)
        y = net.denoise_frames_u8(d, sigma)
        out_h.copy_(y, non_blocking=True)

    with torch.no_grad():
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        import time
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print(json.dumps({
        "metric": "denoised frames/sec at 540x960 (c=64)", "value": T * args.steps / dt, "unit": "frames/s",
        "n_gpus": 1, "steps": args.steps, "ms_per_step": dt / args.steps * 1e3, "dtype": args.precision,
        "config": {"workload": f"uint8 HWC frames end to end (host pinned -> device -> host), 1 clip [{T},{H},{W},3] "
                               "per step, one stream, no overlap between steps; bsvd_denoise_clip_u8"},
        "h2d_bytes_per_step": int(frames.numel()), "d2h_bytes_per_step": int(out_h.numel()),
        "checksum": int(out_h.long().sum())}), flush=True)


def torch_gpu(args):
    """Like-for-like bar: the reference's own formulation — one F.conv2d / pixel_shuffle /
    cat per op through PyTorch + cuDNN (the oracle restatement makes exactly the calls the reference's
    TSN/BSVD classes make) — on the SAME B200, for fp32 (TF32 off / on) and fp16 tensors."""
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    T, H, W = args.frames, 540, 960
    sd = O.make_synthetic_params(0, 0.5)
    x, _ = O.make_synthetic_clip(T, H, W, seed=1)
    for name, dt, tf32 in (("fp32, TF32 off", torch.float32, False), ("fp32, TF32 on (PyTorch default for convs)", torch.float32, True),
                           ("fp16 weights and activations (profile.py mode)", torch.float16, True)):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.benchmark = True
        layers = [(w.to(dev, dt), b.to(dev, dt)) for w, b in O.layers_from_tsn_state(sd)]
        xd = x.to(dev, dt)
        with torch.no_grad():
            for _ in range(2):
                y = O.forward_clip(layers, xd)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                y = O.forward_clip(layers, xd)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        print(json.dumps({
            "metric": "denoised frames/sec at 540x960 (c=64)", "value": T / ms * 1e3, "unit": "frames/s",
            "n_gpus": 1, "steps": args.steps, "ms_per_step": ms, "impl": "PyTorch + cuDNN, op by op (reference formulation)",
            "dtype": name, "config": {"workload": f"BSVD-64 forward, 1 clip [1,{T},4,{H},{W}], clip order"},
            "checksum": float(y.float().abs().sum())}), flush=True)
        del layers, xd, y
        torch.cuda.empty_cache()


def tiles(args):
    """BASELINE.json configs[4]: one 4K clip cut into rows x cols spatial tiles, one per GPU.
    --exchange p2p  (default) neighbour-only strips and output centres through peer-mapped memory with
                    copy engines over NVLink (bsvd_b200.tiling.TileExchange), steps overlapped;
    --exchange nccl the round-1 variant: all_gather of whole input tiles and of the outputs."""
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    assert world == args.rows * args.cols, "world size must equal rows*cols"
    dist.init_process_group("nccl", device_id=dev)
    net = make_net(args.precision, dev)
    T, H, W = args.frames, args.height, args.width
    plan = tiling.tile_plan(H, W, args.rows, args.cols, balance=args.balance)
    t = plan[rank]
    # every rank synthesises the same frame and keeps only its own tile (stands for a decoder that
    # delivers tiles); the 80-px ring then comes from the neighbours
    x, _ = O.make_synthetic_clip(T, H, W, seed=1)
    x_tile = x[:, :, t.y0:t.y1, t.x0:t.x1].contiguous().to(dev)
    fwd = lambda r: net(r[None])[0]  # noqa: E731
    ex, consumer, chk = None, None, None
    if args.exchange == "p2p":
        ex = tiling.TileExchange(T, 4, H, W, args.rows, args.cols, owner=0, balance=args.balance)
        consumer = torch.cuda.Stream(device=dev)
        chk = torch.zeros((), dtype=torch.float64, device=dev)
    ev = torch.cuda.Event()
    counter = [0]
    last_full = [None]

    def step():
        with torch.no_grad():
            if ex is None:
                last_full[0] = tiling.forward_tiled_distributed(fwd, x_tile, H, W, args.rows, args.cols)
                return
            i = counter[0]
            counter[0] += 1
            full = ex.step(fwd, x_tile, i)
            if rank == 0:
                with torch.cuda.stream(consumer):
                    ex.wait_full(i)
                    chk.add_(full[:, 0, ::270, ::480].double().sum())     # touch the assembled frame
                    last_full[0] = full
                    ex.release_full(i)

    def drain():
        if consumer is not None:
            ev.record(consumer)
            torch.cuda.current_stream(dev).wait_event(ev)
        if ex is not None:
            ev.record(ex.pg.side)
            torch.cuda.current_stream(dev).wait_event(ev)

    for _ in range(3):
        step()
    drain()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    drain()
    e1.record()
    dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    exact, single_ms = None, None
    if args.check and rank == 0:
        with torch.no_grad():
            xd = x[None].to(dev)
            whole = net(xd)[0]
            exact = bool(torch.equal(whole, last_full[0]))
            torch.cuda.synchronize()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            for _ in range(2):
                net(xd)
            s1.record()
            torch.cuda.synchronize()
            single_ms = s0.elapsed_time(s1) / 2
    if rank == 0:
        fps = T / (float(ms) / 1e3)
        line = {
            "metric": f"denoised frames/sec at {H}x{W} (c=64)", "value": fps,
            "unit": "frames/s", "n_gpus": world, "steps": args.steps, "ms_per_step": float(ms),
            "scaling": "strong", "dtype": args.precision,
            "config": {"workload": f"BSVD-64 {H}x{W} {T}-frame clip, {args.rows}x{args.cols} spatial "
                                   f"tiles, one per GPU, {tiling.HALO}-px input ring (BASELINE.json configs[4], bit-exact variant)",
                       "exchange": ("neighbour-only strips + output centres: copy-engine puts into peer-mapped buffers over "
                                    "NVLink, double-buffered and overlapped across steps" if ex is not None
                                    else "NCCL all_gather of whole input tiles and of the outputs")},
            "bit_exact_vs_single_gpu": exact}
        if ex is not None:
            line["received_bytes_per_step"] = ex.received_bytes
            line["ring_bytes"] = ex.ring_bytes
            line["balanced_tiles"] = bool(args.balance)
            line["largest_enlarged_tile"] = max((q.hy1 - q.hy0) * (q.hx1 - q.hx0) for q in plan)
            line["redundant_compute"] = sum((q.hy1 - q.hy0) * (q.hx1 - q.hx0) for q in plan) / float(H * W)
        if single_ms is not None:
            line["single_gpu"] = {"ms_per_clip": single_ms, "value": T / (single_ms / 1e3)}
            line["strong_scaling_efficiency"] = fps / (T / (single_ms / 1e3)) / world
        print(json.dumps(line), flush=True)
    if ex is not None:
        ex.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", choices=["stream", "tiles", "c32", "u8", "torch"])
    ap.add_argument("--frames", type=int, default=None)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--precision", default=None)
    ap.add_argument("--rows", type=int, default=4)
    ap.add_argument("--cols", type=int, default=2)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"])
    ap.add_argument("--balance", action="store_true", help="tiles: equalise the enlarged tile sizes (border tiles own more)")
    a = ap.parse_args()
    if a.mode == "torch":
        a.frames = a.frames or 10
        torch_gpu(a)
    elif a.mode == "u8":
        a.frames = a.frames or 10
        a.precision = a.precision or "fp16"
        u8(a)
    elif a.mode == "c32":
        a.frames = a.frames or 10
        a.precision = a.precision or "fp16"
        c32(a)
    elif a.mode == "stream":
        a.frames = a.frames or 100
        a.precision = a.precision or "bf16"
        stream(a)
    else:
        a.frames = a.frames or 10
        a.precision = a.precision or "fp16"
        tiles(a)
