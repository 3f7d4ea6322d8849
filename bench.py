#!/usr/bin/env python
"""bench.py — denoised frames/s of the BSVD-64 forward at 540x960 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # B200 arm (this repo's CUDA path)
    python bench.py --impl reference --steps K --warmup W    # reference arm: CPU forward (oracle port)

A "step" is one pass of the hot path over one synthetic clip [1,10,4,540,960] per GPU
(BASELINE.json configs[1]); N>1 shards independent clips one per GPU (weak scaling) and gathers
the outputs over NCCL inside the timed region.  One JSON line on stdout (rank 0).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "denoised frames/sec at 540x960 (c=64)"
UNIT = "frames/s"
T_CLIP, H, W = 10, 540, 960
FLOP_PER_PX = 2367360          # algorithmic FLOP per output pixel per frame
ELEMS_PER_PX = 2445            # algorithmic activation elements per pixel per frame


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "tflops_burst": d["bf16_tflops"],
                "tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0,
            "source": "fallback (B200_PROFILING.md)"}


# ------------------------------------------------------------------------------------------------
# clock sampling during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                c, m = float(parts[0]), float(parts[1])
            except ValueError:
                continue
            mx = m
            if t0 - 0.05 <= ts <= t1 + 0.15:
                sm.append(c)
                for n, v in zip(names, parts[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        if not sm:
            sm = [float(l.split(",")[0]) for _, l in self.lines[-3:] if l.split(",")[0].strip().replace(".", "").isdigit()]
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU baseline (oracle port of the reference forward, torch CPU = what the reference runs on CPU)
# ------------------------------------------------------------------------------------------------
def cpu_forward_fps(frames, steps, warmup):
    import torch
    from oracle import bsvd_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    layers = O.layers_from_tsn_state(O.make_synthetic_params(0, 0.5))
    x, _ = O.make_synthetic_clip(frames, H, W, seed=1)
    for _ in range(warmup):
        O.forward_clip(layers, x)
    t0 = time.time()
    for _ in range(steps):
        O.forward_clip(layers, x)
    dt = time.time() - t0
    return frames * steps / dt, dt / steps, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    frames = 1
    fps, s_per_step, cores = cpu_forward_fps(frames, args.steps, args.warmup)
    sample = (f"[1,{frames},4,{H},{W}] fp32 per step (per-frame cost of the forward does not depend "
              f"on clip length), torch CPU conv2d, {cores} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": s_per_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "BSVD-64 forward, 1 clip [1,10,4,540,960], fp32 (bounded sample)",
                   "sample": sample},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def stage_alg(cin, cout, stride, first, final, hw_out):
    """Algorithmic FLOPs and 16-bit activation bytes of one conv stage per frame (SURVEY §8d
    convention: input read once, output written once, skip/residual operand one extra read)."""
    cin_eff = 4 if first else cin
    cout_eff = 3 if final else cout
    flops = 2.0 * 9 * cin_eff * cout_eff * hw_out
    hw_in = hw_out * stride * stride
    elems = cin_eff * hw_in + cout_eff * hw_out
    return flops, elems


def run_b200(args):
    import torch
    import torch.distributed as dist
    from bsvd_b200 import capi
    from bsvd_b200.arch import BSVD
    from oracle import bsvd_oracle as O

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    prec = args.precision

    sd = O.make_synthetic_params(0, 0.5)
    net = BSVD(chns=[64, 128, 256], mid_ch=64, shift_input=False, norm='none', interm_ch=64,
               act='relu6', pretrain_ckpt=None, precision=prec)
    net.load_tsn_state(sd)
    net = net.to(dev).eval()
    x_host, clean_host = O.make_synthetic_clip(T_CLIP, H, W, seed=1 + rank)
    x_host = x_host.pin_memory()
    x_dev = x_host.to(dev)
    gathered = None
    if world > 1:
        gathered = torch.empty((world, T_CLIP, 3, H, W), dtype=torch.float32, device=dev)

    def step():
        with torch.no_grad():
            y = net(x_dev[None])[0]
        if world > 1:
            dist.all_gather_into_tensor(gathered, y.contiguous())
        return y

    lib = capi.load_library()
    for _ in range(max(args.warmup, 3)):
        y = step()
    torch.cuda.synchronize()
    capi.check(lib.bsvd_set_profiling(net._handle, 1))
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e0.record()
    for _ in range(args.steps):
        y = step()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t_wall1 = time.time()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    stage_ms = (C.c_float * capi.NUM_STAGES)()
    passes = C.c_int(0)
    capi.check(lib.bsvd_get_stage_ms(net._handle, stage_ms, capi.NUM_STAGES, C.byref(passes)))
    capi.check(lib.bsvd_set_profiling(net._handle, 0))
    launches = args.steps * net.last_launch_count
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * T_CLIP * args.steps / (ms_total / 1e3)

    # ---- end-to-end through the host-buffer C-ABI entry (pinned host -> device -> pinned host)
    out_host = torch.empty((T_CLIP, 3, H, W), dtype=torch.float32).pin_memory()
    for _ in range(2):
        net.denoise_host(x_host, out_host=out_host)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    # Each step copies this step's input from pinned host memory and its result back; the
    # pipelined entry overlaps step i+1's copy-in and step i-1's copy-out with step i's compute
    # (two host output buffers alternate so no result is overwritten before it is complete).
    out_hosts = [out_host, torch.empty_like(out_host).pin_memory()]
    for w in range(3):                       # warm both staging sets and both host buffers
        net.denoise_host_async(x_host, out_hosts[w & 1])
    net.host_sync()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(args.steps):
        net.denoise_host_async(x_host, out_hosts[i & 1])
    net.host_sync()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_check = float(out_hosts[(args.steps - 1) & 1].abs().sum())   # touch the last result
    # synchronous variant (one clip at a time, no overlap) for reference
    t0s = time.perf_counter()
    for _ in range(max(2, args.steps // 3)):
        net.denoise_host(x_host, out_host=out_host)
    e2e_sync_fps = T_CLIP * max(2, args.steps // 3) / (time.perf_counter() - t0s) * world
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_fps = world * T_CLIP * args.steps / e2e_s

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- parity of what was just timed (bounded: 2 frames of the clip, fp32 oracle on the CPU)
    peaks = load_peaks()
    npx = H * W
    kernels = {}
    ci, co, st, nt, rw = (C.c_int() for _ in range(5))
    for s in range(1, capi.NUM_STAGES):
        capi.check(lib.bsvd_stage_info(net._handle, s, C.byref(ci), C.byref(co), C.byref(st),
                                       C.byref(nt), C.byref(rw)))
        l = (s - 1) % 16
        res_div = {0: 1, 1: 1, 2: 4, 3: 4, 4: 4, 5: 16, 6: 16, 7: 16, 8: 16, 9: 16, 10: 16,
                   11: 4, 12: 4, 13: 4, 14: 1, 15: 1}[l]
        fl, el = stage_alg(ci.value, co.value, st.value, s == 1, s == 32, npx / res_div)
        if l in (10, 13):     # skip-add operand: one extra read of the output-shaped tensor
            el += co.value * (npx / res_div)
        if l == 15:
            el += 3 * npx
        name = ("first_conv_kernel" if s == 1 else "final_conv_kernel" if s == 32
                else f"conv3x3_tc_kernel<{nt.value},{rw.value}>")
        k = kernels.setdefault(name, {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
        k["ms"] += stage_ms[s] / max(passes.value, 1)
        k["flops"] += fl * T_CLIP
        k["bytes"] += el * 2 * T_CLIP
        k["launches"] += 1
    dom_name = max(kernels, key=lambda n: kernels[n]["ms"])
    dom = kernels[dom_name]
    conv_ms = sum(k["ms"] for k in kernels.values())
    peak_tf = peaks["tflops_sustained"]
    ach_tf = dom["flops"] / (dom["ms"] * 1e-3) / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        try:
            with open(tpath) as f:
                traffic = json.load(f).get(dom_name)
        except Exception:  # noqa: BLE001
            traffic = None
    roofline = {
        "bound": "tensor", "kernel": dom_name, "achieved": ach_tf, "peak": peak_tf,
        "unit": "TFLOP/s", "frac": ach_tf / peak_tf, "traffic": traffic,
        "peak_source": peaks["source"] + ", bf16 dense sustained (kernel timed inside a long step)",
        "launches_per_step": dom["launches"], "avg_launch_ms": dom["ms"] / dom["launches"],
        "share_of_step": dom["ms"] / ms_step,
        "hbm": {"achieved": dom["bytes"] / (dom["ms"] * 1e-3) / 1e9, "peak": peaks["hbm_gbs"],
                "unit": "GB/s", "frac": dom["bytes"] / (dom["ms"] * 1e-3) / 1e9 / peaks["hbm_gbs"]},
    }
    whole = {
        "alg_tflops": FLOP_PER_PX * npx * T_CLIP / (ms_step * 1e-3) / 1e12,
        "alg_hbm_gbs_16bit": ELEMS_PER_PX * npx * 2 * T_CLIP / (ms_step * 1e-3) / 1e9,
    }
    whole["tensor_frac"] = whole["alg_tflops"] / peak_tf
    whole["hbm_frac"] = whole["alg_hbm_gbs_16bit"] / peaks["hbm_gbs"]
    per_kernel = {n: {"ms_per_step": k["ms"], "tflops": k["flops"] / (k["ms"] * 1e-3) / 1e12,
                      "hbm_gbs": k["bytes"] / (k["ms"] * 1e-3) / 1e9, "launches": k["launches"]}
                  for n, k in kernels.items()}

    # ---- parity check of the timed configuration against the fp32 oracle (bounded sample)
    from oracle import bsvd_oracle as O2
    layers = O2.layers_from_tsn_state(sd)
    xs = x_host[:2].clone()
    with torch.no_grad():
        ys = net(xs[None].to(dev))[0].float().cpu()
    ref = O2.forward_clip(layers, xs)
    # PSNR delta vs the reference (calculate_psnr_float semantics: clamp to [0,1], crop_border 2)
    dpsnr = [O2.psnr_float(ys[i].clamp(0, 1), clean_host[i]) - O2.psnr_float(ref[i].clamp(0, 1), clean_host[i])
             for i in range(ys.shape[0])]
    parity = {"max_abs": float((ys - ref).abs().max()), "mean_abs": float((ys - ref).abs().mean()),
              "tolerance": 1e-3 if prec != "bf16" else 1e-2,
              "psnr_delta_db": {"mean": float(sum(dpsnr) / len(dpsnr)), "worst": float(min(dpsnr))},
              "sample": f"[1,2,4,{H},{W}] vs fp32 CPU oracle"}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        fps, _, cores = cpu_forward_fps(2, 1, 1)
        cpu = {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"[1,2,4,{H},{W}] fp32 (2 of the 10 frames; per-frame cost is independent "
                         f"of clip length), 1 warm-up + 1 timed pass, torch CPU conv2d"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": ("f16" if prec != "bf16" else "bf16") + " operands, f32 accumulate (tcgen05 kind::f16)",
        "data": "synthetic",
        "config": {"workload": f"BSVD-64 forward, 1 clip [1,{T_CLIP},4,{H},{W}] per GPU "
                               "(BASELINE.json configs[1]); fp32 in/out, tolerance 1e-3 vs fp32 reference",
                   "clips_per_step": world, "frames_per_clip": T_CLIP,
                   "sharding": "one clip per GPU, NCCL all_gather of outputs" if world > 1 else "single GPU",
                   "l2": "per-layer tensors are 0.17-0.66 GB each (working set 4.7 GB per step) >> 126 MB L2; no explicit flush needed",
                   "weights": "seeded synthetic, 0.5 x kaiming (SURVEY 8d); random init, no checkpoint available"},
        "e2e": {"value": e2e_fps, "unit": UNIT, "h2d_bytes_per_step": int(x_host.numel() * 4),
                "d2h_bytes_per_step": int(out_host.numel() * 4),
                "path": "BSVD.denoise_host_async -> bsvd_forward_clip_host_async (pinned host buffers; every step's H2D and D2H copies are inside the timed region, overlapped across steps on copy streams)",
                "unpipelined_value": e2e_sync_fps, "checksum": e2e_check},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roofline,
        "whole_net": whole,
        "kernels": per_kernel,
        "conv_ms_per_step": conv_ms,
        "stage_ms": [round(stage_ms[s] / max(passes.value, 1), 4) for s in range(1, capi.NUM_STAGES)],
        "parity": parity,
        "cpu_baseline": cpu,
        "workspace_bytes": int(lib.bsvd_workspace_bytes(net._handle)),
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="fp16", choices=["fp16", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # convenience: the driver launches torchrun itself; do the same when called bare
        port = 29500 + (os.getpid() % 2000)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
               f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
