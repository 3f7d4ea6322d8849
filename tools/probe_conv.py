"""GPU probe: run every fused-conv kernel variant through bsvd_conv_stage and compare with a plain
fp32 torch restatement of the same stage.  Prints one line per case and writes
gpurun_out/probe_conv.json.  Development tool (tests/test_gpu_conv_stage.py is the real test)."""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bsvd_b200 import capi  # noqa: E402


def shifted(y):
    """out[t][0:f] = y[t+1][0:f]; out[t][f:2f] = y[t-1][f:2f]; rest = y[t] (zeros off-clip)."""
    T, C = y.shape[0], y.shape[1]
    f = C // 8
    o = y.clone()
    o[:, :2 * f] = 0
    o[:-1, :f] = y[1:, :f]
    o[1:, f:2 * f] = y[:-1, f:2 * f]
    return o


def run_case(lib, T, H, W, cin, cout, flags, prec, variant, seed=0):
    g = torch.Generator().manual_seed(seed)
    dt = torch.float16 if prec == capi.PREC_FP16 else torch.bfloat16
    x = torch.rand(T, cin, H, W, generator=g)
    w = torch.randn(cout, cin, 3, 3, generator=g) * (0.7 * (2.0 / (9 * cin)) ** 0.5)
    b = torch.randn(cout, generator=g) * 0.1
    stride = 2 if flags & capi.EPI_STRIDE2 else 1
    x16 = x.to(dt)
    w16 = w.to(dt).float()
    dev = torch.device("cuda")
    ref = F.conv2d(x16.float().to(dev), w16.to(dev), b.to(dev), stride=stride, padding=1)
    if flags & capi.EPI_PIXSHUF:
        ref = F.pixel_shuffle(ref, 2)
    skip16 = None
    if flags & capi.EPI_SKIP_ADD:
        skip16 = torch.rand(ref.shape, generator=g).to(dt)
        ref = ref + skip16.float().to(dev)
    if flags & capi.EPI_RELU6:
        ref = ref.clamp(0, 6)
    if flags & capi.EPI_SHIFT_STORE:
        ref = shifted(ref)
    x_nhwc = x16.permute(0, 2, 3, 1).contiguous().to(dev)
    out = torch.full(ref.permute(0, 2, 3, 1).shape, float("nan"), dtype=dt, device=dev)
    skip_nhwc = skip16.permute(0, 2, 3, 1).contiguous().to(dev) if skip16 is not None else None
    d = capi.BsvdConvDesc(T, H, W, cin, cout, flags, prec, variant)
    wc = w.contiguous()
    bc = b.contiguous()
    rc = lib.bsvd_conv_stage(d, x_nhwc.data_ptr(), wc.data_ptr(), bc.data_ptr(),
                             skip_nhwc.data_ptr() if skip_nhwc is not None else None,
                             out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    if rc:
        return {"error": lib.bsvd_last_error().decode()}
    torch.cuda.synchronize()
    got = out.float().permute(0, 3, 1, 2)
    diff = (got - ref).abs()
    nan = int(torch.isnan(got).sum())
    diff = torch.nan_to_num(diff, nan=1e9)
    return {"max_abs": float(diff.max()), "mean_abs": float(diff.mean()), "nan": nan,
            "ref_absmax": float(ref.abs().max())}


CASES = None


def all_cases():
    R, P, K, S, D = (capi.EPI_RELU6, capi.EPI_PIXSHUF, capi.EPI_SKIP_ADD, capi.EPI_SHIFT_STORE,
                     capi.EPI_STRIDE2)
    base = [
        ("64->64 relu6", 3, 12, 200, 64, 64, R),
        ("64->64 relu6 big", 2, 36, 960, 64, 64, R),
        ("64->128 s2 relu6 shift", 3, 12, 200, 64, 128, R | D | S),
        ("128->128 relu6 shift", 3, 10, 136, 128, 128, R | S),
        ("128->128 relu6", 3, 10, 136, 128, 128, R),
        ("128->256 s2 relu6 shift", 3, 12, 264, 128, 256, R | D | S),
        ("256->256 relu6 shift", 3, 7, 132, 256, 256, R | S),
        ("256->512 ps skip shift", 3, 7, 132, 256, 512, P | K | S),
        ("128->256 ps skip", 3, 10, 136, 128, 256, P | K),
    ]
    variants = [int(v) for v in os.environ.get("PROBE_VARIANTS", "0,1").split(",")]
    out = []
    for variant in variants:
        for c in base:
            out.append((capi.PREC_FP16, variant) + c)
    for c in base:
        out.append((capi.PREC_BF16, variants[0]) + c)
    return out


def child(start):
    """Run cases[start:] in this process; stop at the first error (a device fault poisons the
    CUDA context) and report how far we got."""
    lib = capi.load_library()
    cases = all_cases()
    for i in range(start, len(cases)):
        prec, variant, name, T, H, W, cin, cout, flags = cases[i]
        try:
            r = run_case(lib, T, H, W, cin, cout, flags, prec, variant)
        except Exception as e:  # noqa: BLE001
            r = {"error": repr(e)}
        r.update(case=name, prec=prec, variant=variant, index=i)
        print("RESULT " + json.dumps(r), flush=True)
        if "error" in r:
            return


def main():
    import subprocess
    if len(sys.argv) > 1:
        child(int(sys.argv[1]))
        return
    n = len(all_cases())
    results, start = [], 0
    while start < n:
        pr = subprocess.run([sys.executable, os.path.abspath(__file__), str(start)],
                            capture_output=True, text=True, timeout=600)
        got = [json.loads(l[7:]) for l in pr.stdout.splitlines() if l.startswith("RESULT ")]
        for r in got:
            print(json.dumps(r), flush=True)
        results += got
        if not got:
            print("child produced nothing:", pr.stderr[-2000:], flush=True)
            results.append({"index": start, "error": "child crashed: " + pr.stderr[-500:]})
            start += 1
        else:
            start = got[-1]["index"] + 1
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/probe_conv.json", "w") as f:
        json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
