"""A/B probe of one build (BSVD_B200_LIB selects the library): fixture parity of the fp16 path, then the
benchmarked clip timed three ways — bench.py's burst protocol (pause, 10 steps), 40 steps back to back, and
a profiled pass for the per-stage times of the first / last stages.  One JSON line."""
import glob, json, os, sys, time
import ctypes as C
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bsvd_b200 import capi
from bsvd_b200.arch import BSVD
from oracle import bsvd_oracle as O

dev = torch.device("cuda", 0)
lib = capi.load_library()


def build(seed, scale, prec="fp16"):
    m = BSVD(chns=[64, 128, 256], mid_ch=64, shift_input=False, norm='none', interm_ch=64, act='relu6',
             pretrain_ckpt=None, precision=prec)
    m.load_tsn_state(O.make_synthetic_params(seed, scale))
    return m.to(dev).eval()


res = {"lib": os.path.basename(os.environ.get("BSVD_B200_LIB", "libbsvd_b200.so")), "fixtures": {}}
for path in sorted(glob.glob(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "*.npz"))):
    g = np.load(path)
    if "chns" in g.files and list(g["chns"]) != [64, 128, 256]:
        continue
    if "y_stream" not in g.files or float(g["weight_scale"]) != 0.5:
        continue
    try:
        m = build(int(g["param_seed"]), float(g["weight_scale"]))
        x, _ = O.make_synthetic_clip(int(g["T"]), int(g["H"]), int(g["W"]), int(g["clip_seed"]))
        if x.shape[1] != 4:
            continue
        with torch.no_grad():
            y = m(x[None].to(dev))[0].float().cpu()
        res["fixtures"][os.path.basename(path)] = float((y - torch.from_numpy(g["y_stream"])).abs().max())
    except Exception as e:  # noqa: BLE001
        res["fixtures"][os.path.basename(path)] = "skipped: %s" % str(e)[:80]

net = build(0, 0.5)
x, _ = O.make_synthetic_clip(10, 540, 960, seed=1)
xd = x.to(dev)[None]


def timed(n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    with torch.no_grad():
        for _ in range(n):
            y = net(xd)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, y


with torch.no_grad():
    for _ in range(3):
        net(xd)
torch.cuda.synchronize()
burst = []
for _ in range(3):
    time.sleep(0.25)
    burst.append(timed(10)[0])
sus, y = timed(40)
res["burst_ms"] = [round(b, 4) for b in burst]
res["sustained_ms"] = round(sus, 4)
res["checksum"] = float(y.double().sum())
capi.check(lib.bsvd_set_profiling(net._handle, 1))
timed(5)
st = (C.c_float * capi.NUM_STAGES)()
passes = C.c_int(0)
capi.check(lib.bsvd_get_stage_ms(net._handle, st, capi.NUM_STAGES, C.byref(passes)))
capi.check(lib.bsvd_set_profiling(net._handle, 0))
res["stage_ms"] = [round(float(v), 4) for v in st]
print(json.dumps(res))
