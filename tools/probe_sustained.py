"""ms/step of the clip forward over N back-to-back steps (sustained, power-capped regime) and right after an
idle pause (burst regime).  Environment switches of the library (BSVD_B200_*) are read at import, so A/B
runs are separate processes:  BSVD_B200_NTILE_MAX=128 python tools/probe_sustained.py"""
import os, sys, time, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bsvd_b200.arch import BSVD
from oracle import bsvd_oracle as O
n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
dev = torch.device("cuda", 0)
net = BSVD(chns=[64, 128, 256], mid_ch=64, shift_input=False, norm='none', interm_ch=64, act='relu6', pretrain_ckpt=None)
net.load_tsn_state(O.make_synthetic_params(0, 0.5))
net = net.to(dev).eval()
x, _ = O.make_synthetic_clip(10, 540, 960, seed=1)
xd = x.to(dev)
def timed(k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    with torch.no_grad():
        for _ in range(k): net(xd[None])
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k
with torch.no_grad():
    for _ in range(10): net(xd[None])
torch.cuda.synchronize()
sus = timed(n)
time.sleep(1.0)
burst = timed(10)
tag = {k: v for k, v in os.environ.items() if k.startswith("BSVD_B200_")}
print(json.dumps({"env": tag, "sustained_ms": round(sus, 4), "sustained_fps": round(1e4 / sus, 1), "burst_ms": round(burst, 4), "burst_fps": round(1e4 / burst, 1)}))
