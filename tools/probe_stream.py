"""Steady-state cost of one streaming push (BASELINE configs[2]): wall time per push from CUDA events.
Run under `ncu --metrics gpu__time_duration.sum` to compare with the sum of the 32 kernel durations."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bsvd_b200.arch import BSVD
from oracle import bsvd_oracle as O
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 60
dev = torch.device("cuda", 0)
net = BSVD(chns=[64, 128, 256], mid_ch=64, shift_input=False, norm='none', interm_ch=64, act='relu6', pretrain_ckpt=None, precision=prec)
net.load_tsn_state(O.make_synthetic_params(0, 0.5))
net = net.to(dev).eval()
x, _ = O.make_synthetic_clip(4, 540, 960, seed=1)
pool = [x[i:i + 1].to(dev) for i in range(4)]
with torch.no_grad():
    for i in range(20):
        net.feedin_one_element(pool[i % 4])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        net.feedin_one_element(pool[i % 4])
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(json.dumps({"precision": prec, "pushes": n, "ms_per_push": ms, "fps": 1e3 / ms}))
