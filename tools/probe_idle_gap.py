"""Does an idle gap before the timed region change ms/step?  (round 2: unprofiled 10.01 ms vs profiled 9.63 ms)"""
import os, sys, time, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bsvd_b200 import capi
from bsvd_b200.arch import BSVD
from oracle import bsvd_oracle as O
dev = torch.device("cuda", 0)
net = BSVD(chns=[64, 128, 256], mid_ch=64, shift_input=False, norm='none', interm_ch=64, act='relu6', pretrain_ckpt=None)
net.load_tsn_state(O.make_synthetic_params(0, 0.5))
net = net.to(dev).eval()
x, _ = O.make_synthetic_clip(10, 540, 960, seed=1)
xd = x.to(dev)
lib = capi.load_library()
def timed(k, gap, prof=0):
    capi.check(lib.bsvd_set_profiling(net._handle, prof))
    torch.cuda.synchronize()
    if gap: time.sleep(gap)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    with torch.no_grad():
        for _ in range(k): net(xd[None])
    e1.record(); torch.cuda.synchronize()
    capi.check(lib.bsvd_set_profiling(net._handle, 0))
    return e0.elapsed_time(e1) / k
with torch.no_grad():
    for _ in range(5): net(xd[None])
res = {}
for name, k, gap, prof in (("k10_nogap", 10, 0, 0), ("k10_gap250", 10, 0.25, 0), ("k10_nogap_b", 10, 0, 0), ("k10_prof", 10, 0, 1),
                           ("k10_gap250_prof", 10, 0.25, 1), ("k50_nogap", 50, 0, 0), ("k50_prof", 50, 0, 1), ("k10_gap1s", 10, 1.0, 0),
                           ("k10_nogap_c", 10, 0, 0)):
    res[name] = round(timed(k, gap, prof), 4)
print(json.dumps(res))
