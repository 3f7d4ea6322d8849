"""Per-stage device times of the blind c32 configuration at 540x960 (native pair layout vs zero-padded)."""
import os, sys, json, ctypes as C
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bsvd_b200 import capi
from bsvd_b200.arch import BSVD
from oracle import bsvd_oracle as O
c = O.C32
dev = torch.device("cuda", 0)
sd = O.make_synthetic_params(0, 0.5, in_ch=3, chns=c["chns"], mid_ch=c["mid_ch"], interm_ch=c["interm_ch"])
net = BSVD(chns=list(c["chns"]), mid_ch=c["mid_ch"], shift_input=False, norm='none', interm_ch=c["interm_ch"], act=c["act"], blind=True, pretrain_ckpt=None)
net.load_tsn_state(sd); net = net.to(dev).eval()
x, _ = O.make_synthetic_clip(10, 540, 960, seed=1)
x3 = x[:, :3].contiguous().to(dev)
lib = capi.load_library()
with torch.no_grad():
    for _ in range(3): net(x3[None])
    capi.check(lib.bsvd_set_profiling(net._handle, 1))
    for _ in range(5): net(x3[None])
torch.cuda.synchronize()
ms = (C.c_float * capi.NUM_STAGES)(); n = C.c_int(0)
capi.check(lib.bsvd_get_stage_ms(net._handle, ms, capi.NUM_STAGES, C.byref(n)))
print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("BSVD_B200_")}, "total_ms": round(sum(ms[1:]) / n.value, 3),
                  "stage_ms": [round(ms[i] / n.value, 4) for i in range(1, 33)]}))
