"""One clip forward at the benchmarked size for ncu: `python tools/profile_clip.py [T] [warm] [reps] [prec]`
(32 launches per forward; skip the warm-up with `ncu -s 32*warm`)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bsvd_b200.arch import BSVD
from oracle import bsvd_oracle as O
T = int(sys.argv[1]) if len(sys.argv) > 1 else 10
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 1
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
prec = sys.argv[4] if len(sys.argv) > 4 else "fp16"
dev = torch.device("cuda", 0)
net = BSVD(chns=[64, 128, 256], mid_ch=64, shift_input=False, norm='none', interm_ch=64, act='relu6', pretrain_ckpt=None, precision=prec)
net.load_tsn_state(O.make_synthetic_params(0, 0.5))
net = net.to(dev).eval()
x, _ = O.make_synthetic_clip(T, 540, 960, seed=1)
xd = x.to(dev)
with torch.no_grad():
    for _ in range(warm + reps):
        y = net(xd[None])
torch.cuda.synchronize()
print("done", float(y.abs().mean()))
