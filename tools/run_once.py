"""Run N forwards of a [1,T,4,H,W] clip (for ncu captures)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bsvd_b200.arch import BSVD
from oracle import bsvd_oracle as O
T = int(os.environ.get("T", "3")); H = int(os.environ.get("H", "540")); W = int(os.environ.get("W", "960"))
N = int(os.environ.get("N", "2"))
net = BSVD(chns=[64,128,256], mid_ch=64, shift_input=False, norm='none', interm_ch=64, act='relu6',
           pretrain_ckpt=None, precision=os.environ.get("PREC", "fp16"))
net.load_tsn_state(O.make_synthetic_params(0, 0.5))
net = net.cuda().eval()
x, _ = O.make_synthetic_clip(T, H, W, 1)
xc = x[None].cuda()
with torch.no_grad():
    for _ in range(N):
        y = net(xc)
torch.cuda.synchronize()
print("done", float(y.abs().max()))
