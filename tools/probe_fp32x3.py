"""fp32-grade mode at the benchmarked size: frames/s and max-abs from the fp32 oracle (2 frames)."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bsvd_b200.arch import BSVD
from oracle import bsvd_oracle as O
dev = torch.device("cuda", 0)
sd = O.make_synthetic_params(0, 0.5)
net = BSVD(chns=[64, 128, 256], mid_ch=64, shift_input=False, norm='none', interm_ch=64, act='relu6', pretrain_ckpt=None, precision="fp32x3")
net.load_tsn_state(sd)
net = net.to(dev).eval()
x, _ = O.make_synthetic_clip(10, 540, 960, seed=1)
xd = x.to(dev)
with torch.no_grad():
    for _ in range(3): y = net(xd[None])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): y = net(xd[None])
    e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
ref = O.forward_clip(O.layers_from_tsn_state(sd), x[:2])
with torch.no_grad():
    y2 = net(xd[None, :2])[0].float().cpu()
print(json.dumps({"precision": "fp32x3", "ms_per_clip": ms, "fps": 1e4 / ms, "max_abs_vs_fp32_oracle_2_frames": float((y2 - ref).abs().max())}))
