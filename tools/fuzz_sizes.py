"""Randomised shapes through every mode of the path against the fp32 CPU oracle (and stream == clip):
    python tools/fuzz_sizes.py [cases] [seed]
Modes: BSVD-64 fp16 / bf16 / fp32x3, blind c32 (native pair layout).  Prints one JSON line; exit code 1 on a miss."""
import json, os, random, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bsvd_b200.arch import BSVD
from oracle import bsvd_oracle as O

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 24
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
TOL = {"fp16": 1e-3, "bf16": 1e-2, "fp32x3": 1e-4, "c32": 1e-3}
sd64 = O.make_synthetic_params(0, 0.5)
c = O.C32
sd32 = O.make_synthetic_params(1, 0.5, in_ch=3, chns=c["chns"], mid_ch=c["mid_ch"], interm_ch=c["interm_ch"])
nets = {}
for prec in ("fp16", "bf16", "fp32x3"):
    n = BSVD(chns=[64, 128, 256], mid_ch=64, shift_input=False, norm='none', interm_ch=64, act='relu6', pretrain_ckpt=None, precision=prec)
    n.load_tsn_state(sd64); nets[prec] = n.cuda().eval()
n = BSVD(chns=list(c["chns"]), mid_ch=c["mid_ch"], shift_input=False, norm='none', interm_ch=c["interm_ch"], act=c["act"], blind=True, pretrain_ckpt=None)
n.load_tsn_state(sd32); nets["c32"] = n.cuda().eval()
l64, l32 = O.layers_from_tsn_state(sd64), O.layers_from_tsn_state(sd32)
worst, fails = {}, []
special = [(1, 4, 4), (2, 4, 132), (3, 132, 4), (1, 8, 260), (5, 260, 8), (2, 12, 516), (1, 256, 256), (4, 128, 128), (2, 124, 252)]
for i in range(n_cases):
    T, H, W = special[i] if i < len(special) else (rng.randint(1, 9), 4 * rng.randint(1, 60), 4 * rng.randint(1, 80))
    mode = ("fp16", "bf16", "fp32x3", "c32")[i % 4]
    x, _ = O.make_synthetic_clip(T, H, W, seed=1000 + i)
    net = nets[mode]
    with torch.no_grad():
        if mode == "c32":
            xin = x[:, :3].contiguous()
            ref = O.forward_clip(l32, xin, act=c["act"])
        else:
            xin = x
            ref = O.forward_clip(l64, x)
        y = net(xin[None].cuda())[0]
        # streaming schedule on the same frames
        net.reset()
        outs = []
        for t in range(T):
            o = net.feedin_one_element(xin[t:t + 1].cuda())
            if o is not None: outs.append(o)
        while len(outs) < T:
            o = net.feedin_one_element(None)
            if o is not None: outs.append(o)
        net.reset()
        same = bool(torch.equal(torch.cat(outs), y))
    err = float((y.float().cpu() - ref).abs().max())
    worst[mode] = max(worst.get(mode, 0.0), err)
    if err > TOL[mode] or not same or net.overflowed():
        fails.append({"case": i, "mode": mode, "shape": [T, H, W], "err": err, "stream_equals_clip": same})
print(json.dumps({"cases": n_cases, "worst_max_abs": worst, "failures": fails}))
sys.exit(1 if fails else 0)
