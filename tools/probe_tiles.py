import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bsvd_b200 import tiling
from bsvd_b200.arch import BSVD
from oracle import bsvd_oracle as O
net = BSVD(chns=[64,128,256], mid_ch=64, shift_input=False, norm='none', interm_ch=64, act='relu6', pretrain_ckpt=None)
net.load_tsn_state(O.make_synthetic_params(0, 0.5)); net = net.cuda().eval()
x, _ = O.make_synthetic_clip(3, 184, 408, seed=12); xc = x.cuda()
with torch.no_grad():
    full = net(xc[None])[0]
    fwd = lambda t: net(t[None])[0]
    for halo in (76, 96, 128):
        for rc in ((1, 2), (2, 1), (2,2)):
            t = tiling.forward_tiled_local(fwd, xc, rc[0], rc[1], halo=halo)
            d = (t - full).abs()
            cols = torch.nonzero(d.amax(dim=(0, 1, 2)) > 0).flatten()
            rows = torch.nonzero(d.amax(dim=(0, 1, 3)) > 0).flatten()
            print(halo, rc, float(d.max()), int((d > 0).sum()), cols[:6].tolist(), cols[-6:].tolist(), rows[:4].tolist(), rows[-4:].tolist())
    # determinism of a plain crop forward
    a = net(xc[None, :, :, :, :264].contiguous())[0]
    b = net(xc[None, :, :, :, :264].contiguous())[0]
    print('repeat equal', torch.equal(a, b))
