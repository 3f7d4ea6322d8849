"""GPU probe: whole-network parity against golden fixtures and the CPU oracle."""
import glob, os, sys, time, json
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bsvd_b200.arch import BSVD
from oracle import bsvd_oracle as O

def build(seed, scale, prec):
    m = BSVD(chns=[64,128,256], mid_ch=64, shift_input=False, norm='none', interm_ch=64, act='relu6', pretrain_ckpt=None, precision=prec)
    sd = O.make_synthetic_params(seed, scale)
    m.load_tsn_state(sd)
    return m.cuda().eval(), O.layers_from_tsn_state(sd)

for path in sorted(glob.glob('tests/golden/*.npz')):
    g = np.load(path)
    for prec in ('fp16','bf16'):
        m, layers = build(int(g['param_seed']), float(g['weight_scale']), prec)
        x,_ = O.make_synthetic_clip(int(g['T']), int(g['H']), int(g['W']), int(g['clip_seed']))
        with torch.no_grad():
            y = m(x[None].cuda())[0].float().cpu()
        ref = torch.from_numpy(g['y_stream'])
        d = (y-ref).abs()
        print(os.path.basename(path), prec, 'max', float(d.max()), 'mean', float(d.mean()), 'launches', m.last_launch_count, flush=True)

# medium size vs oracle
m, layers = build(0, 0.5, 'fp16')
x,_ = O.make_synthetic_clip(3, 136, 264, 5)
t0=time.time(); ref = O.forward_clip(layers, x); print('oracle s', time.time()-t0)
with torch.no_grad():
    y = m(x[None].cuda())[0].float().cpu()
d=(y-ref).abs(); print('136x264 fp16 max', float(d.max()), 'mean', float(d.mean()))
# timing at 540x960
x,_ = O.make_synthetic_clip(10, 540, 960, 1)
xc = x[None].cuda()
for prec in ('fp16',):
    with torch.no_grad():
        for _ in range(3): y = m(xc)
        torch.cuda.synchronize()
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): y = m(xc)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)/5
        print('540x960 T=10: %.2f ms/clip, %.1f fps' % (ms, 10/(ms/1e3)))
