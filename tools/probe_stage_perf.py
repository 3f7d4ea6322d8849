"""Time single conv stages at full size with debug variants (skip MMA / skip epilogue / A stages)."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bsvd_b200 import capi
lib = capi.load_library()
R, P, K, S, D = (capi.EPI_RELU6, capi.EPI_PIXSHUF, capi.EPI_SKIP_ADD, capi.EPI_SHIFT_STORE, capi.EPI_STRIDE2)
def run(name, T, H, W, cin, cout, flags, variant=0, a_stages=0, reps=4, w_stages=0):
    dev = torch.device("cuda")
    x = torch.rand(T, H, W, cin, device=dev).half()
    w = (torch.randn(cout, cin, 3, 3) * 0.05).contiguous(); b = torch.zeros(cout)
    s = 2 if flags & D else 1
    Ho, Wo, Co = H // s, W // s, cout
    if flags & P: Ho, Wo, Co = 2 * H, 2 * W, cout // 4
    out = torch.empty(T, Ho, Wo, Co, device=dev, dtype=torch.half)
    skip = torch.rand(T, Ho, Wo, Co, device=dev).half() if flags & K else None
    d = capi.BsvdConvDesc(T, H, W, cin, cout, flags, 0, (variant & 0xff) | (variant & (1 << 20)) | (a_stages << 8) | ((reps - 1) << 12) | (w_stages << 16))
    rc = lib.bsvd_conv_stage(d, x.data_ptr(), w.data_ptr(), b.data_ptr(), skip.data_ptr() if skip is not None else None, out.data_ptr(), None)
    if rc: print(name, "ERR", lib.bsvd_last_error().decode()); return
    ms = lib.bsvd_last_stage_ms()
    fl = 2.0 * 9 * cin * cout * T * (H // s) * (W // s)
    by = 2.0 * T * (H * W * cin + Ho * Wo * Co * (2 if flags & K else 1))
    print(f"{name:34s} var={variant} a_st={a_stages} w_st={w_stages}: {ms*1e3:8.1f} us  {fl/ms/1e9:7.1f} TF/s  {by/ms/1e6:7.1f} GB/s", flush=True)
T = 10
which = sys.argv[1] if len(sys.argv) > 1 else "up1"
NOPF = 1 << 21
if which == "64":
    for var in (0, 4, 20, 1 << 20, (1 << 20) | 4, (1 << 20) | 20):
        run("64->64 full", T, 540, 960, 64, 64, R, var)
elif which == "up1":
    # upc1.convblock.0: 128 -> 256 at half res, PixelShuffle + skip add to full res
    for var in (0, NOPF, 64, 128, 64 | 128, 4, 2, 16, 2 | 16):
        run("upc1.conv 128->256 PS+skip", T, 270, 480, 128, 256, P | K, var)
    for var in (0, 4):
        run("128->256 PS only", T, 270, 480, 128, 256, P, var)
        run("128->256 plain", T, 270, 480, 128, 256, 0, var)
elif which == "up2":
    for var in (0, NOPF, 64, 128, 4, 2):
        run("upc2.conv 256->512 PS+skip+shift", T, 135, 240, 256, 512, P | K | S, var)
elif which == "s2":
    for var in (0, 4, 2, 16):
        run("downc0.conv 64->128 s2 shift", T, 540, 960, 64, 128, R | D | S, var)
    for var in (0, 4, 2, 16):
        run("downc1.conv 128->256 s2 shift", T, 270, 480, 128, 256, R | D | S, var)
