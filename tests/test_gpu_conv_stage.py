"""Every fused conv-stage kernel variant (through the C ABI hook bsvd_conv_stage) against a plain
fp32 torch restatement of the same stage on identically rounded 16-bit operands.
Tolerance: the result is stored in 16 bits, so half an ulp of the storage type at the output
magnitude (|y| < 4: fp16 1e-3, bf16 8e-3) plus fp32 summation-order noise."""
import pytest
import torch
import torch.nn.functional as F

from bsvd_b200 import capi

pytestmark = pytest.mark.gpu

R, P, K, S, D = (capi.EPI_RELU6, capi.EPI_PIXSHUF, capi.EPI_SKIP_ADD, capi.EPI_SHIFT_STORE,
                 capi.EPI_STRIDE2)

CASES = [
    ("64to64_relu6", 3, 12, 200, 64, 64, R),
    ("64to64_fullwidth", 2, 36, 960, 64, 64, R),
    ("64to64_odd_rows_T1", 1, 7, 132, 64, 64, R),
    ("64to128_s2_shift", 3, 12, 200, 64, 128, R | D | S),
    ("128to128_shift", 3, 10, 136, 128, 128, R | S),
    ("128to128_shift_T1", 1, 10, 136, 128, 128, R | S),
    ("128to128_shift_T2", 2, 6, 40, 128, 128, R | S),
    ("128to128_plain", 3, 10, 136, 128, 128, R),
    ("128to256_s2_shift", 3, 12, 264, 128, 256, R | D | S),
    ("256to256_shift", 3, 7, 132, 256, 256, R | S),
    ("256to256_noact", 2, 5, 60, 256, 256, 0),
    ("256to512_ps_skip_shift", 3, 7, 132, 256, 512, P | K | S),
    ("128to256_ps_skip", 3, 10, 136, 128, 256, P | K),            # skip on the tensor core + TMA stores
    ("128to256_ps_skip_wide", 2, 6, 480, 128, 256, P | K),
    ("64to128_ps_skip_epilogue", 2, 6, 72, 64, 128, P | K),       # skip added in the epilogue (N tile 128)
    ("128to256_ps_only", 2, 6, 72, 128, 256, P),
    ("tiny_4x4", 2, 4, 4, 64, 64, R),
    # stride 2 through the sub-plane boxes (conv_tc.cuh PIPE 4): ragged widths / heights, one-tile images
    ("64to128_s2_plain", 2, 10, 260, 64, 128, R | D),
    ("64to128_s2_shift_wide", 2, 6, 516, 64, 128, R | D | S),
    ("64to128_s2_shift_tiny", 3, 4, 4, 64, 128, R | D | S),
    ("128to256_s2_shift_odd_rows", 2, 14, 100, 128, 256, R | D | S),
    ("128to256_s2_plain_tiny", 1, 2, 2, 128, 256, R | D),
]


def shifted(y):
    f = y.shape[1] // 8
    o = y.clone()
    o[:, :2 * f] = 0
    o[:-1, :f] = y[1:, :f]
    o[1:, f:2 * f] = y[:-1, f:2 * f]
    return o


@pytest.mark.parametrize("prec", [capi.PREC_FP16, capi.PREC_BF16], ids=["fp16", "bf16"])
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_conv_stage(case, prec):
    _, T, H, W, cin, cout, flags = case
    lib = capi.load_library()
    g = torch.Generator().manual_seed(1234)
    dt = torch.float16 if prec == capi.PREC_FP16 else torch.bfloat16
    x = torch.rand(T, cin, H, W, generator=g).to(dt)
    w = torch.randn(cout, cin, 3, 3, generator=g) * (0.7 * (2.0 / (9 * cin)) ** 0.5)
    b = torch.randn(cout, generator=g) * 0.1
    dev = torch.device("cuda")
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        ref = F.conv2d(x.float().to(dev), w.to(dt).float().to(dev), b.to(dev),
                       stride=2 if flags & D else 1, padding=1)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    if flags & P:
        ref = F.pixel_shuffle(ref, 2)
    skip = None
    if flags & K:
        skip = torch.rand(ref.shape, generator=g).to(dt)
        ref = ref + skip.float().to(dev)
    if flags & R:
        ref = ref.clamp(0, 6)
    if flags & S:
        ref = shifted(ref)
    x_nhwc = x.permute(0, 2, 3, 1).contiguous().to(dev)
    out = torch.full(tuple(ref.permute(0, 2, 3, 1).shape), float("nan"), dtype=dt, device=dev)
    skip_nhwc = skip.permute(0, 2, 3, 1).contiguous().to(dev) if skip is not None else None
    d = capi.BsvdConvDesc(T, H, W, cin, cout, flags, prec, 0)
    capi.check(lib.bsvd_conv_stage(d, x_nhwc.data_ptr(), w.contiguous().data_ptr(),
                                   b.contiguous().data_ptr(),
                                   skip_nhwc.data_ptr() if skip_nhwc is not None else None,
                                   out.data_ptr(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    got = out.float().permute(0, 3, 1, 2)
    assert not torch.isnan(got).any(), "some output elements were never written"
    err = float((got - ref).abs().max())
    tol = (1.2e-3 if prec == capi.PREC_FP16 else 9e-3) * max(1.0, float(ref.abs().max()) / 4.0)
    assert err <= tol, (err, tol)


def test_conv_stage_rejects_bad_shapes():
    lib = capi.load_library()
    d = capi.BsvdConvDesc(1, 8, 8, 48, 64, 0, 0, 0)
    x = torch.zeros(8, device="cuda")
    assert lib.bsvd_conv_stage(d, x.data_ptr(), x.data_ptr(), None, None, x.data_ptr(), None) != 0
    assert b"cin" in lib.bsvd_last_error()
