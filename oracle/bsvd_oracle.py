"""CPU oracle for the BSVD-64 bidirectional-buffer forward.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this file; the product (bsvd_b200/) never does and has no CPU path.

It is a plain fp32 torch-CPU restatement of the reference's algorithm, written from the
reference's behaviour, each function citing the file:line it follows (paths relative to the
reference checkout, ChenyangQiQi/BSVD @ 29a6f05):

  * clip order   — Experimental_root/archs/tsm_arch.py:59-72 (TSN.forward),
                   archs_2d/wnet_models.py:126-183 (DenBlock), :233-278 (WNet),
                   temporal_shift_ops/temporal_shift.py:53-80 (batch_shift, eval mode)
  * stream order — Experimental_root/archs/bsvd_arch.py:53-114 (BiBufferConv), :308-322 (MemSkip),
                   :374-414 (DenBlock.forward), :485-552 (feedin_one_element / streaming_forward)
  * checkpoint   — bsvd_arch.py:462-474 (BSVD.load re-keying of the TSN layout)

The arithmetic itself (conv2d) lives in PyTorch (requirements.txt:14 pins torch>=1.7; this image
has 2.11.0), reached from every nn.Conv2d call site in bsvd_arch.py:31-38, 208-213, 238-239, 265,
295-298.  The reference ships no tests, golden vectors or checkpoint for this path, so the oracle
is pinned against the reference ITSELF: tests/golden/make_golden.py imports the unmodified
reference from /root/reference in the build container, runs both its TSN (clip) and BSVD
(stream) classes on seeded synthetic weights/inputs and commits the outputs as fixtures under
tests/golden/; tests/test_oracle.py checks this file against them.  Parity status: PINNED by
reference-generated fixtures (not by reference-owned test vectors — there are none).
"""
from __future__ import annotations

import hashlib

import numpy as np
import torch
import torch.nn.functional as F

CHNS = (64, 128, 256)
MID_CH = 64
INTERM_CH = 64
IN_CH = 4
OUT_CH = 3
FOLD_DIV = 8  # bsvd_arch.py:43 / temporal_shift.py:53 (shift_div=8)

# (name inside a DenBlock in TSN/WNet layout, stride); execution order of DenBlock.forward
_TSN_LAYER_KEYS = (
    "inc.convblock.0", "inc.convblock.3",
    "downc0.convblock.0", "downc0.convblock.3.c1.net", "downc0.convblock.3.c2.net",
    "downc1.convblock.0", "downc1.convblock.3.c1.net", "downc1.convblock.3.c2.net",
    "upc2.convblock.0.c1.net", "upc2.convblock.0.c2.net", "upc2.convblock.1",
    "upc1.convblock.0.c1.net", "upc1.convblock.0.c2.net", "upc1.convblock.1",
    "outc.convblock.0", "outc.convblock.3",
)
# The same 16 convs in the streaming class' own state-dict names (bsvd_arch.py:143-145, 252-255,
# 280-282: MemCvBlock.load maps 'net.' -> 'op.conv.', UpBlock.load maps convblock.1 -> convblock.0)
_BSVD_LAYER_KEYS = (
    "inc.convblock.0", "inc.convblock.3",
    "downc0.convblock.0", "downc0.memconv.c1.op.conv", "downc0.memconv.c2.op.conv",
    "downc1.convblock.0", "downc1.memconv.c1.op.conv", "downc1.memconv.c2.op.conv",
    "upc2.memconv.c1.op.conv", "upc2.memconv.c2.op.conv", "upc2.convblock.0",
    "upc1.memconv.c1.op.conv", "upc1.memconv.c2.op.conv", "upc1.convblock.0",
    "outc.convblock.0", "outc.convblock.3",
)
# layers whose input goes through the temporal shift (TemporalShift wraps CvBlock.c1/c2,
# tsm_arch.py:49-57; BiBufferConv in MemCvBlock, bsvd_arch.py:123-129)
SHIFT_LAYERS = (3, 4, 6, 7, 8, 9, 11, 12)


# the other configuration the reference trains (options/train/0402_*_blind_c32.yml:63-68: chns
# [32,64,128], mid_ch 32, blind; interm_ch and act keep their defaults 30 / 'relu',
# wnet_models.py:51,135,234)
C32 = dict(chns=(32, 64, 128), mid_ch=32, interm_ch=30, in_ch=3, act="relu")


def layer_shapes(block: int, in_ch: int = IN_CH, chns=CHNS, mid_ch: int = MID_CH,
                 interm_ch: int = INTERM_CH):
    """[(cout, cin, stride)] of the 16 convs of DenBlock `block` (0 = temp1, 1 = temp2).
    in_ch=3 is the blind variant (InputCvBlock(blind=True), bsvd_arch.py:204-205)."""
    c0, c1, c2 = chns
    cin = in_ch if block == 0 else mid_ch
    cout = mid_ch if block == 0 else OUT_CH
    INTERM_CH = interm_ch  # noqa: N806
    return [
        (INTERM_CH, cin, 1), (c0, INTERM_CH, 1),
        (c1, c0, 2), (c1, c1, 1), (c1, c1, 1),
        (c2, c1, 2), (c2, c2, 1), (c2, c2, 1),
        (c2, c2, 1), (c2, c2, 1), (c1 * 4, c2, 1),
        (c1, c1, 1), (c1, c1, 1), (c0 * 4, c1, 1),
        (c0, c0, 1), (cout, c0, 1),
    ]


def tsn_keys(prefix: str = "base_model."):
    """State-dict keys of the reference TSN(WNet_multistage) in execution order."""
    keys = []
    for blk in range(2):
        for name in _TSN_LAYER_KEYS:
            keys.append(f"{prefix}nets_list.{blk}.{name}")
    return keys


def bsvd_keys():
    keys = []
    for blk in range(2):
        for name in _BSVD_LAYER_KEYS:
            keys.append(f"temp{blk + 1}.{name}")
    return keys


def make_synthetic_params(seed: int = 0, weight_scale: float = 0.5, prefix: str = "base_model.",
                          in_ch: int = IN_CH, chns=CHNS, mid_ch: int = MID_CH,
                          interm_ch: int = INTERM_CH, act: str = "relu6"):
    """Seeded synthetic checkpoint in the TSN key layout (SURVEY §8d recipe).

    kaiming_normal_(nonlinearity='relu') statistics (wnet_models.py:155-162: std = sqrt(2/fan_in))
    scaled by `weight_scale` ("trained-like": keeps ReLU6 unsaturated); biases uniform in
    +-1/sqrt(fan_in) like nn.Conv2d's default.  numpy PCG64 so the stream is identical on the GPU
    box.  Returns {key.weight / key.bias: torch.float32 tensor}.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    sd = {}
    for blk in range(2):
        for name, (co, ci, _s) in zip(_TSN_LAYER_KEYS, layer_shapes(blk, in_ch, chns, mid_ch, interm_ch)):
            fan_in = ci * 9
            w = rng.standard_normal((co, ci, 3, 3), dtype=np.float32) * np.float32(
                weight_scale * np.sqrt(2.0 / fan_in))
            b = (rng.random(co, dtype=np.float32) * 2 - 1) * np.float32(1.0 / np.sqrt(fan_in))
            k = f"{prefix}nets_list.{blk}.{name}"
            sd[k + ".weight"] = torch.from_numpy(w)
            sd[k + ".bias"] = torch.from_numpy(b.astype(np.float32))
    return sd


def params_digest(sd) -> str:
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def layers_from_tsn_state(sd):
    """BSVD.load (bsvd_arch.py:462-474): TSN-layout state dict (optional 'module.' prefix) ->
    list of 32 (weight, bias) in execution order."""
    first = next(iter(sd))
    base = "module.base_model." if "module" in first else "base_model."
    out = []
    for k in tsn_keys(base):
        out.append((sd[k + ".weight"].float(), sd[k + ".bias"].float()))
    return out


def layers_from_bsvd_state(sd):
    """state_dict() of the streaming class (reference BSVD or our drop-in) -> 32 (weight, bias)."""
    return [(sd[k + ".weight"].float(), sd[k + ".bias"].float()) for k in bsvd_keys()]


def make_synthetic_clip(T: int, H: int, W: int, seed: int = 1, sigma: float = 20.0 / 255.0):
    """Seeded noisy clip [T,4,H,W]: smooth random field drifting in time + AWGN of
    std sigma (video_dali_dataset.py:231-235), 4th channel = constant sigma map
    (validation_seq_infer.py:18-21).  Returns (noisy_with_map, clean)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    gh, gw = max(H // 8, 2), max(W // 8, 2)
    base = rng.random((1, 3, gh, gw), dtype=np.float32)
    drift = rng.standard_normal((T, 3, gh, gw), dtype=np.float32) * 0.03
    low = torch.from_numpy(np.clip(base + np.cumsum(drift, axis=0), 0.0, 1.0))
    clean = F.interpolate(low, size=(H, W), mode="bilinear", align_corners=False).clamp(0, 1)
    noise = torch.from_numpy(rng.standard_normal((T, 3, H, W), dtype=np.float32)) * sigma
    noisy = clean + noise
    nmap = torch.full((T, 1, H, W), sigma, dtype=torch.float32)
    return torch.cat([noisy, nmap], dim=1).contiguous(), clean.contiguous()


# ------------------------------------------------------------------------------------------------
# optional operand rounding (used only to study/declare tolerances of the 16-bit tensor-core path)
# ------------------------------------------------------------------------------------------------
def _rnd(x, dt):
    return x if dt is None else x.to(dt).float()


def _conv(x, w, b, stride, op_dtype):
    return F.conv2d(_rnd(x, op_dtype), _rnd(w, op_dtype), b, stride=stride, padding=1)


def _relu6(x):
    return x.clamp(0.0, 6.0)  # nn.ReLU6, bsvd_arch.py:185-192 with act='relu6'


def _act_fn(act: str):
    """get_act_function (bsvd_arch.py:185-192): 'relu' -> nn.ReLU, 'relu6' -> nn.ReLU6."""
    if act == "relu6":
        return _relu6
    if act == "relu":
        return lambda t: t.clamp_min(0.0)
    raise ValueError(act)


def temporal_shift(x):
    """batch_shift with an empty past buffer (temporal_shift.py:53-80, batch_index=-1) ==
    ShiftConv's concat (bsvd_arch.py:42-50) over a whole clip [T,C,H,W]."""
    fold = x.shape[1] // FOLD_DIV
    out = torch.zeros_like(x)
    out[:-1, :fold] = x[1:, :fold]                    # channels [0,f)  <- frame t+1
    out[1:, fold:2 * fold] = x[:-1, fold:2 * fold]    # channels [f,2f) <- frame t-1
    out[:, 2 * fold:] = x[:, 2 * fold:]
    return out


def denblock_clip(layers, x, op_dtype=None, store_dtype=None, act: str = "relu6"):
    """DenBlock.forward over all frames at once (wnet_models.py:164-183 with TemporalShift in front
    of every CvBlock conv, tsm_arch.py:49-57).  layers: 16 (w,b).  x: [T,Cin,H,W]."""
    L = lambda i, t, stride=1: _conv(t, layers[i][0], layers[i][1], stride, op_dtype)  # noqa: E731
    st = lambda t: _rnd(t, store_dtype)  # noqa: E731
    _relu6 = _act_fn(act)  # noqa: F811  (the block's activation, 'relu6' for BSVD-64)
    x0 = st(_relu6(L(0, x)))
    x0 = st(_relu6(L(1, x0)))
    x1 = st(_relu6(L(2, x0, 2)))
    x1 = st(_relu6(L(3, temporal_shift(x1))))
    x1 = st(_relu6(L(4, temporal_shift(x1))))
    x2 = st(_relu6(L(5, x1, 2)))
    x2 = st(_relu6(L(6, temporal_shift(x2))))
    x2 = st(_relu6(L(7, temporal_shift(x2))))
    x2 = st(_relu6(L(8, temporal_shift(x2))))
    x2 = st(_relu6(L(9, temporal_shift(x2))))
    x2 = F.pixel_shuffle(L(10, x2), 2)
    y = st(x1 + x2)
    y = st(_relu6(L(11, temporal_shift(y))))
    y = st(_relu6(L(12, temporal_shift(y))))
    y = F.pixel_shuffle(L(13, y), 2)
    y = st(x0 + y)
    y = st(_relu6(L(14, y)))
    y = L(15, y)
    y[:, :3] = x[:, :3] - y[:, :3]    # wnet_models.py:181 / bsvd_arch.py:408-414
    return y


def forward_clip(layers, x, op_dtype=None, store_dtype=None, act: str = "relu6"):
    """BSVD.forward on ONE stream x: [T,4,H,W] -> [T,3,H,W] (clip order; same arithmetic as the
    streaming order).  op_dtype/store_dtype=None is the exact fp32 oracle."""
    with torch.no_grad():
        mid = _rnd(denblock_clip(layers[:16], x, op_dtype, store_dtype, act), store_dtype)
        return denblock_clip(layers[16:], mid, op_dtype, store_dtype, act)


# ------------------------------------------------------------------------------------------------
# streaming order (the reference's own schedule), kept separate so the latency / None protocol of
# feedin_one_element can be checked too
# ------------------------------------------------------------------------------------------------
class _BiBuffer:
    """BiBufferConv (bsvd_arch.py:53-114)."""

    def __init__(self, w, b):
        self.w, self.b = w, b
        self.left = None
        self.center = None

    def reset(self):
        self.left = None
        self.center = None

    def __call__(self, right):
        if right is not None:
            self.shape = right.shape
            self.fold = right.shape[1] // FOLD_DIV
        if self.center is None:                          # :86-100
            self.center = right
            if right is not None and self.left is None:
                n, _, h, w = self.shape
                self.left = torch.zeros((n, self.fold, h, w))
            return None
        if right is None:                                # :102-105
            n, _, h, w = self.shape
            r = torch.zeros((n, self.fold, h, w))
        else:
            r = right[:, :self.fold]
        f = self.fold
        out = F.conv2d(torch.cat([r, self.left, self.center[:, 2 * f:]], dim=1), self.w, self.b,
                       padding=1)                        # ShiftConv.forward :42-50
        self.left = self.center[:, f:2 * f]              # :112-113
        self.center = right
        return out


class _StreamDenBlock:
    """DenBlock.forward with None propagation and MemSkip FIFOs (bsvd_arch.py:308-322, 374-414)."""

    def __init__(self, layers, act: str = "relu6"):
        self.l = layers
        self.act = _act_fn(act)
        self.bib = {i: _BiBuffer(*layers[i]) for i in SHIFT_LAYERS}
        self.skip1, self.skip2, self.skip3 = [], [], []

    def reset(self):
        for b in self.bib.values():
            b.reset()

    def _mem(self, i, x):   # MemCvBlock :133-142
        x = self.bib[i](x)
        if x is not None:
            x = self.act(x)
        x = self.bib[i + 1](x)
        if x is not None:
            x = self.act(x)
        return x

    def __call__(self, in1):
        _relu6 = self.act  # noqa: F811  (the block's activation)
        conv = lambda i, t, s=1: F.conv2d(t, self.l[i][0], self.l[i][1], stride=s, padding=1)  # noqa: E731
        if in1 is not None:
            self.skip1.insert(0, in1[:, 0:3])
        x0 = None
        if in1 is not None:
            x0 = _relu6(conv(1, _relu6(conv(0, in1))))
            self.skip2.insert(0, x0)
        x1 = self._mem(3, _relu6(conv(2, x0, 2)) if x0 is not None else None)
        if x1 is not None:
            self.skip3.insert(0, x1)
        x2 = self._mem(6, _relu6(conv(5, x1, 2)) if x1 is not None else None)
        x2 = self._mem(8, x2)
        if x2 is not None:
            x2 = F.pixel_shuffle(conv(10, x2), 2)
        y = None
        if x2 is not None:
            y = x2 + self.skip3.pop()
        y = self._mem(11, y)
        if y is not None:
            y = F.pixel_shuffle(conv(13, y), 2)
            y = y + self.skip2.pop()
            y = conv(15, _relu6(conv(14, y)))
            s1 = self.skip1.pop()
            y[:, :3] = s1[:, :3] - y[:, :3]
        return y


class StreamOracle:
    """feedin_one_element / streaming_forward protocol (bsvd_arch.py:485-552)."""

    shift_num = 16   # count_shift(), bsvd_arch.py:554-560

    def __init__(self, layers, act: str = "relu6"):
        self.t1 = _StreamDenBlock(layers[:16], act)
        self.t2 = _StreamDenBlock(layers[16:], act)

    def reset(self):
        self.t1.reset()
        self.t2.reset()

    def feedin_one_element(self, x):
        with torch.no_grad():
            return self.t2(self.t1(x))

    def streaming_forward(self, seq):
        frames = [seq[i:i + 1] for i in range(seq.shape[0])]
        outs = [self.feedin_one_element(f) for f in frames]
        outs.append(self.feedin_one_element(None))
        while True:
            end = self.feedin_one_element(None)
            if len(outs) == self.shift_num + len(frames):
                break
            outs.append(end)
        self.reset()
        return torch.cat(outs[self.shift_num:], dim=0)


def psnr_float(img, ref, crop_border: int = 2) -> float:
    """calculate_psnr_float semantics (BasicSR/basicsr/metrics/psnr_ssim.py:130-168): CHW float in
    [0,1], crop, -10*log10(mse)."""
    a = img[..., crop_border:-crop_border, crop_border:-crop_border].double()
    b = ref[..., crop_border:-crop_border, crop_border:-crop_border].double()
    mse = float(((a - b) ** 2).mean())
    return float("inf") if mse == 0 else float(-10.0 * np.log10(mse))


def ssim(img, ref, crop_border: int = 0, data_range: float = 1.0) -> float:
    """calculate_ssim / _ssim (BasicSR/basicsr/metrics/psnr_ssim.py:49-128) on CHW arrays, restated without
    cv2: 11x11 Gaussian window (cv2.getGaussianKernel(11, 1.5) = normalised exp(-(i-5)^2 / (2*1.5^2))),
    'valid' correlation (the [5:-5, 5:-5] slice of cv2.filter2D), float64, mean over positions, then over
    channels.  The reference works on [0,255] images (data_range 255); [0,1] floats with data_range 1 give
    the same value."""
    import numpy as np
    a = np.asarray(img, dtype=np.float64)
    b = np.asarray(ref, dtype=np.float64)
    if crop_border:
        a = a[:, crop_border:-crop_border, crop_border:-crop_border]
        b = b[:, crop_border:-crop_border, crop_border:-crop_border]
    k = np.exp(-((np.arange(11) - 5.0) ** 2) / (2 * 1.5 * 1.5))
    k /= k.sum()
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2

    def filt(x):          # separable 'valid' correlation with outer(k, k)
        h = sum(k[i] * x[:, i:x.shape[1] - 10 + i] for i in range(11))
        return sum(k[i] * h[i:h.shape[0] - 10 + i, :] for i in range(11))

    vals = []
    for ch in range(a.shape[0]):
        x, y = a[ch], b[ch]
        mu1, mu2 = filt(x), filt(y)
        s11, s22, s12 = filt(x * x) - mu1 * mu1, filt(y * y) - mu2 * mu2, filt(x * y) - mu1 * mu2
        m = ((2 * mu1 * mu2 + c1) * (2 * s12 + c2)) / ((mu1 * mu1 + mu2 * mu2 + c1) * (s11 + s22 + c2))
        vals.append(m.mean())
    return float(np.mean(vals))
