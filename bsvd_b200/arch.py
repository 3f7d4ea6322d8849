"""Drop-in replacement of the reference's streaming denoiser class.

Mirrors Experimental_root/archs/bsvd_arch.py::BSVD (:441-560): same constructor kwargs (the
`network_g` block of options/test/bsvd_c64.yml:85-93 is splatted into it by
basicsr.archs.build_network), same parameter names in state_dict() (temp1.inc.convblock.0.weight,
temp1.downc0.memconv.c1.op.conv.weight, ...), same public methods (forward, streaming_forward,
feedin_one_element, reset, load, shift_num) and the same None protocol.  All arithmetic runs in
libbsvd_b200.so through the C ABI (bsvd_b200/capi.py); torch only owns parameters, device buffers
and the stream.  There is no CPU path: constructing the module for any configuration other than
BSVD-64, or calling it without a B200 and the built library, raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch
import torch.nn as nn

from . import capi

# execution order inside one DenBlock -> (state-dict stem, TSN-checkpoint stem)
# (re-keying of BSVD.load / DenBlock.load_from, bsvd_arch.py:143-145, 252-255, 280-282, 349-355)
_LAYERS = (
    ("inc.convblock.0", "inc.convblock.0"),
    ("inc.convblock.3", "inc.convblock.3"),
    ("downc0.convblock.0", "downc0.convblock.0"),
    ("downc0.memconv.c1.op.conv", "downc0.convblock.3.c1.net"),
    ("downc0.memconv.c2.op.conv", "downc0.convblock.3.c2.net"),
    ("downc1.convblock.0", "downc1.convblock.0"),
    ("downc1.memconv.c1.op.conv", "downc1.convblock.3.c1.net"),
    ("downc1.memconv.c2.op.conv", "downc1.convblock.3.c2.net"),
    ("upc2.memconv.c1.op.conv", "upc2.convblock.0.c1.net"),
    ("upc2.memconv.c2.op.conv", "upc2.convblock.0.c2.net"),
    ("upc2.convblock.0", "upc2.convblock.1"),
    ("upc1.memconv.c1.op.conv", "upc1.convblock.0.c1.net"),
    ("upc1.memconv.c2.op.conv", "upc1.convblock.0.c2.net"),
    ("upc1.convblock.0", "upc1.convblock.1"),
    ("outc.convblock.0", "outc.convblock.0"),
    ("outc.convblock.3", "outc.convblock.3"),
)


class _Node(nn.Module):
    """Pure container used to reproduce the reference's dotted parameter names."""

    def forward(self, *a, **k):  # pragma: no cover - never called
        raise RuntimeError("container module; the arithmetic lives in libbsvd_b200.so")


def _attach(root: nn.Module, dotted: str, leaf: nn.Module) -> None:
    parts = dotted.split(".")
    node = root
    for p in parts[:-1]:
        if p not in node._modules:
            node.add_module(p, _Node())
        node = node._modules[p]
    node.add_module(parts[-1], leaf)


def _layer_shapes(block, chns, mid_ch, in_ch, out_ch, interm_ch):
    """(out_ch, in_ch, stride) of the 16 convs of DenBlock `block`.

    blind=True: only temp1 drops the noise-map channel (in_ch 4 -> 3); temp2 always reads temp1's mid_ch
    channels.  This is the TSN/WNet training model's layout (wnet_models.py:233-278, the class the
    published blind checkpoints were trained with).  The reference's streaming class passes `blind` to
    BOTH DenBlocks (bsvd_arch.py:451-452), which makes its temp2.inc.convblock.0 a Conv2d(3, interm) and
    lets it run only when mid_ch == 3; for mid_ch == 3 the two layouts coincide, for any other mid_ch the
    reference class cannot execute, so state_dict() shapes follow the TSN layout here."""
    c0, c1, c2 = chns
    cin = in_ch if block == 0 else mid_ch
    cout = mid_ch if block == 0 else out_ch
    return [(interm_ch, cin, 1), (c0, interm_ch, 1), (c1, c0, 2), (c1, c1, 1), (c1, c1, 1),
            (c2, c1, 2), (c2, c2, 1), (c2, c2, 1), (c2, c2, 1), (c2, c2, 1), (c1 * 4, c2, 1),
            (c1, c1, 1), (c1, c1, 1), (c0 * 4, c1, 1), (c0, c0, 1), (cout, c0, 1)]


class BSVD(nn.Module):
    """Bidirectional-buffer streaming video denoiser, B200-native (see module docstring)."""

    # process-wide counters (the plugin prints them when it ran a reference script, so a harness can
    # tell that ARCH_REGISTRY['BSVD'] really resolved to this class and that its kernels ran)
    stats = {"instances": 0, "forward_calls": 0, "kernel_launches": 0}

    def __init__(self, chns=[32, 64, 128], mid_ch=3, shift_input=False, in_ch=4, out_ch=3,
                 norm='bn', act='relu', interm_ch=30, blind=False,
                 pretrain_ckpt='./experiments/pretrained_ckpt/bsvd-64.pth', precision=None):
        super().__init__()
        chns = list(chns)
        if not (chns in ([64, 128, 256], [32, 64, 128]) and 3 <= mid_ch <= 64 and 1 <= interm_ch <= 64
                and in_ch == 4 and out_ch == 3 and norm == 'none' and act in ('relu6', 'relu')
                and not shift_input):
            raise NotImplementedError(
                "bsvd_b200 implements chns=[64,128,256] (options/test/bsvd_c64.yml) and [32,64,128] "
                "(options/train/0402_*_c32.yml; zero-padded to the 64-channel kernels) with "
                "mid_ch<=64, interm_ch<=64, norm='none', act='relu6'|'relu', shift_input=False "
                "(blind=True is supported); there is no CPU/PyTorch fallback")
        if blind:
            in_ch = 3      # InputCvBlock(blind=True) drops the noise-map channel (bsvd_arch.py:204-205)
        self.blind = bool(blind)
        self.cfg = dict(chns=chns, mid_ch=mid_ch, in_ch=in_ch, out_ch=out_ch, interm_ch=interm_ch,
                        act=act)
        self.temp1 = _Node()
        self.temp2 = _Node()
        self._param_names = []
        for blk, root in enumerate((self.temp1, self.temp2)):
            shapes = _layer_shapes(blk, chns, mid_ch, in_ch, out_ch, interm_ch)
            for (stem, _tsn), (co, ci, st) in zip(_LAYERS, shapes):
                conv = nn.Conv2d(ci, co, kernel_size=3, stride=st, padding=1, bias=True)
                _attach(root, stem, conv)
                self._param_names.append(f"temp{blk + 1}.{stem}")
        self.shift_num = 16          # count_shift(): 8 BiBufferConv per DenBlock (bsvd_arch.py:554-560)
        self.independent_clips = False   # True: forward([N,F,..]) treats the N clips separately
        # 'fp16' | 'bf16' | 'fp32x3' (fp32-grade: three fp16 tensor-core products per contraction, ~1e-5 from
        # the fp32 reference at a third of the speed; for val.fp16 = False callers) | None = from dtype / autocast
        self.precision = precision or os.environ.get("BSVD_B200_PRECISION")
        self._handle = None
        self._handle_prec = None
        self._handle_dev = None
        self._weights_sig = None
        self._stream_out = None
        self.reset_params()
        BSVD.stats["instances"] += 1
        if pretrain_ckpt is not None:
            self.load(pretrain_ckpt)

    # ---------------------------------------------------------------- parameters / checkpoints
    @staticmethod
    def weight_init(m):
        if isinstance(m, nn.Conv2d):
            nn.init.kaiming_normal_(m.weight, nonlinearity='relu')   # bsvd_arch.py:476-479

    def reset_params(self):
        for m in self.modules():
            self.weight_init(m)

    def _convs(self):
        mods = dict(self.named_modules())
        return [mods[n] for n in self._param_names]

    def load(self, path):
        """BSVD.load (bsvd_arch.py:462-474): {'params': TSN-layout state dict}, optional 'module.'."""
        ckpt = torch.load(path, map_location="cpu")
        print("load from %s" % path)
        self.load_tsn_state(ckpt['params'])

    def load_tsn_state(self, ckpt_state):
        first = list(ckpt_state.keys())[0]
        base = 'module.base_model.' if 'module' in first else 'base_model.'
        convs = self._convs()
        with torch.no_grad():
            for blk in range(2):
                for i, (_stem, tsn) in enumerate(_LAYERS):
                    k = f"{base}nets_list.{blk}.{tsn}"
                    conv = convs[blk * 16 + i]
                    conv.weight.copy_(ckpt_state[k + ".weight"])
                    conv.bias.copy_(ckpt_state[k + ".bias"])
        self._weights_sig = None

    # ---------------------------------------------------------------- native handle
    def _select_precision(self):
        if self.precision is not None:
            p = str(self.precision).lower()
            if p not in ("fp16", "bf16", "fp32x3"):
                raise ValueError("precision must be 'fp16', 'bf16' or 'fp32x3'")
            return p
        w = self._first_weight()
        if w.dtype == torch.bfloat16:
            return "bf16"
        if torch.is_autocast_enabled() and torch.get_autocast_dtype('cuda') == torch.bfloat16:
            return "bf16"
        # fp32 weights, fp16 weights (profile.py:79) and fp16 autocast all use fp16 operands with
        # fp32 accumulation: 11-bit significands, the same operand precision as TF32.
        return "fp16"

    def _first_weight(self):
        return self.temp1.inc.convblock._modules['0'].weight

    def _out_dtype(self):
        w = self._first_weight()
        if torch.is_autocast_enabled():
            return torch.get_autocast_dtype('cuda')
        return w.dtype if w.dtype in (torch.float16, torch.bfloat16) else torch.float32

    def _ensure_handle(self, device):
        lib = capi.load_library()
        if not torch.cuda.is_available():
            raise capi.BsvdError("bsvd_b200 needs a CUDA device (B200, sm_100a); no CPU fallback")
        prec = self._select_precision()
        dev_index = device.index if device.index is not None else torch.cuda.current_device()
        # a native handle (packed weights, workspaces, tensor maps) is bound to ONE device: after
        # net.to('cuda:1'), or for an input on another GPU, it is rebuilt there
        if self._handle is not None and (self._handle_prec != prec or self._handle_dev != dev_index):
            self._destroy()
        if self._handle is None:
            cfg = capi.BsvdConfig()
            cfg.chns[0], cfg.chns[1], cfg.chns[2] = self.cfg["chns"]
            cfg.mid_ch, cfg.interm_ch = self.cfg["mid_ch"], self.cfg["interm_ch"]
            cfg.in_ch, cfg.out_ch = self.cfg["in_ch"], self.cfg["out_ch"]
            cfg.act_relu6, cfg.norm_none = (1 if self.cfg["act"] == "relu6" else 0), 1
            cfg.precision = {"fp16": capi.PREC_FP16, "bf16": capi.PREC_BF16, "fp32x3": capi.PREC_FP32X3}[prec]
            cfg.device = dev_index
            h = C.c_void_p()
            capi.check(lib.bsvd_create(C.byref(cfg), C.byref(h)))
            self._handle, self._handle_prec, self._handle_dev, self._weights_sig = h, prec, dev_index, None
        sig = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if sig != self._weights_sig:
            for i, conv in enumerate(self._convs()):
                w = conv.weight.detach().float().cpu().contiguous()
                b = conv.bias.detach().float().cpu().contiguous()
                capi.check(lib.bsvd_set_weights(self._handle, i, w.data_ptr(), b.data_ptr(),
                                                w.shape[0], w.shape[1]))
            self._weights_sig = sig
        return lib

    def _module_device(self):
        """Device of the parameters when they are on a GPU, else the current CUDA device."""
        w = self._first_weight()
        return w.device if w.is_cuda else torch.device("cuda", torch.cuda.current_device())

    def _destroy(self):
        if getattr(self, "_handle", None) is not None:
            try:
                with torch.cuda.device(self._handle_dev):
                    capi.load_library().bsvd_destroy(self._handle)
            except Exception:  # noqa: BLE001
                pass
            self.__dict__["_handle"] = None

    def __del__(self):
        try:
            self._destroy()
        except Exception:  # noqa: BLE001  (interpreter shutdown)
            pass

    def _apply(self, fn, *a, **k):
        self._weights_sig = None
        return super()._apply(fn, *a, **k)

    # ---------------------------------------------------------------- clip mode
    def _run_stream(self, x, noise_map=None, clip_len=None):
        """x: [T,C,H,W] on a CUDA device -> [T,3,H,W]: one continuous stream, or (clip_len given) T/clip_len
        independent clips of clip_len frames each in one pass (bsvd_forward_clips)."""
        if not x.is_cuda:
            x = x.cuda()          # the reference does x.cuda() per frame (bsvd_arch.py:520)
        dev = x.device
        with torch.cuda.device(dev):
            lib = self._ensure_handle(dev)
            xf = x.detach().float().contiguous()
            nm = None
            if noise_map is not None:
                nm = noise_map.detach().to(dev).float().contiguous()
            T, Cc, H, W = xf.shape
            out = torch.empty((T, 3, H, W), dtype=torch.float32, device=dev)
            if clip_len is None or clip_len == T:
                capi.check(lib.bsvd_forward_clip(
                    self._handle, xf.data_ptr(), nm.data_ptr() if nm is not None else None,
                    out.data_ptr(), T, Cc, H, W, torch.cuda.current_stream(dev).cuda_stream))
            else:
                if T % clip_len:
                    raise ValueError(f"{T} frames cannot be split into clips of {clip_len}")
                capi.check(lib.bsvd_forward_clips(
                    self._handle, xf.data_ptr(), nm.data_ptr() if nm is not None else None,
                    out.data_ptr(), T // clip_len, clip_len, Cc, H, W, torch.cuda.current_stream(dev).cuda_stream))
            BSVD.stats["forward_calls"] += 1
            BSVD.stats["kernel_launches"] += lib.bsvd_last_launch_count(self._handle)
        od = self._out_dtype()
        return out if od == torch.float32 else out.to(od)

    def forward(self, input, noise_map=None):
        """[N,F,C,H,W] (+ optional noise_map [N,F,1,H,W]) -> [N,F,3,H,W]  (bsvd_arch.py:490-499).
        Like the reference, N>1 clips form ONE continuous stream unless `independent_clips`."""
        N, Fr, Cc, H, W = input.shape
        nm = None if noise_map is None else noise_map.reshape(N * Fr, 1, H, W)
        # independent clips: all N clips in ONE pass of the 32 stages, folds confined to each clip
        out = self._run_stream(input.reshape(N * Fr, Cc, H, W), nm,
                               clip_len=Fr if (self.independent_clips and N > 1) else None)
        return out.reshape(N, Fr, 3, H, W)

    def streaming_forward(self, input_seq):
        """Pipeline-style inference over a whole sequence (bsvd_arch.py:501-552): list of
        [1,C,H,W] tensors or one [n,C,H,W] tensor -> [n,3,H,W].  State is reset at entry and exit."""
        if isinstance(input_seq, (list, tuple)):
            input_seq = torch.cat([t.cuda() for t in input_seq], dim=0)
        assert isinstance(input_seq, torch.Tensor), "convert the input into a sequence"
        self.reset()
        with torch.no_grad():
            return self._run_stream(input_seq)

    def denoise_sequence(self, noisy, sigma):
        """noisy [F,3,H,W] in [0,1] (any H, W), sigma = noise std in [0,1] (None for a blind model)
        -> denoised [F,3,H,W] clamped to [0,1]: temp_denoise + DenoisingModel.padding_input /
        crop_output (validation_seq_infer.py:10-31, denoising_model.py:133-168) in one C-ABI call;
        reflect padding, the constant noise map, the clamp and the crop happen inside the first and
        last kernels (bsvd_denoise_clip)."""
        dev = noisy.device if noisy.is_cuda else torch.device("cuda", torch.cuda.current_device())
        with torch.cuda.device(dev):
            lib = self._ensure_handle(dev)
            x = noisy.detach().to(dev).float().contiguous()
            Fr, Cc, H, W = x.shape
            assert Cc == 3, "denoise_sequence takes RGB frames"
            out = torch.empty((Fr, 3, H, W), dtype=torch.float32, device=dev)
            capi.check(lib.bsvd_denoise_clip(
                self._handle, x.data_ptr(), -1.0 if sigma is None else float(sigma), out.data_ptr(),
                Fr, H, W, torch.cuda.current_stream(dev).cuda_stream))
        return out

    def denoise_frames_u8(self, frames, sigma, bgr=False):
        """Decoded frames in, displayable frames out: uint8 [F,H,W,3] (HWC; bgr=True for cv2's channel
        order) -> uint8 [F,H,W,3].  /255 normalisation (img2tensor) happens in the first kernel's
        loads, clamp + *255 + round (tensor2img) in the last kernel's stores (bsvd_denoise_clip_u8)."""
        dev = frames.device if frames.is_cuda else torch.device("cuda", torch.cuda.current_device())
        with torch.cuda.device(dev):
            lib = self._ensure_handle(dev)
            x = frames.detach().to(dev).contiguous()
            assert x.dtype == torch.uint8 and x.dim() == 4 and x.shape[-1] == 3
            Fr, H, W, _ = x.shape
            out = torch.empty_like(x)
            capi.check(lib.bsvd_denoise_clip_u8(
                self._handle, x.data_ptr(), -1.0 if sigma is None else float(sigma), out.data_ptr(),
                Fr, H, W, 1 if bgr else 0, torch.cuda.current_stream(dev).cuda_stream))
        return out

    def denoise_host(self, input_host, noise_map_host=None, out_host=None):
        """End-to-end entry with HOST buffers (pinned recommended): H2D + forward + D2H inside the
        C ABI (bsvd_forward_clip_host).  input_host: fp32 [T,C,H,W] CPU tensor -> fp32 [T,3,H,W]."""
        assert not input_host.is_cuda and input_host.dtype == torch.float32
        dev = self._module_device()
        torch.cuda.set_device(dev)
        lib = self._ensure_handle(dev)
        x = input_host.contiguous()
        T, Cc, H, W = x.shape
        nm = noise_map_host.contiguous() if noise_map_host is not None else None
        if out_host is None:
            out_host = torch.empty((T, 3, H, W), dtype=torch.float32).pin_memory()
        capi.check(lib.bsvd_forward_clip_host(
            self._handle, x.data_ptr(), nm.data_ptr() if nm is not None else None,
            out_host.data_ptr(), T, Cc, H, W, torch.cuda.current_stream(dev).cuda_stream))
        return out_host

    def denoise_host_async(self, input_host, out_host, noise_map_host=None):
        """Pipelined end-to-end entry (bsvd_forward_clip_host_async): returns after enqueueing;
        copy-in of the next clip, compute of this one and copy-out of the previous overlap.  Call
        host_sync() before reading `out_host`.  Buffers must be pinned and stay alive."""
        assert not input_host.is_cuda and input_host.dtype == torch.float32
        assert input_host.is_contiguous() and out_host.is_contiguous()
        dev = self._module_device()
        torch.cuda.set_device(dev)
        lib = self._ensure_handle(dev)
        T, Cc, H, W = input_host.shape
        nm = noise_map_host
        capi.check(lib.bsvd_forward_clip_host_async(
            self._handle, input_host.data_ptr(), nm.data_ptr() if nm is not None else None,
            out_host.data_ptr(), T, Cc, H, W, torch.cuda.current_stream(dev).cuda_stream))
        self._last_host_shape = (T, 3, H, W)
        return out_host

    def last_device_output(self):
        """Device-side [T,3,H,W] fp32 result of the latest denoise_host_async call (a view of the C
        ABI's staging buffer, bsvd_host_last_output): valid for work ordered behind that call on the
        current stream, until the call after next overwrites it."""
        p = C.c_void_p()
        capi.check(capi.load_library().bsvd_host_last_output(self._handle, C.byref(p)))

        class _Holder:
            pass
        hold = _Holder()
        hold.__cuda_array_interface__ = {"shape": self._last_host_shape, "typestr": "<f4",
                                         "data": (int(p.value), False), "version": 2}
        return torch.as_tensor(hold, device=torch.device("cuda", self._handle_dev))

    def overflowed(self, reset=True):
        """True if, since the last reset, an fp16-stored activation that is not clamped by ReLU6 left the
        fp16 range (inf/NaN) where the reference's fp32 tensors would not (bsvd_overflow_flag; waits for
        the device).  Switch such a model to precision='bf16'."""
        if self._handle is None:
            return False
        flag = C.c_int(0)
        with torch.cuda.device(self._handle_dev):
            capi.check(capi.load_library().bsvd_overflow_flag(self._handle, C.byref(flag), 1 if reset else 0))
        return bool(flag.value)

    def host_sync(self):
        if self._handle is not None:
            with torch.cuda.device(self._handle_dev):
                capi.check(capi.load_library().bsvd_host_sync(self._handle))

    # ---------------------------------------------------------------- streaming mode
    def feedin_one_element(self, x, noise_map=None):
        """One pipeline step (bsvd_arch.py:485-488): x [1,C,H,W] or None -> [1,3,H,W] or None.
        The first `shift_num` (=16) calls of a stream return None."""
        dev = x.device if x is not None and x.is_cuda else torch.device(
            "cuda", torch.cuda.current_device())
        with torch.cuda.device(dev):
            lib = self._ensure_handle(dev)
            stream = torch.cuda.current_stream(dev).cuda_stream
            produced = C.c_int(0)
            if x is not None:
                xf = x.detach().to(dev).float().contiguous()
                _, Cc, H, W = xf.shape
                nm = None
                if noise_map is not None:
                    nm = noise_map.detach().to(dev).float().contiguous()
                self._stream_shape = (Cc, H, W)
                out = torch.empty((1, 3, H, W), dtype=torch.float32, device=dev)
                capi.check(lib.bsvd_stream_push(
                    self._handle, xf.data_ptr(), nm.data_ptr() if nm is not None else None,
                    out.data_ptr(), Cc, H, W, C.byref(produced), stream))
            else:
                if getattr(self, "_stream_shape", None) is None:
                    return None
                Cc, H, W = self._stream_shape
                out = torch.empty((1, 3, H, W), dtype=torch.float32, device=dev)
                capi.check(lib.bsvd_stream_push(self._handle, None, None, out.data_ptr(), Cc, H, W,
                                                C.byref(produced), stream))
        if not produced.value:
            return None
        od = self._out_dtype()
        return out if od == torch.float32 else out.to(od)

    def reset(self):
        """Clear all streaming buffers (bsvd_arch.py:459-461)."""
        self._stream_shape = None
        if getattr(self, "_handle", None) is not None:
            capi.check(capi.load_library().bsvd_reset(self._handle))

    def stream_graph_replays(self):
        """Streaming pushes served by replaying a captured CUDA graph (steady state) so far."""
        return 0 if self._handle is None else int(capi.load_library().bsvd_stream_graph_replays(self._handle))

    def count_shift(self):
        return self.shift_num

    def extra_repr(self):
        return "B200-native BSVD-64 (tcgen05/TMEM/TMA fused conv stages), precision=%s" % (
            self.precision or "auto")

    @property
    def last_launch_count(self):
        return 0 if self._handle is None else capi.load_library().bsvd_last_launch_count(
            self._handle)


class TSN(BSVD):
    """Drop-in for the reference's training twin `TSN` (Experimental_root/archs/tsm_arch.py:10-72:
    WNet_multistage + TemporalShift, shift_type 'TSM', shift_div 8) — FORWARD only, on the same kernels.

    * parameters carry the TSN names (`base_model.nets_list.{0,1}.<block>...`), so the reference's TSN
      checkpoints load with `load_state_dict` / `{'params': ...}` files as they are;
    * eval mode = `batch_shift` (temporal_shift.py:53-80) over the whole [N*F] batch: one continuous stream
      (the state a fresh `global_queue_buffer._init(0)` gives: no past buffer is consumed);
    * train mode = `shift(x, n_segment)` (temporal_shift.py:27-49): the batch is `N*F / num_segments`
      independent clips of `num_segments` frames, zero folds at every clip boundary (bsvd_forward_clips).

    There is no backward pass: calling it with autograd enabled on parameters that require gradients
    raises instead of silently returning a constant w.r.t. the weights."""

    def __init__(self, num_segments=11, base_model='WNet_multistage', shift_type='TSM', shift_div=8,
                 inplace=False, net2d_opt=None, enable_past_buffer=True, precision=None, **kwargs):
        if base_model != 'WNet_multistage' or shift_type != 'TSM' or shift_div != 8 or inplace:
            raise NotImplementedError("bsvd_b200.TSN implements base_model='WNet_multistage', shift_type='TSM', "
                                      "shift_div=8, inplace=False (options/train/*.yml); no CPU/PyTorch fallback")
        opt = dict(net2d_opt or {})
        opt.pop("pretrain_ckpt", None)
        super().__init__(pretrain_ckpt=None, precision=precision, **opt)
        self.num_segments = int(num_segments)
        # re-home the 32 convs under the TSN names (same Conv2d objects, same execution order)
        convs = self._convs()
        del self.temp1, self.temp2
        self.base_model = _Node()
        self.base_model.add_module("nets_list", _Node())
        names = []
        for blk in range(2):
            root = _Node()
            self.base_model.nets_list.add_module(str(blk), root)
            for i, (_stem, tsn) in enumerate(_LAYERS):
                _attach(root, tsn, convs[blk * 16 + i])
                names.append(f"base_model.nets_list.{blk}.{tsn}")
        self._param_names = names

    def _first_weight(self):
        return self.base_model.nets_list._modules['0'].inc.convblock._modules['0'].weight

    def load_tsn_state(self, ckpt_state):
        first = list(ckpt_state.keys())[0]
        sd = {(k[len('module.'):] if first.startswith('module.') else k): v for k, v in ckpt_state.items()}
        self.load_state_dict(sd, strict=True)
        self._weights_sig = None

    def forward(self, input, noise_map=None):
        """[N,F,C,H,W] (or [N*F,C,H,W]) -> same leading shape with 3 channels (tsm_arch.py:59-72)."""
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()) and self.training:
            raise NotImplementedError("bsvd_b200.TSN has no backward pass: run the training-mode forward under "
                                      "torch.no_grad() (validation inside a training loop), train with the reference")
        if noise_map is not None:
            input = torch.cat([input, noise_map], dim=2 if input.dim() == 5 else 1)
        five = input.dim() == 5
        x = input.reshape(-1, *input.shape[-3:]) if five else input
        clip = None
        if self.training:
            if x.shape[0] % self.num_segments:
                raise RuntimeError(f"shape '[-1, {self.num_segments}, ...]' is invalid for {x.shape[0]} frames")
            clip = self.num_segments
        out = self._run_stream(x, None, clip_len=clip)
        return out.reshape(*input.shape[:2], 3, *input.shape[-2:]) if five else out


def params_to_numpy(module: BSVD):
    return {k: v.detach().float().cpu().numpy().astype(np.float32)
            for k, v in module.state_dict().items()}
