"""Generate tests/golden/*.npz from the UNMODIFIED reference (ChenyangQiQi/BSVD at /root/reference).

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py

For every case it builds the reference's TSN (clip order, tsm_arch.py) and BSVD (stream order,
bsvd_arch.py) classes, loads the same seeded synthetic checkpoint (oracle.make_synthetic_params —
the reference ships no checkpoint), runs both on the same seeded noisy clip on the CPU and stores
input recipe + outputs.  The reference BSVD hard-codes 'cuda' (bsvd_arch.py:94,104,520); the
harness-only shim below redirects those three call sites to the CPU without touching the source.
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = os.environ.get("BSVD_REFERENCE", "/root/reference")

from oracle import bsvd_oracle as O  # noqa: E402

CASES = [
    # name, T, H, W, param seed, weight scale, clip seed
    ("trained_like_T6_32x48", 6, 32, 48, 0, 0.5, 1),
    ("trained_like_T1_16x20", 1, 16, 20, 0, 0.5, 2),
    ("trained_like_T3_20x36", 3, 20, 36, 3, 0.5, 3),
    ("default_init_T4_24x32", 4, 24, 32, 0, 1.0, 4),
]


def import_reference():
    # never put the reference at the FRONT of sys.path: its profile.py shadows the stdlib module
    sys.path.append(os.path.join(REF, "BasicSR"))
    sys.path.append(REF)
    ver = types.ModuleType("basicsr.version")
    ver.__version__ = "1.3.4.2"
    ver.__gitsha__ = "unknown"
    ver.version_info = (1, 3, 4, 2)
    sys.modules["basicsr.version"] = ver
    import basicsr  # noqa: F401
    import Experimental_root.archs  # noqa: F401  (registers TSN and BSVD)
    from basicsr.utils.registry import ARCH_REGISTRY
    from Experimental_root.models import global_queue_buffer
    return ARCH_REGISTRY, global_queue_buffer


class cpu_shim:
    """Redirect the reference's hard-coded CUDA placement to the CPU (harness only)."""

    def __enter__(self):
        self._zeros = torch.zeros
        self._cuda = torch.Tensor.cuda

        def zeros(*a, **k):
            k.pop("device", None)
            return self._zeros(*a, **k)

        torch.zeros = zeros
        torch.Tensor.cuda = lambda self, *a, **k: self
        return self

    def __exit__(self, *exc):
        torch.zeros = self._zeros
        torch.Tensor.cuda = self._cuda


def main_c32(REG, gqb, out_dir):
    """The blind c32 configuration of options/train/0402_*_blind_c32.yml (net2d_opt :63-68), clip
    order through the reference's TSN class (its streaming class cannot run blind models whose
    mid_ch is not 3, bsvd_arch.py:449-452)."""
    cfg = O.C32
    net2d = dict(chns=list(cfg["chns"]), mid_ch=cfg["mid_ch"], shift_input=False, norm="none", blind=True)
    for name, T, H, W, pseed, scale, cseed in [("c32_blind_T5_24x40", 5, 24, 40, 7, 0.5, 9),
                                               ("c32_blind_T2_16x132", 2, 16, 132, 8, 0.5, 10)]:
        sd = O.make_synthetic_params(pseed, scale, in_ch=3, chns=cfg["chns"], mid_ch=cfg["mid_ch"],
                                     interm_ch=cfg["interm_ch"])
        tsn = REG.get("TSN")(num_segments=11, base_model="WNet_multistage", shift_type="TSM",
                             shift_div=8, inplace=False, net2d_opt=net2d).eval()
        assert sorted(tsn.state_dict().keys()) == sorted(sd.keys()), "TSN key layout mismatch"
        tsn.load_state_dict(sd, strict=True)
        x, _ = O.make_synthetic_clip(T, H, W, cseed)
        gqb._init(0)
        with torch.no_grad():
            y_tsn = tsn(x[None, :, :3])[0]
        gqb._clean()
        mine = O.forward_clip(O.layers_from_tsn_state(sd), x[:, :3], act=cfg["act"])
        d = float((y_tsn - mine).abs().max())
        print(f"{name}: reference TSN vs oracle max-abs {d:.3e}; |y|max {float(y_tsn.abs().max()):.3f}")
        assert d < 2e-4, d
        np.savez_compressed(
            os.path.join(out_dir, name + ".npz"), T=T, H=H, W=W, param_seed=pseed, weight_scale=scale,
            clip_seed=cseed, params_digest=O.params_digest(sd), x_digest=O.params_digest({"x": x}),
            y_clip=y_tsn.numpy().astype(np.float32),
            n_params=sum(p.numel() for p in tsn.parameters()),
            reference_commit="29a6f05", torch_version=torch.__version__)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 1)
    REG, gqb = import_reference()
    net2d = dict(chns=[64, 128, 256], mid_ch=64, shift_input=False, norm="none", interm_ch=64,
                 act="relu6")
    out_dir = os.path.dirname(os.path.abspath(__file__))
    if os.environ.get("GOLDEN_ONLY") == "c32":
        return main_c32(REG, gqb, out_dir)
    main_c32(REG, gqb, out_dir)
    for name, T, H, W, pseed, scale, cseed in CASES:
        sd = O.make_synthetic_params(pseed, scale)
        tsn = REG.get("TSN")(num_segments=11, base_model="WNet_multistage", shift_type="TSM",
                             shift_div=8, inplace=False, net2d_opt=net2d).eval()
        assert sorted(tsn.state_dict().keys()) == sorted(sd.keys()), "TSN key layout mismatch"
        tsn.load_state_dict(sd, strict=True)
        with tempfile.TemporaryDirectory() as td:
            ck = os.path.join(td, "bsvd-64.pth")
            torch.save({"params": sd}, ck)
            bsvd = REG.get("BSVD")(pretrain_ckpt=ck, **net2d).eval()
        # our restatement of the streaming-class key names must match the reference's
        assert sorted(O.bsvd_keys()) == sorted({k.rsplit(".", 1)[0] for k in bsvd.state_dict()})
        x, clean = O.make_synthetic_clip(T, H, W, cseed)
        gqb._init(0)
        with torch.no_grad():
            y_tsn = tsn(x[None, :, :3], noise_map=x[None, :, 3:4])[0]
        gqb._clean()
        with cpu_shim(), torch.no_grad():
            y_bsvd = bsvd(x[None, :, :3], noise_map=x[None, :, 3:4])[0]
            # None protocol: first non-None output index of feedin_one_element
            first = None
            outs = []
            for i in range(T):
                outs.append(bsvd.feedin_one_element(x[i:i + 1]))
            n_calls = T
            while sum(o is not None for o in outs) < T:
                outs.append(bsvd.feedin_one_element(None))
                n_calls += 1
            first = next(i for i, o in enumerate(outs) if o is not None)
            bsvd.reset()
        d = float((y_tsn - y_bsvd).abs().max())
        print(f"{name}: TSN vs BSVD max-abs {d:.3e}; |y|max {float(y_bsvd.abs().max()):.3f}; "
              f"first output at call {first}, calls to drain {n_calls}")
        assert d < 2e-4, d
        np.savez_compressed(
            os.path.join(out_dir, name + ".npz"),
            T=T, H=H, W=W, param_seed=pseed, weight_scale=scale, clip_seed=cseed,
            params_digest=O.params_digest(sd), x_digest=O.params_digest({"x": x}),
            y_stream=y_bsvd.numpy().astype(np.float32), y_clip=y_tsn.numpy().astype(np.float32),
            first_output_call=first, calls_to_drain=n_calls, shift_num=int(bsvd.shift_num),
            n_params=sum(p.numel() for p in bsvd.parameters()),
            reference_commit="29a6f05", torch_version=torch.__version__)


if __name__ == "__main__":
    main()
