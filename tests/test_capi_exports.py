"""CPU-side checks of the boundary: the C-ABI library loads here (no GPU) and exports every symbol
include/bsvd_b200.h declares; the nn.Module mirrors the reference's names and error behaviour."""
import os
import re

import pytest
import torch

from bsvd_b200 import capi
from oracle import bsvd_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "bsvd_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bsvd_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = capi.load_library()
    names = _declared()
    assert len(names) >= 13
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(capi.EXPORTS) == names


def test_version_and_error_string():
    lib = capi.load_library()
    assert b"sm_100a" in lib.bsvd_version()
    assert isinstance(lib.bsvd_last_error(), bytes)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_gpu():
    import ctypes as C
    lib = capi.load_library()
    cfg = capi.BsvdConfig()
    cfg.chns[0], cfg.chns[1], cfg.chns[2] = 64, 128, 256
    cfg.mid_ch = cfg.interm_ch = 64
    cfg.in_ch, cfg.out_ch, cfg.act_relu6, cfg.norm_none, cfg.device = 4, 3, 1, 1, -1
    h = C.c_void_p()
    assert lib.bsvd_create(C.byref(cfg), C.byref(h)) != 0
    assert b"no CPU fallback" in lib.bsvd_last_error()


def test_create_rejects_other_configs():
    import ctypes as C
    lib = capi.load_library()
    cfg = capi.BsvdConfig()
    for chns, norm_none in (((16, 32, 64), 1), ((64, 128, 256), 0), ((64, 128, 512), 1)):
        cfg.chns[0], cfg.chns[1], cfg.chns[2] = chns
        cfg.mid_ch, cfg.interm_ch = 64, 30
        cfg.in_ch, cfg.out_ch, cfg.act_relu6, cfg.norm_none, cfg.device = 4, 3, 1, norm_none, -1
        h = C.c_void_p()
        assert lib.bsvd_create(C.byref(cfg), C.byref(h)) != 0
        assert b"no CPU fallback" in lib.bsvd_last_error()


def _net(**kw):
    from bsvd_b200.arch import BSVD
    args = dict(chns=[64, 128, 256], mid_ch=64, shift_input=False, norm='none', interm_ch=64,
                act='relu6', pretrain_ckpt=None)
    args.update(kw)
    return BSVD(**args)


def test_module_state_dict_uses_reference_names():
    net = _net()
    keys = {k.rsplit(".", 1)[0] for k in net.state_dict()}
    assert keys == set(O.bsvd_keys())          # names checked against the reference in make_golden.py
    assert sum(p.numel() for p in net.parameters()) == 9815683
    assert net.shift_num == 16
    assert "B200-native" in str(net)


def test_module_loads_tsn_checkpoint(tmp_path):
    sd = O.make_synthetic_params(0)
    for prefix in ("", "module."):
        ck = tmp_path / f"ck{len(prefix)}.pth"
        torch.save({"params": {prefix + k: v for k, v in sd.items()}}, ck)
        net = _net(pretrain_ckpt=str(ck))
        got = O.layers_from_bsvd_state(net.state_dict())
        want = O.layers_from_tsn_state(sd)
        assert all(torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) for a, b in zip(got, want))


def test_module_state_dict_roundtrip_and_half():
    a, b = _net(), _net()
    b.load_state_dict(a.state_dict(), strict=True)
    assert all(torch.equal(p, q) for p, q in zip(a.parameters(), b.parameters()))
    assert next(a.half().parameters()).dtype == torch.float16


def test_module_rejects_unsupported_configs():
    with pytest.raises(NotImplementedError):
        _net(chns=[16, 32, 64], interm_ch=30)
    with pytest.raises(NotImplementedError):
        _net(act='leaky')
    with pytest.raises(NotImplementedError):
        _net(norm='bn')
    with pytest.raises(NotImplementedError):
        _net(shift_input=True)
    # the c32 configurations of options/train/0402_*_c32.yml construct (same parameter names)
    n32 = _net(chns=[32, 64, 128], mid_ch=32, interm_ch=30, act='relu', blind=True)
    assert n32.state_dict()["temp1.inc.convblock.0.weight"].shape == (30, 3, 3, 3)
    assert n32.state_dict()["temp2.inc.convblock.0.weight"].shape == (30, 32, 3, 3)
    assert n32.state_dict()["temp1.upc1.convblock.0.weight"].shape == (128, 64, 3, 3)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_module_forward_fails_loudly_without_gpu():
    net = _net()
    with pytest.raises(Exception):
        net(torch.zeros(1, 2, 4, 8, 8))


def test_header_is_valid_c_and_a_plain_c_client_links(tmp_path):
    """include/bsvd_b200.h compiles as C99 and examples/c_client.c links against the library with
    nothing but gcc (no torch, no CUDA headers); without a GPU the client reports the loud failure."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    exe = str(tmp_path / "c_client")
    libdir = os.path.join(ROOT, "bsvd_b200", "lib")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "c_client.c"), "-L" + libdir, "-lbsvd_b200",
                           "-Wl,-rpath," + libdir, "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert "sm_100a" in out.stdout
    if not torch.cuda.is_available():
        assert "no CPU fallback" in out.stdout
    else:
        assert "layer 31: weight [3,64,3,3]" in out.stdout


def test_frame_io_names_and_png_roundtrip(tmp_path):
    """bsvd_b200.frames: the reference's file ordering rule (utils_common.py:78-95: by the integer made of all
    digits of the path) and a lossless PNG write / read round trip."""
    import numpy as np
    from bsvd_b200 import frames
    rng = np.random.RandomState(0)
    seq = rng.randint(0, 256, size=(12, 20, 28, 3), dtype=np.uint8)
    d = tmp_path / "seq"
    paths = frames.write_sequence_u8(seq, str(d))
    assert [os.path.basename(p) for p in paths][:3] == ["00000000.png", "00000001.png", "00000002.png"]
    names = frames.get_imagenames(str(d))
    assert [os.path.basename(n) for n in names] == [f"{i:08d}.png" for i in range(12)]
    back = frames.read_sequence_u8(str(d))
    assert back.dtype == np.uint8 and np.array_equal(back, seq)
    assert frames.read_sequence_u8(str(d), max_num_fr=5).shape[0] == 5
    with pytest.raises(FileNotFoundError):
        frames.read_sequence_u8(str(tmp_path / "missing"))


def test_tsn_twin_uses_the_reference_tsn_parameter_names():
    """bsvd_b200.arch.TSN (forward-only twin of tsm_arch.TSN): its state_dict keys are the TSN checkpoint's."""
    from bsvd_b200.arch import TSN
    from oracle import bsvd_oracle as O
    net = TSN(num_segments=4, net2d_opt=dict(chns=[64, 128, 256], mid_ch=64, shift_input=False, norm='none',
                                             interm_ch=64, act='relu6'))
    sd = O.make_synthetic_params(0, 0.5)
    assert sorted(net.state_dict().keys()) == sorted(sd.keys())
    net.load_state_dict(sd, strict=True)
    net.load_tsn_state({"module." + k: v for k, v in sd.items()})
    net.train()
    with pytest.raises(NotImplementedError):
        net(torch.zeros(1, 4, 4, 8, 8))            # no backward pass: refuses under autograd


def test_production_kernel_instances_do_not_spill():
    """Register budget of the built library (cuobjdump -res-usage): a 320-thread CTA leaves 168 registers per
    thread, the skip-in-the-epilogue instances sit right under that ceiling, and a spill inside their epilogue
    loop doubled a stage's time once (c32 upc1.convblock.0, 0.22 -> 0.45 ms).  Only the general debug instance
    (all epilogue features compiled in, mask 543) may use local memory; the first conv at most 16 bytes."""
    import shutil
    import subprocess
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([exe, "-res-usage", capi.LIB_PATH], capture_output=True, text=True, timeout=300).stdout
    funcs = re.findall(r"Function (\S+?):\s*\n\s*REG:(\d+) STACK:(\d+)", out)
    conv = [(f, int(r), int(s)) for f, r, s in funcs if "conv3x3_tc_kernel" in f]
    first = [(f, int(r), int(s)) for f, r, s in funcs if "first_conv_kernel" in f]
    final = [(f, int(r), int(s)) for f, r, s in funcs if "final_conv_kernel" in f]
    assert len(conv) >= 100 and len(first) >= 6 and len(final) >= 4, (len(conv), len(first), len(final))
    general = "ELi%dE" % 543          # kMaskAll: RELU6 | RELU | SHIFT | PIXSHUF | SKIP | RESID_IN
    bad = [(f, r, s) for f, r, s in conv if s > 0 and general not in f]
    assert not bad, bad
    assert all(r <= 168 for _, r, _ in conv)
    assert all(s <= 16 for _, _, s in first), first
    assert all(s == 0 for _, _, s in final), final
