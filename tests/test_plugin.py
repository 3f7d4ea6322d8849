"""The registry hook: the reference's build_network must hand the yml's network_g block to the
B200 class.  Uses a faithful miniature of basicsr.utils.registry.Registry (registry.py:1-82) so the
test runs where the reference checkout is absent; when /root/reference exists (build container)
the real registry, the real yml and the real build_network are exercised as well."""
import os
import sys
import types

import pytest

from bsvd_b200 import plugin
from bsvd_b200.arch import BSVD

REF = "/root/reference"


class MiniRegistry:
    """Same contract as basicsr.utils.registry.Registry: _obj_map, assert-on-duplicate, get()."""

    def __init__(self, name):
        self._name, self._obj_map = name, {}

    def _do_register(self, name, obj):
        assert name not in self._obj_map, f"An object named '{name}' was already registered"
        self._obj_map[name] = obj

    def register(self, obj=None):
        self._do_register(obj.__name__, obj)
        return obj

    def get(self, name):
        ret = self._obj_map.get(name)
        if ret is None:
            raise KeyError(name)
        return ret


YML_NETWORK_G = dict(type="BSVD", chns=[64, 128, 256], mid_ch=64, shift_input=False, norm="none",
                     interm_ch=64, act="relu6", pretrain_ckpt=None)


def test_install_overrides_existing_entry():
    reg = MiniRegistry("arch")

    class BSVD_ref:   # stands for the reference class registered at import time
        pass
    BSVD_ref.__name__ = "BSVD"
    reg.register(BSVD_ref)
    with pytest.raises(AssertionError):
        reg.register(BSVD)                      # what a naive second registration would hit
    prev = plugin.install(registry=reg)
    assert prev is BSVD_ref and reg.get("BSVD") is BSVD
    opt = dict(YML_NETWORK_G)
    net = reg.get(opt.pop("type"))(**opt)       # build_network, archs/__init__.py:19-22
    assert isinstance(net, BSVD) and net.shift_num == 16


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout only exists in the build container")
def test_real_registry_and_yml(tmp_path, monkeypatch):
    import yaml
    import torch
    from oracle import bsvd_oracle as O
    sys.path.append(os.path.join(REF, "BasicSR"))
    sys.path.append(REF)
    plugin.stub_optional_dependencies()
    prev = plugin.install()
    assert prev is not None and prev.__module__.endswith("bsvd_arch")
    from basicsr.archs import build_network
    from basicsr.utils.options import ordered_yaml
    with open(os.path.join(REF, "options/test/bsvd_c64.yml")) as f:
        opt = yaml.load(f, Loader=ordered_yaml()[0])
    ck = tmp_path / "experiments" / "pretrained_ckpt"
    ck.mkdir(parents=True)
    torch.save({"params": O.make_synthetic_params(0)}, ck / "bsvd-64.pth")
    monkeypatch.chdir(tmp_path)                 # the yml's pretrain_ckpt path is relative
    net = build_network(dict(opt["network_g"]))
    assert isinstance(net, BSVD)
    got = O.layers_from_bsvd_state(net.state_dict())
    want = O.layers_from_tsn_state(O.make_synthetic_params(0))
    assert all(torch.equal(a[0], b[0]) for a, b in zip(got, want))


def test_bench_reference_arm_prints_one_json_line(tmp_path):
    """`bench.py --impl reference` (the driver's reference arm) on a tiny shape: exactly one stdout line, the
    contract's keys, and — when the unmodified reference is staged under baseline/_ref — kind 'reference'."""
    import json
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "2",
                        "--warmup", "1", "--shape", "3,16,24"], capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["steps"] == 2
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["cpu_baseline"]["kind"] in ("reference", "port")
    staged = os.path.isdir(os.path.join(root, "baseline", "_ref", "Experimental_root"))
    assert d["cpu_baseline"]["kind"] == ("reference" if staged else "port")
