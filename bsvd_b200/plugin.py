"""Registration under the reference's plugin API.

The reference builds its network with
    basicsr.archs.build_network(opt['network_g'])  ->  ARCH_REGISTRY.get('BSVD')(**opt)
(BasicSR/basicsr/archs/__init__.py:19-25, basicsr/utils/registry.py:38-66,79) and registers its own
`BSVD` class as an import side effect of `Experimental_root.archs` (bsvd_arch.py:440-442).
`Registry._do_register` asserts that a name is new, so the replacement cannot call `.register()`
again; `install()` lets the reference register first and then overwrites the registry entry, which
is all `build_network` looks at.  options/test/bsvd_c64.yml and profile.py stay byte-identical.
"""
from __future__ import annotations

import importlib
import sys
import types


def install(registry=None, import_reference_archs: bool = True, install_tsn: bool = False):
    """Make ARCH_REGISTRY['BSVD'] resolve to the B200-native class.  Returns the previous entry."""
    from .arch import BSVD
    if registry is None:
        if import_reference_archs:
            try:
                importlib.import_module("Experimental_root.archs")   # reference registers itself
            except Exception as e:  # noqa: BLE001
                raise RuntimeError(
                    "could not import Experimental_root.archs; put the BSVD checkout and its "
                    "BasicSR directory on sys.path (append, do not prepend: the checkout's "
                    "profile.py shadows the stdlib module)") from e
        registry = importlib.import_module("basicsr.utils.registry").ARCH_REGISTRY
    prev = registry._obj_map.get("BSVD")
    registry._obj_map["BSVD"] = BSVD
    if install_tsn:
        from .arch import TSN
        registry._obj_map["TSN"] = TSN      # forward-only twin (validation inside a training run)
    return prev


def stub_optional_dependencies():
    """profile.py / run_test.py import a few packages that are irrelevant to the forward pass and
    absent from this image: nvidia.dali, torchstat, ptflops, thop, line_profiler,
    and the generated basicsr/version.py.  Provide empty stand-ins so the scripts load unchanged."""
    def mod(name, **attrs):
        if name in sys.modules:
            return sys.modules[name]
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    try:
        importlib.import_module("basicsr.version")
    except Exception:  # noqa: BLE001
        mod("basicsr.version", __version__="1.3.4.2", __gitsha__="unknown", version_info=(1, 3, 4, 2))
    for name in ("torchstat", "ptflops", "thop", "line_profiler"):
        try:
            importlib.import_module(name)
        except Exception:  # noqa: BLE001
            mod(name, stat=None, get_model_complexity_info=None, profile=None, LineProfiler=object)
    try:
        importlib.import_module("nvidia.dali")
    except Exception:  # noqa: BLE001
        class _Pipeline:  # noqa: D401
            def __init__(self, *a, **k):
                pass
        nv = sys.modules.get("nvidia") or mod("nvidia")
        dali = mod("nvidia.dali", ops=mod("nvidia.dali.ops"), types=mod("nvidia.dali.types"))
        mod("nvidia.dali.pipeline", Pipeline=_Pipeline)
        plug = mod("nvidia.dali.plugin")
        mod("nvidia.dali.plugin.pytorch", DALIGenericIterator=object)
        nv.dali = dali
        dali.plugin = plug


def ensure_legacy_nvidia_smi():
    """profile.py:5-6 picks its GPU with `nvidia-smi -q -d Memory | grep -A4 GPU | grep Free`, which relies on
    the pre-R515 layout (Total / Used / Free directly under the GPU line).  Current drivers print a `Reserved`
    line as well, the pipe matches nothing and the script dies in np.argmax before importing torch.  When that
    is the case, put a pass-through `nvidia-smi` wrapper in front of PATH that answers exactly this query in the
    old layout (real numbers from --query-gpu) and forwards everything else to the real binary."""
    import os
    import shutil
    import stat
    import subprocess
    import tempfile
    real = shutil.which("nvidia-smi")
    if real is None:
        return None
    try:
        out = subprocess.run("nvidia-smi -q -d Memory | grep -A4 GPU | grep Free", shell=True, capture_output=True,
                             text=True, timeout=60).stdout
        if any(len(l.split()) >= 3 and l.split()[2].isdigit() for l in out.splitlines()):
            return None                      # the script's own parse works here
    except Exception:  # noqa: BLE001
        pass
    d = tempfile.mkdtemp(prefix="bsvd_b200_smi_")
    w = os.path.join(d, "nvidia-smi")
    with open(w, "w") as f:
        f.write(f"""#!/bin/sh
if [ "$1" = "-q" ] && [ "$2" = "-d" ] && [ "$3" = "Memory" ]; then
  {real} --query-gpu=pci.bus_id,memory.total,memory.used,memory.free --format=csv,noheader,nounits | \\
  while IFS=, read bus total used free; do
    echo "GPU $bus"; echo "    FB Memory Usage"
    echo "        Total                             :$total MiB"
    echo "        Used                              :$used MiB"
    echo "        Free                              :$free MiB"
  done
else
  exec {real} "$@"
fi
""")
    os.chmod(w, os.stat(w).st_mode | stat.S_IXUSR | stat.S_IXGRP | stat.S_IXOTH)
    os.environ["PATH"] = d + os.pathsep + os.environ.get("PATH", "")
    return w


def reseed_after_model_build():
    """A/B harness only.  The reference seeds the CPU generator once (options.py: set_random_seed), then builds
    the model (weight init draws from that generator) and only then synthesises the validation noise
    (ValFolderDataset.__getitem__).  Two model classes that draw a different amount of random numbers in their
    constructors therefore see DIFFERENT noise.  Re-applying the seed right after build_model makes runs with
    different classes denoise the same frames; nothing in the reference tree is touched."""
    import torch
    models = importlib.import_module("basicsr.models")
    orig = models.build_model

    def build_model(opt):
        m = orig(opt)
        seed = opt.get("manual_seed", 0) or 0
        import random
        import numpy as np
        random.seed(seed)
        np.random.seed(seed)
        torch.manual_seed(seed)
        torch.cuda.manual_seed_all(seed)
        return m

    models.build_model = build_model
    for name in ("basicsr.test", "basicsr.train", "basicsr"):
        mod = sys.modules.get(name)
        if mod is not None and getattr(mod, "build_model", None) is orig:
            mod.build_model = build_model


def run_reference_script(path: str, reference_root: str, install_b200: bool = True, reseed: bool = False):
    """`python -m bsvd_b200.plugin <reference_root> profile.py`: run an unmodified reference entry
    point (profile.py, run_test.py) with the B200 class installed under ARCH_REGISTRY['BSVD'].
    `--reference-only` runs the same script in the same harness with the reference's own class."""
    import os
    import runpy
    sys.path.append(os.path.join(reference_root, "BasicSR"))
    sys.path.append(reference_root)
    import atexit
    import json
    reference_root = os.path.abspath(reference_root)
    stub_optional_dependencies()
    ensure_legacy_nvidia_smi()
    if install_b200:
        install()
        from .arch import BSVD
        atexit.register(lambda: print("bsvd_b200.plugin: " + json.dumps(BSVD.stats), flush=True))
    else:
        # same harness (stubs, cwd, argv), the reference's own BSVD class: the A side of an A/B run
        importlib.import_module("Experimental_root.archs")
        atexit.register(lambda: print("bsvd_b200.plugin: reference class left in place", flush=True))
    if reseed:
        importlib.import_module("basicsr.test")
        reseed_after_model_build()
    os.chdir(reference_root)
    sys.argv = [os.path.join(reference_root, path)] + sys.argv[3:]
    runpy.run_path(os.path.join(reference_root, path), run_name="__main__")


if __name__ == "__main__":
    # python -m bsvd_b200.plugin [--reference-only] [--reseed-after-build] <reference_root> <script> [script args...]
    flags = set()
    while len(sys.argv) > 1 and sys.argv[1] in ("--reference-only", "--reseed-after-build"):
        flags.add(sys.argv.pop(1))
    run_reference_script(sys.argv[2], sys.argv[1], install_b200="--reference-only" not in flags,
                         reseed="--reseed-after-build" in flags)
