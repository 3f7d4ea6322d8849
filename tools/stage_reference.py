#!/usr/bin/env python
"""Stage the UNMODIFIED reference under baseline/_ref/ so it can travel to the GPU box.

    python tools/stage_reference.py [--src /root/reference]

`/root/reference` only exists in the build container.  The GPU box receives a snapshot of this
repository's working tree, in which `baseline/_ref/` is git-ignored (it never enters history: the
reference's sources are not part of this repository) but NOT gpurun-ignored, so whatever is staged
there is present when `bench.py --impl reference`, the `gpu_reference` block of the B200 bench line
and the `-m gpu` reference tests run.  What is staged is a byte-for-byte copy of the Python the
reference needs for `profile.py` / `run_test.py` / `BSVD.forward`:

    profile.py  run_test.py  run.py            entry points (profile.py:55-83 is the headline harness)
    options/                                   options/test/bsvd_c64.yml, options/train/*.yml
    Experimental_root/                         archs (bsvd_arch.py, tsm_arch.py, ...), models, scripts, data
    BasicSR/basicsr/, BasicSR/VERSION          registry, build_network / build_model, metrics

(no datasets, figures, docs or native op sources are needed).  A MANIFEST.json with the sha256 of every
staged file is written next to them; `verify()` re-hashes it so tests can assert the copy is unmodified.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "baseline", "_ref")
DEFAULT_SRC = os.environ.get("BSVD_REFERENCE", "/root/reference")

TOP_FILES = ("profile.py", "run_test.py", "run.py", "LICENSE.md", "requirements.txt")
TREES = ("options", "Experimental_root", os.path.join("BasicSR", "basicsr"))
EXTRA = (os.path.join("BasicSR", "VERSION"), os.path.join("BasicSR", "LICENSE"))
KEEP_EXT = (".py", ".yml", ".yaml", ".npz", ".txt", ".md")
SKIP_DIRS = ("__pycache__", os.path.join("ops", "dcn", "src"), os.path.join("ops", "fused_act", "src"),
             os.path.join("ops", "upfirdn2d", "src"), os.path.join("data", "meta_info"))


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def _wanted(rel):
    d = "/" + os.path.dirname(rel).replace(os.sep, "/") + "/"
    if any(("/" + s.replace(os.sep, "/") + "/") in d for s in SKIP_DIRS):
        return False
    return rel.endswith(KEEP_EXT)


def stage(src: str = DEFAULT_SRC, dest: str = DEST, quiet: bool = False) -> dict:
    """Copy the reference's Python into `dest`; returns the manifest {relative path: sha256}."""
    if not os.path.isdir(src):
        raise FileNotFoundError(f"reference checkout not found at {src}")
    if os.path.isdir(dest):
        shutil.rmtree(dest)
    os.makedirs(dest)
    manifest = {}

    def copy(rel):
        s, d = os.path.join(src, rel), os.path.join(dest, rel)
        if not os.path.isfile(s) or os.path.islink(s) and not os.path.exists(s):
            return
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        manifest[rel.replace(os.sep, "/")] = _sha(d)

    for f in TOP_FILES + EXTRA:
        copy(f)
    for tree in TREES:
        for base, dirs, files in os.walk(os.path.join(src, tree)):
            dirs[:] = [d for d in dirs if d != "__pycache__"]
            for f in files:
                rel = os.path.relpath(os.path.join(base, f), src)
                if _wanted(rel):
                    copy(rel)
    meta = {"source": src, "files": manifest,
            "note": "byte-for-byte copy of ChenyangQiQi/BSVD Python sources; staged, never committed"}
    with open(os.path.join(dest, "MANIFEST.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    if not quiet:
        print(f"staged {len(manifest)} reference files under {os.path.relpath(dest, ROOT)}/")
    return manifest


def verify(dest: str = DEST) -> int:
    """Re-hash the staged copy against its manifest; returns the number of files checked."""
    with open(os.path.join(dest, "MANIFEST.json")) as f:
        files = json.load(f)["files"]
    for rel, sha in files.items():
        got = _sha(os.path.join(dest, rel))
        if got != sha:
            raise RuntimeError(f"staged reference file {rel} was modified after staging")
    return len(files)


def staged(dest: str = DEST) -> bool:
    return os.path.isfile(os.path.join(dest, "MANIFEST.json")) and os.path.isfile(os.path.join(dest, "profile.py"))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default=DEFAULT_SRC)
    a = ap.parse_args()
    stage(a.src)
    print(f"verified {verify()} files")
    sys.exit(0)
