"""The UNMODIFIED reference in the loop on the B200 (staged under baseline/_ref by
tools/stage_reference.py; the GPU box has no /root/reference).

  * profile.py (the reference's headline harness, profile.py:55-83) runs unchanged, end to end, with
    ARCH_REGISTRY['BSVD'] resolving to the B200 class (`python -m bsvd_b200.plugin baseline/_ref profile.py`)
  * our output against the live reference BSVD through PyTorch/cuDNN with TF32 off on the same GPU, on
    the benchmarked [1,10,4,540,960] clip, every frame
  * the staged copy is byte-identical to what was staged (sha256 manifest)
"""
import json
import os
import re
import subprocess
import sys

import pytest
import torch

from oracle import bsvd_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")

sys.path.insert(0, os.path.join(ROOT, "tools"))
import stage_reference  # noqa: E402
from baseline import reference_runner as R  # noqa: E402

needs_ref = pytest.mark.skipif(not R.available(), reason="baseline/_ref not staged (python tools/stage_reference.py "
                                                        "in the build container)")


@needs_ref
def test_staged_reference_is_unmodified():
    assert stage_reference.verify(REF) > 50


@needs_ref
def test_profile_py_runs_unchanged_through_the_plugin(tmp_path):
    ck = os.path.join(REF, "experiments", "pretrained_ckpt")
    os.makedirs(ck, exist_ok=True)
    torch.save({"params": O.make_synthetic_params(0, 0.5)}, os.path.join(ck, "bsvd-64.pth"))
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    r = subprocess.run([sys.executable, "-m", "bsvd_b200.plugin", REF, "profile.py"], cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=900)
    out = r.stdout + "\n" + r.stderr
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "profile_py_through_plugin.log"), "w") as f:
        f.write(out)
    assert r.returncode == 0, out[-4000:]
    assert stage_reference.verify(REF) > 50                       # profile.py and the yml were not touched
    m = re.search(r"bsvd_b200\.plugin: (\{.*\})", out)
    assert m, out[-2000:]
    stats = json.loads(m.group(1))
    assert stats["instances"] >= 1 and stats["forward_calls"] >= 10, stats
    assert stats["kernel_launches"] == 32 * stats["forward_calls"], stats
    assert "test function name: <class 'bsvd_b200.arch.BSVD'>" in out   # profile.py:41 names what it times
    assert "output shape is torch.Size([1, 10, 3, 540, 960])" in out
    t = re.findall(r"loops, mean of best \d+: ([0-9.]+) sec per loop", out)
    assert t, out[-2000:]
    assert float(t[-1]) < 0.05          # 10 frames at 540x960: ~10 ms; the reference itself needs ~52 ms in fp16


@needs_ref
def test_matches_live_reference_on_gpu_full_size_all_frames():
    from bsvd_b200.arch import BSVD
    sd = O.make_synthetic_params(0, 0.5)
    x, clean = O.make_synthetic_clip(10, 540, 960, seed=1)
    xd = x.cuda()
    ref_net = R.build_reference_bsvd(sd, "cuda")
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            ref = ref_net(xd[None])[0]
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    del ref_net
    for prec, tol in (("fp16", 1e-3), ("bf16", 1e-2)):
        net = BSVD(chns=[64, 128, 256], mid_ch=64, shift_input=False, norm='none', interm_ch=64,
                   act='relu6', pretrain_ckpt=None, precision=prec)
        net.load_tsn_state(sd)
        net = net.cuda().eval()
        with torch.no_grad():
            y = net(xd[None])[0]
        err = float((y - ref).abs().max())
        assert err <= tol, (prec, err)
        assert not net.overflowed()
        for t in (0, 4, 9):
            d = abs(O.psnr_float(y[t].clamp(0, 1).cpu(), clean[t]) - O.psnr_float(ref[t].clamp(0, 1).cpu(), clean[t]))
            assert d < 0.01, (prec, t, d)
        del net


@needs_ref
def test_reference_stream_protocol_live_matches_ours():
    """feedin_one_element call by call against the live reference class: same None pattern, same values."""
    from bsvd_b200.arch import BSVD
    sd = O.make_synthetic_params(3, 0.5)
    x, _ = O.make_synthetic_clip(5, 36, 68, seed=12)
    xd = x.cuda()
    ref_net = R.build_reference_bsvd(sd, "cuda")
    net = BSVD(chns=[64, 128, 256], mid_ch=64, shift_input=False, norm='none', interm_ch=64,
               act='relu6', pretrain_ckpt=None)
    net.load_tsn_state(sd)
    net = net.cuda().eval()
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            feeds = [xd[i:i + 1] for i in range(5)] + [None] * 17
            for i, f in enumerate(feeds):
                a, b = ref_net.feedin_one_element(f), net.feedin_one_element(f)
                assert (a is None) == (b is None), i
                if a is not None:
                    assert float((a - b).abs().max()) <= 1e-3, i
    finally:
        torch.backends.cudnn.allow_tf32 = old
        ref_net.reset()
        net.reset()


def _make_sequence_folder(d, n, h, w, ext, seed):
    import cv2
    import numpy as np
    os.makedirs(d, exist_ok=True)
    _, clean = O.make_synthetic_clip(n, h, w, seed=seed)
    for i in range(n):
        img = (clean[i].clamp(0, 1) * 255).round().byte().permute(1, 2, 0).numpy()[:, :, ::-1]
        assert cv2.imwrite(os.path.join(d, f"{i:05d}.{ext}"), np.ascontiguousarray(img))


@needs_ref
def test_run_test_py_runs_unchanged_through_the_plugin():
    """The reference's whole validation pipeline (run_test.py -> basicsr test_pipeline -> ValFolderDataset ->
    DenoisingModel.validation: pad, denoise_seq, crop, tensor2img, imwrite, calculate_psnr / psnr_float / ssim)
    with options/test/bsvd_c64.yml unchanged and ARCH_REGISTRY['BSVD'] resolving to the B200 class.  The yml's
    dataset folders (Set8, DAVIS) are dangling links in the reference; small synthetic sequences stand in."""
    import glob
    import shutil
    ck = os.path.join(REF, "experiments", "pretrained_ckpt")
    os.makedirs(ck, exist_ok=True)
    torch.save({"params": O.make_synthetic_params(0, 0.5)}, os.path.join(ck, "bsvd-64.pth"))
    shutil.rmtree(os.path.join(REF, "datasets"), ignore_errors=True)
    shutil.rmtree(os.path.join(REF, "results"), ignore_errors=True)
    _make_sequence_folder(os.path.join(REF, "datasets", "Set8", "tractor"), 5, 46, 62, "png", 1)     # odd size: pad / crop
    _make_sequence_folder(os.path.join(REF, "datasets", "Set8", "park"), 4, 48, 64, "png", 2)
    _make_sequence_folder(os.path.join(REF, "datasets", "DAVIS-2017-test-dev-480p", "JPEGImages", "480p", "aerobatics"),
                          5, 48, 80, "jpg", 3)
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    r = subprocess.run([sys.executable, "-m", "bsvd_b200.plugin", REF, "run_test.py", "-opt", "options/test/bsvd_c64.yml"],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=1200)
    out = r.stdout + "\n" + r.stderr
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "run_test_py_through_plugin.log"), "w") as f:
        f.write(out)
    assert r.returncode == 0, out[-4000:]
    assert stage_reference.verify(REF) > 50                       # run_test.py, the yml, the pipeline: untouched
    stats = json.loads(re.search(r"bsvd_b200\.plugin: (\{.*\})", out).group(1))
    # 5 Set8 entries x 2 folders + 5 DAVIS entries x 1 folder = 15 validated sequences, one forward each
    assert stats["instances"] == 1 and stats["forward_calls"] == 15, stats
    assert stats["kernel_launches"] == 32 * 15, stats
    pngs = glob.glob(os.path.join(REF, "results", "bsvd_c64", "visualization", "**", "*.png"), recursive=True)
    assert len(pngs) == 5 * (5 + 4) + 5 * 5, len(pngs)            # one PNG per validated frame
    assert "psnr_float" in out and "ssim" in out
    import cv2
    img = cv2.imread(sorted(p for p in pngs if "tractor" in p)[0])
    assert img is not None and img.shape == (46, 62, 3)           # cropped back to the odd source size
    # A/B/C/D on the same pipeline and seeds (the noise comes from the seeded CPU generator, so every run
    # denoises the same frames):
    # (`--reseed-after-build`: both classes must see the same noise although their constructors draw a different
    #  amount of random numbers for the weight init, see plugin.reseed_after_model_build)
    #   A  ours, yml unchanged (val.fp16: True -> our default fp16 mode)
    #   B  the reference's own class, yml unchanged (its fp16 autocast)
    #   C  the reference's own class in true fp32: the reference CLI's own override `--force_yml val:fp16=false`
    #      plus NVIDIA_TF32_OVERRIDE=0 (PyTorch would otherwise run the "fp32" convs in TF32)
    #   D  ours in the fp32-grade mode (BSVD_B200_PRECISION=fp32x3) with the same override
    pat = re.compile(r"# (psnr_float|ssim): ([0-9.]+)")

    def run(tag, extra_args, extra_env, reference_only):
        shutil.rmtree(os.path.join(REF, "results"), ignore_errors=True)
        cmd = [sys.executable, "-m", "bsvd_b200.plugin", "--reseed-after-build"] + (["--reference-only"] if reference_only else []) + \
              [REF, "run_test.py", "-opt", "options/test/bsvd_c64.yml"] + extra_args
        rr = subprocess.run(cmd, cwd=ROOT, env=dict(env, **extra_env), capture_output=True, text=True, timeout=1200)
        o = rr.stdout + "\n" + rr.stderr
        with open(os.path.join(ROOT, "gpurun_out", f"run_test_py_{tag}.log"), "w") as f:
            f.write(o)
        assert rr.returncode == 0, o[-4000:]
        return o, [(k, float(v)) for k, v in pat.findall(o)]

    _, m_a = run("A_ours_fp16", [], {}, False)
    out_b, m_b = run("B_reference_fp16", [], {}, True)
    assert "reference class left in place" in out_b
    fp32 = ["--force_yml", "val:fp16=false"]
    _, m_c = run("C_reference_fp32", fp32, {"NVIDIA_TF32_OVERRIDE": "0"}, True)
    _, m_d = run("D_ours_fp32x3", fp32, {"BSVD_B200_PRECISION": "fp32x3"}, False)
    assert len(m_a) == len(m_b) == len(m_c) == len(m_d) >= 22
    worst = {"A": 0.0, "B": 0.0, "D": 0.0}
    for (k, a), (_, b), (_, c), (_, d) in zip(m_a, m_b, m_c, m_d):
        if k != "psnr_float":
            assert abs(a - c) <= 2e-3 and abs(d - c) <= 2e-4, (k, a, b, c, d)      # SSIM
            continue
        worst["A"] = max(worst["A"], abs(a - c))
        worst["B"] = max(worst["B"], abs(b - c))
        worst["D"] = max(worst["D"], abs(d - c))
    with open(os.path.join(ROOT, "gpurun_out", "run_test_py_abcd.json"), "w") as f:
        json.dump({"worst_abs_psnr_float_delta_vs_reference_fp32_dB": worst, "metrics_compared": len(m_a)}, f)
    assert worst["A"] <= 0.01, worst             # our default mode: within 0.01 dB of the reference's fp32 run
    assert worst["D"] <= 0.001, worst            # our fp32-grade mode: within 0.001 dB
    assert worst["A"] <= worst["B"] + 1e-3, worst   # and never further from fp32 than the reference's own fp16 run
