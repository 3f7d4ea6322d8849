"""Multi-GPU paths on real devices (need >= 2 GPUs: run with `gpurun --gpus 2`): the copy-engine clip
gather over peer-mapped memory against NCCL's all_gather, and the 4K spatial-tile exchange."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
needs2 = pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")


def _torchrun(n, script_args, timeout=600):
    port = 29600 + (os.getpid() % 300)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", str(port)] + script_args
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)
    return r


@needs2
def test_clip_gather_over_peer_memory_matches_nccl():
    r = _torchrun(2, [os.path.join(ROOT, "tests", "multi_gpu_worker.py"), "gather"])
    assert r.returncode == 0, (r.stdout + r.stderr)[-4000:]
    res = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert res["ok"] and res["steps"] == 6 and res["world"] == 2, res


@needs2
def test_spatial_tiles_neighbour_exchange_bit_exact():
    r = _torchrun(2, [os.path.join(ROOT, "tests", "multi_gpu_worker.py"), "tiles"])
    assert r.returncode == 0, (r.stdout + r.stderr)[-4000:]
    res = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert res["bit_exact"] and res["received_bytes"] <= res["ring_bytes"], res


def test_peer_group_single_rank_put_signal_wait_and_large_step_counters():
    """World 1 (no process group): the one-sided primitives against the local buffer, and signal values far
    beyond the 2^20-value window of the device pool (a long-running service's step counter), including a late,
    smaller value after the window has moved."""
    from bsvd_b200.peer import PeerGroup
    torch.cuda.set_device(0)
    pg = PeerGroup(1 << 16, 8)
    try:
        src = torch.arange(1024, dtype=torch.float32, device="cuda")
        pg.put(0, 256, src)
        pg.signal(0, 3, 5)
        pg.wait(3, 5)                                   # current stream waits for the flag behind the put
        got = pg.local_tensor(256, (1024,)).clone()
        torch.cuda.synchronize()
        assert torch.equal(got, src) and pg.read_flag(3) == 5
        for v in (7, (1 << 20) - 1, (1 << 20) + 9, (1 << 20) + 3, 3 * (1 << 20) + 1, 0xFFFFFFF0, 12):
            pg.signal(0, 2, v)
            torch.cuda.synchronize()
            assert pg.read_flag(2) == v, v
    finally:
        pg.close()
