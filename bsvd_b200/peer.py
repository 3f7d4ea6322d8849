"""Multi-GPU plumbing on top of the peer-memory C ABI (bsvd_peer_*, bsvd_b200/csrc/peer.cu).

One process per GPU on one node (torchrun).  torch.distributed (NCCL) is the control plane —
rendezvous, exchanging the CUDA-IPC handles, barriers, the timing reduction; the DATA moves with
copy-engine copies between peer-mapped buffers over NVLink, on a side stream, with flag words for
completion.  No SM is used by a transfer, so it overlaps the persistent conv kernels (one CTA per SM on
all 148 SMs) without taking SMs away from them — an NCCL all_gather issued next to them either waits
for the step to finish or, once resident, makes the stage kernels run in two waves.

    group = PeerGroup(nbytes, nflags)                # collective: every rank calls it
    gather = ClipGather(group_or_None, shape)        # ring of 2 slots of [world, *shape] per rank
    gather.put(y, step)                              # after the forward that produced y (any stream)
    g = gather.wait(step)                            # current stream waits; returns the [world, *shape] view

The reference's counterpart is DataParallel's scatter/gather (BasicSR/basicsr/models/base_model.py:74-75).

`PeerGroup(..., backend="shm")` is the same interface over POSIX shared memory between CPU processes
(numpy memmaps under /dev/shm, flags polled by the waiting process): it exists so that the host-side
protocols — ring slots, ready / free flags, strip geometry — run under world_size-2 gloo tests on a box
without GPUs (tests/test_tiling.py).  It is test infrastructure, never a fallback of the GPU path.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import capi


class _NoStream:
    """Stands for streams / events in the shared-memory backend (everything is program order there)."""
    cuda_stream = 0

    def wait_event(self, ev):
        pass

    def record(self, stream=None):
        pass

    def synchronize(self):
        pass


class PeerGroup:
    """A symmetric buffer (data + flags) on every rank, mapped into every other rank.
    backend "cuda": device memory shared through CUDA IPC, copy-engine transfers (the product path);
    backend "shm" : host shared memory between CPU processes (protocol tests only)."""

    def __init__(self, nbytes: int, nflags: int, group=None, backend: str = "cuda"):
        import torch.distributed as dist
        self.dist_group = group
        self.backend = backend
        if dist.is_available() and dist.is_initialized():
            self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        else:
            self.rank, self.world = 0, 1
        self.nbytes, self.nflags = int(nbytes), int(nflags)
        if backend == "shm":
            self._init_shm(group)
            return
        if backend != "cuda":
            raise ValueError("backend must be 'cuda' or 'shm'")
        self.lib = capi.load_library()
        self.device = torch.device("cuda", torch.cuda.current_device())
        # Every rank goes through the same collectives whatever happens locally (a rank that raised early
        # would leave the others in a mismatched collective): failures are recorded, agreed on with one
        # all_reduce, and then raised on ALL ranks.
        h = C.c_void_p()
        err = None
        if self.lib.bsvd_peer_create(self.rank, self.world, self.nbytes, self.nflags, C.byref(h)) != 0:
            err = self.lib.bsvd_last_error().decode("utf-8", "replace")
            h = None
        self._h = h
        import os
        if err is None and os.environ.get("BSVD_B200_PEER_FAIL") == str(self.rank):
            err = "forced failure (BSVD_B200_PEER_FAIL, tests of the fallback)"
        if self.world > 1:
            hb = self.lib.bsvd_peer_handle_bytes()
            mine = (C.c_ubyte * hb)()
            if err is None and self.lib.bsvd_peer_get_handle(self._h, mine) != 0:
                err = self.lib.bsvd_last_error().decode("utf-8", "replace")
            t = torch.tensor(list(mine), dtype=torch.uint8, device=self.device)
            allh = torch.empty((self.world, hb), dtype=torch.uint8, device=self.device)
            dist.all_gather_into_tensor(allh, t, group=group)
            okt = torch.tensor([0 if err else 1], device=self.device)
            dist.all_reduce(okt, op=dist.ReduceOp.MIN, group=group)
            if int(okt.item()) == 1:
                buf = (C.c_ubyte * (hb * self.world)).from_buffer_copy(bytes(allh.cpu().numpy().tobytes()))
                if self.lib.bsvd_peer_open(self._h, buf) != 0:
                    err = self.lib.bsvd_last_error().decode("utf-8", "replace")
                okt = torch.tensor([0 if err else 1], device=self.device)
                dist.all_reduce(okt, op=dist.ReduceOp.MIN, group=group)   # also: nobody puts before everybody has mapped
            if int(okt.item()) == 0:
                if self._h is not None:
                    self.lib.bsvd_peer_destroy(self._h)
                    self._h = None
                raise capi.BsvdError("peer memory could not be set up on every rank: " + (err or "another rank failed"))
        elif err:
            raise capi.BsvdError(err)
        self.data_ptr = int(self.lib.bsvd_peer_local_data(self._h))
        self.side = torch.cuda.Stream(device=self.device)      # transfers are enqueued here

    # ---- shared-memory backend (CPU processes; tests) ---------------------------------------------
    def _init_shm(self, group):
        import os
        import uuid
        import numpy as np
        import torch.distributed as dist
        self.device = torch.device("cpu")
        self.side = _NoStream()
        self._flag_bytes = (self.nflags * 4 + 4095) // 4096 * 4096
        name = f"/dev/shm/bsvd_peer_{os.getpid()}_{uuid.uuid4().hex[:8]}_{self.rank}"
        mm = np.memmap(name, dtype=np.uint8, mode="w+", shape=(self._flag_bytes + self.nbytes,))
        mm[:] = 0
        mm.flush()
        names = [None] * self.world
        if self.world > 1:
            dist.all_gather_object(names, name, group=group)
        else:
            names = [name]
        self._names = names
        self._maps = [mm if r == self.rank else np.memmap(names[r], dtype=np.uint8, mode="r+",
                                                           shape=(self._flag_bytes + self.nbytes,))
                      for r in range(self.world)]
        self._flags = [m[:self.nflags * 4].view(np.uint32) for m in self._maps]
        self._h = "shm"
        if self.world > 1:
            dist.barrier(group=group)

    def _shm_data(self, r):
        return self._maps[r][self._flag_bytes:]

    # ---- views ---------------------------------------------------------------------------------
    def local_tensor(self, offset: int, shape, dtype=torch.float32) -> torch.Tensor:
        """A torch view of the local data area (no copy; lives as long as the group)."""
        n = 1
        for s in shape:
            n *= int(s)
        esz = torch.empty((), dtype=dtype).element_size()
        if offset < 0 or offset + n * esz > self.nbytes:
            raise ValueError("view outside the peer buffer")
        if self.backend == "shm":
            import numpy as np
            npdt = {torch.float32: np.float32, torch.float16: np.float16, torch.uint8: np.uint8, torch.int32: np.int32}[dtype]
            arr = self._shm_data(self.rank)[offset:offset + n * esz].view(npdt).reshape(tuple(int(s) for s in shape))
            return torch.from_numpy(arr)

        class _Holder:        # __cuda_array_interface__ provider
            pass
        hold = _Holder()
        typestr = {torch.float32: "<f4", torch.float16: "<f2", torch.uint8: "|u1", torch.int32: "<i4"}[dtype]
        hold.__cuda_array_interface__ = {"shape": tuple(int(s) for s in shape), "typestr": typestr,
                                         "data": (self.data_ptr + offset, False), "version": 2}
        hold._keep = self
        return torch.as_tensor(hold, device=self.device)

    # ---- one-sided operations ------------------------------------------------------------------
    def put(self, dst_rank: int, dst_off: int, src: torch.Tensor, stream=None):
        assert src.is_contiguous()
        if self.backend == "shm":
            b = src.numpy().reshape(-1).view("uint8")
            self._shm_data(dst_rank)[dst_off:dst_off + b.size] = b
            return
        assert src.is_cuda
        st = (stream or self.side).cuda_stream
        capi.check(self.lib.bsvd_peer_put(self._h, dst_rank, dst_off, src.data_ptr(),
                                          src.numel() * src.element_size(), st))

    def put2d(self, dst_rank: int, dst_off: int, dst_pitch: int, src_ptr: int, src_pitch: int,
              width_bytes: int, rows: int, stream=None):
        st = (stream or self.side).cuda_stream
        capi.check(self.lib.bsvd_peer_put2d(self._h, dst_rank, dst_off, dst_pitch, src_ptr, src_pitch,
                                            width_bytes, rows, st))

    def put3d(self, dst_rank: int, dst_off: int, dst_pitch: int, dst_plane_rows: int, src,
              src_pitch: int, src_plane_rows: int, width_bytes: int, rows: int, planes: int, stream=None,
              src_off: int = 0):
        """planes x rows x width_bytes block.  `src` is a device pointer (int), or a tensor whose storage
        the block is cut from starting `src_off` bytes in."""
        if self.backend == "shm":
            import numpy as np
            sb = src.numpy().reshape(-1).view(np.uint8)
            dd = self._shm_data(dst_rank)
            for pl in range(planes):
                for r in range(rows):
                    so = src_off + (pl * src_plane_rows + r) * src_pitch
                    do = dst_off + (pl * dst_plane_rows + r) * dst_pitch
                    dd[do:do + width_bytes] = sb[so:so + width_bytes]
            return
        src_ptr = (src.data_ptr() + src_off) if torch.is_tensor(src) else int(src) + src_off
        st = (stream or self.side).cuda_stream
        capi.check(self.lib.bsvd_peer_put3d(self._h, dst_rank, dst_off, dst_pitch, dst_plane_rows, src_ptr,
                                            src_pitch, src_plane_rows, width_bytes, rows, planes, st))

    def signal(self, dst_rank: int, flag: int, value: int, stream=None):
        if self.backend == "shm":
            self._flags[dst_rank][flag] = value
            return
        st = (stream or self.side).cuda_stream
        capi.check(self.lib.bsvd_peer_signal(self._h, dst_rank, flag, value, st))

    def wait(self, flag: int, value: int, stream=None, timeout_s: float = 120.0):
        if self.backend == "shm":
            import time
            t0 = time.time()
            while int(self._flags[self.rank][flag]) < value:
                if time.time() - t0 > timeout_s:
                    raise TimeoutError(f"rank {self.rank}: flag {flag} never reached {value}")
                time.sleep(0.0005)
            return
        st = (stream or torch.cuda.current_stream(self.device)).cuda_stream
        capi.check(self.lib.bsvd_peer_wait(self._h, flag, value, st))

    def read_flag(self, flag: int) -> int:
        if self.backend == "shm":
            return int(self._flags[self.rank][flag])
        v = C.c_uint(0)
        capi.check(self.lib.bsvd_peer_read_flag(self._h, flag, C.byref(v)))
        return int(v.value)

    def current_stream(self):
        return _NoStream() if self.backend == "shm" else torch.cuda.current_stream(self.device)

    def new_event(self):
        return _NoStream() if self.backend == "shm" else torch.cuda.Event()

    def close(self):
        if getattr(self, "_h", None) is None:
            return
        if self.world > 1:
            import torch.distributed as dist
            if self.backend == "cuda":
                torch.cuda.synchronize(self.device)
            dist.barrier(group=self.dist_group)     # peers may still be writing into / reading from us
        if self.backend == "shm":
            import os
            self._flags = None
            self._maps = None
            try:
                os.unlink(self._names[self.rank])
            except OSError:
                pass
        else:
            if self.world == 1:
                torch.cuda.synchronize(self.device)
            self.lib.bsvd_peer_destroy(self._h)
        self._h = None


def gather_layout(world: int, slot_bytes: int, depth: int = 2):
    """Offsets of the gather ring inside a rank's symmetric buffer and its flag indices.

    data : [depth][world][slot_bytes]            slot (k, r) = clip of rank r for steps with step % depth == k
    flags: ready[k][r]  (index k*world + r)      rank r's clip of slot k has landed   (value = step + 1)
           free [k][r]  (index depth*world + k*world + r)   rank r has consumed slot k (value = step + 1)
    """
    return {"bytes": depth * world * slot_bytes, "nflags": 2 * depth * world,
            "data": lambda k, r: (k * world + r) * slot_bytes,
            "ready": lambda k, r: k * world + r,
            "free": lambda k, r: depth * world + k * world + r}


class ClipGather:
    """All ranks end up with every rank's output clip: an all-gather done with copy engines.

    put(y, step): the producing stream's work so far is awaited by the side stream, which then (1) waits
    until every peer has released slot step % depth (their `free` flag from step - depth), (2) copies y
    into slot (step % depth, rank) of EVERY rank (its own included) and (3) raises ready[k][rank] there.
    wait(step): the calling stream waits for all `world` ready flags of the slot and gets the view.
    release(step): the calling stream's reads of the slot are done — tell the producers.
    """

    def __init__(self, shape, dtype=torch.float32, depth: int = 2, group=None, backend: str = "cuda"):
        import torch.distributed as dist
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.shape, self.dtype, self.depth = tuple(int(s) for s in shape), dtype, depth
        n = 1
        for s in self.shape:
            n *= s
        self.slot_bytes = n * torch.empty((), dtype=dtype).element_size()
        self.lay = gather_layout(world, self.slot_bytes, depth)
        self.pg = PeerGroup(self.lay["bytes"], self.lay["nflags"], group, backend=backend)
        self.world, self.rank = self.pg.world, self.pg.rank
        self.ev = [self.pg.new_event() for _ in range(depth)]
        self.ev_rel = self.pg.new_event()
        self.ev_put = [self.pg.new_event() for _ in range(depth)]

    def put(self, y: torch.Tensor, step: int):
        assert tuple(y.shape) == self.shape and y.dtype == self.dtype and y.is_contiguous()
        pg, k = self.pg, step % self.depth
        ev = self.ev[k]
        ev.record(pg.current_stream())
        pg.side.wait_event(ev)
        if step >= self.depth:
            for r in range(self.world):              # slot k was last used by step - depth
                pg.wait(self.lay["free"](k, r), step - self.depth + 1, stream=pg.side)
        for i in range(self.world):                  # start with the next rank: spreads the NVSwitch load
            r = (self.rank + 1 + i) % self.world
            pg.put(r, self.lay["data"](k, self.rank), y, stream=pg.side)
            pg.signal(r, self.lay["ready"](k, self.rank), step + 1, stream=pg.side)
        if y.is_cuda:
            y.record_stream(pg.side)
        done = self.ev_put[k]
        done.record(pg.side)          # y has been read: its owner may overwrite it after this event
        return done

    def wait(self, step: int) -> torch.Tensor:
        k = step % self.depth
        for r in range(self.world):
            self.pg.wait(self.lay["ready"](k, r), step + 1)
        return self.pg.local_tensor(self.lay["data"](k, 0), (self.world,) + self.shape, self.dtype)

    def release(self, step: int):
        """Reads of slot `step` enqueued on the current stream so far are the last ones."""
        pg, k = self.pg, step % self.depth
        self.ev_rel.record(pg.current_stream())
        pg.side.wait_event(self.ev_rel)
        for r in range(self.world):
            pg.signal(r, self.lay["free"](k, self.rank), step + 1, stream=pg.side)

    def close(self):
        self.pg.close()
