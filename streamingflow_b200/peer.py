"""NVLink peer-memory arena of the row-sharded rollout (csrc/sf_peer.cuh, include/sf_b200.h ``sf_peer_*``).

Every rank allocates one device buffer through the C ABI, exports its 64-byte CUDA IPC handle, and maps the buffers of the other
ranks of the node; the exchange kernels then store halo rows / partial sums straight into the neighbours' memory and signal with
system-scope counters -- no NCCL call on the per-event path.  torch.distributed only carries the handles (once) and the barriers
around set-up and tear-down.

Arena layout (identical on every rank, offsets in bytes):
    0     flags      64 x u32 : [0] halo rows from the upper neighbour arrived, [1] from the lower neighbour,
                               [8 + r] all-reduce contribution of rank r arrived            (written by PEERS)
    256   local      64 x u32 : [0,1] push sequence / ticket, [2,3] pull sequence / ticket, [4] all-reduce sequence, [8] error word
    512   trace      3 x [64][4] u64 : %globaltimer stamps of the last 64 push / pull / all-reduce launches (SF_PEER_TRACE=1)
    6656  slots      [2][world][n_max] f32 : all-reduce contributions
    ...   recv_up    [2][halo_bytes]       : rows pushed by the upper neighbour (its band's last rows)
    ...   recv_dn    [2][halo_bytes]       : rows pushed by the lower neighbour
"""
from __future__ import annotations

import ctypes as C
import os

import torch
import torch.distributed as dist

from . import _lib as L

FLAG_UP, FLAG_DN, FLAG_AR0 = 0, 1, 8
LOC_PUSH, LOC_PULL, LOC_AR, LOC_ERR = 0, 2, 4, 8


def _align(n: int, a: int = 256) -> int:
    return (n + a - 1) // a * a


class PeerArena:
    def __init__(self, lib, rank: int, world: int, group, halo_bytes: int, n_max: int):
        if world > 8:
            raise L.SfError("peer-memory exchange is built for the <= 8 GPUs of one NVSwitch node")
        self.lib, self.rank, self.world, self.group = lib, rank, world, group
        self.n_max = int(n_max)
        self.halo_stride = _align(int(halo_bytes))
        self.off_trace, self.off_slots = 512, 512 + 3 * 64 * 4 * 8
        self.tracing = os.environ.get("SF_PEER_TRACE", "0") == "1"
        self.off_up = _align(self.off_slots + 2 * world * self.n_max * 4)
        self.off_dn = self.off_up + 2 * self.halo_stride
        self.nbytes = self.off_dn + 2 * self.halo_stride
        # set-up never raises before the ranks have agreed on its outcome (a rank that bailed out early would leave the others
        # hanging in a collective): allocate -> exchange handles -> map -> all-reduce(min) of "ok"
        own, handle, ok, why = C.c_void_p(), (C.c_ubyte * 64)(), True, ""
        if lib.sf_peer_alloc(self.nbytes, C.byref(own), handle) < 0:
            ok, why = False, "sf_peer_alloc: " + (lib.sf_last_error() or b"").decode()
        self.own = own.value if ok else None
        handles = [None] * world
        dist.all_gather_object(handles, bytes(handle) if ok else None, group=group)
        self.base = []
        if ok and all(h is not None for h in handles):
            for r in range(world):
                if r == rank:
                    self.base.append(self.own)
                    continue
                p, h = C.c_void_p(), (C.c_ubyte * 64).from_buffer_copy(handles[r])
                if lib.sf_peer_open(h, C.byref(p)) < 0:
                    ok, why = False, f"sf_peer_open(rank {r}): " + (lib.sf_last_error() or b"").decode()
                    break
                self.base.append(p.value)
        else:
            ok, why = False, why or "another rank could not allocate its arena"
        agreed = torch.tensor([1 if ok else 0], dtype=torch.int32, device=torch.device("cuda", torch.cuda.current_device()))
        dist.all_reduce(agreed, op=dist.ReduceOp.MIN, group=group)
        if int(agreed.item()) == 0:
            for r, b in enumerate(self.base):
                if r != rank:
                    lib.sf_peer_close(C.c_void_p(b))
            dist.barrier(group=group)
            if self.own is not None:
                lib.sf_peer_free(C.c_void_p(self.own))
            self.own = None
            raise L.SfError("peer-memory set-up failed: " + (why or "on another rank"))
        self.slots = (C.c_void_p * world)(*[b + self.off_slots for b in self.base])
        self.ar_flags = (C.c_void_p * world)(*[b + 4 * FLAG_AR0 for b in self.base])
        self._err_dev = self.words(256 + 4 * LOC_ERR, 1)
        self._err_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        dist.barrier(group=group)          # every arena is mapped everywhere before the first store into it

    # pointers ------------------------------------------------------------------------------------------------------------
    def flag(self, r: int, which: int) -> C.c_void_p:
        return C.c_void_p(self.base[r] + 4 * which)

    def local(self, which: int) -> C.c_void_p:
        return C.c_void_p(self.own + 256 + 4 * which)

    def recv_up(self, r: int) -> C.c_void_p:
        return C.c_void_p(self.base[r] + self.off_up)

    def recv_dn(self, r: int) -> C.c_void_p:
        return C.c_void_p(self.base[r] + self.off_dn)

    def trace(self, which: int):
        """Device pointer of the stamp ring of kernel `which` (0 push, 1 pull, 2 all-reduce), or NULL when tracing is off."""
        return C.c_void_p(self.own + self.off_trace + which * 64 * 4 * 8) if self.tracing else None

    def trace_report(self):
        """Per kernel kind, over the last <= 64 launches: mean microseconds from launch start to wait done / to launch end, and the
        mean gap between the end of a pull and the start of the next push (= the event's own compute)."""
        torch.cuda.synchronize()
        t = self.words(self.off_trace, 3 * 64 * 4 * 2).view(torch.int64).view(3, 64, 4).cpu().double()
        out = {}
        for k, name in enumerate(("push", "pull", "allreduce")):
            v = t[k][t[k][:, 2] > 0]
            if len(v):
                out[name] = dict(n=len(v), us_total=round(float((v[:, 2] - v[:, 0]).mean()) / 1e3, 2),
                                 us_wait=round(float((v[:, 1] - v[:, 0]).clamp(min=0).mean()) / 1e3, 2) if name != "push" else None)
        push, pull = t[0], t[1]
        gaps = [float(push[(i + 1) % 64, 0] - pull[i, 2]) for i in range(64) if pull[i, 2] > 0 and push[(i + 1) % 64, 0] > pull[i, 2]
                and push[(i + 1) % 64, 0] - pull[i, 2] < 5e6]
        if gaps:
            out["us_between_pull_end_and_next_push"] = round(sum(gaps) / len(gaps) / 1e3, 2)
        return out

    def words(self, offset: int, n: int) -> torch.Tensor:
        """n int32 words of the OWN arena at a byte offset, as a zero-copy tensor (debugging, the error word)."""
        class _Raw:
            pass
        raw = _Raw()
        raw.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (self.own + offset, False), "version": 2}
        return torch.as_tensor(raw, device=torch.device("cuda", torch.cuda.current_device()))

    # error word ------------------------------------------------------------------------------------------------------------
    _WHAT = {1: "halo rows of the upper neighbour", 2: "halo rows of the lower neighbour", 3: "an all-reduce contribution"}

    def _raise(self, code: int):
        raise L.SfError(f"rank {self.rank}: timed out waiting for {self._WHAT.get(code, code)} (peer-memory exchange)")

    def poll(self):
        """Non-blocking: raises if an EARLIER rollout's exchange timed out, then queues a copy of the error word behind the work
        already in the stream (read by the next poll / check)."""
        if int(self._err_host[0]) != 0:
            self._raise(int(self._err_host[0]))
        self._err_host.copy_(self._err_dev, non_blocking=True)

    def check(self):
        """Synchronises the device and raises if any exchange kernel timed out waiting for a neighbour."""
        torch.cuda.synchronize()
        code = int(self._err_dev[0].item())
        if code:
            self._raise(code)

    def close(self):
        if getattr(self, "own", None) is None:
            return
        torch.cuda.synchronize()
        dist.barrier(group=self.group)     # nobody still stores into a buffer that is about to be unmapped
        for r, b in enumerate(self.base):
            if r != self.rank:
                self.lib.sf_peer_close(C.c_void_p(b))
        dist.barrier(group=self.group)
        self.lib.sf_peer_free(C.c_void_p(self.own))
        self.own = None
