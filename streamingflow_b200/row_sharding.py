"""Row sharding of ONE large BEV grid across the GPUs of a node (BASELINE config 5).

Each rank owns a contiguous band of latent rows and keeps HALO = 12 extra rows above and below it:
  * the dual-GRU cell sees the state within 7 rows and its x input within 6 rows of an output pixel (6x 3x3 + 7x7 + 3x3 in
    series), the prior network p_model sees the new state within 5 rows; an event therefore yields a correct
    sampled input x' on a band only if (state, x) were correct within 7 + 5 = 12 rows of it;
  * per event every rank runs the UNCHANGED stage kernels on its local image (band + halos), then swaps 12 boundary rows of
    the new state (fp32 master + bf16 operand planes) and of x' with its two neighbours:
    one exchange per event instead of one per conv stage, paid for with (24 / band) redundant rows of compute;
  * the two squeeze-excite layers need WHOLE-image channel means: each rank reduces over its own band only
    (sf_plan_se_reduce_totals with a pixel window), the [B, 2C] partial sums are summed over the ranks, then sf_plan_se_finish
    folds the scales into the consuming convs' weights;
  * transport (default "peer"): the exchanges are KERNELS over NVLink peer memory (csrc/sf_peer.cuh, peer.py): sf_halo_push
    stores the boundary rows straight into the neighbours' receive buffers and releases their arrival counters, sf_halo_pull
    waits for both neighbours and fills the halos, sf_peer_allreduce_f32 sums the SE totals through per-rank slots -- three
    launches per event, no NCCL call, and the whole rollout (stages + exchanges) replays as ONE CUDA graph.  "nccl" keeps the
    send/recv + all-reduce calls (graph segments between them, or captured inside one graph with SF_ROWSHARD_GRAPH=whole).
Zero padding at the true image border comes from TMA's out-of-bounds fill on the first / last rank; on interior ranks the
local image's outer rows are halo data and whatever the fill corrupts stays inside the discarded part of the halo.

Noise must be identical on all ranks (it is indexed by GLOBAL pixel): rank 0 broadcasts a seed, every rank draws the full
tensor from a generator seeded with it and keeps its rows.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import _lib as L
from .rollout import compile_rollout
from .schedule import plan_sample
from .sharding import shard_bounds

HALO = 12          # 7 (cell w.r.t. state) + 5 (p_model w.r.t. the new state)


def local_window(h: int, rank: int, world: int, halo: int = HALO) -> Tuple[int, int, int, int]:
    """(own_lo, own_hi, lo, hi): the rank's band [own_lo, own_hi) and its local image [lo, hi) of global latent rows."""
    own_lo, own_hi = shard_bounds(h, rank, world)
    if world > 1 and own_hi - own_lo < halo:
        raise ValueError(f"band of {own_hi - own_lo} rows is thinner than the {halo}-row halo; use fewer ranks")
    return own_lo, own_hi, max(0, own_lo - halo), min(h, own_hi + halo)


def exchange_halo_rows(ts, own_lo: int, own_hi: int, lo: int, hi: int, rank: int, world: int, group=None, halo: int = HALO):
    """ts: one tensor or a list of tensors [B, hi - lo, W, C] (NHWC local images, any dtypes).  Sends the band's first / last
    ``halo`` rows of ALL of them to the upper / lower neighbour in ONE message per direction (byte-packed) and fills the
    local halos with the neighbours' rows.  Works with NCCL (CUDA tensors) and gloo (CPU tensors)."""
    if world == 1:
        return
    if isinstance(ts, torch.Tensor):
        ts = [ts]
    a, b = own_lo - lo, own_hi - lo              # band inside the local image
    peer = (lambda r: dist.get_global_rank(group, r)) if group is not None else (lambda r: r)

    def pack(rows):
        return torch.cat([t[:, rows].contiguous().view(torch.uint8).reshape(-1) for t in ts])

    def unpack(buf, rows):
        off = 0
        for t in ts:
            dst = t[:, rows]
            n = dst.numel() * dst.element_size()
            dst.copy_(buf[off:off + n].view(t.dtype).view(dst.shape))
            off += n

    ops, recvs = [], []
    if rank > 0:
        send_up = pack(slice(a, a + halo))
        recv_up = torch.empty_like(send_up)
        ops += [dist.P2POp(dist.isend, send_up, peer(rank - 1), group), dist.P2POp(dist.irecv, recv_up, peer(rank - 1), group)]
        recvs.append((slice(a - halo, a), recv_up))
    if rank < world - 1:
        send_dn = pack(slice(b - halo, b))
        recv_dn = torch.empty_like(send_dn)
        ops += [dist.P2POp(dist.isend, send_dn, peer(rank + 1), group), dist.P2POp(dist.irecv, recv_dn, peer(rank + 1), group)]
        recvs.append((slice(b, b + halo), recv_dn))
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    for rows, buf in recvs:
        unpack(buf, rows)


def _recorded_event():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


class RowShardedOde:
    """Integrates the latent rollout of a NNFOwithBayesianJumps module with the grid's rows split over the ranks of ``group``."""

    def __init__(self, ode, h: int, w: int, batch: int, group=None, use_graphs: bool = True, transport: Optional[str] = None):
        """transport: how the per-event exchanges travel -- "peer": kernels that store into the neighbours' memory over NVLink
        (csrc/sf_peer.cuh; the whole rollout incl. the exchanges is ONE CUDA graph), "nccl": send/recv + all-reduce calls.
        Default: $SF_ROWSHARD_TRANSPORT or "peer", falling back to "nccl" (on all ranks together) if the arenas cannot be mapped."""
        import os

        from .engine import OdeEngine

        self.ode, self.h, self.w, self.B, self.group = ode, h, w, batch, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.own_lo, self.own_hi, self.lo, self.hi = local_window(h, self.rank, self.world)
        dev = next(ode.parameters()).device
        if ode.training:
            raise L.SfError("eval mode only")
        # squeeze-excite: band totals (sf_plan_se_reduce_totals) -> all-reduce across the ranks -> scales folded into the consuming
        # convs' weights (sf_plan_se_finish), like the single-GPU engine: no pass over the activation tensor
        self.eng = OdeEngine(ode._hot_state_dict(), "", self.hi - self.lo, w, batch, ode.precision, dev, se_fold=True)
        P, CH = L.SE_MAX_PARTIALS, 2 * self.eng.C
        off = 2 * batch * P * CH
        self.se_totals = [self.eng.se_scratch[off + k * batch * CH: off + (k + 1) * batch * CH].view(batch, CH) for k in range(2)]
        self.device = dev
        self.launches = 0
        self.use_graphs = bool(use_graphs)      # replay captured graphs (eager stage launches if False)
        self.graph_mode = "eager"
        self.transport = (transport or os.environ.get("SF_ROWSHARD_TRANSPORT", "peer")) if self.world > 1 else "none"
        if self.transport not in ("peer", "nccl", "none"):
            raise L.SfError(f"unknown transport {self.transport!r}")
        self.arena = None
        self.peer_error = None
        self.profile = None          # a list here receives (label, CUDA event) marks from integrate()
        if self.transport == "peer":
            from .peer import PeerArena

            try:
                self.arena = PeerArena(self.eng.lib, self.rank, self.world, group, self._halo_bytes_max(), batch * CH)
            except L.SfError as e:          # raised on every rank together (the set-up agrees on its outcome first)
                self.peer_error, self.transport = str(e)[:200], "nccl"

    def _halo_bytes_max(self) -> int:
        return self.B * HALO * self.w * self.eng.C * (4 + 4 * 2)        # fp32 state + up to 4 bf16 planes (state hi/lo, x hi/lo)

    # ------------------------------------------------------------------ noise shared by all ranks
    def draw_noise(self, n: int) -> torch.Tensor:
        seed = torch.randint(0, 2 ** 62, (1,), dtype=torch.int64, device=self.device)
        if self.world > 1:
            dist.broadcast(seed, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0, group=self.group)
        g = torch.Generator(device=self.device).manual_seed(int(seed.item()))
        rows = self.hi - self.lo
        eps = torch.empty((max(n, 1), self.eng.C, rows, self.w), dtype=torch.float32, device=self.device)
        full = torch.empty((self.eng.C, self.h, self.w), dtype=torch.float32, device=self.device)
        for i in range(n):
            full.normal_(generator=g)
            eps[i].copy_(full[:, self.lo:self.hi])
        return eps

    _KERNEL_OPS = ("stage", "se_reduce", "se_apply", "pack", "unpack", "push", "pull", "peer_allreduce")      # ops that are launches of OUR kernels

    # ------------------------------------------------------------------ one engine event with the collectives in place
    def _run_event(self, ev, tdev):
        """Eager form: the event's op list (_event_ops) executed in order."""
        self._ensure_exchange_buffers()
        for op in self._event_ops(ev):
            self._run_op(op, ev, tdev)
            self.launches += op[0] in self._KERNEL_OPS

    # ------------------------------------------------------------------ the same event as replayed CUDA-graph segments
    def _event_ops(self, ev):
        """Flat op list of one event; the NCCL calls ('allreduce', 'p2p') split it into capturable segments."""
        eng = self.eng
        ops = []
        if ev.run_cell:
            ops += [("stage", st) for st in eng.cell_slots[ev.kind]]
        if ev.run_prior:
            for item in eng.prior_items:
                if item < L.SE_ITEM_BASE:
                    ops.append(("stage", item))
                else:
                    w = item % 1000          # SE_ITEM_BASE + which (apply form) or SE_FOLD_ITEM_BASE + which (folded form)
                    reduce = [("peer_allreduce" if self.transport == "peer" else "allreduce", w)] if self.world > 1 else []
                    ops += [("se_reduce", w)] + reduce + [("se_apply", w)]
        if self.transport == "peer":
            ops += [("push", 0), ("pull", 0)]
        elif self.world > 1:
            ops += [("pack", 0), ("p2p", 0), ("unpack", 0)]
        return ops

    def _halo_tensors(self, ev):
        from .engine import BUF_S0, BUF_X

        eng, s = self.eng, ev.s_out
        xs = [eng.state32[s]] + [p for p in eng.act[BUF_S0 + s] if p is not None]
        if ev.run_prior:
            xs += [p for p in eng.act[BUF_X] if p is not None]
        return xs

    def _halo_args(self, ev):
        key = (ev.s_out, ev.run_prior)
        cache = self.__dict__.setdefault("_halo_cache", {})
        if key not in cache:
            ts = self._halo_tensors(ev)
            n = len(ts)
            cache[key] = ((C.c_void_p * n)(*[t.data_ptr() for t in ts]),
                          (C.c_longlong * n)(*[t.stride(0) * t.element_size() for t in ts]),
                          (C.c_longlong * n)(*[t.stride(1) * t.element_size() for t in ts]), n)
        return cache[key]

    def _halo_peer(self, ev, push: bool):
        """Peer transport.  push: the band's boundary rows of all halo tensors stored straight into the neighbours' receive
        buffers + their arrival counters released (sf_halo_push); pull: wait for both neighbours' rows, unpack them into the
        local halos (sf_halo_pull).  One launch each."""
        from . import peer as P

        ptrs, bstride, rbytes, n = self._halo_args(ev)
        ar, r, lib = self.arena, self.rank, self.eng.lib
        a, b = self.own_lo - self.lo, self.own_hi - self.lo
        up, dn = r > 0, r < self.world - 1
        if push:
            L.check(lib.sf_halo_push(ptrs, bstride, rbytes, n, self.B, HALO,
                                     ar.recv_dn(r - 1) if up else None, a, ar.recv_up(r + 1) if dn else None, b - HALO, ar.halo_stride,
                                     ar.flag(r - 1, P.FLAG_DN) if up else None, ar.flag(r + 1, P.FLAG_UP) if dn else None,
                                     ar.local(P.LOC_PUSH), ar.trace(0), self.eng._stream()), "sf_halo_push")
        else:
            L.check(lib.sf_halo_pull(ptrs, bstride, rbytes, n, self.B, HALO,
                                     ar.recv_up(r) if up else None, a - HALO, ar.recv_dn(r) if dn else None, b, ar.halo_stride,
                                     ar.flag(r, P.FLAG_UP) if up else None, ar.flag(r, P.FLAG_DN) if dn else None,
                                     ar.local(P.LOC_PULL), ar.local(P.LOC_ERR), ar.trace(1), self.eng._stream()), "sf_halo_pull")

    def _halo_launch(self, ev, to_flat: bool):
        """NCCL transport: all halo tensors of the event, both directions, in ONE kernel (sf_halo_copy): pack the band's boundary
        rows into the send buffers, or unpack the received rows into the halos."""
        ptrs, bstride, rbytes, n = self._halo_args(ev)
        a, b = self.own_lo - self.lo, self.own_hi - self.lo
        up, dn = self.rank > 0, self.rank < self.world - 1
        if to_flat:
            fa, ra, fb, rb = (self.send_up if up else None), a, (self.send_dn if dn else None), b - HALO
        else:
            fa, ra, fb, rb = (self.recv_up if up else None), a - HALO, (self.recv_dn if dn else None), b
        ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        L.check(self.eng.lib.sf_halo_copy(ptrs, bstride, rbytes, n, self.B, HALO, ptr(fa), ra, ptr(fb), rb, int(to_flat),
                                          self.eng._stream()), "sf_halo_copy")

    def _run_op(self, op, ev, tdev):
        eng, lib = self.eng, self.eng.lib
        stream = eng._stream()
        kind, arg = op
        n, ch = ev.n_active, 2 * eng.C
        a, b = self.own_lo - self.lo, self.own_hi - self.lo
        if kind == "stage":
            L.check(lib.sf_plan_run_stage(eng.plan, arg, C.byref(ev), tdev.data_ptr(), stream), "run_stage")
        elif kind == "se_reduce":         # this rank's band totals per (sample, channel) -> self.se_totals[arg][:n]
            L.check(lib.sf_plan_se_reduce_totals(eng.plan, arg, C.byref(ev), tdev.data_ptr(), a * self.w, b * self.w, stream), "se_reduce_totals")
        elif kind == "allreduce":
            dist.all_reduce(self.se_totals[arg][:n], op=dist.ReduceOp.SUM, group=self.group)
        elif kind == "se_apply":          # totals -> scales (whole-image mean) -> folded into the consumers' weights
            L.check(lib.sf_plan_se_finish(eng.plan, arg, C.byref(ev), tdev.data_ptr(), C.c_float(1.0 / (self.h * self.w)), stream), "se_finish")
        elif kind == "pack":
            self._halo_launch(ev, True)
        elif kind == "p2p":
            peer = (lambda r: dist.get_global_rank(self.group, r)) if self.group is not None else (lambda r: r)
            nbytes = sum(t[:, :HALO].numel() * t.element_size() for t in self._halo_tensors(ev))
            ops = []
            if self.rank > 0:
                ops += [dist.P2POp(dist.isend, self.send_up[:nbytes], peer(self.rank - 1), self.group),
                        dist.P2POp(dist.irecv, self.recv_up[:nbytes], peer(self.rank - 1), self.group)]
            if self.rank < self.world - 1:
                ops += [dist.P2POp(dist.isend, self.send_dn[:nbytes], peer(self.rank + 1), self.group),
                        dist.P2POp(dist.irecv, self.recv_dn[:nbytes], peer(self.rank + 1), self.group)]
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        elif kind == "unpack":
            self._halo_launch(ev, False)
        elif kind == "peer_allreduce":    # the band totals summed over the ranks by ONE kernel storing into every rank's arena
            from . import peer as P

            ar = self.arena
            L.check(lib.sf_peer_allreduce_f32(C.c_void_p(self.se_totals[arg].data_ptr()), n * ch, ar.n_max, self.rank, self.world, ar.slots,
                                              ar.ar_flags, ar.local(P.LOC_AR), ar.local(P.LOC_ERR), ar.trace(2), stream), "sf_peer_allreduce_f32")
        elif kind == "push":
            self._halo_peer(ev, True)
        elif kind == "pull":
            self._halo_peer(ev, False)

    def _ensure_exchange_buffers(self):
        eng = self.eng
        if self.transport == "nccl" and getattr(self, "send_up", None) is None:
            nbytes = self._halo_bytes_max()
            mk = lambda: torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self.send_up, self.recv_up, self.send_dn, self.recv_dn = mk(), mk(), mk(), mk()


    def _run_rollout_graphed(self, evs, tdev, key):
        """Replays the rollout as CUDA graphs.  Preferred: ONE graph for the whole rollout with the NCCL calls (two [B, 2C]
        all-reduces and one send/recv group per event) captured inside it -- a single host launch per rollout, no stream
        hand-over between the stage kernels and the collectives (at 8 ranks the ~60 kernels of an event on a 74-row band are
        shorter than their launch overhead, and each eager NCCL call costs a host round trip).  If the installed NCCL / torch
        refuses to capture collectives, falls back to graph segments between the NCCL calls (4 graph launches + 3 NCCL calls per
        event on the host).  The whole-rollout capture is opt-in (SF_ROWSHARD_GRAPH=whole) until it is proven on this NCCL build."""
        import os

        self._ensure_exchange_buffers()
        cache = self.__dict__.setdefault("_graph_cache", {})
        ent = cache.get(key)
        if ent is None or ent["gen"] != self.eng.alloc_gen:
            if len(cache) >= 4:
                cache.clear()
            torch.cuda.synchronize(self.device)
            n_launch = sum(1 for ev in evs for o in self._event_ops(ev) if o[0] in self._KERNEL_OPS)
            prog, mode = None, "segments"
            if self.transport == "peer":
                # every exchange is a plain kernel: the whole rollout is one graph, nothing for the host to do per event
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, capture_error_mode="thread_local"):
                    for ev in evs:
                        for op in self._event_ops(ev):
                            self._run_op(op, ev, tdev)
                prog, mode = [("graph", g, None)], "whole rollout incl. peer-memory exchanges (one graph)"
                dist.barrier(group=self.group)      # capture times differ per rank: start the first replay together (the waits are bounded)
            elif self.world > 1 and os.environ.get("SF_ROWSHARD_GRAPH", "segments") == "whole" and self.__dict__.get("_whole_graph_ok", True):
                try:
                    # warm the communicator outside the capture (the first collective of a process group initialises it)
                    # and so does the first send/recv with each neighbour; both only touch scratch buffers
                    dist.all_reduce(self.se_totals[0][:1], group=self.group)
                    self._run_op(("p2p", 0), evs[0], tdev)
                    torch.cuda.synchronize(self.device)
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, capture_error_mode="thread_local"):
                        for ev in evs:
                            for op in self._event_ops(ev):
                                self._run_op(op, ev, tdev)
                    prog, mode = [("graph", g, None)], "whole rollout incl. NCCL"
                except Exception as e:      # capture of collectives unsupported here: remember and use segments
                    self.__dict__["_whole_graph_ok"] = False
                    self.__dict__["_whole_graph_error"] = str(e)[:200]
                    torch.cuda.synchronize(self.device)
            if prog is None:
                prog = []                                    # ("graph", g) | ("nccl", op, ev)
                for ev in evs:
                    seg = []
                    for op in self._event_ops(ev) + [("flush", 0)]:
                        if op[0] in ("allreduce", "p2p", "flush"):
                            if seg:
                                g = torch.cuda.CUDAGraph()
                                with torch.cuda.graph(g, capture_error_mode="thread_local"):     # the NCCL watchdog thread keeps polling events
                                    for o in seg:
                                        self._run_op(o, ev, tdev)
                                prog.append(("graph", g, None))
                                seg = []
                            if op[0] != "flush":
                                prog.append(("nccl", op, ev))
                        else:
                            seg.append(op)
            ent = cache[key] = dict(prog=prog, gen=self.eng.alloc_gen, launches=n_launch, mode=mode)
        self.graph_mode = ent["mode"]
        for kind, x, ev in ent["prog"]:
            if kind == "graph":
                x.replay()
            else:
                self._run_op(x, ev, tdev)
        self.launches += ent["launches"]

    def release_graphs(self):
        """Destroys the captured graphs.  A graph that contains NCCL kernels must be gone BEFORE its process group is destroyed
        (``dist.destroy_process_group()`` otherwise blocks forever -- observed with NCCL 2.28.9): call this (or drop the
        object) before tearing the group down."""
        import gc

        self.__dict__.pop("_graph_cache", None)
        gc.collect()
        torch.cuda.synchronize(self.device)
        if self.arena is not None:          # peer transport: a timed-out exchange surfaces here at the latest; then unmap
            arena, self.arena = self.arena, None
            try:
                arena.check()
            finally:
                arena.close()

    def integrate(self, hx_obs: torch.Tensor, obs_counts: Sequence[int], times, targets, delta_t: float,
                  noise: Optional[torch.Tensor] = None, local_rows: bool = False):
        """hx_obs: the [sum(obs_counts), C, h, w] encoded observations.  By default every rank passes the FULL grid and only its
        rows are used; with ``local_rows=True`` hx_obs (and noise) already hold just this rank's local image, rows [lo, hi) --
        how a sharded producer would deliver them (no per-rollout slicing pass).  noise: optional tape [n, C, rows, w] (tests,
        benchmarks).  Returns the selected latents restricted to this rank's band: [B, T, C, own rows, w], plus the Rollout.
        Nothing in here synchronises the host with the device: schedules, event tables and gather indices are cached per
        schedule, so the host prepares rollout i + 1 while the device runs rollout i."""
        ode, eng = self.ode, self.eng
        B = len(obs_counts)
        prof = self.profile                          # optional: list that receives (label, cuda event) marks of this call
        mark = (lambda label: prof.append((label, _recorded_event()))) if prof is not None else (lambda label: None)
        mark("start")
        if self.transport == "peer" and self.arena is None:          # released earlier: map the arenas again
            from .peer import PeerArena

            self.arena = PeerArena(eng.lib, self.rank, self.world, self.group, self._halo_bytes_max(), self.B * 2 * eng.C)
            eng.alloc_gen += 1
        key = (B, tuple(obs_counts), tuple(tuple(float(x) for x in t) for t in times),
               tuple(tuple(float(x) for x in t) for t in targets), float(delta_t), ode.solver, bool(ode.impute), eng.precision,
               bool(ode.skip_dead_prior))
        sched = self.__dict__.setdefault("_schedules", {}).get(key)
        if sched is None:
            if len(self._schedules) >= 8:
                self._schedules.clear()
            plans = [plan_sample(times[b], targets[b], delta_t, ode.use_variable_ode_step, ode.solver) for b in range(B)]
            base = np.concatenate([[0], np.cumsum(obs_counts)[:-1]]).astype(int).tolist()
            ro = compile_rollout(plans, base, ode.solver, bool(ode.impute), skip_dead_prior=ode.skip_dead_prior)
            flat = torch.tensor([s for slots in ro.out_slots for s in slots], dtype=torch.int32, device=self.device)
            sched = self._schedules[key] = (ro, flat)
        ro, flat = sched
        rows = slice(None) if local_rows else slice(self.lo, self.hi)
        if local_rows and hx_obs.shape[2] != self.hi - self.lo:
            raise L.SfError(f"local_rows: expected {self.hi - self.lo} rows, got {hx_obs.shape[2]}")
        eng.bind_observations(hx_obs[:, :, rows].contiguous())
        eng.zero_state(0)
        eng.ensure_path_slots(ro.n_path)
        eps = noise[:, :, rows] if noise is not None else self.draw_noise(ro.n_eps)
        mark("inputs bound")
        if self.use_graphs:
            # static noise buffer and event table: the captured segments hold their addresses
            if getattr(self, "_eps_static", None) is None or self._eps_static.shape[0] < max(ro.n_eps, 1):
                self._eps_static = torch.empty((max(ro.n_eps, 1),) + tuple(eps.shape[1:]), dtype=torch.float32, device=self.device)
                eng.alloc_gen += 1
            k = min(eps.shape[0], ro.n_eps)          # a caller's tape may hold more slots than this schedule consumes
            self._eps_static[:k].copy_(eps[:k])
            eng.bind_eps(self._eps_static)
            plan = self.__dict__.setdefault("_tables", {}).get(key)
            if plan is None:
                table, evs = eng.build_table(ro.events)          # one upload for the whole rollout
                plan = self._tables[key] = (evs, eng.upload_table(table))
            evs, tdev = plan
            mark("noise bound")
            self._run_rollout_graphed(evs, tdev, key)
        else:
            eng.bind_eps(eps[: max(ro.n_eps, 1)].contiguous())
            table, evs = eng.build_table(ro.events)          # one upload for the whole rollout
            tdev = eng.upload_table(table)
            mark("noise bound")
            for ev in evs:
                self._run_event(ev, tdev)
        mark("events done")
        if self.arena is not None:
            self.arena.poll()
        T = len(targets[0])
        sel = eng.unpack_path(flat).view(B, T, eng.C, self.hi - self.lo, self.w)
        ro.launches = self.launches
        band = sel[:, :, :, self.own_lo - self.lo:self.own_hi - self.lo].contiguous()
        mark("outputs gathered")
        return band, ro

    def gather_rows(self, band: torch.Tensor) -> torch.Tensor:
        """all_gather of the bands along the row axis -> the full [B, T, C, h, w] tensor on every rank."""
        if self.world == 1:
            return band
        sizes = [shard_bounds(self.h, r, self.world) for r in range(self.world)]
        parts = [torch.empty(band.shape[:3] + (b - a, self.w), dtype=band.dtype, device=band.device) for a, b in sizes]
        if len({b - a for a, b in sizes}) == 1:
            dist.all_gather(parts, band, group=self.group)
        else:
            for r, buf in enumerate(parts):
                if r == self.rank:
                    buf.copy_(band)
                dist.broadcast(buf, src=dist.get_global_rank(self.group, r) if self.group is not None else r, group=self.group)
        return torch.cat(parts, dim=3)
