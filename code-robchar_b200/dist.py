"""Multi-GPU sharding of the sweep: one process per GPU.

Every (sigma, controller, draw) evaluation is independent and statistics are per (sigma, controller); only a
global ranking needs all controllers' statistics.  Two shardings:

* **controllers** (default): each rank runs the sweep on its contiguous controller block; the only exchange is
  "every rank gets every rank's [15][S][C_local] block".  On GPUs this is `PeerStatsExchange`: each rank PUSHES
  its block into the column range of every peer's [15][S][C_total] tensor over NVLink with the copy engines, on a
  side stream, underneath the evolution kernel of the next step (csrc/rc_peer.cu) — no pad, no concatenation, no
  collective kernel competing for SMs.  `all_gather_stats` is the torch.distributed form of the same exchange
  (NCCL, or gloo in the CPU tests of the host logic).
* **draws** (fewer controllers than ranks — the single-controller sweeps of
  gen_fig_8_arim_fcall_scaling.py:121-132 and qnewton.py:447-455): ranks take disjoint draw ranges (whole merge
  blocks, see rc_draw_shard_range), the per-rank block results are all-gathered and merged in a fixed order, so the
  statistics equal the single-GPU run bit for bit.

Philox counters use GLOBAL (sigma, controller, draw) indices, so results do not depend on the world size.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as td

MERGE_BLOCKS = 8          # csrc/rc_fidelity.cu: fixed merge order of the chunk partials of a segment
PART_DOUBLES = 17         # csrc/rc_stats.cuh: one streaming partial


def shard_bounds(n: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of `n` items for `rank`; the first n % world ranks get one extra."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _world(group=None) -> tuple[int, int]:
    if td.is_available() and td.is_initialized():
        return td.get_world_size(group), td.get_rank(group)
    return 1, 0


def all_gather_stats(local: torch.Tensor, C_total: int, group=None, out: torch.Tensor | None = None) -> torch.Tensor:
    """local: [K][S][C_local] on this rank -> [K][S][C_total] on every rank (controller axis gathered in rank
    order) through torch.distributed.  Equal shards are gathered straight into `out` viewed as
    [world][K][S][C_local] blocks and re-laid out with one strided copy; uneven shards are padded to the largest
    block for the collective.  `out` (optional, [K][S][C_total]) is reused across calls."""
    world, rank = _world(group)
    if world == 1:
        return local
    sizes = [shard_bounds(C_total, world, r)[1] - shard_bounds(C_total, world, r)[0] for r in range(world)]
    cmax = max(sizes)
    K, S = local.shape[0], local.shape[1]
    send = local.contiguous()
    if local.shape[2] != cmax:
        send = torch.zeros((K, S, cmax), dtype=local.dtype, device=local.device)
        send[:, :, :local.shape[2]] = local
    blocks = torch.empty((world, K, S, cmax), dtype=local.dtype, device=local.device)
    td.all_gather_into_tensor(blocks.view(world * K, S, cmax), send, group=group)   # rank-major blocks
    if out is None:
        out = torch.empty((K, S, C_total), dtype=local.dtype, device=local.device)
    lo = 0
    for r in range(world):
        out[:, :, lo:lo + sizes[r]] = blocks[r, :, :, :sizes[r]]
        lo += sizes[r]
    return out


class _DeviceArray:
    """A raw device allocation seen as an array through __cuda_array_interface__ (torch.as_tensor wraps it without
    copying)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class PeerExchangeUnavailable(RuntimeError):
    """CUDA IPC mapping of the peers' exchange buffers failed on at least one rank (agreed collectively)."""


class PeerStatsExchange:
    """Exchange of the per-rank statistics blocks by direct NVLink writes (csrc/rc_peer.cu).

    Every rank owns one buffer holding `depth` (=2) tensors [rows][C_total] plus a flag array; its peers map it
    through CUDA IPC.  push(seq) copies this rank's dense block [rows][C_local] into the column range of the
    tensor `seq % depth` of every rank with copy-engine transfers on a side stream and then publishes `seq` in
    every rank's flag array; gathered(seq) makes the current stream wait until all ranks have published `seq` and
    returns this rank's tensor.  Nothing here runs on the SMs the evolution kernel occupies, and nothing blocks the
    host: step k's exchange rides underneath step k+1's evolution.

    Contract: the tensor returned by gathered(seq) stays valid until push(seq + depth) is issued on any rank, and a
    rank must have enqueued its readers of gathered(seq) on the stream it later uses for the compute of step
    seq + 1 (push(seq + 1) is ordered after them) — i.e. consume a step's result before launching the next step."""

    FLAG_BYTES = 4096

    def __init__(self, rows: int, C_total: int, group=None, depth: int = 2, timeout_s: float = 20.0):
        from . import engine
        from ._lib import check, lib
        self._check, self._lib = check, lib()
        self.dev = engine.require_cuda()
        self.world, self.rank = _world(group)
        if self.world > 64:
            raise ValueError("PeerStatsExchange supports up to 64 ranks")
        self.rows, self.C_total, self.depth, self.timeout_s = int(rows), int(C_total), int(depth), float(timeout_s)
        self.lo, self.hi = shard_bounds(C_total, self.world, self.rank)
        self.tensor_bytes = self.rows * self.C_total * 8
        nbytes = self.FLAG_BYTES + self.depth * self.tensor_bytes
        ptr, handle = C.c_void_p(0), (C.c_ubyte * 64)()
        check(self._lib.rc_peer_alloc(nbytes, C.byref(ptr), handle))
        self.base = int(ptr.value)
        handles = [None] * self.world
        if self.world > 1:
            td.all_gather_object(handles, bytes(handle), group=group)
        else:
            handles[0] = bytes(handle)
        self.peer_base = [0] * self.world
        self.peer_base[self.rank] = self.base
        err = None
        for r in range(self.world):
            if r == self.rank:
                continue
            p = C.c_void_p(0)
            code = self._lib.rc_peer_open((C.c_ubyte * 64).from_buffer_copy(handles[r]), C.byref(p))
            if code:
                from ._lib import last_error
                err = last_error()
                break
            self.peer_base[r] = int(p.value)
        if self.world > 1:
            # the mapping either works on EVERY rank or the exchange is abandoned on every rank (the caller falls back to
            # the torch.distributed all-gather): agree before anybody pushes
            okflag = torch.tensor([0 if err else 1], dtype=torch.int32, device=self.dev)
            td.all_reduce(okflag, op=td.ReduceOp.MIN, group=group)
            if int(okflag.item()) == 0:
                for r, b in enumerate(self.peer_base):
                    if r != self.rank and b:
                        self._lib.rc_peer_close(C.c_void_p(b))
                td.barrier(group=group)
                self._lib.rc_peer_free(C.c_void_p(self.base))
                raise PeerExchangeUnavailable(err or "a peer could not map this rank's exchange buffer")
            td.barrier(group=group)            # every buffer is mapped (and zeroed) before anyone pushes
        VP = C.c_void_p * self.world
        self._flag_ptrs = VP(*[C.c_void_p(b) for b in self.peer_base])
        self._tensor_ptrs = [VP(*[C.c_void_p(b + self.FLAG_BYTES + k * self.tensor_bytes) for b in self.peer_base])
                             for k in range(self.depth)]
        self.tensors = [torch.as_tensor(_DeviceArray(self.base + self.FLAG_BYTES + k * self.tensor_bytes,
                                                     (self.rows, self.C_total), "<f8"), device=self.dev)
                        for k in range(self.depth)]
        self._timed_out_ptr = self.base + 8 * 64           # uint64 behind the 64 flag slots
        self._timed_out = torch.as_tensor(_DeviceArray(self._timed_out_ptr, (1,), "<i8"), device=self.dev)
        self.local = [torch.empty((self.rows, self.hi - self.lo), dtype=torch.float64, device=self.dev)
                      for _ in range(self.depth)]
        self.side = torch.cuda.Stream(device=self.dev)
        self._ready = [torch.cuda.Event() for _ in range(self.depth)]
        self._pushed = [None] * self.depth
        self.seq = 0
        self._closed = False

    def local_block(self, seq: int) -> torch.Tensor:
        """Dense [rows][C_local] tensor the compute of step `seq` writes its statistics into.  Waits (stream-side) for
        the push that last read this slot."""
        k = seq % self.depth
        if self._pushed[k] is not None:
            torch.cuda.current_stream().wait_event(self._pushed[k])
        return self.local[k]

    def push(self, seq: int) -> None:
        """Publish local_block(seq) (written on the current stream) to every rank; returns immediately."""
        if seq < 1:
            raise ValueError("sequence numbers start at 1")
        k = seq % self.depth
        cur = torch.cuda.current_stream()
        self._ready[k].record(cur)
        self.side.wait_event(self._ready[k])
        st = C.c_void_p(self.side.cuda_stream)
        if self.world > 1 and seq > 1:
            # peers have finished with the tensor this push overwrites once they have published seq - 1
            self._check(self._lib.rc_peer_wait(C.c_void_p(self.base), self.world, seq - 1, self.timeout_s,
                                               C.c_void_p(self._timed_out_ptr), st))
        self._check(self._lib.rc_peer_push_columns(self._tensor_ptrs[k], self.world, C.c_void_p(self.local[k].data_ptr()),
                                                   self.rows, self.hi - self.lo, self.C_total, self.lo, st))
        self._check(self._lib.rc_peer_signal(self._flag_ptrs, self.world, self.rank, seq, st))
        ev = torch.cuda.Event()
        ev.record(self.side)
        self._pushed[k] = ev
        self.seq = seq

    def gathered(self, seq: int | None = None) -> torch.Tensor:
        """[rows][C_total] of step `seq` (default: the last pushed), valid on the current stream."""
        seq = self.seq if seq is None else seq
        self._check(self._lib.rc_peer_wait(C.c_void_p(self.base), self.world, seq, self.timeout_s,
                                           C.c_void_p(self._timed_out_ptr), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return self.tensors[seq % self.depth]

    def raise_if_timed_out(self) -> None:
        if int(self._timed_out.item()):
            raise RuntimeError("PeerStatsExchange: a peer did not publish its block within the timeout")

    def close(self) -> None:
        if self._closed:
            return
        self._closed = True
        torch.cuda.synchronize(self.dev)
        for r, b in enumerate(self.peer_base):
            if r != self.rank:
                self._lib.rc_peer_close(C.c_void_p(b))
        if self.world > 1:
            td.barrier()                        # nobody frees while a peer still has the buffer mapped
        self.tensors, self._timed_out = [], None
        self._lib.rc_peer_free(C.c_void_p(self.base))


class ShardedRobustnessSweep:
    """The whole fig-4/5 sweep of a controller set sharded by controller block over the ranks of `group`: per rank one
    RobustnessSweepPlan (evolution, statistics, per-group top-k / Kendall matrices, ARIM bootstrap: groups never
    straddle ranks) plus the exchange of the statistics blocks, so that every rank ends up with the [15][S][C_total]
    tensor of the whole set (what MCDataSim.get_metrics_dict holds for the whole controller set, mcsim.py:463-510).

    exchange = "peer" (default on CUDA): PeerStatsExchange, overlapped with the next step; "nccl": all_gather_stats."""

    def __init__(self, C_total: int, S: int, B: int, nspin: int, inspin: int, outspin: int, *, groups_per_rank: int = 1,
                 topk: int = 100, alpha_cluster: float = 0.05, dkw_eps: float = 0.0, model: int = 0, zz: bool = False,
                 fused: bool = False, nboot: int = 100, group=None, exchange: str = "peer"):
        from . import engine
        self.engine = engine
        self.world, self.rank = _world(group)
        self.group = group
        self.lo, self.hi = shard_bounds(C_total, self.world, self.rank)
        self.C_total, self.C_local, self.S, self.B = C_total, self.hi - self.lo, S, B
        self.nspin, self.inspin, self.outspin = nspin, inspin, outspin
        self.kw = dict(groups=groups_per_rank, topk=topk, alpha_cluster=alpha_cluster, dkw_eps=dkw_eps, model=model, zz=zz,
                       fused=fused, nboot=nboot)
        self.plan = engine.RobustnessSweepPlan(self.C_local, S, B, nspin, inspin, outspin, **self.kw)
        self.mode = exchange if self.world > 1 else "none"
        self.xchg = None
        if self.mode == "peer":
            try:
                self.xchg = PeerStatsExchange(engine.NUM_STATS * S, C_total, group=group)
            except PeerExchangeUnavailable as e:      # decided by all ranks together: use the collective instead
                if self.rank == 0:
                    print(f"[robchar_b200.dist] peer exchange unavailable ({e}); using the torch.distributed all-gather")
                self.mode = "nccl"
        self.seq = 0
        self._gathered = None
        self._host = None

    def step(self, ctrl: torch.Tensor, sigmas: torch.Tensor, *, seed: int = 0, evolution_events=None):
        """Device-resident step on this rank's controller block ctrl [C_local][N+1]; the exchange is started and NOT
        waited for.  Returns (local stats [15][S][C_local], tau [G][S][S])."""
        self.seq += 1
        S, Cl = self.S, self.C_local
        if self.mode == "peer":
            stats = self.xchg.local_block(self.seq).view(self.engine.NUM_STATS, S, Cl)
            st, tau = self.plan.run(ctrl, sigmas, seed=seed, c_offset=self.lo, evolution_events=evolution_events, stats=stats)
            self.xchg.push(self.seq)
        else:
            st, tau = self.plan.run(ctrl, sigmas, seed=seed, c_offset=self.lo, evolution_events=evolution_events)
            if self.mode == "nccl":
                self._gathered = all_gather_stats(st, self.C_total, self.group, out=self._gathered)
        return st, tau

    def step_host(self, ctrl_host: np.ndarray, sigmas_host: np.ndarray, *, seed: int = 0):
        """End-to-end step with HOST buffers for this rank's block (ctrl_host [C_local][N+1], ideally pinned): H2D,
        sweep, D2H of the local statistics / Kendall matrices / selection / ARIM into pinned host arrays, one C call
        (rc_robustness_sweep_host_keep); the statistics also stay on the device and their exchange with the peers is
        started (not waited for).  Returns (stats, tau, sel, arim, arim_std) numpy views of pinned buffers."""
        eng = self.engine
        self.seq += 1
        S, Cl = self.S, self.C_local
        k = min(self.kw["topk"], Cl // self.kw["groups"])
        if self._host is None:
            pin = lambda shape, dt=torch.float64: torch.empty(shape, dtype=dt).pin_memory().numpy()
            G = self.kw["groups"]
            self._host = (pin((eng.NUM_STATS, S, Cl)), pin((G, S, S)), pin((G, k), torch.int64), pin((G, S)), pin((G, S)))
            self._keep = torch.empty((eng.NUM_STATS, S, Cl), dtype=torch.float64, device=eng.require_cuda())
        st, tau, sel, ar, ars = self._host
        keep = self.xchg.local_block(self.seq).view(eng.NUM_STATS, S, Cl) if self.mode == "peer" else self._keep
        ctrl_host = np.ascontiguousarray(ctrl_host, dtype=np.float64)
        sig = np.ascontiguousarray(sigmas_host, dtype=np.float64).reshape(-1)
        vp = lambda a: C.c_void_p(a.ctypes.data)
        eng.check(eng.lib().rc_robustness_sweep_host_keep(
            vp(ctrl_host), Cl, self.nspin, self.inspin, self.outspin, vp(sig), S, self.B, self.kw["model"], int(bool(self.kw["zz"])),
            C.c_uint64(seed & (2**64 - 1)), self.lo, 0, float(self.kw["dkw_eps"]), int(bool(self.kw["fused"])), self.kw["groups"],
            self.kw["topk"], float(self.kw["alpha_cluster"]), vp(st), vp(tau), vp(sel), int(self.kw["nboot"]), vp(ar), vp(ars),
            C.c_void_p(keep.data_ptr()), eng._stream()))
        if self.mode == "peer":
            self.xchg.push(self.seq)
        elif self.mode == "nccl":
            self._gathered = all_gather_stats(keep, self.C_total, self.group, out=self._gathered)
        return st, tau, sel, ar, ars

    def gathered(self) -> torch.Tensor:
        """[15][S][C_total] statistics of the last step on the current stream (waits for the peers' blocks)."""
        if self.mode == "peer":
            return self.xchg.gathered(self.seq).view(self.engine.NUM_STATS, self.S, self.C_total)
        if self.mode == "nccl":
            return self._gathered
        return self.plan.stats

    def finish(self) -> None:
        """Wait for the last exchange, synchronise, and raise on non-convergence / illegal samples / exchange timeout."""
        if self.mode == "peer" and self.seq:
            self.xchg.gathered(self.seq)
        torch.cuda.synchronize()
        self.plan.counters.raise_if_set()
        if self.xchg is not None:
            self.xchg.raise_if_timed_out()

    def close(self) -> None:
        if self.xchg is not None:
            self.xchg.close()
            self.xchg = None


def draw_shard_range(B: int, world: int, rank: int) -> tuple[int, int]:
    """Draw range [b_lo, b_hi) of `rank` in a draw-sharded sweep of B draws (rc_draw_shard_range): whole merge blocks
    of the fixed chunk grid, so the merged statistics do not depend on the world size."""
    from ._lib import check, lib
    lo, hi = C.c_int64(0), C.c_int64(0)
    check(lib().rc_draw_shard_range(B, world, rank, C.byref(lo), C.byref(hi), None, None))
    return int(lo.value), int(hi.value)


def gather_blocks(local_blocks: torch.Tensor, group=None) -> torch.Tensor:
    """local [MERGE_BLOCKS/world][nseg][17] -> [MERGE_BLOCKS][nseg][17] in rank order (one all-gather)."""
    world, _ = _world(group)
    if world == 1:
        return local_blocks
    out = torch.empty((world * local_blocks.shape[0],) + tuple(local_blocks.shape[1:]), dtype=local_blocks.dtype,
                      device=local_blocks.device)
    td.all_gather_into_tensor(out, local_blocks.contiguous(), group=group)
    return out


def sharded_rim_sweep(ctrl, sigmas, B: int, nspin: int, inspin: int, outspin: int, *, dkw_eps: float = 0.0,
                      seed: int = 0, model: int = 0, zz: bool = False, fused: bool = True, group=None,
                      compute_fn=None, shard: str = "auto", blocks_fn=None, finalize_fn=None) -> torch.Tensor:
    """[15][S][C] statistics of the whole controller set on every rank.

    shard = "controllers": this rank's contiguous controller block, statistics all-gathered;
    shard = "draws": this rank's draw range of EVERY controller (Philox mode, fused streaming statistics), block
    results all-gathered and merged in a fixed order — bit-identical to the single-GPU run;
    shard = "auto": draws when there are fewer controllers than ranks, controllers otherwise.
    compute_fn(ctrl_block, c_offset) -> [15][S][C_local], blocks_fn(world, rank) -> [8/world][S*C][17] and
    finalize_fn(blocks [8][S*C][17]) -> [15][S][C] override the device calls (gloo host-logic tests)."""
    ctrl = np.asarray(ctrl, dtype=np.float64) if not isinstance(ctrl, torch.Tensor) else ctrl
    C_total = ctrl.shape[0]
    world, rank = _world(group)
    if shard == "auto":
        shard = "draws" if (C_total < world and MERGE_BLOCKS % world == 0) else "controllers"
    if shard == "draws":
        if MERGE_BLOCKS % world:
            raise ValueError(f"draw sharding needs a world size that divides {MERGE_BLOCKS}")
        if blocks_fn is not None:
            local = blocks_fn(world, rank)
        else:
            from . import engine
            local = engine.fidelity_stats_blocks(ctrl, sigmas, B, nspin, inspin, outspin, world=world, rank=rank,
                                                 dkw_eps=dkw_eps, seed=seed, model=model, zz=zz)
        blocks = gather_blocks(local, group)
        if finalize_fn is not None:
            return finalize_fn(blocks)
        from . import engine
        S = int(np.asarray(sigmas.cpu() if isinstance(sigmas, torch.Tensor) else sigmas).reshape(-1).shape[0])
        return engine.stats_from_blocks(blocks, B, dkw_eps).view(engine.NUM_STATS, S, C_total)
    if shard != "controllers":
        raise ValueError(f"unknown shard mode {shard!r}")
    lo, hi = shard_bounds(C_total, world, rank)
    block = ctrl[lo:hi]
    if compute_fn is not None:
        local = compute_fn(block, lo)
    else:
        from . import engine
        if fused:
            local = engine.fidelity_stats(block, sigmas, B, nspin, inspin, outspin, dkw_eps=dkw_eps, seed=seed,
                                          model=model, zz=zz, c_offset=lo)
        else:
            f = engine.fidelity_mc(block, sigmas, B, nspin, inspin, outspin, seed=seed, model=model, zz=zz, c_offset=lo)
            local = engine.stats(f, dkw_eps)
    return all_gather_stats(local, C_total, group)
