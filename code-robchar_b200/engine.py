"""Device-level operations: torch tensors in HBM handed to the C-ABI by raw pointer.

PyTorch is used only for device memory, streams and (in dist.py) torch.distributed.  Every
function here requires a CUDA device and the built native library; nothing is computed on the
CPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import MODEL_COMPLEX3, MODEL_REAL2, NUM_STATS, ObjectiveFrame, check, lib

# keys of the reference's metric registry, in .mcm order (mcsim.py:178-183, 496-498)
METRIC_W = r'$W(.,\delta(x-1))$'
METRIC_NAMES = [METRIC_W, "Q th. 0.95", "Q th. 0.98", "std", "worst case fid"]
STAT_KEYS = [m + s for m in METRIC_NAMES for s in ("", " upper", " lower")]


def launch_count() -> int:
    """Kernels of librobchar_b200.so launched by this process so far, counted at the launch sites inside the library
    (rc_launch_count; bench.py reports the difference over its timed region as gpu_launches)."""
    return int(lib().rc_launch_count())


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.RobcharLibraryError("robchar_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
    return torch.device("cuda", torch.cuda.current_device())


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f64(t, dev):
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(np.asarray(t, dtype=np.float64))
    return t.to(device=dev, dtype=torch.float64).contiguous()


def draws_per_eval(nspin: int, model: int = MODEL_COMPLEX3) -> int:
    return (3 if model == MODEL_COMPLEX3 else 2) * nspin


def compute_dkw_error(alpha, nobs):
    """wd_sortof_fast_implementation.py:38-39 (host scalar)."""
    return np.sqrt(np.log(2 / alpha) / (2 * nobs))


class Counters:
    """Device counters [nonconv, illegal] checked lazily (one D2H) by raise_if_set()."""

    def __init__(self, dev):
        self.t = torch.zeros(2, dtype=torch.int64, device=dev)

    @property
    def nonconv_ptr(self):
        return C.c_void_p(self.t.data_ptr())

    @property
    def illegal_ptr(self):
        return C.c_void_p(self.t.data_ptr() + 8)

    def raise_if_set(self):
        nc, il = (int(v) for v in self.t.tolist())
        if nc:
            raise _lib.EigensolverNonConvergence(f"eigensolver did not converge for {nc} evaluations")
        if il:
            raise AssertionError("illegal fids values - must be in [0,1]")


def fidelity_mc(ctrl, sigmas, B: int, nspin: int, inspin: int, outspin: int, *, model: int = MODEL_COMPLEX3,
                zz: bool = False, seed: int = 0, c_offset: int = 0, b_offset: int = 0, replay=None, out=None,
                counters: Counters | None = None, check_convergence: bool = True, topo: str = "chain") -> torch.Tensor:
    """Fidelity tensor [S][C][B] (mcsim.py:422-456).  replay: standard normals [S][C][B][K] or None (Philox).
    topo="ring" (noise_model.py:83-85): the batched dense path (rc_dense_fidelity_mc), same layout and draws."""
    if topo not in ("chain", "linear", "ring"):
        raise ValueError(f"unknown topology {topo!r}")
    if topo == "ring":
        return dense_fidelity_mc(ctrl, sigmas, B, nspin, inspin, outspin, model=model, zz=zz, seed=seed, c_offset=c_offset,
                                 b_offset=b_offset, replay=replay, out=out, ring=True)
    dev = require_cuda()
    ctrl = _f64(ctrl, dev)
    sigmas = _f64(sigmas, dev).reshape(-1)
    if ctrl.dim() != 2 or ctrl.shape[1] != nspin + 1:
        raise ValueError(f"ctrl must be [C][{nspin + 1}], got {tuple(ctrl.shape)}")
    Cn, S = ctrl.shape[0], sigmas.shape[0]
    if replay is not None:
        replay = _f64(replay, dev)
        K = draws_per_eval(nspin, model)
        if replay.numel() != S * Cn * B * K:
            raise ValueError(f"replay must hold S*C*B*{K} standard normals")
    if out is None:
        out = torch.empty((S, Cn, B), dtype=torch.float64, device=dev)
    own = counters is None
    if own:
        counters = Counters(dev)
    check(lib().rc_fidelity_mc(_ptr(ctrl), Cn, nspin, inspin, outspin, _ptr(sigmas), S, B, model, int(bool(zz)),
                               C.c_uint64(seed & (2**64 - 1)), c_offset, b_offset, _ptr(replay), _ptr(out),
                               counters.nonconv_ptr, _stream()))
    if own and check_convergence:
        counters.raise_if_set()
    return out


def dense_fidelity_mc(ctrl, sigmas, B: int, nspin: int, inspin: int, outspin: int, *, model: int = MODEL_COMPLEX3,
                      zz: bool = False, seed: int = 0, c_offset: int = 0, b_offset: int = 0, replay=None, out=None,
                      ring: bool = True, tile: int = 1 << 17) -> torch.Tensor:
    """Fidelity tensor [S][C][B] of a sweep through the dense path (rc_dense_fidelity_mc): ring topology under the
    structured perturbation, `tile` evaluations per pass of the batched matrix exponential."""
    dev = require_cuda()
    ctrl = _f64(ctrl, dev)
    sigmas = _f64(sigmas, dev).reshape(-1)
    if ctrl.dim() != 2 or ctrl.shape[1] != nspin + 1:
        raise ValueError(f"ctrl must be [C][{nspin + 1}], got {tuple(ctrl.shape)}")
    Cn, S = ctrl.shape[0], sigmas.shape[0]
    if replay is not None:
        replay = _f64(replay, dev)
        if replay.numel() != S * Cn * B * draws_per_eval(nspin, model):
            raise ValueError("replay must hold S*C*B*K standard normals")
    if out is None:
        out = torch.empty((S, Cn, B), dtype=torch.float64, device=dev)
    total = S * Cn * B
    wb = lib().rc_dense_fidelity_mc_workspace_bytes(nspin, max(1, min(tile, total)))
    ws = torch.empty(wb, dtype=torch.uint8, device=dev)
    check(lib().rc_dense_fidelity_mc(_ptr(ctrl), Cn, nspin, inspin, outspin, _ptr(sigmas), S, B, model, int(bool(zz)),
                                     int(bool(ring)), C.c_uint64(seed & (2**64 - 1)), c_offset, b_offset, _ptr(replay),
                                     _ptr(out), _ptr(ws), wb, _stream()))
    return out


def directional_fidelity_mc(ctrl, sigmas, B: int, nspin: int, inspin: int, outspin: int, *, zz: bool = False,
                            ring: bool = False, seed: int = 0, c_offset: int = 0, b_offset: int = 0, replay=None, out=None,
                            return_draws: bool = False, tile: int = 1 << 17):
    """Fidelity tensor [S][C][B] of a sweep under directional_perturbation (noise_model.py:150-201) through
    rc_directional_fidelity_mc.  replay: [S][C][B][3] = (direction index, n0, n1) with standard normals — what
    np.random.randint(0, 3N) and rng(size=2) / sigma give upstream — or None (in-kernel Philox draws;
    return_draws=True also returns the [S][C][B][3] draws that were used)."""
    dev = require_cuda()
    ctrl = _f64(ctrl, dev)
    sigmas = _f64(sigmas, dev).reshape(-1)
    if ctrl.dim() != 2 or ctrl.shape[1] != nspin + 1:
        raise ValueError(f"ctrl must be [C][{nspin + 1}], got {tuple(ctrl.shape)}")
    Cn, S = ctrl.shape[0], sigmas.shape[0]
    if replay is not None:
        replay = _f64(replay, dev)
        if replay.numel() != S * Cn * B * 3:
            raise ValueError("replay must hold S*C*B*3 values: (direction index, n0, n1) per evaluation")
    if out is None:
        out = torch.empty((S, Cn, B), dtype=torch.float64, device=dev)
    draws = torch.empty((S, Cn, B, 3), dtype=torch.float64, device=dev) if (return_draws and replay is None) else None
    total = S * Cn * B
    wb = lib().rc_dense_fidelity_mc_workspace_bytes(nspin, max(1, min(tile, total)))
    ws = torch.empty(wb, dtype=torch.uint8, device=dev)
    check(lib().rc_directional_fidelity_mc(_ptr(ctrl), Cn, nspin, inspin, outspin, _ptr(sigmas), S, B, int(bool(zz)),
                                           int(bool(ring)), C.c_uint64(seed & (2**64 - 1)), c_offset, b_offset,
                                           _ptr(replay), _ptr(out), _ptr(draws), _ptr(ws), wb, _stream()))
    if return_draws:
        return out, (replay.reshape(S, Cn, B, 3) if replay is not None else draws)
    return out


def fidelity_mc_stats(ctrl, sigmas, B: int, nspin: int, inspin: int, outspin: int, *, dkw_eps: float = 0.0,
                      model: int = MODEL_COMPLEX3, zz: bool = False, seed: int = 0, c_offset: int = 0, b_offset: int = 0,
                      replay=None, out=None, counters: Counters | None = None, check: bool = True):
    """Fidelity tensor [S][C][B] (mcsim.py:422-456) AND its [15][S][C] statistics (mcsim.py:463-510) in one C
    call (rc_fidelity_mc_stats: evolution kernel + the sort-free statistics pass).  Returns (fids, stats)."""
    dev = require_cuda()
    ctrl = _f64(ctrl, dev)
    sigmas = _f64(sigmas, dev).reshape(-1)
    if ctrl.dim() != 2 or ctrl.shape[1] != nspin + 1:
        raise ValueError(f"ctrl must be [C][{nspin + 1}], got {tuple(ctrl.shape)}")
    Cn, S = ctrl.shape[0], sigmas.shape[0]
    if replay is not None:
        replay = _f64(replay, dev)
        K = draws_per_eval(nspin, model)
        if replay.numel() != S * Cn * B * K:
            raise ValueError(f"replay must hold S*C*B*{K} standard normals")
    if out is None:
        out = torch.empty((S, Cn, B), dtype=torch.float64, device=dev)
    st = torch.empty((NUM_STATS, S, Cn), dtype=torch.float64, device=dev)
    own = counters is None
    if own:
        counters = Counters(dev)
    _lib.check(lib().rc_fidelity_mc_stats(_ptr(ctrl), Cn, nspin, inspin, outspin, _ptr(sigmas), S, B, model,
                                          int(bool(zz)), C.c_uint64(seed & (2**64 - 1)), c_offset, b_offset, _ptr(replay),
                                          float(dkw_eps), _ptr(out), _ptr(st), counters.nonconv_ptr, counters.illegal_ptr,
                                          _stream()))
    if own and check:
        counters.raise_if_set()
    return out, st


def stats_unsorted(fids: torch.Tensor, dkw_eps: float = 0.0, *, check_legal: bool = True) -> torch.Tensor:
    """[15, *lead] statistics of fids[*lead, B] without sorting (rc_stats_unsorted): same values as stats()
    up to rounding of W and std; counts and minimum identical.  fids is left untouched."""
    dev = require_cuda()
    fids = _f64(fids, dev) if not (isinstance(fids, torch.Tensor) and fids.is_cuda and fids.dtype == torch.float64
                                   and fids.is_contiguous()) else fids
    lead, B = tuple(fids.shape[:-1]), fids.shape[-1]
    nseg = int(np.prod(lead)) if lead else 1
    out = torch.empty((NUM_STATS,) + lead, dtype=torch.float64, device=dev)
    cnt = Counters(dev)
    _lib.check(lib().rc_stats_unsorted(_ptr(fids), nseg, B, float(dkw_eps), _ptr(out), cnt.illegal_ptr, _stream()))
    if check_legal:
        cnt.raise_if_set()
    return out


def evolution_kernel_name(nspin: int, replay: bool = False, fused: bool = False) -> str:
    buf = C.create_string_buffer(128)
    check(lib().rc_evolution_kernel_name(nspin, int(bool(replay)), int(bool(fused)), buf, 128))
    return buf.value.decode()


def spectral_fallbacks(reset: bool = False) -> int:
    """Evaluations of the N >= 13 kernels that were recomputed with accumulated eigenvector rows because the
    spectral-weights error estimate rejected them (rc_spectral_fallbacks), since the last reset."""
    require_cuda()
    v = C.c_ulonglong(0)
    check(lib().rc_spectral_fallbacks(C.byref(v), int(bool(reset)), _stream()))
    return int(v.value)


def rim_p(fids, p: float = 2, *, check_legal: bool = True) -> torch.Tensor:
    """p-RIM (mean((1-f)^p))^(1/p) of fids[*lead, B] per segment on the device (rc_rim_p): [*lead]."""
    dev = require_cuda()
    fids = _f64(fids, dev) if not (isinstance(fids, torch.Tensor) and fids.is_cuda and fids.dtype == torch.float64
                                   and fids.is_contiguous()) else fids
    lead, B = tuple(fids.shape[:-1]), fids.shape[-1]
    nseg = int(np.prod(lead)) if lead else 1
    out = torch.empty(lead if lead else (1,), dtype=torch.float64, device=dev)
    cnt = Counters(dev)
    check(lib().rc_rim_p(_ptr(fids), nseg, B, float(p), _ptr(out), cnt.illegal_ptr, _stream()))
    if check_legal:
        cnt.raise_if_set()
    return out if lead else out.reshape(())


def philox_normals(C_: int, nspin: int, S: int, B: int, *, model: int = MODEL_COMPLEX3, seed: int = 0,
                   c_offset: int = 0, b_offset: int = 0) -> torch.Tensor:
    dev = require_cuda()
    K = draws_per_eval(nspin, model)
    out = torch.empty((S, C_, B, K), dtype=torch.float64, device=dev)
    check(lib().rc_philox_normals(C_, nspin, S, B, model, C.c_uint64(seed & (2**64 - 1)), c_offset, b_offset, _ptr(out),
                                  _stream()))
    return out


def stats(fids: torch.Tensor, dkw_eps: float = 0.0, *, sort_inplace: bool = False, check_legal: bool = True) -> torch.Tensor:
    """[15, *lead] statistics of fids[*lead, B] (mcsim.py:482-498).  sort_inplace mimics wd_from_ideal's
    in-place sort of its argument (wd_sortof_fast_implementation.py:105)."""
    dev = require_cuda()
    fids = _f64(fids, dev) if not (isinstance(fids, torch.Tensor) and fids.is_cuda and fids.dtype == torch.float64
                                   and fids.is_contiguous()) else fids
    lead, B = tuple(fids.shape[:-1]), fids.shape[-1]
    nseg = int(np.prod(lead)) if lead else 1
    out = torch.empty((NUM_STATS,) + lead, dtype=torch.float64, device=dev)
    wb = lib().rc_stats_workspace_bytes(nseg, B)
    if wb == 0:
        raise ValueError("segment length too large for the sort path; use fidelity_stats (streaming)")
    ws = torch.empty(wb, dtype=torch.uint8, device=dev)
    cnt = Counters(dev)
    check(lib().rc_stats(_ptr(fids), nseg, B, float(dkw_eps), _ptr(out), _ptr(fids) if sort_inplace else C.c_void_p(0),
                         cnt.illegal_ptr, _ptr(ws), wb, _stream()))
    if check_legal:
        cnt.raise_if_set()
    return out


def fidelity_stats(ctrl, sigmas, B: int, nspin: int, inspin: int, outspin: int, *, dkw_eps: float = 0.0,
                   model: int = MODEL_COMPLEX3, zz: bool = False, seed: int = 0, c_offset: int = 0, b_offset: int = 0,
                   replay=None, check_convergence: bool = True) -> torch.Tensor:
    """Fused evolution + streaming statistics, [15][S][C]; never materialises the fidelity tensor."""
    dev = require_cuda()
    ctrl = _f64(ctrl, dev)
    sigmas = _f64(sigmas, dev).reshape(-1)
    Cn, S = ctrl.shape[0], sigmas.shape[0]
    if replay is not None:
        replay = _f64(replay, dev)
    out = torch.empty((NUM_STATS, S, Cn), dtype=torch.float64, device=dev)
    wb = lib().rc_fidelity_stats_workspace_bytes(S * Cn, B)
    ws = torch.empty(wb, dtype=torch.uint8, device=dev)
    cnt = Counters(dev)
    check(lib().rc_fidelity_stats(_ptr(ctrl), Cn, nspin, inspin, outspin, _ptr(sigmas), S, B, model, int(bool(zz)),
                                  C.c_uint64(seed & (2**64 - 1)), c_offset, b_offset, _ptr(replay), float(dkw_eps),
                                  _ptr(out), cnt.nonconv_ptr, _ptr(ws), wb, _stream()))
    if check_convergence:
        cnt.raise_if_set()
    return out


def fidelity_stats_blocks(ctrl, sigmas, B: int, nspin: int, inspin: int, outspin: int, *, world: int, rank: int,
                          dkw_eps: float = 0.0, model: int = MODEL_COMPLEX3, zz: bool = False, seed: int = 0,
                          c_offset: int = 0, b_offset: int = 0, counters: Counters | None = None) -> torch.Tensor:
    """Draw-sharded sweep, this rank's part: fused evolution + streaming statistics over the draw range
    rc_draw_shard_range(B, world, rank) of every controller, merged into the rank's block results
    [8/world][S*C][17] (rc_fidelity_stats_blocks).  All-gathered in rank order they feed stats_from_blocks."""
    dev = require_cuda()
    ctrl = _f64(ctrl, dev)
    sigmas = _f64(sigmas, dev).reshape(-1)
    Cn, S = ctrl.shape[0], sigmas.shape[0]
    if world < 1 or 8 % world:
        raise ValueError("draw sharding needs a world size that divides 8")
    out = torch.empty((8 // world, S * Cn, 17), dtype=torch.float64, device=dev)
    wb = lib().rc_fidelity_stats_blocks_workspace_bytes(S * Cn, B, world)
    ws = torch.empty(wb, dtype=torch.uint8, device=dev)
    cnt = counters or Counters(dev)
    check(lib().rc_fidelity_stats_blocks(_ptr(ctrl), Cn, nspin, inspin, outspin, _ptr(sigmas), S, B, model, int(bool(zz)),
                                         C.c_uint64(seed & (2**64 - 1)), c_offset, b_offset, world, rank, float(dkw_eps),
                                         _ptr(out), cnt.nonconv_ptr, _ptr(ws), wb, _stream()))
    if counters is None:
        cnt.raise_if_set()
    return out


def stats_from_blocks(blocks: torch.Tensor, B: int, dkw_eps: float = 0.0) -> torch.Tensor:
    """[15][nseg] statistics from the gathered block results [8][nseg][17] (rc_stats_from_blocks): the fixed-order
    finish of a draw-sharded sweep, bit-identical to fidelity_stats on one GPU."""
    dev = require_cuda()
    blocks = _f64(blocks, dev)
    if blocks.dim() != 3 or blocks.shape[0] != 8 or blocks.shape[2] != 17:
        raise ValueError("blocks must be [8][nseg][17]")
    nseg = blocks.shape[1]
    out = torch.empty((NUM_STATS, nseg), dtype=torch.float64, device=dev)
    check(lib().rc_stats_from_blocks(_ptr(blocks), nseg, B, float(dkw_eps), _ptr(out), _stream()))
    return out


def ranks(values) -> torch.Tensor:
    """Ordinal ranks per row, ascending, NaN last, ties by index (mcsim.py:513-518)."""
    dev = require_cuda()
    v = _f64(values, dev)
    one_d = v.dim() == 1
    v2 = v.reshape(1, -1) if one_d else v.reshape(-1, v.shape[-1])
    R, n = v2.shape
    out = torch.empty((R, n), dtype=torch.int64, device=dev)
    wb = lib().rc_ranks_workspace_bytes(R, n)
    if wb == 0:
        raise ValueError("ranking problem too large")
    ws = torch.empty(wb, dtype=torch.uint8, device=dev)
    check(lib().rc_ranks(_ptr(v2), R, n, _ptr(out), _ptr(ws), wb, _stream()))
    return out.reshape(v.shape)


def clustered_ranks(values, alpha: float | None = 0.05, r: float | None = None) -> torch.Tensor:
    """get_ranks_clustered_little (generate_fig4_kendallrankanalysis.py:146-164); radius r, or
    alpha*(max-min) per row when r is None (…fig4…py:97)."""
    dev = require_cuda()
    v = _f64(values, dev)
    v2 = v.reshape(1, -1) if v.dim() == 1 else v.reshape(-1, v.shape[-1])
    R, n = v2.shape
    out = torch.empty((R, n), dtype=torch.float64, device=dev)
    wb = lib().rc_ranks_workspace_bytes(R, n)
    ws = torch.empty(wb, dtype=torch.uint8, device=dev)
    a, rf = (-1.0, float(r)) if r is not None else (float(alpha), 0.0)
    check(lib().rc_clustered_ranks(_ptr(v2), R, n, a, rf, _ptr(out), _ptr(ws), wb, _stream()))
    return out.reshape(v.shape)


def kendall_tau_b(x, y) -> torch.Tensor:
    """tau[j][i] between x rows (double ranks) and y rows (int64 ranks); scipy.stats.kendalltau variant b."""
    dev = require_cuda()
    x = _f64(x, dev)
    y = (y if isinstance(y, torch.Tensor) else torch.as_tensor(np.asarray(y))).to(device=dev, dtype=torch.int64).contiguous()
    x2 = x.reshape(1, -1) if x.dim() == 1 else x
    y2 = y.reshape(1, -1) if y.dim() == 1 else y
    Rx, n = x2.shape
    Ry = y2.shape[0]
    if y2.shape[1] != n:
        raise ValueError("x and y rows must have the same length")
    if n > KENDALL_PAIRCOUNT_MAX_N:
        return kendall_tau_b_batched(x2[None], y2[None])[0]
    tau = torch.empty((Rx, Ry), dtype=torch.float64, device=dev)
    counts = torch.empty((Rx, Ry, 4), dtype=torch.int64, device=dev)
    check(lib().rc_kendall_tau_b(_ptr(x2), Rx, _ptr(y2), Ry, n, _ptr(tau), _ptr(counts), _stream()))
    return tau


KENDALL_PAIRCOUNT_MAX_N = 4096     # above: rc_kendall_tau_b_large (O(n log^2 n)) instead of the O(n^2) pair count


def kendall_tau_b_batched(x: torch.Tensor, y: torch.Tensor, force_large: bool = False) -> torch.Tensor:
    """x [G][Rx][n] double ranks, y [G][Ry][n] int64 ranks -> tau [G][Rx][Ry] (one launch for all groups; long rank
    vectors, n > 4096, take the sort + merge-pass path, bit-identical integer counts)."""
    dev = require_cuda()
    x = _f64(x, dev)
    y = y.to(device=dev, dtype=torch.int64).contiguous()
    G, Rx, n = x.shape
    Ry = y.shape[1]
    if y.shape[0] != G or y.shape[2] != n:
        raise ValueError("x and y must share group count and row length")
    tau = torch.empty((G, Rx, Ry), dtype=torch.float64, device=dev)
    counts = torch.empty((G, Rx, Ry, 4), dtype=torch.int64, device=dev)
    if n > KENDALL_PAIRCOUNT_MAX_N or force_large:      # long rank vectors: sort + merge-pass inversion count
        wb = lib().rc_kendall_large_workspace_bytes(G, Rx, Ry, n)
        if wb == 0:
            raise ValueError("Kendall problem too large (more than 2^31 elements)")
        ws = torch.empty(wb, dtype=torch.uint8, device=dev)
        check(lib().rc_kendall_tau_b_large(_ptr(x), _ptr(y), G, Rx, Ry, n, _ptr(tau), _ptr(counts), _ptr(ws), wb, _stream()))
        return tau
    check(lib().rc_kendall_tau_b_batched(_ptr(x), _ptr(y), G, Rx, Ry, n, _ptr(tau), _ptr(counts), _stream()))
    return tau


def grouped_rank_consistency(W: torch.Tensor, groups: int, topk: int = 100, alpha: float = 0.05):
    """The paper's fig-4 analysis for `groups` controller sets at once (one C call).  W: RIM matrix
    [S][groups*Cg].  Per group: keep the topk controllers with the smallest RIM at sigma_sim index 0
    in their original column order (mcsim.py:651-660), then the S x S Kendall matrix of clustered vs
    ordinal ranks (generate_fig4_kendallrankanalysis.py:94-120).  Returns (tau [G][S][S], selected
    column index [G][k] within the group, W_topk [G][S][k])."""
    dev = require_cuda()
    W = _f64(W, dev)
    S, Ctot = W.shape
    if Ctot % groups:
        raise ValueError("controller count must be a multiple of the group count")
    Cg = Ctot // groups
    k = min(topk, Cg)
    tau = torch.empty((groups, S, S), dtype=torch.float64, device=dev)
    sel = torch.empty((groups, k), dtype=torch.int64, device=dev)
    Wsel = torch.empty((groups, S, k), dtype=torch.float64, device=dev)
    wb = lib().rc_rank_consistency_workspace_bytes(S, groups, Cg, topk)
    if wb == 0:
        raise ValueError("ranking problem too large")
    ws = torch.empty(wb, dtype=torch.uint8, device=dev)
    check(lib().rc_rank_consistency(_ptr(W), S, groups, Cg, topk, float(alpha), _ptr(tau), _ptr(sel), _ptr(Wsel),
                                    _ptr(ws), wb, _stream()))
    return tau, sel, Wsel


def kendall_matrix(wd_data_c, alpha: float = 0.05) -> torch.Tensor:
    """jkt_or_ordinaltau_pairwise (generate_fig4_kendallrankanalysis.py:94-120): clustered ranks of row j
    against ordinal ranks (+1) of row i."""
    dev = require_cuda()
    w = _f64(wd_data_c, dev)
    cr = clustered_ranks(w, alpha=alpha)
    rk = ranks(w) + 1
    return kendall_tau_b(cr, rk)


def mc_sweep_host(ctrl: np.ndarray, sigmas: np.ndarray, B: int, nspin: int, inspin: int, outspin: int, *,
                  dkw_eps: float = 0.0, model: int = MODEL_COMPLEX3, zz: bool = False, seed: int = 0, c_offset: int = 0,
                  b_offset: int = 0, replay: np.ndarray | None = None, fused: bool = False, want_fids: bool = False,
                  stats_out: np.ndarray | None = None, fids_out: np.ndarray | None = None):
    """Whole sweep through the C-ABI with HOST (numpy) buffers: H2D, evolution, statistics, D2H."""
    require_cuda()
    ctrl = np.ascontiguousarray(ctrl, dtype=np.float64)
    sigmas = np.ascontiguousarray(sigmas, dtype=np.float64).reshape(-1)
    Cn, S = ctrl.shape[0], sigmas.shape[0]
    if stats_out is None:
        stats_out = np.empty((NUM_STATS, S, Cn))
    if want_fids and fids_out is None:
        fids_out = np.empty((S, Cn, B))
    if replay is not None:
        replay = np.ascontiguousarray(replay, dtype=np.float64)
    vp = lambda a: C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(0)
    check(lib().rc_mc_sweep_host(vp(ctrl), Cn, nspin, inspin, outspin, vp(sigmas), S, B, model, int(bool(zz)),
                                 C.c_uint64(seed & (2**64 - 1)), c_offset, b_offset, vp(replay), float(dkw_eps),
                                 int(bool(fused)), vp(fids_out), vp(stats_out), _stream()))
    return stats_out, fids_out


_PINNED = {}


def _pinned(name, shape, dtype):
    """Cached pinned host buffer (numpy view) so D2H copies run at full PCIe rate and asynchronously."""
    key = (name, tuple(shape), dtype)
    if key not in _PINNED:
        _PINNED[key] = torch.empty(tuple(shape), dtype=dtype).pin_memory()
    return _PINNED[key].numpy()


def robustness_sweep_host(ctrl: np.ndarray, sigmas: np.ndarray, B: int, nspin: int, inspin: int, outspin: int, *,
                          groups: int = 1, topk: int = 100, alpha_cluster: float = 0.05, dkw_eps: float = 0.0,
                          model: int = MODEL_COMPLEX3, zz: bool = False, seed: int = 0, c_offset: int = 0,
                          b_offset: int = 0, fused: bool = False, pinned_outputs: bool = True, nboot: int = 100):
    """Evolution + statistics + per-group top-k / Kendall matrices in ONE C call with host buffers
    (rc_robustness_sweep_host).  Returns (stats [15][S][C], tau [G][S][S], sel [G][k], arim [G][S], arim_std [G][S]) as numpy arrays;
    with pinned_outputs they are views of cached pinned buffers (copy them to keep across calls)."""
    require_cuda()
    ctrl = np.ascontiguousarray(ctrl, dtype=np.float64)
    sigmas = np.ascontiguousarray(sigmas, dtype=np.float64).reshape(-1)
    Cn, S = ctrl.shape[0], sigmas.shape[0]
    if Cn % groups:
        raise ValueError("controller count must be a multiple of the group count")
    k = min(topk, Cn // groups)
    if pinned_outputs:
        st = _pinned("stats", (NUM_STATS, S, Cn), torch.float64)
        tau = _pinned("tau", (groups, S, S), torch.float64)
        sel = _pinned("sel", (groups, k), torch.int64)
        ar = _pinned("arim", (groups, S), torch.float64)
        ars = _pinned("arim_std", (groups, S), torch.float64)
    else:
        st, tau, sel = np.empty((NUM_STATS, S, Cn)), np.empty((groups, S, S)), np.empty((groups, k), dtype=np.int64)
        ar, ars = np.empty((groups, S)), np.empty((groups, S))
    vp = lambda a: C.c_void_p(a.ctypes.data)
    check(lib().rc_robustness_sweep_host(vp(ctrl), Cn, nspin, inspin, outspin, vp(sigmas), S, B, model, int(bool(zz)),
                                         C.c_uint64(seed & (2**64 - 1)), c_offset, b_offset, float(dkw_eps),
                                         int(bool(fused)), groups, topk, float(alpha_cluster), vp(st), vp(tau), vp(sel),
                                         int(nboot), vp(ar), vp(ars), _stream()))
    return st, tau, sel, ar, ars


def objective_host(x, rows, nspin: int, inspin: int, outspin: int, *, model: int = MODEL_COMPLEX3, zz: bool = False,
                   want_fids: bool = True, want_stats: bool = False, dkw_eps: float = 0.0, want_amps: bool = False):
    """Fidelities of ONE controller x [N+1] under m explicit perturbation rows [m][K] (replay layout, sigma 1), or
    its nominal fidelity when rows is None: the optimiser-loop entry point (rc_objective_host), host numpy in and
    out, one H2D + one launch (+ one for the statistics) + one D2H, no allocations in steady state.
    Returns fids [m], or (fids, stats [15]) with want_stats (fids is None without want_fids); with want_amps the
    complex amplitudes U_k[out, in] as a complex128 array [m] are appended (real symmetric model only)."""
    require_cuda()
    x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
    if x.shape[0] != nspin + 1:
        raise ValueError(f"x must hold {nspin + 1} values")
    if rows is None:
        m, rp = 1, C.c_void_p(0)
    else:
        rows = np.ascontiguousarray(rows, dtype=np.float64)
        K = draws_per_eval(nspin, model)
        if rows.ndim != 2 or rows.shape[1] != K:
            raise ValueError(f"rows must be [m][{K}]")
        m, rp = rows.shape[0], C.c_void_p(rows.ctypes.data)
    out = np.empty(m) if want_fids else None
    st = np.empty(NUM_STATS) if want_stats else None
    amps = np.empty(m, dtype=np.complex128) if want_amps else None
    check(lib().rc_objective_host(C.c_void_p(x.ctypes.data), nspin, inspin, outspin, rp, m, model, int(bool(zz)),
                                  float(dkw_eps), C.c_void_p(out.ctypes.data if want_fids else 0),
                                  C.c_void_p(st.ctypes.data if want_stats else 0),
                                  C.c_void_p(amps.ctypes.data if want_amps else 0), _stream()))
    res = (out, st) if want_stats else (out,)
    if want_amps:
        res = res + (amps,)
    return res if len(res) > 1 else res[0]


class ObjectiveEvaluator:
    """Preallocated call frame of rc_objective_host for ONE shape (chain, model, number of perturbation rows m, wanted
    outputs): an optimiser calls its objective thousands of times with the same shape, so the numpy staging arrays
    and the ctypes argument tuple are built once and a call is two small copies plus one foreign call (the Python
    side of the per-call latency, tools/latency_bench.py).  m = 0: the nominal evaluation (no rows)."""

    def __init__(self, nspin: int, inspin: int, outspin: int, m: int, *, model: int = MODEL_COMPLEX3, zz: bool = False,
                 want_fids: bool = True, want_stats: bool = False, want_amps: bool = False, dkw_eps: float = 0.0):
        require_cuda()
        self.n, self.m = nspin, max(1, m)
        self.K = draws_per_eval(nspin, model)
        self.x = np.zeros(nspin + 1)
        self.rows = np.zeros((m, self.K)) if m > 0 else None
        self.fids = np.empty(self.m) if want_fids else None
        self.stats = np.empty(NUM_STATS) if want_stats else None
        self.amps = np.empty(self.m, dtype=np.complex128) if want_amps else None
        ptr = lambda a: a.ctypes.data if a is not None else None
        # one pointer argument per call instead of thirteen; stream 0: host buffers in and out, the call is synchronous
        self._frame = ObjectiveFrame(ptr(self.x), ptr(self.rows), ptr(self.fids), ptr(self.stats), ptr(self.amps), None,
                                     self.m, float(dkw_eps), nspin, inspin, outspin, model, int(bool(zz)), 0)
        self._fn = lib().rc_objective_call
        self._arg = C.byref(self._frame)

    def __call__(self, x, rows=None):
        """Evaluate; results are in self.fids / self.stats / self.amps (overwritten by the next call)."""
        self.x[:] = x
        if self.rows is not None:
            self.rows[:] = rows
        code = self._fn(self._arg)
        if code:
            check(code)
        return self


def objective_release() -> None:
    """Ask the calling thread's resident objective evaluator (rc_objective_host keeps one CTA on the device between
    calls, see include/robchar_b200.h) to leave now rather than after its idle time; waits for it."""
    check(lib().rc_objective_release())


class RobustnessSweepPlan:
    """Device buffers + workspace of rc_robustness_sweep for one problem shape, allocated once; run() issues the
    whole fig-4/5 sweep (evolution, statistics, top-k, Kendall matrices, ARIM bootstrap) from ONE C call with no
    allocation and no synchronisation — the device-resident twin of robustness_sweep_host."""

    def __init__(self, C_: int, S: int, B: int, nspin: int, inspin: int, outspin: int, *, groups: int = 1, topk: int = 100,
                 alpha_cluster: float = 0.05, dkw_eps: float = 0.0, model: int = MODEL_COMPLEX3, zz: bool = False,
                 fused: bool = False, nboot: int = 100, fids: torch.Tensor | None = None):
        dev = require_cuda()
        if C_ % groups:
            raise ValueError("controller count must be a multiple of the group count")
        self.shape = (C_, S, B, nspin, inspin, outspin)
        self.groups, self.topk, self.alpha, self.eps = groups, topk, float(alpha_cluster), float(dkw_eps)
        self.model, self.zz, self.fused, self.nboot = model, int(bool(zz)), int(bool(fused)), int(nboot)
        k = min(topk, C_ // groups)
        f64 = dict(dtype=torch.float64, device=dev)
        self.fids = None if fused else (fids if fids is not None else torch.empty((S, C_, B), **f64))
        self.stats = torch.empty((NUM_STATS, S, C_), **f64)
        self.tau = torch.empty((groups, S, S), **f64)
        self.sel = torch.empty((groups, k), dtype=torch.int64, device=dev)
        self.wsel = torch.empty((groups, S, k), **f64)
        self.arim = torch.empty((groups, S), **f64)
        self.arim_std = torch.empty((groups, S), **f64)
        self.counters = Counters(dev)
        self.wb = lib().rc_robustness_sweep_workspace_bytes(C_, S, B, self.fused, groups, topk)
        if self.wb == 0:
            raise ValueError("ranking problem too large")
        self.ws = torch.empty(self.wb, dtype=torch.uint8, device=dev)

    def run(self, ctrl: torch.Tensor, sigmas: torch.Tensor, *, seed: int = 0, c_offset: int = 0, b_offset: int = 0,
            evolution_events=None, stats: torch.Tensor | None = None):
        """ctrl [C][N+1], sigmas [S]: contiguous float64 CUDA tensors.  Returns (stats, tau); the other outputs are
        the plan's attributes (sel, wsel, arim, arim_std, fids, counters).  evolution_events: optional pair of
        torch.cuda.Event(enable_timing=True) recorded around the evolution launch (they must have been recorded
        once before so that their CUDA handles exist).  stats: optional contiguous float64 CUDA tensor [15][S][C] that
        receives the statistics instead of the plan's own (the multi-GPU exchange hands in its staging block)."""
        C_, S, B, nspin, inspin, outspin = self.shape
        if stats is None:
            stats = self.stats
        elif not (stats.is_cuda and stats.dtype == torch.float64 and stats.is_contiguous() and stats.numel() == NUM_STATS * S * C_):
            raise ValueError(f"stats must be a contiguous float64 CUDA tensor [{NUM_STATS}][{S}][{C_}]")
        if not (ctrl.is_cuda and ctrl.dtype == torch.float64 and ctrl.is_contiguous() and tuple(ctrl.shape) == (C_, nspin + 1)):
            raise ValueError(f"ctrl must be a contiguous float64 CUDA tensor [{C_}][{nspin + 1}]")
        if not (sigmas.is_cuda and sigmas.dtype == torch.float64 and sigmas.is_contiguous() and sigmas.numel() == S):
            raise ValueError(f"sigmas must be a contiguous float64 CUDA tensor [{S}]")
        check(lib().rc_robustness_sweep(_ptr(ctrl), C_, nspin, inspin, outspin, _ptr(sigmas), S, B, self.model, self.zz,
                                        C.c_uint64(seed & (2**64 - 1)), c_offset, b_offset, self.eps, self.fused,
                                        self.groups, self.topk, self.alpha, _ptr(self.fids), _ptr(stats),
                                        _ptr(self.tau), _ptr(self.sel), _ptr(self.wsel), self.nboot, _ptr(self.arim),
                                        _ptr(self.arim_std), self.counters.nonconv_ptr, _ptr(self.ws), self.wb,
                                        C.c_void_p(evolution_events[0].cuda_event if evolution_events else 0),
                                        C.c_void_p(evolution_events[1].cuda_event if evolution_events else 0), _stream()))
        return stats, self.tau


def arim_bootstrap_device(rims, nboot: int = 100, seed: int = 0):
    """(ARIM [R], bootstrap std [R]) of RIM rows [R][k] entirely on the device (rc_arim_bootstrap, Philox
    resampling indices)."""
    dev = require_cuda()
    r = _f64(rims, dev)
    r2 = r.reshape(-1, r.shape[-1])
    R, k = r2.shape
    a = torch.empty(R, dtype=torch.float64, device=dev)
    s = torch.empty(R, dtype=torch.float64, device=dev)
    check(lib().rc_arim_bootstrap(_ptr(r2), R, k, int(nboot), C.c_uint64(seed & (2**64 - 1)), _ptr(a), _ptr(s), _stream()))
    return a.reshape(r.shape[:-1]), s.reshape(r.shape[:-1])


def expm_batch(A) -> torch.Tensor:
    """expm of a batch of dense complex matrices [batch][M][M] (M <= 32) on the device (Pade-13 scaling and
    squaring): the generality path for non-tridiagonal / non-Hermitian Hamiltonians."""
    dev = require_cuda()
    if not isinstance(A, torch.Tensor):
        A = torch.as_tensor(np.asarray(A, dtype=np.complex128))
    A = A.to(device=dev, dtype=torch.complex128).contiguous()
    if A.dim() == 2:
        A = A[None]
    batch, M, M2 = A.shape
    if M != M2:
        raise ValueError("square matrices expected")
    out = torch.empty_like(A)
    check(lib().rc_expm_batch(_ptr(A), batch, M, _ptr(out), _stream()))
    return out


def dense_fidelity(H, T, inspin: int, outspin: int) -> torch.Tensor:
    """|expm(-1j*T*H)[out, in]|^2 for a batch of dense complex Hamiltonians [batch][N][N] and times [batch]
    (noise_model.py:105-109 verbatim, for Hamiltonians outside the tridiagonal fast path)."""
    dev = require_cuda()
    H = torch.as_tensor(np.asarray(H, dtype=np.complex128)) if not isinstance(H, torch.Tensor) else H
    H = H.to(device=dev, dtype=torch.complex128)
    if H.dim() == 2:
        H = H[None]
    Tt = torch.as_tensor(np.abs(np.asarray(T, dtype=np.float64)).reshape(-1)) if not isinstance(T, torch.Tensor) else T.abs().reshape(-1)
    Tt = Tt.to(device=dev, dtype=torch.float64)
    U = expm_batch(-1j * Tt[:, None, None] * H)
    phi = U[:, outspin, inspin]
    return phi.real * phi.real + phi.imag * phi.imag


def fidelity_grad(X, nspin: int, inspin: int, outspin: int, *, rows=None, zz: bool = False):
    """(err [C], grad [C][N+1]) of eval_static_fidelity_gradient (qnewton.py:162-212) for the controllers X [C][N+1]
    (host arrays in and out, rc_fidelity_grad_host): infidelity and its analytic gradient w.r.t. biases and time from
    the eigendecomposition, any N <= 32.  rows [C][2N]: optional explicit perturbations (real 2-draw replay layout)."""
    require_cuda()
    X = np.ascontiguousarray(np.asarray(X, dtype=np.float64).reshape(-1, nspin + 1))
    Cn = X.shape[0]
    rp = C.c_void_p(0)
    if rows is not None:
        rows = np.ascontiguousarray(np.asarray(rows, dtype=np.float64).reshape(Cn, 2 * nspin))
        rp = C.c_void_p(rows.ctypes.data)
    err, grad = np.empty(Cn), np.empty((Cn, nspin + 1))
    check(lib().rc_fidelity_grad_host(C.c_void_p(X.ctypes.data), Cn, nspin, inspin, outspin, rp, int(bool(zz)),
                                      C.c_void_p(err.ctypes.data), C.c_void_p(grad.ctypes.data), _stream()))
    return err, grad


def fp64_peak_tflops() -> float:
    require_cuda()
    v = C.c_double(0.0)
    check(lib().rc_fp64_peak_tflops(C.byref(v), _stream()))
    return v.value


def device_info():
    sm, maj, mnr = C.c_int(0), C.c_int(0), C.c_int(0)
    check(lib().rc_device_info(C.byref(sm), C.byref(maj), C.byref(mnr)))
    return sm.value, maj.value, mnr.value
