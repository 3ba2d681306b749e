"""The RIM / Wasserstein robustness call named by the north star.

Upstream ``rim_analysis.py`` is a toy plotting script that executes at import
(rim_analysis.py:59,94-96); the real robustness functions live in
``wd_sortof_fast_implementation.py:82-174``.  This module exposes those, plus the order-p helper
of the toy script (rim_analysis.py:62-81) and the sweep-level entry point.
"""
from __future__ import annotations

import numpy as np

from . import engine
from .wd_sortof_fast_implementation import (RIM_p, compute_dkw_error, dkw_ecdf_bounds, wd_from_ideal,  # noqa: F401
                                            wd_from_ideal_batch, wd_from_ideal_zero)


def p_order_rim(fids, orders=(1, 2, 3, 4)):
    """rim_analysis.py:62-81: RIM_p over a list of orders."""
    return [RIM_p(np.asarray(fids), p) for p in orders]


def rim_sweep(controllers, noises, bootreps: int, Nspin: int, inspin: int, outspin: int, *, alpha: float = 0.05,
              seed: int = 0, fused: bool = True, model: int = engine.MODEL_COMPLEX3, zz: bool = False):
    """RIM (and the other 14 statistics) of every controller at every simulation noise level:
    returns a dict key -> ndarray [S][C] with the reference's metric names (mcsim.py:178-183)."""
    eps = float(compute_dkw_error(alpha, bootreps))
    if fused:
        st = engine.fidelity_stats(controllers, noises, bootreps, Nspin, inspin, outspin, dkw_eps=eps, seed=seed,
                                   model=model, zz=zz)
    else:
        f = engine.fidelity_mc(controllers, noises, bootreps, Nspin, inspin, outspin, seed=seed, model=model, zz=zz)
        st = engine.stats(f, eps)
    st = st.cpu().numpy()
    return {k: st[i] for i, k in enumerate(engine.STAT_KEYS)}


def robustness_sweep(controllers: np.ndarray, noises: np.ndarray, bootreps: int, Nspin: int, inspin: int, outspin: int,
                     *, groups: int = 1, topk: int = 100, alpha_dkw: float = 0.05, alpha_cluster: float = 0.05,
                     seed: int = 0, fused: bool = False, model: int = engine.MODEL_COMPLEX3, zz: bool = False,
                     nboot: int = 100, c_offset: int = 0, b_offset: int = 0, copy: bool = True) -> dict:
    """The paper's fig-4/5 sweep for `groups` controller sets given as HOST arrays: evolution,
    the 15 statistics, per-group top-k selection and Kendall matrices.  Host in, host out: the
    controllers travel to the device and the statistics / tau matrices come back (the end-to-end
    call bench.py times).  Returns {"stats": {key: [S][C]}, "tau": [G][S][S], "topk_idx": [G][k], "arim": [G][S],
    "arim_std": [G][S]} (ARIM and its bootstrap error bar over the top-k controllers, fig 5).
    c_offset / b_offset: global index of the first controller / draw (Philox counters), for callers that shard a larger
    sweep themselves.  The arrays are the caller's own (copies); copy=False returns views of the engine's cached pinned
    staging buffers instead, which the NEXT call with the same shapes overwrites (benchmark loops only)."""
    eps = float(compute_dkw_error(alpha_dkw, bootreps))
    st, tau, sel, ar, ars = engine.robustness_sweep_host(np.asarray(controllers), np.asarray(noises), bootreps, Nspin,
                                                         inspin, outspin, groups=groups, topk=topk,
                                                         alpha_cluster=alpha_cluster, dkw_eps=eps, seed=seed, fused=fused,
                                                         model=model, zz=zz, nboot=nboot, c_offset=c_offset, b_offset=b_offset)
    if copy:
        st, tau, sel, ar, ars = (np.array(a) for a in (st, tau, sel, ar, ars))
    return {"stats": {k: st[i] for i, k in enumerate(engine.STAT_KEYS)}, "tau": tau, "topk_idx": sel, "arim": ar,
            "arim_std": ars}
