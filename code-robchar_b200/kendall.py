"""Rank-consistency analysis (Kendall tau between simulation-noise levels) on the device.

Mirrors the compute parts of upstream ``generate_fig4_kendallrankanalysis.py``: the nested helpers
``get_top_k_by_fid`` (:61-70), ``jkt_or_ordinaltau`` (:72-92), ``jkt_or_ordinaltau_pairwise``
(:94-120) and ``get_ranks_clustered_little`` (:146-164) of ``KTRConsitency.plot_kendalltaus``.
Plotting is out of scope.
"""
from __future__ import annotations

import numpy as np

from . import engine
from .mcsim import MCDataSim


def get_ranks_clustered_little(infids, r: float = -1e-15) -> np.ndarray:
    """1-d cluster ranks with discrepancy radius r (…fig4…py:146-164)."""
    return engine.clustered_ranks(np.asarray(infids, dtype=np.float64), r=r).cpu().numpy()


def jkt_or_ordinaltau(wd_data_c, r: float = 1e-3) -> list:
    """tau between clustered ranks of row 0 and ordinal ranks of every row (…fig4…py:72-92;
    the VN independence test there only prints)."""
    w = np.asarray(wd_data_c, dtype=np.float64)
    cr = engine.clustered_ranks(w[:1], r=r)
    rk = engine.ranks(w) + 1
    return engine.kendall_tau_b(cr, rk).reshape(-1).cpu().numpy().tolist()


def jkt_or_ordinaltau_pairwise(wd_data_c, alpha: float = 0.05) -> list:
    """S x S tau matrix: clustered ranks (radius alpha*(max-min)) of row j vs ordinal ranks of row i
    (…fig4…py:94-120)."""
    return engine.kendall_matrix(np.asarray(wd_data_c, dtype=np.float64), alpha=alpha).cpu().numpy().tolist()


class KTRConsitency(MCDataSim):
    """Compute-only counterpart of upstream's KTRConsitency (…fig4…py:12): for one controller group,
    the top-k selection at sigma_sim index 0 and the Kendall matrix."""

    def kendall_taus(self, training_noise, algoname: str, topk: int = None, fid_thres=None, alpha: float = 0.05):
        topk = self.topk if topk is None else topk
        wd = self.get_metrics_dict(training_noise, self.noises, algoname=algoname)[algoname]
        c = np.array(wd[engine.METRIC_W]); u = np.array(wd[engine.METRIC_W + " upper"]); l = np.array(wd[engine.METRIC_W + " lower"])
        c, u, l = self.get_top_k_by_fid(c, u, l, topk, fid_thres=fid_thres)
        return np.array(jkt_or_ordinaltau_pairwise(c, alpha=alpha)), c
