"""ARIM — algorithm-level robustness infidelity measure, with bootstrap error bars, on the device.

Mirrors the compute parts of upstream ``generate_arim_all_fig5.py`` (ARIM_generator.get_ARIM
:54-196: ARIM(sigma) = wd_from_ideal_zero of the RIM vector of the top-k controllers, :119/:166;
error bar = bootstrap std over 100 resamples, :124/:169 via MCDataSim.bootstrap_resampling_std,
mcsim.py:267-275) and of ``gen_fig_8_arim_fcall_scaling.py`` (NStochOpt.get_rims :121-132).
"""
from __future__ import annotations

import numpy as np
import torch

from . import engine
from .mcsim import MCDataSim


def arim(rims) -> np.ndarray:
    """ARIM per row of a RIM matrix [S][k]: wd_from_ideal_zero(row) = 1 - W(row, delta(x-1)), through
    the device sort + W reduction (generate_arim_all_fig5.py:119)."""
    st = engine.stats(np.ascontiguousarray(np.asarray(rims, dtype=np.float64)), 0.0)
    return (1 - st[0]).cpu().numpy()


def bootstrap_indices_numpy(rows: int, k: int, bootsamples: int) -> np.ndarray:
    """The resampling indices upstream draws, in upstream's order (row by row, one
    np.random.randint(0, k, size=k) per bootstrap sample; mcsim.py:270-271)."""
    idx = np.empty((rows, bootsamples, k), dtype=np.int64)
    for j in range(rows):
        for i in range(bootsamples):
            idx[j, i] = np.random.randint(0, k, size=k)
    return idx


def arim_bootstrap(rims, bootsamples: int = 100, rng_mode: str = "numpy", seed: int = 0):
    """(ARIM [S], bootstrap std [S]) for a RIM matrix [S][k].
    rng_mode="numpy": consumes the global np.random stream exactly like upstream's
    bootstrap_resampling_std loop (indices drawn on the host, resampled statistics on the device);
    "device" (alias "torch"): Philox indices generated inside the kernel (rc_arim_bootstrap), seeded, one launch."""
    dev = engine.require_cuda()
    r = torch.as_tensor(np.ascontiguousarray(np.asarray(rims, dtype=np.float64))).to(dev)
    S, k = r.shape
    if rng_mode in ("device", "torch"):      # resampling indices from in-kernel Philox: one launch (rc_arim_bootstrap)
        a, s = engine.arim_bootstrap_device(r, nboot=bootsamples, seed=seed)
        return a.cpu().numpy(), s.cpu().numpy()
    if rng_mode != "numpy":
        raise ValueError("rng_mode must be 'numpy' or 'device'")
    centre = 1 - engine.stats(r.clone(), 0.0)[0]
    idx = torch.as_tensor(bootstrap_indices_numpy(S, k, bootsamples)).to(dev)
    res = torch.gather(r[:, None, :].expand(S, bootsamples, k), 2, idx).contiguous()   # index plumbing
    boot = 1 - engine.stats(res, 0.0)[0]                        # [S][bootsamples] ARIM of every resample
    std = engine.stats(boot.contiguous(), 0.0)[9]               # population std over the resamples (np.std)
    return centre.cpu().numpy(), std.cpu().numpy()


class ARIM_generator(MCDataSim):
    """Compute-only counterpart of upstream's ARIM_generator (generate_arim_all_fig5.py:40)."""

    def get_ARIM(self, algo: str, training_noise=None, plot_noises=None, bootsamples: int = 100, rng_mode: str = "numpy"):
        """ARIM(sigma_sim) and its bootstrap std for one controller group: top-k at sigma index 0,
        NaN columns dropped (generate_arim_all_fig5.py:101-126 / 151-171)."""
        if plot_noises is None:
            plot_noises = self.noises
        tn = None if algo == "lbfgs" else training_noise
        wd = self.get_metrics_dict(tn, plot_noises, algoname=algo)[algo]
        c = np.array(wd[engine.METRIC_W]); u = np.array(wd[engine.METRIC_W + " upper"]); l = np.array(wd[engine.METRIC_W + " lower"])
        if self.topk:
            c, u, l = self.get_top_k_by_fid(c, u, l, self.topk, fid_thres=None)
        wdd = c[~np.isnan(c)].reshape((len(plot_noises), -1))
        return arim_bootstrap(wdd, bootsamples, rng_mode=rng_mode)


class NStochOpt(MCDataSim):
    """gen_fig_8_arim_fcall_scaling.py:121-132: RIM(sigma) = 1 - mean fidelity over bootreps draws for
    one controller — row 0 of the fused device statistics."""

    def get_rims(self, cont, seed: int = 0):
        st = engine.fidelity_stats(np.asarray(cont, dtype=np.float64).reshape(1, -1), np.asarray(self.noises), self.bootreps,
                                   self.Nspin, self.inspin, self.outspin, seed=seed)
        self.noise_model.rng.args["scale"] = self.noises[-1]
        return st[0, :, 0].cpu().numpy()

    def get_rims_batch(self, conts, seed: int = 0):
        """All controllers at once: [C][S]."""
        st = engine.fidelity_stats(np.asarray(conts, dtype=np.float64), np.asarray(self.noises), self.bootreps, self.Nspin,
                                   self.inspin, self.outspin, seed=seed)
        return st[0].T.contiguous().cpu().numpy()
