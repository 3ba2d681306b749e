"""RIM / 1-Wasserstein robustness measures with the reference's API, computed on the GPU.

Mirrors upstream ``wd_sortof_fast_implementation.py`` (check_fidtype :13-30, compute_dkw_error
:38-39, dkw_ecdf_bounds :41-79, wd_from_ideal :82-116, wd_from_ideal_zero :119-142, RIM_p
:147-174).  The scalar calls keep the upstream signatures (numpy array / list / scalar in, Python
float out, in-place sort of the argument); ``wd_from_ideal_batch`` is the segment-batched device
call the sweep uses.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch

from . import engine


def _as_fids(fids):
    """check_fidtype's coercion (wd_sortof_fast_implementation.py:15-20)."""
    if isinstance(fids, torch.Tensor):
        return fids
    if not isinstance(fids, np.ndarray):
        fids = np.array(fids) if isinstance(fids, list) else np.array([fids])
    return fids


def compute_dkw_error(alpha, nobs):
    return np.sqrt(np.log(2 / alpha) / (2 * nobs))


def dkw_ecdf_bounds(cdf, conf_level: float, visualize: bool = False) -> Tuple[np.ndarray, np.ndarray]:
    """wd_sortof_fast_implementation.py:41-79 (plotting dropped: no matplotlib on the path)."""
    cdf = _as_fids(cdf)
    if (np.abs(cdf - 1e-8) > 1).any():
        raise AssertionError("illegal fids values - must be in [0,1]")
    epsilon = compute_dkw_error(1 - conf_level, cdf.shape[-1])
    return np.clip(cdf - epsilon, 0, 1), np.clip(cdf + epsilon, 0, 1)


def wd_from_ideal_batch(fids, dkw_eps: float = 0.0, sort_inplace: bool = False) -> torch.Tensor:
    """[15, *lead] statistics for fids[*lead, B] on the device (row 0 is W(., delta(x-1)))."""
    return engine.stats(fids, dkw_eps, sort_inplace=sort_inplace)


def wd_from_ideal(fids, sort_fids: bool = True):
    """1-Wasserstein distance of the sample to delta(x-1) (wd_sortof_fast_implementation.py:82-116).
    Raises AssertionError for values outside [0,1]; sorts a numpy argument in place like upstream.
    ``sort_fids=False`` (caller promises sorted input) gives the same value."""
    f = _as_fids(fids)
    if isinstance(f, torch.Tensor):
        return float(engine.stats(f.reshape(1, -1), 0.0, sort_inplace=sort_fids)[0, 0].item())
    dev = engine.require_cuda()
    dev_f = torch.as_tensor(np.ascontiguousarray(f, dtype=np.float64).reshape(1, -1)).to(dev)
    st = engine.stats(dev_f, 0.0, sort_inplace=True)
    if sort_fids and isinstance(fids, np.ndarray) and fids.ndim == 1 and fids.flags.writeable:
        fids[...] = dev_f.reshape(-1).cpu().numpy().astype(fids.dtype, copy=False)
    return float(st[0, 0].item())


def wd_from_ideal_zero(fids, sort_fids: bool = True):
    """wd_sortof_fast_implementation.py:119-142."""
    return 1 - wd_from_ideal(fids, sort_fids)


def RIM_p(fids, p=2) -> float:
    """(mean((1-f)^p))^(1/p) (wd_sortof_fast_implementation.py:147-174), one device pass (rc_rim_p); p = 0 gives 1.
    Raises AssertionError for values outside [0,1] like upstream's check_fidtype."""
    f = _as_fids(fids)
    if p == 0:
        return 1
    t = f if isinstance(f, torch.Tensor) else torch.as_tensor(np.asarray(f, dtype=np.float64))
    return float(engine.rim_p(t.reshape(1, -1), p)[0].item())


def RIM_p_batch(fids, p=2) -> torch.Tensor:
    """RIM_p of every segment of fids[*lead, B] on the device: [*lead]."""
    return engine.rim_p(fids, p)
