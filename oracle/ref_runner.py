"""Runs the UNMODIFIED reference (qyber-black/Code-RobChar) on the host CPU, for timing and cross-checks.

THIS IS TEST / BENCHMARK INFRASTRUCTURE, NOT PRODUCT CODE.  The reference's own .py files are never committed:
``__graft_entry__.build()`` packs them UNMODIFIED, where ``/root/reference`` is mounted (the build container), into
one archive ``oracle/_ref/reference_modules.zip`` (imported with zipimport) inside the git-ignored ``oracle/_ref/`` —
which travels to the GPU box with the snapshot like the built ``.so`` does.  Only
``bench.py`` (``cpu_baseline`` / ``--impl reference`` legs) and ``tests/`` import this module.

The reference imports plotting / optimiser packages that are not installed here (matplotlib, seaborn, IPython,
skquant, SQSnobFit); they are stubbed with MagicMock exactly as tests/golden/make_golden.py does — none of them is
on the path that is timed (noise_model.py:98-147, mcsim.py:422-510, wd_sortof_fast_implementation.py:82-116).
"""
from __future__ import annotations

import json
import os
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
REF_ZIP = os.path.join(REF_DIR, "reference_modules.zip")
STAGED_DATA = ["noisy_analysis/lbfgs_spin_4_0-2_in", "noisy_analysis/lbfgs_spin_7_0-6_in"]
_STUBS = ["matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.ticker", "seaborn", "IPython",
          "IPython.display", "skquant", "skquant.opt", "SQSnobFit"]


def stage(reference_root: str = "/root/reference") -> bool:
    """Pack the reference's top-level .py files into oracle/_ref/reference_modules.zip and copy two small controller
    files next to it (build container only).  Returns True when the staged tree is usable."""
    import shutil
    import zipfile
    if not os.path.isdir(reference_root):
        return available()
    os.makedirs(os.path.join(REF_DIR, "noisy_analysis"), exist_ok=True)
    tmp = REF_ZIP + ".tmp"
    with zipfile.ZipFile(tmp, "w", zipfile.ZIP_DEFLATED) as zf:
        for fn in sorted(os.listdir(reference_root)):
            if fn.endswith(".py") and fn != "rim_analysis.py":    # rim_analysis.py plots at import; never needed
                zf.write(os.path.join(reference_root, fn), fn)
    os.replace(tmp, REF_ZIP)
    for fn in os.listdir(REF_DIR):                                 # loose copies of an earlier staging layout
        if fn.endswith(".py"):
            os.remove(os.path.join(REF_DIR, fn))
    for rel in STAGED_DATA:
        src = os.path.join(reference_root, rel)
        if os.path.exists(src):
            shutil.copyfile(src, os.path.join(REF_DIR, rel))
    return available()


def available() -> bool:
    return os.path.exists(REF_ZIP)


def _import_reference(full: bool = False):
    """noise_model (always) and, with `full`, mcsim / wd_sortof_fast_implementation of the staged reference."""
    sys.dont_write_bytecode = True
    if REF_ZIP not in sys.path:
        sys.path.insert(0, REF_ZIP)
    import noise_model as ref_nm
    if not full:
        return ref_nm, None, None
    from unittest.mock import MagicMock
    for name in _STUBS:
        sys.modules.setdefault(name, MagicMock())
    import mcsim as ref_mc
    import wd_sortof_fast_implementation as ref_wd
    return ref_nm, ref_mc, ref_wd


def synthetic_controllers(C, nspin, seed=20221):
    rs = np.random.RandomState(seed)
    ctrl = np.empty((C, nspin + 1))
    ctrl[:, :nspin] = rs.uniform(-10, 10, (C, nspin))
    ctrl[:, nspin] = rs.uniform(1, 30, C)
    return ctrl


def eval_worker(args):
    """One worker of the CPU arm: `nevals` calls of the reference's own
    structured_perturbation.evaluate_noisy_fidelity(x, True) (noise_model.py:98-147), after a 200-call warm-up."""
    nspin, inspin, outspin, sigma, nevals, seed = args
    os.environ["OMP_NUM_THREADS"] = "1"
    ref_nm, _, _ = _import_reference()
    np.random.seed(seed)
    model = ref_nm.structured_perturbation(Nspin=nspin, inspin=inspin, outspin=outspin, noise=sigma)
    ctrl = synthetic_controllers(8, nspin, seed=seed)
    for k in range(200):
        model.evaluate_noisy_fidelity(ctrl[k % 8], True)
    t0 = time.perf_counter()
    acc = 0.0
    for k in range(nevals):
        acc += model.evaluate_noisy_fidelity(ctrl[k % 8], True)
    return time.perf_counter() - t0, acc


def mcdatasim_wall(nspin, inspin, outspin, controllers, noises, bootreps, seed=1):
    """Wall time of the reference's MCDataSim.get_metrics_dict from scratch (mcsim.py:463-510: the triple loop, the
    JSON dumps of .mc / .mcm, the metric maps) in a temporary experiments/ tree, single process as a user runs it.
    Returns (seconds, W matrix [S][C])."""
    _, ref_mc, _ = _import_reference(full=True)
    cwd = os.getcwd()
    C = len(controllers)
    with tempfile.TemporaryDirectory() as td:
        os.makedirs(f"{td}/experiments/bench")
        json.dump({"lbfgs": {str(nspin): {"controller": [list(map(float, c)) for c in controllers]}}},
                  open(f"{td}/experiments/bench/ppo_spin_{nspin}_{inspin}-{outspin}_c_{C}", "w"))
        os.chdir(td)
        try:
            sim = ref_mc.MCDataSim(experiment_name="bench", Nspin=nspin, inspin=inspin, outspin=outspin, noises=noises,
                                   bootreps=bootreps, numcontrollers=C, topk=min(100, C))
            np.random.seed(seed)
            t0 = time.perf_counter()
            metrics = sim.get_metrics_dict(None, noises, algoname="lbfgs")
            wall = time.perf_counter() - t0
        finally:
            os.chdir(cwd)
    W = np.array(metrics["lbfgs"][r'$W(.,\delta(x-1))$'], dtype=np.float64)
    return wall, W


def stage_timings(B=100, n_rank=100):
    """wd_from_ideal on B samples and scipy.stats.kendalltau on n_rank ranks, microseconds per call (the statistics /
    ranking stages of the reference, wd_sortof_fast_implementation.py:82-116, generate_fig4_kendallrankanalysis.py:117)."""
    from scipy.stats import kendalltau
    _, _, ref_wd = _import_reference(full=True)
    rs = np.random.RandomState(0)
    x = rs.uniform(0.5, 1.0, (2000, B))
    t0 = time.perf_counter()
    for row in x:
        ref_wd.wd_from_ideal(row)
    t_wd = (time.perf_counter() - t0) / len(x)
    a = rs.permutation(n_rank); b = rs.permutation(n_rank)
    t0 = time.perf_counter()
    for _ in range(200):
        kendalltau(a, b)
    t_k = (time.perf_counter() - t0) / 200
    return {"wd_from_ideal_us": t_wd * 1e6, "wd_samples": B, "kendalltau_us": t_k * 1e6, "kendall_n": n_rank}


def lbfgs_controllers(nspin, inspin, outspin):
    path = os.path.join(REF_DIR, f"noisy_analysis/lbfgs_spin_{nspin}_{inspin}-{outspin}_in")
    rec = json.load(open(path, "rb"))["lbfgs"][str(nspin)]["controller"]
    return [c for c in rec if c is not None]
