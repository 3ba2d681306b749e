"""CPU-only tests: the C-ABI library loads and exports every declared symbol, the host-side mirror of
the reference interface behaves like the reference, and nothing computes without a GPU."""
import json
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden
from oracle import robchar_oracle as orc

import robchar_b200 as rb


def test_library_exports_every_declared_symbol():
    syms = rb._lib.declared_symbols()
    assert len(syms) >= 16 and "rc_fidelity_mc" in syms and "rc_mc_sweep_host" in syms
    lib = rb._lib.lib()
    for s in syms:
        assert hasattr(lib, s), s
    assert set(rb._lib._SIGNATURES) == set(syms)          # every declared entry point is bound
    assert lib.rc_version() >= 100
    assert isinstance(rb._lib.last_error(), str)


def test_header_cites_reference_interfaces():
    text = open(os.path.join(ROOT, "include", "robchar_b200.h")).read()
    for cite in ["mcsim.py:422", "noise_model.py:98", "wd_sortof_fast_implementation.py:82", "mcsim.py:513",
                 "generate_fig4_kendallrankanalysis.py:146", "qnewton.py:366"]:
        assert cite in text, cite
    assert "#include <torch" not in text and "at::Tensor" not in text     # plain pointers and sizes only


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(rb._lib.RobcharLibraryError):
        rb.engine.fidelity_mc(np.zeros((1, 5)), [0.0], 1, 4, 0, 2)
    with pytest.raises(rb._lib.RobcharLibraryError):
        rb.wd_from_ideal(np.array([0.5, 0.6]))
    m = rb.structured_perturbation(Nspin=4, inspin=0, outspin=2)
    with pytest.raises(rb._lib.RobcharLibraryError):
        m.evaluate_noisy_fidelity(np.zeros(5), True)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "code-robchar_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), fn
                assert "/root/reference" not in src, fn


def test_status_codes_map_to_reference_exceptions():
    lib = rb._lib
    with pytest.raises(AssertionError):
        lib.check(lib.RC_ERR_ILLEGAL_FIDS)            # wd_sortof_fast_implementation.py:25
    with pytest.raises(ValueError):
        lib.check(lib.RC_ERR_BAD_ARG)
    with pytest.raises(lib.EigensolverNonConvergence):
        lib.check(lib.RC_ERR_NONCONV)
    with pytest.raises(lib.RobcharLibraryError):
        lib.check(lib.RC_ERR_CUDA)
    lib.check(lib.RC_OK)


def test_noise_function_semantics():
    """noise_model.py:21-46: every call updates the stored kwargs and draws."""
    calls = []
    nf = rb.noise_function(lambda **kw: calls.append(dict(kw)) or kw.get("scale", 0), scale=0.02)
    assert nf() == 0.02
    assert nf(scale=0.07) == 0.07 and nf.args == {"scale": 0.07}
    assert nf(size=2) == 0.07 and calls[-1] == {"scale": 0.07, "size": 2}
    np.random.seed(3)
    g = rb.noise_function(np.random.normal, scale=0.05)
    a = [g() for _ in range(3)]
    np.random.seed(3)
    assert a == list(0.05 * np.random.standard_normal(3))


def test_structured_perturbation_draw_order_matches_reference_port():
    """noise_model.py:135-147: (z_ii, nn_i, nn2_i) per site, site-0 couplings discarded."""
    n = 6
    m = rb.structured_perturbation(Nspin=n, inspin=0, outspin=3, noise=0.05)
    assert m.HH.dtype == np.complex128 and len(m.CC) == n and m.CC[2][2, 2] == 1
    np.random.seed(9)
    z = m.perturbation()
    stream = np.random.RandomState(9).standard_normal(3 * n) * 0.05
    want = orc.perturbation_from_draws(stream, n)
    assert np.array_equal(z, want)
    row = m._replay_row_from_matrix(z)
    used = np.ones(3 * n, bool); used[[1, 2]] = False
    assert np.array_equal(row[used], stream[used]) and np.all(row[~used] == 0)
    with pytest.raises(NotImplementedError):
        bad = z.copy(); bad[0, 3] = 1.0
        m._replay_row_from_matrix(bad)
    ring = rb.structured_perturbation(Nspin=4, topo="ring")
    assert ring.HH[3, 0] == 1 and ring.HH[0, 3] == 1
    d = rb.directional_perturbation(Nspin=5, outspin=4)
    assert len(d.directions) == 2 + 3 * 3 + 4
    np.random.seed(1)
    zz = d.perturbation()
    assert np.count_nonzero(zz) in (1, 2)


def test_experiment_namer_and_mc_file_names(tmp_path, monkeypatch):
    """File-name grammar of mcsim.py:351-356 / noise_analysis.py:33-49 resolves the reference's files."""
    monkeypatch.chdir(tmp_path)
    nm = rb.noise_analysis.ExperimentNamer(experiment_name="pipeline_nmplus2", Nspin=4, inspin=0, outspin=2,
                                           numcontrollers=1000)
    assert nm() == "experiments/pipeline_nmplus2/ppo_spin_4_0-2_c_1000"
    assert os.path.isdir("experiments/pipeline_nmplus2")
    json.dump({"lbfgs": {"4": {"controller": [[0.0] * 5]}}, "empty": {}},
              open("experiments/pipeline_nmplus2/ppo_spin_4_0-2_c_1000.le", "w"))
    sim = rb.MCDataSim(experiment_name="pipeline_nmplus2", Nspin=4, inspin=0, outspin=2, bootreps=1,
                       numcontrollers=1000, filemarker=".le")
    assert sim.algos == ["lbfgs"]                                      # empty containers purged (mcsim.py:339-343)
    # the exact name of a cache file shipped with the reference (experiments/pipeline_nmplus2/)
    want = ("experiments/pipeline_nmplus2/ppo_spin_4_0-2_c_1000.le_tn0.01_br_1_nlvl"
            "[0.   0.01 0.02 0.03 0.04 0.05 0.06 0.07 0.08 0.09 0.1 ].mc")
    assert sim.get_mcname(0.01, np.linspace(0, 0.1, 11)) == want
    assert sim.get_mcname(None, np.linspace(0, 0.1, 11)).startswith("experiments/pipeline_nmplus2/ppo_spin_4_0-2_c_1000.le_tnNone_br_1_")
    c = sim._controller_matrix("lbfgs", None)
    assert c.shape == (1000, 5) and np.isnan(c[1:]).all() and not np.isnan(c[0]).any()   # NaN padding (mcsim.py:436-443)
    missing = rb.MCDataSim(experiment_name="nothing_here", Nspin=4)
    assert missing.controllers is None and missing.algos is None       # FileNotFoundError swallowed (mcsim.py:231-237)
    with pytest.raises(rb.noise_analysis.DirectoryDoesNotExistError):
        sim.get_path("does_not_exist")
    with pytest.raises(TypeError):
        sim.ctrlnames(3)


def test_numpy_stream_replay_layout(tmp_path, monkeypatch):
    """rng_mode='numpy' consumes np.random like mcsim.py:424-456 (one discarded draw per level, none for NaN rows)."""
    monkeypatch.chdir(tmp_path)
    sim = rb.MCDataSim(experiment_name="x", Nspin=4, inspin=0, outspin=2, bootreps=3, numcontrollers=3, rng_mode="numpy")
    ctrl = np.zeros((3, 5)); ctrl[1] = np.nan
    np.random.seed(5)
    z = sim._numpy_stream_replay(ctrl, np.array([0.0, 0.1]))
    rs = np.random.RandomState(5)
    for s in range(2):
        rs.standard_normal()
        for c in (0, 2):
            assert np.array_equal(z[s, c], rs.standard_normal((3, 12)))
    assert np.all(z[:, 1] == 0)
    assert sim.noise_model.rng.args["scale"] == 0.1


def test_metric_registry_names_and_helpers():
    assert list(rb.mcsim.__metric_name_to_metric__) == orc.METRIC_NAMES
    assert rb.engine.STAT_KEYS[:3] == [orc.METRIC_W, orc.METRIC_W + " upper", orc.METRIC_W + " lower"]
    g = load_golden("kat_mcm_pairs.npz")
    assert [str(s) for s in g["names"]] == rb.engine.STAT_KEYS       # key order of the reference's .mcm files
    assert rb.compute_dkw_error(0.05, 100) == orc.compute_dkw_error(0.05, 100)
    cdf, srt = rb.mcsim.get_cdf(np.array([0.3, 0.1, 0.6]))
    assert np.allclose(cdf, [0.1, 0.4, 1.0]) and np.array_equal(srt, [0.1, 0.3, 0.6])
    with pytest.raises(TypeError):
        rb.mcsim.get_cdf([0.1, 0.2])
    assert rb.mcsim.Q(np.array([0.9, 0.96, 0.99]), 0.95) == 2 / 3
    rs = np.random.RandomState(0)
    assert rb.mcsim.vn_test(rs.normal(0, 1, 50000), verbose=False)[0] is True       # mcsim.py:126-130
    assert rb.mcsim.vn_test(np.arange(1000.0), verbose=False)[0] is False
    lo, hi = rb.wd_sortof_fast_implementation.dkw_ecdf_bounds(np.array([0.2, 0.5, 0.9]), 0.95)
    eps = orc.compute_dkw_error(0.05, 3)
    assert np.allclose(lo, np.clip(np.array([0.2, 0.5, 0.9]) - eps, 0, 1)) and np.all(hi <= 1)


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 19000, 100001):
        for w in (1, 2, 3, 8):
            spans = [rb.dist.shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_reference_module_aliases():
    import sys
    saved = {k: sys.modules.get(k) for k in ("mcsim", "noise_model", "wd_sortof_fast_implementation", "noise_analysis")}
    try:
        rb.install_reference_module_aliases()
        from mcsim import MCDataSim            # noqa: the unmodified analysis scripts' import lines
        from noise_model import structured_perturbation  # noqa
        from wd_sortof_fast_implementation import wd_from_ideal, compute_dkw_error  # noqa
        assert MCDataSim is rb.MCDataSim
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def _oracle_objective_host(x, rows, nspin, inspin, outspin, *, model=0, zz=False, want_fids=True, want_stats=False,
                           dkw_eps=0.0, want_amps=False):
    """Stand-in for engine.objective_host built on the CPU oracle's Hamiltonian + scipy expm (test infrastructure):
    lets the HOST logic of the optimiser-facing mirrors (draw order, action wrapping, shot noise, propagator
    averaging) be checked against the reference goldens without a GPU."""
    import scipy.linalg
    x = np.asarray(x, dtype=float)
    m = 1 if rows is None else rows.shape[0]
    amps = np.zeros(m, dtype=complex)
    for k in range(m):
        H = orc.base_hamiltonian(nspin, "chain", zz)
        if rows is not None:
            H = H + orc.perturbation_from_draws(rows[k], nspin, model)
        H = H + np.diag(x[:nspin])
        amps[k] = scipy.linalg.expm(-1j * abs(x[nspin]) * H)[outspin, inspin]
    f = np.abs(amps) ** 2
    st = None
    if want_stats:
        st = np.full(15, np.nan)
        st[0] = orc.wd_from_ideal(f.copy())
    res = (f if want_fids else None,) + ((st,) if want_stats else ()) + ((amps,) if want_amps else ())
    return res if len(res) > 1 else res[0]


class _OracleObjectiveEvaluator:
    """Stand-in for engine.ObjectiveEvaluator (the preallocated call frame of rc_objective_host) on the CPU oracle."""

    def __init__(self, nspin, inspin, outspin, m, *, model=0, zz=False, want_fids=True, want_stats=False, want_amps=False,
                 dkw_eps=0.0):
        self.a = (nspin, inspin, outspin)
        self.kw = dict(model=model, zz=zz, want_fids=True, want_stats=want_stats, dkw_eps=dkw_eps, want_amps=want_amps)
        self.want_fids, self.want_stats, self.want_amps = want_fids, want_stats, want_amps
        self.fids = self.stats = self.amps = None

    def __call__(self, x, rows=None):
        res = _oracle_objective_host(x, None if rows is None else np.asarray(rows), *self.a, **self.kw)
        res = res if isinstance(res, tuple) else (res,)
        self.fids = res[0]
        k = 1
        if self.want_stats:
            self.stats = res[k]; k += 1
        if self.want_amps:
            self.amps = res[k]
        return self


def test_rl_environment_host_logic_matches_reference(monkeypatch):
    """The Environment mirror's host side (RNG draw order incl. reset's hidden draws, bias accumulation and
    wrapping into the bounds, time wrapping, binomial / adaptive shot noise, mean-propagator reward, transfer-
    learning couplings) against the goldens recorded from the unmodified reference; the device evaluation is
    replaced by the oracle stand-in above (the GPU version of this test is test_rl_environment_matches_reference)."""
    import scipy.linalg
    import torch
    monkeypatch.setattr(rb.engine, "objective_host", _oracle_objective_host)
    monkeypatch.setattr(rb.engine, "expm_batch", lambda A: torch.as_tensor(np.array([scipy.linalg.expm(a) for a in np.asarray(A)])))
    g = load_golden("rl_env.npz")
    acts, times = g["env_acts"], g["env_times"]
    n, i, o = (int(v) for v in g["env_meta"])
    Env = rb.RLreinforceXXchain_actionedtime.Environment

    def drive(env, seed):
        np.random.seed(seed)
        env.reset()
        rew, tf, act = [], [], []
        for a, t in zip(acts, times):
            env.timestep = t
            ao, r, d = env.step(np.diag(a))
            rew.append(np.real(r)); tf.append(np.real(env.tf)); act.append(np.diag(ao).copy())
        return np.array(rew), np.array(tf), np.array(act)

    for name, kw in (("plain", {}), ("hamnoisy", dict(ham_noisy=True)), ("shot", dict(fid_noisy=True, draws=20)),
                     ("adaptive", dict(fid_noisy=True, adaptive=True, draws=20)), ("heis", dict(heisenberg_int=True)),
                     ("fixed", dict(use_fixed_ham=True)), ("ring", dict(topo="ring"))):
        env = Env(n, i, o, noise=0.05, opt_train_size=12, opt_test_size=50, **kw)
        r, tf, a = drive(env, 17)
        assert np.abs(r - g[f"env_{name}_reward"]).max() < 1e-12, name
        assert np.abs(tf - g[f"env_{name}_tf"]).max() < 1e-12, name
        assert np.abs(a - g[f"env_{name}_action"]).max() < 1e-12, name
    np.random.seed(5)
    env = Env(n, i, o, noise=0.05, opt_train_size=4, opt_test_size=10, transfer_learning=True)
    assert np.array_equal(env.sys, g["env_tl_sys"])
    r, tf, _ = drive(env, 18)
    assert np.abs(r - g["env_tl_reward"]).max() < 1e-12 and np.abs(tf - g["env_tl_tf"]).max() < 1e-12


def test_lbfgs_objective_host_logic_matches_reference(monkeypatch):
    """Host side of the qnewton.LBFGS mirror (fixed Hamiltonian sets from seed 4, perturbation draw order, binomial
    and adaptive shot noise, W1 objective) against the goldens recorded from the unmodified reference, with the
    device evaluation replaced by the oracle stand-in (GPU version: test_optimiser_objectives_match_reference)."""
    monkeypatch.setattr(rb.engine, "objective_host", _oracle_objective_host)
    monkeypatch.setattr(rb.engine, "ObjectiveEvaluator", _OracleObjectiveEvaluator)
    g = load_golden("objective_arim.npz")
    n, i, o, train = (int(v) for v in g["obj_meta"])
    env = rb.qnewton.LBFGS(n, i, o, noise=0.05, opt_train_size=train, opt_test_size=200)
    assert np.array_equal(env.randH[0], g["obj_randH0"])
    X = g["obj_X"]
    av10 = np.array([env.fidelity_ss_av(x, reps=10) for x in X])
    assert np.abs(av10 - g["obj_av10"]).max() < 1e-12
    np.random.seed(11)
    w = np.array([env.wass_cost(x, 7) for x in X])
    assert np.abs(w - g["obj_wass7"]).max() < 1e-12
    np.random.seed(12)
    shot = np.array([env.fidelity_ss(x, noisy=True, ham_noisy=True) for x in X])
    assert np.array_equal(shot, g["obj_shot"])
    env.adaptive = True
    np.random.seed(13)
    ad = np.array([env.fidelity_ss(x, noisy=True, ham_noisy=False) for x in X])
    assert np.abs(ad - g["obj_adaptive"]).max() < 1e-12


def test_oracle_port_equals_the_staged_unmodified_reference():
    """The oracle's per-sample port (bench.py's fallback CPU arm) against the UNMODIFIED reference staged by build()
    in oracle/_ref (the default CPU arm): same numpy seed, same controller -> the same fidelities, draw for draw."""
    from oracle import ref_runner
    if not ref_runner.available():
        pytest.skip("reference not staged (no /root/reference at build time)")
    ref_nm, _, _ = ref_runner._import_reference()
    for n, i, o in [(4, 0, 2), (7, 0, 6)]:
        x = orc.synthetic_controllers(3, n, seed=n)
        np.random.seed(5)
        ref = ref_nm.structured_perturbation(Nspin=n, inspin=i, outspin=o, noise=0.05)
        want = [ref.evaluate_noisy_fidelity(x[k % 3], True) for k in range(12)]
        np.random.seed(5)
        port = orc.ReferencePathPort(n, i, o, 0.05)
        got = [port.evaluate_noisy_fidelity(x[k % 3], True) for k in range(12)]
        assert np.abs(np.array(got) - np.array(want)).max() < 1e-13
