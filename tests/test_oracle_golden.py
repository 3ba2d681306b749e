"""Pins the CPU oracle (oracle/robchar_oracle.py) against the reference's own stored results and
against outputs of the unmodified reference run on seeded streams (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import scipy.stats

from oracle import robchar_oracle as orc
from conftest import load_golden


def test_bestfid_kats():
    """noisy_analysis/lbfgs_spin_*_in: best_fid == nominal fidelity of controller (757 KATs)."""
    g = load_golden("kat_bestfid.npz")
    total = 0
    for key, n, i, o in g["meta"]:
        n, i, o = int(n), int(i), int(o)
        ctrl = g[key + "_ctrl"]
        f = orc.fidelity_batch(ctrl, n, i, o)
        assert np.abs(f - g[key + "_best_fid"]).max() < 5e-13, key
        total += len(ctrl)
    assert total >= 700


def test_mc_zero_noise_rows():
    """sigma_sim = 0 rows of the reference's .mc caches vs the .le controllers."""
    g = load_golden("kat_mc_zero.npz")
    for key, n, i, o in g["meta"]:
        n, i, o = int(n), int(i), int(o)
        ctrl = g[key + "_ctrl"]
        f = orc.fidelity_batch(ctrl, n, i, o)
        ref = g[key + "_fid0"]
        ok = ~np.isnan(ref)
        assert np.array_equal(np.isnan(f), np.isnan(ref)), key
        assert np.abs(f[ok] - ref[ok]).max() < 5e-13, key


def test_single_sample_matches_batch():
    rs = np.random.RandomState(3)
    ctrl = orc.synthetic_controllers(5, 6)
    z = rs.standard_normal((5, 18)) * 0.05
    fb = orc.fidelity_batch(ctrl, 6, 0, 3, z)
    for c in range(5):
        fs = orc.evaluate_fidelity(ctrl[c], 6, 0, 3, z[c])
        assert abs(fs - fb[c]) < 1e-13


@pytest.mark.parametrize("name", ["replay_n4_0_2", "replay_n5_0_4", "replay_n6_0_3", "replay_n7_0_6"])
def test_replay_against_reference_run(name):
    """Unmodified MCDataSim.get_metrics_dict on a seeded stream: fidelities, 15 metrics, ranks,
    clustered ranks and the Kendall matrix."""
    g = load_golden(name + ".npz")
    n, i, o = (int(v) for v in g["nio"])
    fids = orc.fidelity_mc_replay(g["ctrl"], g["sigmas"], g["normals"], n, i, o)
    ref = g["fids"]
    assert np.array_equal(np.isnan(fids), np.isnan(ref))
    ok = ~np.isnan(ref)
    assert np.abs(fids[ok] - ref[ok]).max() < 1e-12
    m = orc.metrics(ref.copy(), float(g["alpha"]))
    for k, name_k in enumerate(g["metric_names"]):
        a, b = m[str(name_k)], g["metrics"][k]
        assert np.array_equal(np.isnan(a), np.isnan(b)), name_k
        okm = ~np.isnan(b)
        assert np.abs(a[okm] - b[okm]).max() < 1e-14, name_k
    W = g["metrics"][list(g["metric_names"]).index(orc.METRIC_W)]
    Wc = W[:, ~np.isnan(W[0])]
    assert np.array_equal(orc.get_ranks(Wc[0]), g["ranks_row0"])
    r3 = 0.05 * (Wc[3].max() - Wc[3].min())
    assert np.array_equal(orc.get_ranks_clustered_little(Wc[3], r=r3), g["clustered_row3"])
    K = orc.kendall_matrix(Wc, alpha=0.05)
    assert np.array_equal(K, g["kendall"])  # integer pair counts + IEEE sqrt/div: bit-exact


def test_real2_and_zz_variants():
    g = load_golden("replay_real2_zz.npz")
    for tag in ("n5", "n6zz", "n16zz"):
        n, i, o, zz = (int(v) for v in g[tag + "_meta"])
        ctrl = g[tag + "_ctrl"]
        scaled = g[tag + "_normals"] * float(g[tag + "_sigma"])
        f = orc.fidelity_batch(ctrl[:, None, :], n, i, o, scaled, orc.MODEL_REAL2, bool(zz))
        assert np.abs(f - g[tag + "_fids"]).max() < 1e-12, tag
        f0 = orc.fidelity_batch(ctrl, n, i, o, None, orc.MODEL_REAL2, bool(zz))
        assert np.abs(f0 - g[tag + "_nominal"]).max() < 1e-12, tag


def test_large_n():
    g = load_golden("replay_large_n.npz")
    for n in (10, 16, 32):
        ctrl = g[f"n{n}_ctrl"]
        scaled = g[f"n{n}_normals"] * float(g["sigma"])
        f = orc.fidelity_batch(ctrl[:, None, :], n, 0, n - 1, scaled)
        assert np.abs(f - g[f"n{n}_fids"]).max() < 1e-12


def test_wd_kats():
    g = load_golden("kat_wd.npz")
    for k in g["names"]:
        k = str(k)
        v = g[k]
        assert orc.wd_from_ideal(v.copy()) == g[k + "_wd"]
        assert orc.wd_from_ideal_zero(v.copy()) == g[k + "_wd0"]
        for p in (1, 2, 3):
            assert orc.RIM_p(v.copy(), p) == g[f"{k}_rim{p}"]
    # wd_sortof_fast_implementation.py:184-198: fixed vector, and equality with scipy's W1
    X = g["X"]
    assert abs(orc.wd_from_ideal(X.copy()) - 0.507069833) < 1e-9
    assert abs(orc.wd_from_ideal(X.copy()) - scipy.stats.wasserstein_distance(X, np.ones_like(X))) < 1e-14
    assert orc.compute_dkw_error(1 - 0.95, 100) == g["dkw_0.05_100"]
    with pytest.raises(AssertionError):
        orc.wd_from_ideal(np.array([0.5, 2.5]))


def test_mcm_pairs():
    """Reference-stored N=7 .mc/.mcm pair (B=1): statistics stage."""
    g = load_golden("kat_mcm_pairs.npz")
    m = orc.metrics(g["fids"].copy(), 1 - 0.95)
    for k, nm in enumerate(g["names"]):
        a, b = m[str(nm)], g["metrics"][k]
        ok = ~np.isnan(b)
        assert np.array_equal(np.isnan(a), np.isnan(b))
        assert np.abs(a[ok] - b[ok]).max() < 1e-15, nm


def test_kendall_against_scipy():
    rs = np.random.RandomState(11)
    for n in (2, 5, 37, 100):
        for _ in range(5):
            x = rs.randint(0, max(2, n // 3), n).astype(float)
            y = rs.permutation(n) + 1
            a = orc.kendall_tau_b(x, y)
            b = scipy.stats.kendalltau(x, y).correlation
            assert (np.isnan(a) and np.isnan(b)) or a == b
    assert np.isnan(orc.kendall_tau_b(np.ones(5), np.arange(5)))


RL_ENV_KATS = [
    # (biases, T, N, in, out, expected, places)  RLreinforceXXchain_actionedtime.py:298-341
    ([9.76909983, 10.65815206, 10.65467358, 9.71995292, -12., 8.69457352, 12., -11.77314325, -11.29782006,
      5.27449319], 25.13468797, 10, 0, 3, 0.995, 2),
    ([-0.20574245, 4.3713235, -0.30473375], 22.035034, 3, 0, 2, 0.90, 2),
    ([2.9160861365962774, 4.385934774763882, 2.9311789427883923, 9.826275581493974, 9.276727781863883,
      5.071161912055686], 3.6651542489416897, 6, 0, 2, 0.9025, 2),
]


def test_rl_env_kats():
    """RLreinforceXXchain_actionedtime.py:298-341: 2-decimal fidelity known answers, plus the N=3
    closed form (zero biases: perfect 0->2 transfer at T = pi/sqrt(2))."""
    for b, T, n, i, o, want, places in RL_ENV_KATS:
        f = orc.evaluate_fidelity(np.array(b + [T]), n, i, o)
        assert round(abs(f - want), places) == 0
    bad = np.array([3.86111206, -0.8067965, 3.86887524, 5.8814842, -3.03354326, 7.42084848, 24.83387072])
    assert orc.evaluate_fidelity(bad, 6, 0, 2) < 0.9025
    x = np.array([0.0, 0.0, 0.0, np.pi / np.sqrt(2)])
    assert abs(orc.evaluate_fidelity(x, 3, 0, 2) - 1.0) < 1e-12


def test_arim_golden():
    """ARIM = wd_from_ideal_zero of the RIM vector (generate_arim_all_fig5.py:119) from the reference."""
    g = load_golden("objective_arim.npz")
    for j in range(g["arim_rims"].shape[0]):
        assert abs(orc.arim(g["arim_rims"][j]) - g["arim_centre"][j]) < 1e-15


def _directional_draws_of_reference_run(seed, n, sigma, count):
    """The draws directional_perturbation.perturbation (noise_model.py:165-201) consumes per evaluation under
    np.random.seed(seed): np.random.randint(0, 3N), then rng(size=2) = np.random.normal(scale=sigma, size=2)."""
    np.random.seed(seed)
    d = np.empty((count, 3))
    for k in range(count):
        d[k, 0] = np.random.randint(low=0, high=3 * n)
        d[k, 1:] = np.random.normal(scale=sigma, size=2) / sigma
    return d


def test_directional_oracle_against_reference_run():
    """The oracle's directional-perturbation sweep on the replayed numpy stream == the unmodified reference's
    directional_perturbation.evaluate_noisy_fidelity under the same seed (tests/golden/make_golden.py:make_dense_path)."""
    g = load_golden("dense_path.npz")
    n, sigma = 5, 0.1
    ref = g["dir_noisy"]
    d = _directional_draws_of_reference_run(42, n, sigma, len(ref))
    got = orc.directional_fidelity_mc_replay(g["dir_x"][None, :], np.array([sigma]), d[None, None], n, 0, 4)[0, 0]
    assert np.abs(got - ref).max() < 1e-12
    assert set(d[:, 0].astype(int)) & {0, 1, 3, 6, 9}            # the run contains complex-diagonal (non-Hermitian) draws
    assert orc.directional_directions(5)[:5] == [(0, 0), (4, 4), (1, 0), (1, 1), (1, 2)] and len(orc.directional_directions(2)) == 6
