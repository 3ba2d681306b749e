#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Runs only in the build container (needs /root/reference; the GPU box never runs it).
It imports the reference's own modules (noise_model, mcsim, wd_sortof_fast_implementation,
generate_fig4_kendallrankanalysis' nested helpers) with sys.modules stubs for the plotting /
optimiser packages that are not installed, runs them on seeded numpy legacy-RNG streams and
stores inputs + outputs as small .npz files.  Nothing here is product code.

    python tests/golden/make_golden.py
"""
import ast
import json
import os
import sys
import tempfile
import types
from unittest.mock import MagicMock

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True

for name in ["matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.ticker", "seaborn",
             "IPython", "IPython.display", "skquant", "skquant.opt", "SQSnobFit"]:
    sys.modules[name] = MagicMock()
sys.path.insert(0, REF)

import noise_model as ref_nm  # noqa: E402
import wd_sortof_fast_implementation as ref_wd  # noqa: E402
import mcsim as ref_mc  # noqa: E402
from scipy.stats import kendalltau  # noqa: E402


def load_json(path):
    return json.load(open(path, "rb"))


# ---------------------------------------------------------------------------------------------
# 1. Known-answer fixtures stored by the reference itself
# ---------------------------------------------------------------------------------------------
def make_kat_bestfid():
    """noisy_analysis/lbfgs_spin_*_in: best_fid[i] is the nominal fidelity of controller[i]."""
    out = {}
    meta = []
    for fn in sorted(os.listdir(f"{REF}/noisy_analysis")):
        if not fn.startswith("lbfgs_spin_"):
            continue
        _, _, n, io, _ = fn.split("_")
        n = int(n); i, o = (int(v) for v in io.split("-"))
        rec = load_json(f"{REF}/noisy_analysis/{fn}")["lbfgs"][str(n)]
        ctrl = np.array(rec["controller"], dtype=np.float64)
        best = np.array(rec["best_fid"], dtype=np.float64)
        key = f"n{n}_{i}_{o}"
        out[key + "_ctrl"] = ctrl
        out[key + "_best_fid"] = best
        meta.append((key, n, i, o))
    out["meta"] = np.array(meta)
    np.savez_compressed(f"{OUT}/kat_bestfid.npz", **out)
    print("kat_bestfid:", [m[0] for m in meta])


def make_kat_mc_zero(per_group=60):
    """sigma_sim = 0 rows of the stored .mc caches against the .le controllers."""
    exp = f"{REF}/experiments/pipeline_nmplus2"
    nl = "[0.   0.01 0.02 0.03 0.04 0.05 0.06 0.07 0.08 0.09 0.1 ]"
    out = {}
    meta = []
    for n, i, o in [(4, 0, 2), (4, 0, 3), (5, 0, 2), (5, 0, 4), (6, 0, 3)]:
        le = load_json(f"{exp}/ppo_spin_{n}_{i}-{o}_c_1000.le")
        for tn in ["0.0", "0.03"]:
            mc = load_json(f"{exp}/ppo_spin_{n}_{i}-{o}_c_1000.le_tn{tn}_br_1_nlvl{nl}.mc")
            for algo in mc:
                keyname = str(n) if algo == "lbfgs" else tn
                if keyname not in le.get(algo, {}):
                    continue
                conts = le[algo][keyname]["controller"][:per_group]
                ctrl = np.array([c if c is not None else [np.nan] * (n + 1) for c in conts], dtype=np.float64)
                fid0 = np.array(mc[algo], dtype=np.float64)[0, :len(conts), 0]
                key = f"n{n}_{i}_{o}_{algo}_tn{tn}"
                out[key + "_ctrl"] = ctrl
                out[key + "_fid0"] = fid0
                meta.append((key, n, i, o))
    out["meta"] = np.array(meta)
    np.savez_compressed(f"{OUT}/kat_mc_zero.npz", **out)
    print("kat_mc_zero groups:", len(meta))


def make_kat_mcm_pairs(ncol=200):
    """N=7 .mc/.mcm pairs (B=1) pin the statistics stage exactly."""
    exp = f"{REF}/experiments/pipeline_nmplus2"
    nl = "[0.   0.01 0.02 0.03 0.04 0.05 0.06 0.07 0.08 0.09 0.1 ]"
    base = f"{exp}/ppo_spin_7_0-6_c_1000.le_tn0.02_br_1_nlvl{nl}"
    if not (os.path.exists(base + ".mc") and os.path.exists(base + ".mcm")):
        print("kat_mcm_pairs: files missing, skipped")
        return
    mc = load_json(base + ".mc"); mcm = load_json(base + ".mcm")
    out = {}
    algo = "nmplus"
    out["fids"] = np.array(mc[algo], dtype=np.float64)[:, :ncol, :]
    names = list(mcm[algo].keys())
    out["names"] = np.array(names)
    out["metrics"] = np.array([np.array(mcm[algo][k], dtype=np.float64)[:, :ncol] for k in names])
    np.savez_compressed(f"{OUT}/kat_mcm_pairs.npz", **out)
    print("kat_mcm_pairs:", out["fids"].shape, names)


# ---------------------------------------------------------------------------------------------
# 2. Seeded runs of the unmodified reference (sigma > 0 has no stored goldens)
# ---------------------------------------------------------------------------------------------
def run_reference_mcsim(n, i, o, controllers, noises, bootreps, seed, numcontrollers, algo="lbfgs"):
    """Drive the real MCDataSim.get_metrics_dict in a temp experiments/ tree and recover the
    standard-normal stream it consumed (np.random.normal(scale=s) == s*standard_normal())."""
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as td:
        os.makedirs(f"{td}/experiments/golden")
        keyname = str(n)
        conts = [None if (isinstance(c, float) and np.isnan(c)) else list(map(float, c)) for c in controllers]
        json.dump({algo: {keyname: {"controller": conts}}},
                  open(f"{td}/experiments/golden/ppo_spin_{n}_{i}-{o}_c_{numcontrollers}", "w"))
        os.chdir(td)
        try:
            sim = ref_mc.MCDataSim(experiment_name="golden", Nspin=n, inspin=i, outspin=o, noises=noises,
                                   bootreps=bootreps, numcontrollers=numcontrollers, topk=10)
            np.random.seed(seed)
            metrics = sim.get_metrics_dict(None, noises, algoname=algo)
            mcname = sim.get_mcname(None, noises)
            fids = np.array(load_json(mcname)[algo], dtype=np.float64)
        finally:
            os.chdir(cwd)
    # reconstruct the stream: one discarded draw per sigma level (mcsim.py:425), then 3N per
    # evaluation; NaN-padded controllers (index >= len) consume nothing (mcsim.py:369-374).
    # JSON `null` controllers are *not* np.nan -> the reference would crash; we never feed those.
    K = 3 * n
    S = len(noises); C = numcontrollers; B = bootreps
    rs = np.random.RandomState(seed)
    normals = np.zeros((S, C, B, K))
    for s in range(S):
        rs.standard_normal()
        for c in range(C):
            if c < len(controllers):
                normals[s, c] = rs.standard_normal((B, K))
    ctrl = np.full((C, n + 1), np.nan)
    ctrl[:len(controllers)] = np.array(controllers, dtype=np.float64)
    names = list(metrics[algo].keys())
    mt = np.array([np.array(metrics[algo][k], dtype=np.float64) for k in names])
    return dict(ctrl=ctrl, sigmas=np.array(noises, dtype=np.float64), normals=normals, fids=fids,
                metric_names=np.array(names), metrics=mt, nio=np.array([n, i, o]), seed=seed,
                alpha=1 - 0.95)


def extract_nested(src_path, names):
    """Pull nested helper functions out of generate_fig4_kendallrankanalysis.py and exec them
    unchanged (they are closures inside plot_kendalltaus and cannot be imported)."""
    tree = ast.parse(open(src_path).read())
    found = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name in names:
            found[node.name] = node
    ns = {"np": np, "kendalltau": kendalltau, "vn_test": ref_mc.vn_test,
          "self": types.SimpleNamespace(get_ranks=ref_mc.MCDataSim.get_ranks)}
    for nm in names:
        mod = ast.Module(body=[found[nm]], type_ignores=[])
        exec(compile(mod, src_path, "exec"), ns)
    return ns


def make_replay():
    noises = np.linspace(0, 0.1, 11)
    lb = lambda n, i, o: load_json(f"{REF}/noisy_analysis/lbfgs_spin_{n}_{i}-{o}_in")["lbfgs"][str(n)]["controller"]
    cases = {
        # config 1: N=4 0->2, the 100 LBFGS controllers, padded with 4 NaN rows
        "replay_n4_0_2": dict(n=4, i=0, o=2, controllers=lb(4, 0, 2), bootreps=6, seed=101, numcontrollers=104),
        "replay_n5_0_4": dict(n=5, i=0, o=4, controllers=lb(5, 0, 4)[:40], bootreps=10, seed=102, numcontrollers=40),
        "replay_n6_0_3": dict(n=6, i=0, o=3, controllers=lb(6, 0, 3)[:40], bootreps=10, seed=103, numcontrollers=40),
        # config 3: N=7 0->6, all 57 LBFGS controllers in the mount
        "replay_n7_0_6": dict(n=7, i=0, o=6, controllers=lb(7, 0, 6), bootreps=12, seed=104, numcontrollers=57),
    }
    ns = extract_nested(f"{REF}/generate_fig4_kendallrankanalysis.py",
                        ["get_ranks_clustered_little", "jkt_or_ordinaltau_pairwise"])
    for name, kw in cases.items():
        d = run_reference_mcsim(noises=noises, **kw)
        # ranking / Kendall stage with the reference's own functions on the reference's RIM matrix
        W = d["metrics"][list(d["metric_names"]).index(r'$W(.,\delta(x-1))$')]
        ok = ~np.isnan(W[0])
        Wc = W[:, ok]
        d["ranks_row0"] = ref_mc.MCDataSim.get_ranks(Wc[0])
        d["clustered_row3"] = ns["get_ranks_clustered_little"](Wc[3], r=0.05 * (Wc[3].max() - Wc[3].min()))
        d["kendall"] = np.array(ns["jkt_or_ordinaltau_pairwise"](Wc, alpha=0.05))
        np.savez_compressed(f"{OUT}/{name}.npz", **d)
        print(name, d["fids"].shape, "kendall", d["kendall"].shape)


def make_wd_kats():
    """Embedded unit-test vectors of wd_sortof_fast_implementation.py:182-311, evaluated by the
    reference's own functions."""
    rs = np.random.RandomState(7)
    vecs = {
        "X": np.array(ref_wd.testwdimplementation.X),
        "normal_hi": rs.normal(0.85, 0.02, size=10000),
        "normal_clip": rs.normal(0.85, 0.8, size=10000).clip(min=0, max=1),
        "normal_10": rs.normal(0.67, 0.02, size=10),
        "uniform_10": rs.uniform(size=10),
        "ones": np.array([1., 1, 1, 1, 1]),
        "mixed": np.array([1., 0, 1, 1, 0]),
        "zeros": np.array([0., 0, 0, 0, 0]),
        "scalar": np.array([0.76]),
    }
    out = {}
    for k, v in vecs.items():
        out[k] = v
        out[k + "_wd"] = ref_wd.wd_from_ideal(v.copy())
        out[k + "_wd0"] = ref_wd.wd_from_ideal_zero(v.copy())
        out[k + "_rim1"] = ref_wd.RIM_p(v.copy(), 1)
        out[k + "_rim2"] = ref_wd.RIM_p(v.copy(), 2)
        out[k + "_rim3"] = ref_wd.RIM_p(v.copy(), 3)
    out["dkw_0.05_100"] = ref_wd.compute_dkw_error(1 - 0.95, 100)
    out["names"] = np.array(list(vecs.keys()))
    np.savez_compressed(f"{OUT}/kat_wd.npz", **out)
    print("kat_wd:", list(vecs.keys()))


def make_real2_and_zz():
    """Optimiser-side variant: LBFGS.fidelity_ss with the real 2-draw noise (qnewton.py:366-423)
    and the Heisenberg/Z diagonal (qnewton.py:148-150), run through the real qnewton.LBFGS."""
    import qnewton as ref_q
    out = {}
    for tag, n, i, o, zz in [("n5", 5, 0, 4, False), ("n6zz", 6, 0, 3, True), ("n16zz", 16, 0, 15, True)]:
        env = ref_q.LBFGS(n, i, o, noise=0.05, heisenberg_int=zz, opt_train_size=2)
        rs = np.random.RandomState(55 + n)
        C, B = 12, 8
        ctrl = np.concatenate([rs.uniform(-10, 10, (C, n)), rs.uniform(1, 30, (C, 1))], axis=1)
        np.random.seed(900 + n)
        fids = np.zeros((C, B))
        for c in range(C):
            for b in range(B):
                fids[c, b] = env.fidelity_ss(ctrl[c], ham_noisy=True)
        normals = np.random.RandomState(900 + n).standard_normal((C, B, 2 * n))
        out[tag + "_ctrl"] = ctrl; out[tag + "_fids"] = fids; out[tag + "_normals"] = normals
        out[tag + "_meta"] = np.array([n, i, o, int(zz)]); out[tag + "_sigma"] = 0.05
        out[tag + "_nominal"] = np.array([env.fidelity_ss(ctrl[c]) for c in range(C)])
    np.savez_compressed(f"{OUT}/replay_real2_zz.npz", **out)
    print("replay_real2_zz done")


def make_large_n():
    """N=16 / N=32 (scaled configs 4/5) and the RL environment KATs (N=10, 3, 6) through the
    reference's structured_perturbation.evaluate_noisy_fidelity."""
    out = {}
    for n in (10, 16, 32):
        rs = np.random.RandomState(20221 + n)
        C, B = 10, 6
        ctrl = np.concatenate([rs.uniform(-10, 10, (C, n)), rs.uniform(1, 30, (C, 1))], axis=1)
        model = ref_nm.structured_perturbation(Nspin=n, inspin=0, outspin=n - 1)
        np.random.seed(300 + n)
        model.rng(scale=0.05)
        fids = np.zeros((C, B))
        for c in range(C):
            for b in range(B):
                fids[c, b] = model.evaluate_noisy_fidelity(ctrl[c], True)
        rs2 = np.random.RandomState(300 + n); rs2.standard_normal()
        out[f"n{n}_ctrl"] = ctrl; out[f"n{n}_fids"] = fids
        out[f"n{n}_normals"] = rs2.standard_normal((C, B, 3 * n))
        out[f"n{n}_nominal"] = np.array([model.evaluate_noisy_fidelity(ctrl[c], False) for c in range(C)])
    out["sigma"] = 0.05
    np.savez_compressed(f"{OUT}/replay_large_n.npz", **out)
    print("replay_large_n done")


def make_objective_and_arim():
    """Optimiser-side objectives (qnewton.py:425-455, :407) and ARIM + bootstrap std
    (generate_arim_all_fig5.py:119-126, mcsim.py:267-275) from the unmodified reference."""
    import qnewton as ref_q
    out = {}
    n, i, o = 5, 0, 4
    env = ref_q.LBFGS(n, i, o, noise=0.05, opt_train_size=20)
    rs = np.random.RandomState(77)
    X = np.concatenate([rs.uniform(-10, 10, (4, n)), rs.uniform(1, 30, (4, 1))], axis=1)
    out["obj_X"] = X; out["obj_meta"] = np.array([n, i, o, 20])
    out["obj_av10"] = np.array([env.fidelity_ss_av(x, reps=10) for x in X])
    out["obj_av_test"] = np.array([env.fidelity_ss_av(x, test=True) for x in X[:2]])
    np.random.seed(11)
    out["obj_wass7"] = np.array([env.wass_cost(x, 7) for x in X])
    np.random.seed(12)
    out["obj_shot"] = np.array([env.fidelity_ss(x, noisy=True, ham_noisy=True) for x in X])
    env.adaptive = True
    np.random.seed(13)
    out["obj_adaptive"] = np.array([env.fidelity_ss(x, noisy=True, ham_noisy=False) for x in X])
    out["obj_randH0"] = env.randH[0]
    # ARIM
    rs = np.random.RandomState(5)
    wdd = rs.uniform(0, 0.5, (11, 100)) * np.linspace(0.2, 1, 11)[:, None]
    out["arim_rims"] = wdd
    out["arim_centre"] = np.array([ref_wd.wd_from_ideal_zero(wdd[j].copy()) for j in range(11)])
    np.random.seed(21)
    out["arim_std"] = np.array([ref_mc.MCDataSim.bootstrap_resampling_std(ref_wd.wd_from_ideal_zero, wdd[j].copy(), 100)
                                for j in range(11)])
    np.savez_compressed(f"{OUT}/objective_arim.npz", **out)
    print("objective_arim done")


def make_dense_path():
    """Reference outputs for the generality path: analytic gradient (qnewton.py:162-212), ring topology
    (noise_model.py:83-85) and directional_perturbation incl. its complex-diagonal draws (:150-201)."""
    import qnewton as ref_q
    out = {}
    env = ref_q.LBFGS(5, 0, 4, noise=0.05, opt_train_size=2)
    rs = np.random.RandomState(31)
    X = np.concatenate([rs.uniform(-10, 10, (3, 5)), rs.uniform(1, 30, (3, 1))], axis=1)
    out["grad_X"] = X
    eg = [env.eval_static_fidelity_gradient(x) for x in X]
    out["grad_err"] = np.array([e for e, g in eg]); out["grad_g"] = np.array([g for e, g in eg])
    ring = ref_nm.structured_perturbation(Nspin=6, inspin=0, outspin=3, noise=0.05, topo="ring")
    Xr = np.concatenate([rs.uniform(-10, 10, (4, 6)), rs.uniform(1, 30, (4, 1))], axis=1)
    out["ring_X"] = Xr
    out["ring_nominal"] = np.array([ring.evaluate_noisy_fidelity(x, False) for x in Xr])
    np.random.seed(41)
    out["ring_noisy"] = np.array([ring.evaluate_noisy_fidelity(x, True) for x in Xr])
    dp = ref_nm.directional_perturbation(Nspin=5, inspin=0, outspin=4, noise=0.1)
    Xd = np.concatenate([rs.uniform(-10, 10, (1, 5)), rs.uniform(1, 30, (1, 1))], axis=1)[0]
    out["dir_x"] = Xd
    np.random.seed(42)
    out["dir_noisy"] = np.array([dp.evaluate_noisy_fidelity(Xd, True) for _ in range(40)])
    rs2 = np.random.RandomState(9)
    mats = []
    for M in (2, 5, 14, 32):
        A = (rs2.standard_normal((3, M, M)) + 1j * rs2.standard_normal((3, M, M))) * np.array([0.01, 1.0, 20.0])[:, None, None] / np.sqrt(M)
        out[f"expm_A{M}"] = A
        import scipy.linalg
        out[f"expm_E{M}"] = np.array([scipy.linalg.expm(a) for a in A])
    np.savez_compressed(f"{OUT}/dense_path.npz", **out)
    print("dense_path done")


def make_gradient():
    """eval_static_fidelity_gradient of the UNMODIFIED reference (qnewton.py:162-212) at N = 4, 7, 16, 32, interior and
    end-to-end targets, Heisenberg term, and with ham_noisy=True under a fixed numpy seed (the perturbation draw is
    recorded so the device path can be fed the same matrix)."""
    import qnewton as ref_q
    out = {}
    meta = []
    rs = np.random.RandomState(77)
    for n, i, o, scale, hz in [(4, 0, 2, 10.0, False), (7, 0, 6, 10.0, False), (7, 0, 3, 3.0, True), (16, 0, 15, 1.0, False),
                               (16, 3, 9, 1.0, True), (32, 0, 31, 0.2, False)]:
        env = ref_q.LBFGS(n, i, o, noise=0.05, opt_train_size=2, heisenberg_int=hz)
        C = 4 if n <= 16 else 2
        X = np.concatenate([rs.uniform(-scale, scale, (C, n)), rs.uniform(1, 30, (C, 1)) if n <= 16 else rs.uniform(18, 30, (C, 1))], axis=1)
        X[0, n] *= -1                                            # abs(T)
        eg = [env.eval_static_fidelity_gradient(x) for x in X]
        key = f"n{n}_{i}_{o}"
        out[key + "_X"] = X
        out[key + "_err"] = np.array([e for e, g in eg]); out[key + "_grad"] = np.array([g for e, g in eg])
        env.ham_noisy = True
        np.random.seed(1000 + n)
        eg = [env.eval_static_fidelity_gradient(x) for x in X]
        out[key + "_noisy_err"] = np.array([e for e, g in eg]); out[key + "_noisy_grad"] = np.array([g for e, g in eg])
        meta.append((key, n, i, o, int(hz), 1000 + n))
    out["meta"] = np.array(meta)
    np.savez_compressed(f"{OUT}/gradient.npz", **out)
    print("gradient:", [m[0] for m in meta])


def make_rl_env():
    """Rewards / noise-free fidelities of the UNMODIFIED RL environment (RLreinforceXXchain_actionedtime.py:260-279)
    driven the way ppo.py:338-363 drives it: reset, advance timestep, step(diag(bias increments))."""
    import warnings
    from RLreinforceXXchain_actionedtime import Environment
    warnings.simplefilter("ignore")
    out = {}
    rs = np.random.RandomState(123)
    n, i, o = 5, 0, 4
    acts = rs.uniform(-9, 9, (6, n)); acts[3] += 14 * np.sign(acts[3])      # step 4 leaves the +-20 bounds: wrap
    times = np.array([2.5, 7.25, 11.0, 19.5, 28.0, 33.5])                    # last one exceeds max_time: wrap
    good = np.array(load_json(f"{REF}/noisy_analysis/lbfgs_spin_5_0-4_in")["lbfgs"]["5"]["controller"][0], dtype=float)
    acts[0], times[0] = good[:n], good[n]                                   # a real optimised controller (fidelity ~ 1)
    out["env_acts"], out["env_times"], out["env_meta"] = acts, times, np.array([n, i, o])

    def drive(env, seed):
        np.random.seed(seed)
        env.reset()
        rew, tf, done, act = [], [], [], []
        for a, t in zip(acts, times):
            env.timestep = t
            ao, r, d = env.step(np.diag(a))
            rew.append(np.real(r)); tf.append(np.real(env.tf)); done.append(d); act.append(np.diag(ao).copy())
        return np.array(rew), np.array(tf), np.array(done), np.array(act)

    for name, kw in (("plain", {}), ("hamnoisy", dict(ham_noisy=True)), ("shot", dict(fid_noisy=True, draws=20)),
                     ("adaptive", dict(fid_noisy=True, adaptive=True, draws=20)), ("heis", dict(heisenberg_int=True)),
                     ("fixed", dict(use_fixed_ham=True)), ("ring", dict(topo="ring"))):
        env = Environment(n, i, o, noise=0.05, opt_train_size=12, **kw)
        r, tf, d, a = drive(env, 17)
        out[f"env_{name}_reward"], out[f"env_{name}_tf"], out[f"env_{name}_done"], out[f"env_{name}_action"] = r, tf, d, a
        if name == "fixed":
            out["env_fixed_truefid"] = np.array([np.real(env.true_fid(np.diag(acts[0]), timestep_n=times[0]))])
        if name == "adaptive":
            out["env_adaptive_calls"] = np.array([env.adp_func_calls_increment])
    np.random.seed(5)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        env = Environment(n, i, o, noise=0.05, opt_train_size=4, transfer_learning=True)
    out["env_tl_sys"] = env.sys
    r, tf, d, a = drive(env, 18)
    out["env_tl_reward"], out["env_tl_tf"] = r, tf
    np.savez_compressed(f"{OUT}/rl_env.npz", **out)
    print("rl_env done")


if __name__ == "__main__":
    make_kat_bestfid()
    make_kat_mc_zero()
    make_kat_mcm_pairs()
    make_wd_kats()
    make_replay()
    make_real2_and_zz()
    make_large_n()
    make_objective_and_arim()
    make_dense_path()
    make_gradient()
    make_rl_env()
