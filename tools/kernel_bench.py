"""Development aid: kernel-only timings of the evolution kernel for several chain lengths."""
import argparse, sys, os, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import robchar_b200 as rb
from bench import synthetic_controllers, f_alg

ap = argparse.ArgumentParser()
ap.add_argument("--ns", default="4,5,6,7,16,32")
ap.add_argument("--evals", type=float, default=2e7)
ap.add_argument("--B", type=int, default=100)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--fused", type=int, default=0)
ap.add_argument("--stats", type=int, default=0)
ap.add_argument("--replay", type=int, default=0, help="1: stream standard normals from HBM (replay mode) instead of Philox")
a = ap.parse_args()
peak = rb.engine.fp64_peak_tflops()
print("fp64 peak TFLOP/s", peak)
for n in [int(v) for v in a.ns.split(",")]:
    S = 11
    scale = 1.0 if n <= 8 else (0.25 if n <= 16 else 0.06)
    C = max(1, int(a.evals * scale / (S * a.B)))
    ctrl = torch.as_tensor(synthetic_controllers(C, n)).cuda()
    sig = torch.linspace(0, 0.1, S, dtype=torch.float64).cuda()
    out = torch.empty((S, C, a.B), dtype=torch.float64, device="cuda")
    rep = torch.randn((S, C, a.B, 3 * n), dtype=torch.float64, device="cuda") if a.replay else None
    ts = []
    for r in range(a.reps + 2):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        if a.fused:
            rb.engine.fidelity_stats(ctrl, sig, a.B, n, 0, n - 1, dkw_eps=0.01, seed=r, check_convergence=False)
        else:
            if a.stats == 2:     # evolution + sort-free statistics pass
                rb.engine.fidelity_mc_stats(ctrl, sig, a.B, n, 0, n - 1, dkw_eps=0.01, seed=r, out=out, check=False)
                e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
                continue
            rb.engine.fidelity_mc(ctrl, sig, a.B, n, 0, n - 1, seed=r, out=out, check_convergence=False, replay=rep)
            if a.stats:
                rb.engine.stats(out, 0.01, check_legal=False)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts[2:]))
    ev = S * C * a.B
    extra = {"replay_GB/s": ev * (24 * n + 8) / ms / 1e6} if a.replay else {}
    print(json.dumps({**extra, "n": n, "evals": ev, "ms": ms, "evals_per_s": ev / ms * 1e3, "alg_tflops": ev * f_alg(n) / ms * 1e3 / 1e12,
                      "frac_fp64_peak": ev * f_alg(n) / ms * 1e3 / 1e12 / peak}))
