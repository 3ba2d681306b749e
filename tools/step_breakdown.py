"""Development aid: where the time of one paper_n7 step goes (CUDA events around sub-sequences, median of 20):
evolution only / + statistics pass / whole rc_robustness_sweep, each with and without the 160 MiB L2 flush."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import robchar_b200 as rb
from bench import synthetic_controllers
eng = rb.engine
n, S, B, G, cg = 7, 11, 100, 19, 1000
C = G * cg
ctrl = torch.as_tensor(synthetic_controllers(C, n)).cuda()
sig = torch.linspace(0, 0.1, S, dtype=torch.float64).cuda()
fids = torch.empty((S, C, B), dtype=torch.float64, device="cuda")
flush = torch.empty(160 << 20, dtype=torch.uint8, device="cuda")
plan = eng.RobustnessSweepPlan(C, S, B, n, 0, n - 1, groups=G, topk=100, dkw_eps=float(eng.compute_dkw_error(0.05, B)))

def timed(fn, with_flush, reps=20):
    for k in range(3):
        fn(k); flush.zero_()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    for k in range(reps):
        ev[k].record()
        fn(10 + k)
        if with_flush:
            flush.zero_()
    ev[reps].record(); torch.cuda.synchronize()
    return float(np.median([ev[k].elapsed_time(ev[k + 1]) for k in range(reps)]))

res = {}
for wf in (False, True):
    tag = "+flush" if wf else ""
    res["flush_only" + tag] = timed(lambda k: None, wf)
    res["evolution" + tag] = timed(lambda k: eng.fidelity_mc(ctrl, sig, B, n, 0, n - 1, seed=k, out=fids, check_convergence=False), wf)
    res["evolution+stats" + tag] = timed(lambda k: eng.fidelity_mc_stats(ctrl, sig, B, n, 0, n - 1, dkw_eps=0.01, seed=k, out=fids, check=False), wf)
    res["whole_step" + tag] = timed(lambda k: plan.run(ctrl, sig, seed=k), wf)
print(json.dumps(res))
