"""Development aid: replay-mode vs Philox-mode timing (isolates the noise-generation cost)."""
import sys, os, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import robchar_b200 as rb
from bench import synthetic_controllers
n, S, B = 7, 11, 100
C = int(sys.argv[1]) if len(sys.argv) > 1 else 18181
ctrl = torch.as_tensor(synthetic_controllers(C, n)).cuda()
sig = torch.linspace(0, 0.1, S, dtype=torch.float64).cuda()
out = torch.empty((S, C, B), dtype=torch.float64, device="cuda")
z = rb.engine.philox_normals(C, n, S, B, seed=1)
def t(fn, reps=5):
    ts = []
    for r in range(reps + 2):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(r); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts[2:]))
ev = S * C * B
a = t(lambda r: rb.engine.fidelity_mc(ctrl, sig, B, n, 0, 6, seed=1, out=out, check_convergence=False))
b = t(lambda r: rb.engine.fidelity_mc(ctrl, sig, B, n, 0, 6, replay=z, out=out, check_convergence=False))
c = t(lambda r: rb.engine.philox_normals(C, n, S, B, seed=r))
sig0 = torch.zeros(S, dtype=torch.float64).cuda()
d = t(lambda r: rb.engine.fidelity_mc(ctrl, sig0, B, n, 0, 6, seed=1, out=out, check_convergence=False))
print(json.dumps({"evals": ev, "philox_ms": a, "replay_ms": b, "normals_dump_ms": c, "philox_sigma0_ms": d}))
