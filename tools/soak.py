"""Development aid: long sweeps with the convergence / legality counters checked (no NaN, no non-convergence)."""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import robchar_b200 as rb
from bench import synthetic_controllers
for n, C, B, zz in [(4, 2000, 20000, False), (5, 2000, 20000, False), (6, 2000, 20000, False), (7, 4000, 20000, False),
                    (8, 1000, 20000, True), (12, 500, 10000, False), (16, 500, 10000, True), (32, 200, 2000, False)]:
    ctrl = synthetic_controllers(C, n, seed=n)
    sig = np.linspace(0, 0.1, 11)
    t0 = time.time()
    st = rb.engine.fidelity_stats(ctrl, sig, B, n, 0, n - 1, dkw_eps=0.01, seed=123, zz=zz, check_convergence=True)
    torch.cuda.synchronize()
    dt = time.time() - t0
    s = st.cpu().numpy()
    ok = np.isfinite(s).all() and (s[0] >= -1e-12).all() and (s[0] <= 1 + 1e-12).all() and (-s[12] >= -1e-12).all()
    print(f"N={n} evals={11*C*B:.2e} time={dt:.2f}s finite={ok} W range [{s[0].min():.3e},{s[0].max():.3f}] min fid {(-s[12]).min():.3e}")
    assert ok
print("soak ok")
