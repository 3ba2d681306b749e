"""Development aid: per-call latency of the optimiser-facing evaluator (qnewton.LBFGS mirror) — the small-batch,
launch-latency-bound end of the path — next to the oracle's CPU evaluation of the same call."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import robchar_b200 as rb
from oracle import robchar_oracle as orc


def timeit(fn, n=200, warm=20):
    for _ in range(warm):
        fn()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    return (time.perf_counter() - t0) / n * 1e6


for nspin, i, o in ((5, 0, 4), (7, 0, 6)):
    ev = rb.qnewton.LBFGS(nspin, i, o, noise=0.05, opt_train_size=100, opt_test_size=10000)
    x = orc.synthetic_controllers(1, nspin)[0]
    res = {"nspin": nspin}
    res["fidelity_ss_us"] = timeit(lambda: ev.fidelity_ss(x))
    res["fidelity_ss_ham_noisy_us"] = timeit(lambda: ev.fidelity_ss(x, ham_noisy=True))
    res["fidelity_ss_av_100_us"] = timeit(lambda: ev.fidelity_ss_av(x, reps=100))
    res["fidelity_ss_av_test_10000_us"] = timeit(lambda: ev.fidelity_ss_av(x, test=True), n=50, warm=5)
    res["wass_cost_5_us"] = timeit(lambda: ev.wass_cost(x, 5))
    res["wass_cost_30_us"] = timeit(lambda: ev.wass_cost(x, 30))
    rows30 = ev._noise_rows(30)
    res["objective_host_30_rows_us"] = timeit(lambda: rb.engine.objective_host(x, rows30, nspin, i, o, model=1, want_fids=False, want_stats=True))
    res["objective_host_nominal_us"] = timeit(lambda: rb.engine.objective_host(x, None, nspin, i, o, model=1))
    # the C entry point alone (ctypes call with prebuilt arguments): what a compiled optimiser would pay per call
    import ctypes as C
    lib = rb._lib.lib()
    xs = np.ascontiguousarray(x); out = np.empty(1); st15 = np.empty(15); out30 = np.empty(30)
    a_nom = (C.c_void_p(xs.ctypes.data), nspin, i, o, C.c_void_p(0), 1, 1, 0, 0.0, C.c_void_p(out.ctypes.data), C.c_void_p(0), C.c_void_p(0), C.c_void_p(0))
    a_30 = (C.c_void_p(xs.ctypes.data), nspin, i, o, C.c_void_p(rows30.ctypes.data), 30, 1, 0, 0.0, C.c_void_p(0), C.c_void_p(st15.ctypes.data), C.c_void_p(0), C.c_void_p(0))
    res["c_abi_nominal_us"] = timeit(lambda: lib.rc_objective_host(*a_nom), n=2000, warm=100)
    res["c_abi_30_rows_stats_us"] = timeit(lambda: lib.rc_objective_host(*a_30), n=2000, warm=100)
    res["cpu_expm_single_us"] = timeit(lambda: orc.fidelity_batch(x[None], nspin, i, o), n=200)
    res["cpu_expm_av_100_us"] = 100 * res["cpu_expm_single_us"]
    print(json.dumps(res))
