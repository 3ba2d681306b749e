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
    res["cpu_expm_single_us"] = timeit(lambda: orc.fidelity_batch(x[None], nspin, i, o), n=200)
    res["cpu_expm_av_100_us"] = 100 * res["cpu_expm_single_us"]
    print(json.dumps(res))
