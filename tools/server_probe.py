"""Development aid: per-call latency of the resident objective evaluator vs CTA size (RC_OBJECTIVE_MIN_THREADS)."""
import ctypes as C, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import robchar_b200 as rb
from oracle import robchar_oracle as orc

def timeit(fn, n=3000, warm=200):
    for _ in range(warm): fn()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    return (time.perf_counter() - t0) / n * 1e6

lib = rb._lib.lib()
res = {"min_threads": os.environ.get("RC_OBJECTIVE_MIN_THREADS", "0"), "server": os.environ.get("RC_OBJECTIVE_SERVER", "1")}
for nspin in (5, 7):
    x = np.ascontiguousarray(orc.synthetic_controllers(1, nspin)[0]); out = np.empty(1); st15 = np.empty(15)
    rows30 = np.random.RandomState(1).standard_normal((30, 2 * nspin)) * 0.05
    a_nom = (C.c_void_p(x.ctypes.data), nspin, 0, nspin - 1, C.c_void_p(0), 1, 1, 0, 0.0, C.c_void_p(out.ctypes.data), C.c_void_p(0), C.c_void_p(0), C.c_void_p(0))
    a_30 = (C.c_void_p(x.ctypes.data), nspin, 0, nspin - 1, C.c_void_p(rows30.ctypes.data), 30, 1, 0, 0.0, C.c_void_p(0), C.c_void_p(st15.ctypes.data), C.c_void_p(0), C.c_void_p(0))
    res[f"n{nspin}_nominal_us"] = round(timeit(lambda: lib.rc_objective_host(*a_nom)), 2)
    res[f"n{nspin}_30rows_us"] = round(timeit(lambda: lib.rc_objective_host(*a_30)), 2)
    def spaced():
        lib.rc_objective_host(*a_nom)
        t = time.perf_counter()
        while time.perf_counter() - t < 8e-6: pass
    res[f"n{nspin}_nominal_spaced8_us"] = round(timeit(spaced) - 8.0, 2)
print(json.dumps(res))
