"""Development aid: device time of the statistics stage alone (sort-based rc_stats vs sort-free
rc_stats_unsorted) on a fidelity tensor of the headline shape, against the HBM roofline."""
import argparse, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import robchar_b200 as rb

ap = argparse.ArgumentParser()
ap.add_argument("--S", type=int, default=11)
ap.add_argument("--C", type=int, default=19000)
ap.add_argument("--Bs", default="100")
ap.add_argument("--reps", type=int, default=10)
a = ap.parse_args()
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) \
    if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for B in [int(v) for v in a.Bs.split(",")]:
    C = max(1, a.C * 100 // B)
    f = torch.rand((a.S, C, B), dtype=torch.float64, device="cuda")
    for name, fn in (("rc_stats (sort)", lambda: rb.engine.stats(f, 0.1, check_legal=False)),
                     ("rc_stats_unsorted", lambda: rb.engine.stats_unsorted(f, 0.1, check_legal=False))):
        ts = []
        for r in range(a.reps + 2):
            flush.zero_()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts[2:]))
        gbs = f.numel() * 8 / ms / 1e6
        print(json.dumps({"kernel": name, "B": B, "segments": a.S * C, "ms": ms, "GB/s": gbs, "frac_hbm_peak": gbs / peak}))
