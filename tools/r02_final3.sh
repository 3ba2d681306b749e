#!/bin/bash
# Closing validation of the round-2 final build (pinned-end register solver, register kernels up to N = 12).
O=gpurun_out/r02f3; mkdir -p $O
python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -3 $O/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log; tail -2 $O/smoke.log
python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err
python bench.py > $O/bench_default.json 2> $O/bench_default.err
for wl in n4_paper n5_paper n6_paper cfg1_n4 cfg2_n5 cfg2_n6 cfg1_full_n4; do
  python bench.py --workload $wl --steps 10 --warmup 3 --no-mcdatasim --cpu-evals 2000 > $O/bench_1gpu_$wl.json 2> $O/bench_1gpu_$wl.err || echo "FAILED $wl"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02f3/bench_*.json")):
    try:
        d = [json.loads(l) for l in open(f) if l.startswith("{")][-1]
        print(f.split("bench_")[1][:-5], "value %.3e e2e %.3e ms/step %.3f kernel %s frac %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], (d.get("roofline") or {}).get("kernel"), (d.get("roofline") or {}).get("frac", 0)))
    except Exception as ex: print(f, "??", ex)
PY
python tools/kernel_bench.py --ns 2,3,4,5,6,7,8,9,10,11,12 --reps 3 > $O/kb.txt 2>&1
python tools/kernel_bench.py --ns 4,5,6,7,8,9,10,11,12 --fused 1 --B 100000 --evals 4e7 > $O/kbf.txt 2>&1
grep -h evals_per_s $O/kb.txt $O/kbf.txt | cut -c1-150
python tools/latency_bench.py > $O/latency.txt 2>&1; tail -4 $O/latency.txt | cut -c1-400
