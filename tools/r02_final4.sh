#!/bin/bash
# Closing validation of the round-2 final build + A/B of the shift-at-top loop form (variant lib_shiftend.so = shift at the end of the trip)
O=gpurun_out/r02f4; mkdir -p $O
python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -3 $O/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log; tail -2 $O/smoke.log
KB="python tools/kernel_bench.py"
$KB --ns 4,5,6,7,8,12 --reps 7 > $O/kb_top.txt 2>&1
RC_LIB_PATH=build/variants/lib_shiftend.so $KB --ns 4,5,6,7,8,12 --reps 7 > $O/kb_end.txt 2>&1
for f in kb_top kb_end; do echo "== $f"; grep evals_per_s $O/$f.txt | python -c "
import sys, json
for l in sys.stdin:
    j = json.loads(l); print(j['n'], '%.5g' % j['evals_per_s'], '%.4f' % j['frac_fp64_peak'])"; done
python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err
python bench.py > $O/bench_default.json 2> $O/bench_default.err
for wl in cfg1_n4 cfg2_n5 cfg2_n6; do
  python bench.py --workload $wl --steps 10 --warmup 3 --no-mcdatasim --cpu-evals 2000 > $O/bench_1gpu_$wl.json 2> $O/bench_1gpu_$wl.err || echo "FAILED $wl"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02f4/bench_*.json")):
    try:
        d = [json.loads(l) for l in open(f) if l.startswith("{")][-1]
        print(f.split("bench_")[1][:-5], "value %.4e e2e %.4e ms/step %.3f kernel %s frac %.4f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], (d.get("roofline") or {}).get("kernel"), (d.get("roofline") or {}).get("frac", 0)))
    except Exception as ex: print(f, "??", ex)
PY
