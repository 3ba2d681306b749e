#!/bin/bash
# Closing validation of the final build (shared ziggurat table + shortened rotation chain in the spectral kernels)
OUT=gpurun_out/r02w
mkdir -p $OUT
python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
tail -3 $OUT/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/smoke.log
tail -2 $OUT/smoke.log
for wl in n16_paper cfg5_n32 cfg4_n16; do
  python bench.py --workload $wl --steps 10 --warmup 3 --no-mcdatasim --cpu-evals 2000 > $OUT/bench_1gpu_$wl.json 2> $OUT/bench_1gpu_$wl.err || echo "FAILED $wl"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02w/bench_1gpu_*.json")):
    d = [json.loads(l) for l in open(f) if l.startswith("{")][-1]
    print(f.split("bench_1gpu_")[1][:-5], "value %.3e e2e %.3e ms/step %.2f kernel %s frac %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["kernel"], d["roofline"]["frac"]))
PY
python tools/kernel_bench.py --ns 11,12,14,16,18,20,24,28,32 --reps 3 > $OUT/kb.txt 2>&1
python tools/kernel_bench.py --ns 16,32 --fused 1 --B 100000 --evals 4e7 > $OUT/kbf.txt 2>&1
grep -h evals_per_s $OUT/kb.txt $OUT/kbf.txt | cut -c1-140
