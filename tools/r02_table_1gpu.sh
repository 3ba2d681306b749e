#!/bin/bash
# 1-GPU cells of the throughput table (profiles/README.md): paper sweep shape at N = 4, 5, 6, 7, 16 and the BASELINE
# configurations at full size where one GPU is enough.
OUT=gpurun_out/r02d
mkdir -p $OUT
for wl in n4_paper n5_paper n6_paper n16_paper cfg5_n32 cfg4_n16 cfg1_full_n4; do
  python bench.py --workload $wl --steps 10 --warmup 3 > $OUT/bench_1gpu_$wl.json 2> $OUT/bench_1gpu_$wl.err || echo "FAILED $wl"
done
for wl in cfg4_full cfg2_full_n5 cfg2_full_n6; do
  python bench.py --workload $wl --steps 1 --warmup 3 > $OUT/bench_1gpu_$wl.json 2> $OUT/bench_1gpu_$wl.err || echo "FAILED $wl"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02d/bench_1gpu_*.json")):
    try:
        d = [json.loads(l) for l in open(f) if l.startswith("{")][-1]
        print(f.split("bench_1gpu_")[1][:-5], "value %.3e e2e %.3e ms/step %.2f kernel %s frac %.3f cpu %.3e" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["kernel"], d["roofline"]["frac"], d["cpu_baseline"]["value"]))
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -3 $OUT/*.err | tail -30
