#!/bin/bash
# Multi-GPU cells of the throughput table: usage r02_table_multi.sh N [workloads...]
N=$1; shift
OUT=gpurun_out/r02e
mkdir -p $OUT
PORT=29600
for wl in "$@"; do
  PORT=$((PORT+1))
  steps=10; [[ $wl == *_full* ]] && steps=1
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
      bench.py --gpus $N --workload $wl --steps $steps --warmup 3 --cpu-evals 5000 > $OUT/bench_${N}gpu_$wl.out 2> $OUT/bench_${N}gpu_$wl.err || echo "FAILED $wl"
  grep "^{" $OUT/bench_${N}gpu_$wl.out | tail -1 > $OUT/bench_${N}gpu_$wl.json
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02e/bench_*gpu_*.json")):
    try:
        d = json.loads(open(f).read())
        ps = d["per_step_ms"]
        print(f.split("r02e/bench_")[1][:-5], "value %.3e e2e %.3e ms/step %.3f (min %.3f med %.3f max %.3f) tails %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], ps["min"], ps["median"], ps["max"], [round(p["exchange_tail"], 3) for p in ps["per_rank"]]))
    except Exception as e:
        print(f, "unreadable", e)
PY
