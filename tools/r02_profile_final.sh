#!/bin/bash
# ncu evidence of the FINAL round-2 step: launch list + full capture of the four-lane statistics kernel
O=gpurun_out/r02pf; mkdir -p $O
B="python bench.py --steps 2 --warmup 3 --no-mcdatasim --cpu-evals 200 --skip-e2e"
$B > $O/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_paper_n7.csv $B > $O/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stats_unsorted_group -s 2 -c 1 -o $O/prof_stats_group $B > $O/ncu_full_stats.log 2>&1
ls -la $O
