#!/bin/bash
# Round-2 ncu evidence: launch list of the default bench step, full capture of the dominant kernel, DRAM traffic of
# the evolution kernel of every bench workload, full captures of the spectral shared-memory kernel at N=16 / 32.
OUT=gpurun_out/r02p
mkdir -p $OUT
B="python bench.py --steps 2 --warmup 3 --no-mcdatasim --cpu-evals 200"
$B > $OUT/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_paper_n7.csv $B > $OUT/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fidelity_reg_kernel -s 3 -c 1 -o $OUT/prof_paper_n7 $B > $OUT/ncu_full_n7.log 2>&1
ncu --set full --clock-control none -k regex:stats_unsorted -s 3 -c 1 -o $OUT/prof_stats_n7 $B > $OUT/ncu_full_stats.log 2>&1
for wl in n4_paper n5_paper n6_paper n16_paper cfg5_n32 cfg4_n16 cfg2_n5 cfg2_n6 cfg1_n4; do
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:fidelity_ -s 3 -c 1 --csv \
      --log-file $OUT/traffic_$wl.csv $B --workload $wl > $OUT/ncu_traffic_$wl.log 2>&1
done
python tools/kernel_bench.py --ns 16,32 --reps 1 > $OUT/plain_kb.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fidelity_smem_kernel -s 2 -c 1 -o $OUT/prof_smem_n16 python tools/kernel_bench.py --ns 16 --reps 1 > $OUT/ncu16.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fidelity_smem_kernel -s 2 -c 1 -o $OUT/prof_smem_n32 python tools/kernel_bench.py --ns 32 --reps 1 > $OUT/ncu32.log 2>&1
ls -la $OUT | head -40
