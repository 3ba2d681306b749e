#!/bin/bash
OUT=gpurun_out/r02b
mkdir -p $OUT
python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
tail -4 $OUT/pytest.log
python tools/kernel_bench.py --ns 7,8,9,10,11,12,16,24,32 --reps 3 > $OUT/kb.txt 2>&1
grep -h evals_per_s $OUT/kb.txt | cut -c1-120
python tools/kernel_bench.py --ns 16,32 --reps 1 > $OUT/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fidelity_smem_kernel -s 2 -c 1 -o $OUT/prof_n16 python tools/kernel_bench.py --ns 16 --reps 1 > $OUT/ncu16.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fidelity_smem_kernel -s 2 -c 1 -o $OUT/prof_n32 python tools/kernel_bench.py --ns 32 --reps 1 > $OUT/ncu32.log 2>&1
ls -la $OUT
