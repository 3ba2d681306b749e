#!/bin/bash
# A/B: ziggurat table staged in shared memory (spectral kernels) vs read through L1
OUT=gpurun_out/r02y
mkdir -p $OUT
python -m pytest tests -m gpu -q -x > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
tail -3 $OUT/pytest.log
for z in 0 1; do
  RC_SPEC_ZIGS=$z python tools/kernel_bench.py --ns 11,12,14,16,17,18,20,24,28,32 --reps 3 > $OUT/kb_z$z.txt 2>&1
  RC_SPEC_ZIGS=$z python tools/kernel_bench.py --ns 16,32 --fused 1 --B 100000 --evals 4e7 > $OUT/kbf_z$z.txt 2>&1
done
grep -h evals_per_s $OUT/kb_z0.txt | cut -c1-110
echo ---
grep -h evals_per_s $OUT/kb_z1.txt | cut -c1-110
echo --- fused
grep -h evals_per_s $OUT/kbf_z0.txt $OUT/kbf_z1.txt | cut -c1-110
