#!/bin/bash
# A/B of the shared-memory family: spectral weights vs eigenvector rows, CTA-size sweeps (development aid)
OUT=gpurun_out/r02a
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
tail -5 $OUT/pytest.log
for algo in vectors spectral; do
  RC_SMEM_ALGO=$algo python tools/kernel_bench.py --ns 9,10,12,16,24,32 > $OUT/kb_$algo.txt 2>&1
  RC_SMEM_ALGO=$algo python tools/kernel_bench.py --ns 16,32 --fused 1 --B 100000 --evals 4e7 > $OUT/kb_fused_$algo.txt 2>&1
done
for t in 384 416 448; do RC_SMEM_THREADS=$t python tools/kernel_bench.py --ns 32 > $OUT/kb_spec_n32_t$t.txt 2>&1; done
for t in 512 640 768; do RC_SMEM_THREADS=$t python tools/kernel_bench.py --ns 16,12 > $OUT/kb_spec_n16_t$t.txt 2>&1; done
grep -h evals_per_s $OUT/kb_*.txt | cut -c1-150
