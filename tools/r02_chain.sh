#!/bin/bash
# A/B: dependent-chain length of the eigenvalue-only QL rotation (RC_QL_CHAIN = 0 / 1 / 2, variant libraries)
OUT=gpurun_out/r02x
mkdir -p $OUT
for v in 0 1 2; do
  L=""; [ $v != 0 ] && L="$PWD/build/variants/lib_chain$v.so"
  RC_LIB_PATH=$L python tools/kernel_bench.py --ns 11,12,16,20,24,28,32 --reps 3 > $OUT/kb_c$v.txt 2>&1
  echo "--- chain $v"; grep -h evals_per_s $OUT/kb_c$v.txt | cut -c1-100
done
RC_LIB_PATH=$PWD/build/variants/lib_chain1.so python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "spectral or smem or replay or golden or sweep" > $OUT/pytest_c1.log 2>&1; tail -2 $OUT/pytest_c1.log
RC_LIB_PATH=$PWD/build/variants/lib_chain2.so python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "spectral or smem or replay or golden or sweep" > $OUT/pytest_c2.log 2>&1; tail -2 $OUT/pytest_c2.log
