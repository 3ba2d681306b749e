#!/bin/bash
# Round-2: register kernels now serve N <= 12 (512-lane CTAs at N = 9..12); does the register family also win at N = 13..16?
O=gpurun_out/r02s2; mkdir -p $O
KB="python tools/kernel_bench.py"
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -3 $O/pytest.log
$KB --ns 9,10,11,12,13,14,15,16 > $O/kb_main.txt 2>&1
RC_REG_MAX_N=16 RC_LIB_PATH=build/variants/lib_reg384.so $KB --ns 13,14,15,16 > $O/kb_reg384.txt 2>&1
$KB --ns 9,10,11,12 --fused 1 --B 100000 --evals 4e7 > $O/kbf_main.txt 2>&1
RC_REG_MAX_N=8 $KB --ns 9,10,11,12 --fused 1 --B 100000 --evals 4e7 > $O/kbf_smem.txt 2>&1
$KB --ns 9,12 --replay 1 --evals 1e7 > $O/kbr_main.txt 2>&1
RC_REG_MAX_N=8 $KB --ns 9,12 --replay 1 --evals 1e7 > $O/kbr_smem.txt 2>&1
for f in kb_main kb_reg384 kbf_main kbf_smem kbr_main kbr_smem; do echo "== $f"; grep evals_per_s $O/$f.txt | python -c "
import sys, json
for l in sys.stdin:
    j = json.loads(l); print(j['n'], '%.4g' % j['evals_per_s'], '%.3f' % j['frac_fp64_peak'])"; done
