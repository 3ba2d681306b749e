#!/bin/bash
# Round-2: register kernels for N = 9..12 with the pinned-end QL (variants), 896-lane CTAs at N = 6..8, ncu evidence of the new N=7 kernel.
O=gpurun_out/r02q; mkdir -p $O
KB="python tools/kernel_bench.py"
$KB --ns 9,10,11,12 > $O/kb_base.txt 2>&1
RC_REG_MAX_N=12 $KB --ns 9,10,11,12 > $O/kb_reg128.txt 2>&1
RC_REG_MAX_N=12 RC_LIB_PATH=build/variants/lib_reg512.so $KB --ns 9,10,11,12 > $O/kb_reg512.txt 2>&1
RC_REG_MAX_N=12 RC_LIB_PATH=build/variants/lib_reg640.so $KB --ns 9,10,11 > $O/kb_reg640.txt 2>&1
RC_LIB_PATH=build/variants/lib_t896.so $KB --ns 6,7,8 > $O/kb_t896.txt 2>&1
$KB --ns 6,7,8 > $O/kb_t768.txt 2>&1
for f in kb_base kb_reg128 kb_reg512 kb_reg640 kb_t896 kb_t768; do echo "== $f"; grep evals_per_s $O/$f.txt | python -c "
import sys, json
for l in sys.stdin:
    j = json.loads(l); print(j['n'], '%.4g' % j['evals_per_s'], '%.3f' % j['frac_fp64_peak'])"; done
B="python bench.py --steps 2 --warmup 3 --no-mcdatasim --cpu-evals 200"
$B > $O/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_paper_n7.csv $B > $O/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fidelity_reg_kernel -s 3 -c 1 -o $O/prof_paper_n7 $B > $O/ncu_full_n7.log 2>&1
ls -la $O
