#!/bin/bash
# Round-2: eigenvalue-only register solver with the pinned-end chase (tuning build lib_rspec.so, RC_REG_SPECTRAL=1) vs the shipped kernels.
O=gpurun_out/r02rs; mkdir -p $O
KB="python tools/kernel_bench.py"
V=build/variants/lib_rspec.so
RC_LIB_PATH=$V python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "every_chain_length or philox_mode_matches or split_matrices or replay or kat or full_size" > $O/pytest_rspec.log 2>&1; echo "pytest rc=$?" >> $O/pytest_rspec.log; tail -4 $O/pytest_rspec.log
$KB --ns 4,5,6,7,8,9,10,12,13,14,16 > $O/kb_main.txt 2>&1
RC_LIB_PATH=$V $KB --ns 4,5,6,7,8,9,10,12,13,14,16 > $O/kb_rspec.txt 2>&1
$KB --ns 7,16 --fused 1 --B 100000 --evals 4e7 > $O/kbf_main.txt 2>&1
RC_LIB_PATH=$V $KB --ns 7,16 --fused 1 --B 100000 --evals 4e7 > $O/kbf_rspec.txt 2>&1
for f in kb_main kb_rspec kbf_main kbf_rspec; do echo "== $f"; grep evals_per_s $O/$f.txt | python -c "
import sys, json
for l in sys.stdin:
    j = json.loads(l); print(j['n'], '%.4g' % j['evals_per_s'], '%.3f' % j['frac_fp64_peak'])"; done
