#!/bin/bash
# Round-2 A/B of the pinned-end register QL against the block-at-0 form (build/variants/lib_oldql.so).
mkdir -p gpurun_out/r02p
O=gpurun_out/r02p
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -3 $O/pytest.log
python tools/kernel_bench.py --ns 3,4,5,6,7,8 > $O/kb_new.txt 2>&1
RC_LIB_PATH=build/variants/lib_oldql.so python tools/kernel_bench.py --ns 3,4,5,6,7,8 > $O/kb_old.txt 2>&1
python tools/kernel_bench.py --ns 4,5,6,7,8 --fused 1 --B 100000 --evals 4e7 > $O/kbf_new.txt 2>&1
RC_LIB_PATH=build/variants/lib_oldql.so python tools/kernel_bench.py --ns 4,5,6,7,8 --fused 1 --B 100000 --evals 4e7 > $O/kbf_old.txt 2>&1
python tools/kernel_bench.py --ns 5,7 --replay 1 --evals 1e7 > $O/kbr_new.txt 2>&1
RC_LIB_PATH=build/variants/lib_oldql.so python tools/kernel_bench.py --ns 5,7 --replay 1 --evals 1e7 > $O/kbr_old.txt 2>&1
python bench.py > $O/bench_default.json 2> $O/bench_default.err
for f in kb_new kb_old kbf_new kbf_old kbr_new kbr_old; do echo "== $f"; grep evals_per_s $O/$f.txt | python -c "
import sys, json
for l in sys.stdin:
    j = json.loads(l); print(j['n'], '%.4g' % j['evals_per_s'], '%.3f' % j['frac_fp64_peak'])"; done
tail -c 600 $O/bench_default.json
