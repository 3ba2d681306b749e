#!/bin/bash
# Round-2: per-lane accumulation in the fused register kernels (A/B against build/variants/lib_oldacc.so) + stats quad kernel validation
O=gpurun_out/r02la; mkdir -p $O
python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -3 $O/pytest.log
KB="python tools/kernel_bench.py"
$KB --ns 4,5,6,7,12 --fused 1 --B 100000 --evals 4e7 > $O/kbf_new.txt 2>&1
RC_LIB_PATH=build/variants/lib_oldacc.so $KB --ns 4,5,6,7,12 --fused 1 --B 100000 --evals 4e7 > $O/kbf_old.txt 2>&1
for f in kbf_new kbf_old; do echo "== $f"; grep evals_per_s $O/$f.txt | python -c "
import sys, json
for l in sys.stdin:
    j = json.loads(l); print(j['n'], '%.4g' % j['evals_per_s'], '%.3f' % j['frac_fp64_peak'])"; done
python bench.py --no-mcdatasim --cpu-evals 2000 > $O/bench_default.json 2> $O/bench_default.err
python -c "
import json; d=[json.loads(l) for l in open('$O/bench_default.json') if l.startswith('{')][-1]; print(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['per_step_ms']['median'])"
