#!/bin/bash
# Round-2 closing validation of HEAD on one B200: GPU suite, smoke(), default bench line (both arms), refreshed full
# captures of the spectral shared-memory kernel (final LD-templated build) at N = 16 / 32.
OUT=gpurun_out/r02z
mkdir -p $OUT
python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
tail -3 $OUT/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/smoke.log
tail -4 $OUT/smoke.log
python bench.py --steps 20 --warmup 5 > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "ref rc=$?"
cut -c1-600 $OUT/bench_default.json; cut -c1-400 $OUT/bench_reference.json
python tools/kernel_bench.py --ns 11,12,16,24,32 --reps 3 > $OUT/kb.txt 2>&1
grep -h evals_per_s $OUT/kb.txt | cut -c1-140
ncu --set full --clock-control none --import-source on -k regex:fidelity_smem_kernel -s 2 -c 1 -o $OUT/prof_smem_n16 python tools/kernel_bench.py --ns 16 --reps 1 > $OUT/ncu16.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fidelity_smem_kernel -s 2 -c 1 -o $OUT/prof_smem_n32 python tools/kernel_bench.py --ns 32 --reps 1 > $OUT/ncu32.log 2>&1
ls -la $OUT | head -30
