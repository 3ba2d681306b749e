#!/bin/bash
OUT=gpurun_out/r02t
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "objective or qnewton or lbfgs or wass or RL or environment or optimis" > $OUT/pytest_obj.log 2>&1; echo "pytest-obj rc=$?" >> $OUT/pytest_obj.log
tail -3 $OUT/pytest_obj.log
RC_OBJECTIVE_SERVER=0 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "objective or qnewton or lbfgs or wass or RL or environment or optimis" > $OUT/pytest_obj0.log 2>&1; echo "pytest-obj0 rc=$?" >> $OUT/pytest_obj0.log
tail -2 $OUT/pytest_obj0.log
for t in 0 128 256; do RC_OBJECTIVE_MIN_THREADS=$t timeout 100 python tools/server_probe.py 2>&1 | tail -1; done | tee $OUT/probe.txt
RC_OBJECTIVE_SERVER=0 timeout 100 python tools/server_probe.py 2>&1 | tail -1 | tee -a $OUT/probe.txt
timeout 200 python tools/latency_bench.py > $OUT/lat_server.txt 2>&1; tail -2 $OUT/lat_server.txt | cut -c1-700
