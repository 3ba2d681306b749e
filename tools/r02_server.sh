#!/bin/bash
# Resident objective evaluator: parity tests, then per-call latency with / without it
OUT=gpurun_out/r02v
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "objective or qnewton or lbfgs or wass or RL or environment or optimis" > $OUT/pytest_obj.log 2>&1; echo "pytest-obj rc=$?" >> $OUT/pytest_obj.log
tail -3 $OUT/pytest_obj.log
timeout 200 python tools/latency_bench.py > $OUT/lat_server.txt 2>&1; echo "rc=$?" >> $OUT/lat_server.txt
RC_OBJECTIVE_SERVER=0 timeout 200 python tools/latency_bench.py > $OUT/lat_oneshot.txt 2>&1; echo "rc=$?" >> $OUT/lat_oneshot.txt
RC_OBJECTIVE_IDLE_US=50 timeout 200 python tools/latency_bench.py > $OUT/lat_server_idle50.txt 2>&1; echo "rc=$?" >> $OUT/lat_server_idle50.txt
tail -4 $OUT/lat_server.txt | cut -c1-900; tail -3 $OUT/lat_oneshot.txt | cut -c1-900; tail -3 $OUT/lat_server_idle50.txt | cut -c1-900
timeout 600 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
tail -3 $OUT/pytest.log
