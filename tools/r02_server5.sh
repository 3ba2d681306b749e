#!/bin/bash
OUT=gpurun_out/r02r
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "objective or qnewton or lbfgs or wass or RL or environment or optimis or resident" > $OUT/pytest_obj.log 2>&1; echo "pytest-obj rc=$?" >> $OUT/pytest_obj.log
tail -3 $OUT/pytest_obj.log
timeout 100 python tools/server_probe.py 2>&1 | tail -1 | tee $OUT/probe.txt
RC_OBJECTIVE_REGQL=0 timeout 100 python tools/server_probe.py 2>&1 | tail -1 | tee -a $OUT/probe.txt
timeout 200 python tools/latency_bench.py > $OUT/lat_server.txt 2>&1; tail -2 $OUT/lat_server.txt | cut -c1-420
RC_OBJECTIVE_REGQL=0 timeout 200 python tools/latency_bench.py > $OUT/lat_server_smemql.txt 2>&1; tail -2 $OUT/lat_server_smemql.txt | cut -c1-420
