#!/bin/bash
OUT=gpurun_out/r02s
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
tail -4 $OUT/pytest.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 200 python tools/latency_bench.py > $OUT/lat_server.txt 2>&1
RC_OBJECTIVE_SERVER=0 timeout 200 python tools/latency_bench.py > $OUT/lat_oneshot.txt 2>&1
tail -2 $OUT/lat_server.txt | cut -c1-300
