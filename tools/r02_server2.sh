#!/bin/bash
OUT=gpurun_out/r02u
mkdir -p $OUT
for t in 0 32 64 128 256; do RC_OBJECTIVE_MIN_THREADS=$t timeout 100 python tools/server_probe.py 2>&1 | tail -1; done | tee $OUT/probe.txt
RC_OBJECTIVE_SERVER=0 timeout 100 python tools/server_probe.py 2>&1 | tail -1 | tee -a $OUT/probe.txt
