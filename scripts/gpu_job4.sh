#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_blocked.py -x -q ) > gpurun_out/pytest_blocked.log 2>&1
tail -3 gpurun_out/pytest_blocked.log
RCHOL_PROBE_CACHE=1 timeout 1200 python scripts/gpu_bc_probe.py 256 8 0,0 0,0,0,0,0,0,8 0,0,0,0,0,0,6 2048,0,0,0,0,0,8 2048,0,0,0,0,0,6 0,0,0,0,0,0,8,3 0,0,0,0,0,0,8,5 2048,0,0,0,0,0,8,5 2048,0,0,0,0,0,8,6 2048,0,0,0,0,0,6,6 0,0,0,1,0,0,8 > gpurun_out/probe256d.log 2>&1
grep -E "^---|pcg it|fwd level 0|bwd level 3|CTA0|fwd plan" gpurun_out/probe256d.log | cut -c1-420
