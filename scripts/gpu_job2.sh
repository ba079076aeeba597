#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_blocked.py -x -q ) > gpurun_out/pytest_blocked.log 2>&1
tail -3 gpurun_out/pytest_blocked.log
timeout 200 python scripts/gpu_bc_probe.py 128 8 0,0 0,0,0,1 0,0,0,0,0,8 0,0,0,0,0,6 > gpurun_out/probe128b.log 2>&1
grep -E "^---|pcg it|level 0|level 3" gpurun_out/probe128b.log
timeout 900 python scripts/gpu_bc_probe.py 256 8 0,0 0,0,0,1 0,0,0,0,0,8 0,0,0,0,0,6 0,0,0,0,0,4 2048,0 8192,0 2048,0,0,0,0,6 0,1 0,1,0,0,0,6 > gpurun_out/probe256.log 2>&1
grep -E "^---|pcg it|fwd level 0|bwd level 3|late-t" gpurun_out/probe256.log
