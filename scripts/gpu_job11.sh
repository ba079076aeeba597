#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_blocked.py -x -q ) > gpurun_out/pytest_blocked.log 2>&1
tail -3 gpurun_out/pytest_blocked.log
RCHOL_PROBE_CACHE=1 RCHOL_PROBE_MAXIT=40 timeout 1200 python scripts/gpu_bc_probe.py 256 8 0,0 0,0,0,0,0,0,0,0,0,1 0,0,0,0,0,0,0,0,0,4 0,0,0,0,0,0,0,0,0,8 0,0,0,0,2048 0,0,0,1 > gpurun_out/probe256h.log 2>&1
grep -E "^---|pcg it|fwd level|bwd level|CTA0" gpurun_out/probe256h.log | cut -c1-420
