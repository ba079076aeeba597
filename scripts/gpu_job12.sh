#!/bin/bash
mkdir -p gpurun_out
RCHOL_PROBE_CACHE=1 RCHOL_PROBE_MAXIT=40 timeout 1200 python scripts/gpu_bc_probe.py 256 8 0,0 0,0,0,0,0,8 0,0,0,0,0,6 0,0,0,0,0,4 0,0,0,0,0,0,6 0,0,0,0,0,6,6 0,0,0,0,0,0,0,3 > gpurun_out/probe256i.log 2>&1
grep -E "^---|pcg it|fwd level|bwd level|CTA0" gpurun_out/probe256i.log | cut -c1-420
