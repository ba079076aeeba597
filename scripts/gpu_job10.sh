#!/bin/bash
mkdir -p gpurun_out
RCHOL_PROBE_CACHE=1 RCHOL_PROBE_MAXIT=8 timeout 1200 python scripts/gpu_bc_probe.py 256 8 0,0,0,1 0,0,0,1,2048 > gpurun_out/probe256g.log 2>&1
grep -E "^---|pcg it|fwd level|bwd level|CTA0" gpurun_out/probe256g.log | cut -c1-420
