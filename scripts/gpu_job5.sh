#!/bin/bash
# timing experiments of the chain (results are wrong by construction with dbg bits 2 / 32): where does the chunk time go?
mkdir -p gpurun_out
RCHOL_PROBE_CACHE=1 RCHOL_PROBE_MAXIT=4 timeout 1200 python scripts/gpu_bc_probe.py 256 8 0,0 0,0,0,2 0,0,0,32 0,0,0,34 0,0,0,16 0,0,3,0 > gpurun_out/probe256e.log 2>&1
grep -E "^---|pcg it|fwd level|bwd level" gpurun_out/probe256e.log | cut -c1-330
