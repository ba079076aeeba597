#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/r02_dp_trace.py 128 512 0 9 > gpurun_out/c23_trace_fwd9.log 2>&1
cat gpurun_out/c23_trace_fwd9.log | tail -45
timeout 600 python scripts/r02_dp_trace.py 128 512 0 6 > gpurun_out/c23_trace_fwd6.log 2>&1
cat gpurun_out/c23_trace_fwd6.log | tail -45
