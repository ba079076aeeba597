#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_dp_solve --launch-skip 8 --launch-count 1 -o gpurun_out/c29_dp128 -f python scripts/r02_ncu_target.py 128 512 1 > gpurun_out/c29_ncu.log 2>&1
tail -2 gpurun_out/c29_ncu.log
