#!/bin/bash
# GPU job: parity suite, cycle breakdown of the chain at 128^3, ncu launch list and one full capture of k_bc_solve
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python scripts/gpu_bc_probe.py 128 8 0,0,0,1 > gpurun_out/probe128.log 2>&1
tail -30 gpurun_out/probe128.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches128.csv python scripts/profile_target.py 128 8 2 > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_bc_solve -c 2 -o gpurun_out/bc_solve_128 -f python scripts/profile_target.py 128 8 1 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
