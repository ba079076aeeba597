#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_sdd.py tests/test_gpu_permutation.py tests/test_gpu_parity.py -x -q -k "sdd or permut or reorder or original or cxx or config0 or config1 or refresh" ) > gpurun_out/pytest_new.log 2>&1
tail -25 gpurun_out/pytest_new.log
