#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_permutation.py "tests/test_gpu_parity.py::test_cxx_driver_runs_like_the_reference_example" -x -q ) > gpurun_out/pytest_perm.log 2>&1
tail -25 gpurun_out/pytest_perm.log
