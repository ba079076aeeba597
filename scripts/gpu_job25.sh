#!/bin/bash
# round 2, last GPU call (3.6 GPU-minutes left): the C-side changes of this session against a fast subset of the parity
# tests, smoke(), then the configs[3] bench leg at 2048^2 (what the default bench runs at 4096^2 in a child process)
mkdir -p gpurun_out
( time timeout 75 python -m pytest tests/test_gpu_parity.py tests/test_gpu_permutation.py -m gpu -x -q --timeout 60 \
    -k "kat or goldens or error_reporting or value_only or matrix_replaced or reorder_matches or permutation_errors or original_ordering" ) > gpurun_out/c43_tests.log 2>&1
tail -3 gpurun_out/c43_tests.log
( time timeout 40 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/c43_smoke.log 2>&1
tail -4 gpurun_out/c43_smoke.log
timeout 150 python bench.py --leg aniso2d --n 2048 --threads 4096 --steps 2 --warmup 2 > gpurun_out/c43_aniso2048.json 2> gpurun_out/c43_aniso2048.err
echo "aniso leg rc=$?"
grep "^\[bench\]" gpurun_out/c43_aniso2048.err | tail -6
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/c43_aniso2048.json").read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("workload", "iterations", "ms_per_iter", "value", "frac_of_peak", "trsv", "spmv_ms", "parity", "factor_s")})
except Exception as e:
    print("no leg line", e)
PY
