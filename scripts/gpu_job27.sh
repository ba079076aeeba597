#!/bin/bash
# round 2, last seconds of the GPU budget: parity subset on the final library (every test runs the changed row sort), then the
# set-up times at 128^3 / T = 512 (before this change: upload 97 ms, analysis 35 ms -- gpurun_out/c44_bench128.err)
mkdir -p gpurun_out
( time timeout 24 python -m pytest tests/test_gpu_parity.py tests/test_gpu_blocked.py tests/test_gpu_dense.py -m gpu -x -q --timeout 20 \
    -k "kat or goldens or error_reporting or value_only or matrix_replaced or kernel_configurations or stock_signature or seeded or chain_mode or empty_separator" ) > gpurun_out/c45_tests.log 2>&1
tail -3 gpurun_out/c45_tests.log | head -2
RCG_TIMING=1 timeout 30 python - > gpurun_out/c45_setup128.log 2>&1 <<'PY'
import sys, time
sys.path.insert(0, ".")
import bench
from rchol_b200 import capi
d, info = bench.build_problem(128, 512)
A = (d["A_rp"], d["A_ci"], d["A_v"]); G = (d["G_rp"], d["G_ci"], d["G_v"])
for rep in range(3):
    with capi.Solver(0) as s:
        t0 = time.time(); s.set_matrix(*A); t1 = time.time(); s.set_factor(*G, d["part"]); t2 = time.time()
        st = s.stats()
        print(f"rep {rep}: set_matrix {1e3*(t1-t0):.1f} ms, set_factor {1e3*(t2-t1):.1f} ms, upload {st['upload_ms']:.1f}, analysis {st['analysis_ms']:.1f}", flush=True)
PY
grep "^rep\|transpose" gpurun_out/c45_setup128.log | tail -8
