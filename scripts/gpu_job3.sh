#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_256.json 2> gpurun_out/bench_256.log
tail -12 gpurun_out/bench_256.log
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_256.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_iter", "iterations", "relres", "e2e", "time_to_solution_ms", "cpu_baseline", "clocks")})
PY
RCHOL_PROBE_CACHE=1 timeout 600 python scripts/gpu_bc_probe.py 256 8 0,0 0,0,0,1 > gpurun_out/probe256c.log 2>&1
grep -E "^---|pcg it|level|CTA0" gpurun_out/probe256c.log
