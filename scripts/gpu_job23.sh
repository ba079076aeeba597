#!/bin/bash
mkdir -p gpurun_out
RCG_TIMING=1 timeout 900 python bench.py --steps 2 --warmup 1 --no-configs1 --no-cpu-baseline --no-parity > gpurun_out/c41_bench.json 2> gpurun_out/c41_bench.err
grep "\[rcg\]" gpurun_out/c41_bench.err | grep -v "k_bc_count:\|after the launch" | tail -42
python - <<'PY'
import json
d = json.loads(open("gpurun_out/c41_bench.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_iter", "e2e", "time_to_solution_ms")})
PY
