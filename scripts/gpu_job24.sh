#!/bin/bash
mkdir -p gpurun_out
RCG_TIMING=1 timeout 900 python bench.py --steps 2 --warmup 1 --no-configs1 --no-cpu-baseline --no-parity > gpurun_out/c42_bench.json 2> gpurun_out/c42_bench.err
grep "one-shot\|allocations, memsets\|dp_build:" gpurun_out/c42_bench.err | tail -12
python - <<'PY'
import json
d = json.loads(open("gpurun_out/c42_bench.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_iter", "e2e", "time_to_solution_ms")})
PY
( time timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 ) > gpurun_out/c42_gputests.log 2>&1
grep -E "passed|failed|error" gpurun_out/c42_gputests.log | tail -2
