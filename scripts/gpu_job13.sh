#!/bin/bash
# round-end style verification: full GPU suite, headline bench, ncu launch list + full capture, level breakdown
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_256.json 2> gpurun_out/bench_256.log
tail -3 gpurun_out/bench_256.log
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_256.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_iter", "iterations", "relres", "e2e", "time_to_solution_ms", "device_reorder", "clocks")})
print(d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline"]["traffic"], d["cpu_baseline"]["value"])
PY
RCHOL_PROBE_CACHE=1 RCHOL_PROBE_MAXIT=40 timeout 600 python scripts/gpu_bc_probe.py 256 8 0,0 0,0,0,0,2048 0,0,0,1 > gpurun_out/probe256j.log 2>&1
grep -E "^---|pcg it|fwd level|bwd level|CTA0" gpurun_out/probe256j.log | cut -c1-420
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches128.csv python scripts/profile_target.py 128 8 2 > gpurun_out/ncu_launches.log 2>&1
tail -1 gpurun_out/ncu_launches.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_bc_solve -c 2 -o gpurun_out/bc_solve_128_final -f python scripts/profile_target.py 128 8 1 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
