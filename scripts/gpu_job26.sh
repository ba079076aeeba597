#!/bin/bash
# round 2: the whole N = 1 bench flow with this session's additions (refresh object, configs[3] leg in a child process) at a
# small size, then the aniso leg at 2048^2 with T = 1024 for the leaf-count comparison
mkdir -p gpurun_out
( time timeout 110 python bench.py --n 128 --threads 512 --aniso-n 1024 --aniso-threads 1024 --no-configs1 --steps 2 --warmup 3 ) > gpurun_out/c44_bench128.json 2> gpurun_out/c44_bench128.err
echo "bench rc=$?"
grep "^\[bench\]\|Traceback\|Error" gpurun_out/c44_bench128.err | tail -12
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/c44_bench128.json").read().splitlines() if l.startswith("{")][-1])
    print({k: d.get(k) for k in ("value", "ms_per_iter", "iterations", "refresh", "parity")})
    c3 = d.get("configs3") or {}
    print("configs3:", {k: c3.get(k) for k in ("workload", "iterations", "ms_per_iter", "value", "frac_of_peak", "error", "leg_wall_s")})
except Exception as e:
    print("no bench line", e)
PY
timeout 60 python bench.py --leg aniso2d --n 2048 --threads 1024 --steps 2 --warmup 2 --no-parity > gpurun_out/c44_aniso2048_T1024.json 2> gpurun_out/c44_aniso2048_T1024.err
echo "aniso T=1024 leg rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/c44_aniso2048_T1024.json").read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("workload", "iterations", "ms_per_iter", "value", "frac_of_peak", "factor_s")})
except Exception as e:
    print("no leg line", e)
PY
