#!/bin/bash
# end-of-round verification: full GPU suite, headline bench (with the one-shot stage times), ncu launch list with DRAM bytes
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q --timeout 600 ) > gpurun_out/c39_gputests.log 2>&1
grep -E "passed|failed|error" gpurun_out/c39_gputests.log | tail -2
RCG_TIMING=1 timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/c39_bench.json 2> gpurun_out/c39_bench.err
grep "one-shot\|dp_build" gpurun_out/c39_bench.err | head -12
python - <<'PY'
import json
d = json.loads(open("gpurun_out/c39_bench.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_iter", "iterations", "relres", "e2e", "time_to_solution_ms", "clocks", "gpu_launches")})
r = d["roofline"]; print(r["achieved"], r["frac"], r.get("traffic"), r.get("traffic_over_algorithmic"), r["iteration"])
print(d.get("configs1")); print(d.get("cpu_baseline")); print(d.get("parity"))
PY
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_T4096_final.csv python scripts/r02_ncu_target.py 256 4096 2 > gpurun_out/c39_ncu_list.log 2>&1
tail -1 gpurun_out/c39_ncu_list.log
