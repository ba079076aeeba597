#!/bin/bash
# bench line at the headline workload, ncu launch list (times + DRAM bytes) and one full capture of the root-level k_dp_solve at
# the same workload (the bench caches the factor under /tmp for the later steps), sanitizer runs on small cases
mkdir -p gpurun_out
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/c36_bench.json 2> gpurun_out/c36_bench.err
tail -3 gpurun_out/c36_bench.err | cut -c1-400
python - <<'PY'
import json
d = json.loads(open("gpurun_out/c36_bench.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_iter", "iterations", "relres", "e2e", "time_to_solution_ms", "clocks", "gpu_launches")})
r = d["roofline"]; print(r["achieved"], r["frac"], r["iteration"])
print(d.get("configs1")); print(d.get("cpu_baseline")); print(d.get("parity"))
PY
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_T4096.csv python scripts/r02_ncu_target.py 256 4096 2 > gpurun_out/c36_ncu_list.log 2>&1
tail -1 gpurun_out/c36_ncu_list.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_dp_solve --launch-skip 6 --launch-count 2 -o gpurun_out/c36_dp256 -f python scripts/r02_ncu_target.py 256 4096 1 > gpurun_out/c36_ncu_full.log 2>&1
tail -1 gpurun_out/c36_ncu_full.log
bash scripts/r02_sanitize.sh > gpurun_out/r02_sanitize.txt 2>&1
tail -30 gpurun_out/r02_sanitize.txt | cut -c1-200
