#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q --timeout 600 ) > gpurun_out/c37_gputests.log 2>&1
tail -4 gpurun_out/c37_gputests.log
RCG_TIMING=1 timeout 600 python scripts/r02_chain_perf.py 128 512 def: def2: nodp:dp_min_rows=-1 > gpurun_out/c37_perf128_T512.jsonl 2> gpurun_out/c37_perf128_T512.err
grep "\[rcg\]" gpurun_out/c37_perf128_T512.err | head -30
python - <<'PY'
import json
for l in open("gpurun_out/c37_perf128_T512.jsonl"):
    d = json.loads(l)
    print(d["variant"], d.get("error"), d.get("iterations"), d.get("ms_per_iter"), "analysis", d.get("analysis_ms"), "setup wall", d.get("setup_wall_s"))
    if "levels" in d: print("   ", {k: v["ms"] for k, v in d["levels"].items()})
PY
