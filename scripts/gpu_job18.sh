#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_gpu_dense.py -x -q --timeout 600 ) > gpurun_out/c35_gputests.log 2>&1
tail -4 gpurun_out/c35_gputests.log
timeout 1500 python scripts/r02_chain_perf.py 256 4096 def: > gpurun_out/c35_perf256_T4096.jsonl 2> gpurun_out/c35_perf256_T4096.err
python - <<'PY'
import json
for l in open("gpurun_out/c35_perf256_T4096.jsonl"):
    d = json.loads(l)
    print(d["variant"], d.get("error"), d.get("iterations"), d.get("ms_per_iter"), d.get("analysis_ms"), d.get("device_gb"), d.get("x_vs_first_variant"), d.get("split_ms"))
    if "levels" in d: print("   ", {k: v["ms"] for k, v in d["levels"].items()})
PY
