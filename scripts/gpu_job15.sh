#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_dense.py -x -q --timeout 300 ) > gpurun_out/c26_dense_tests.log 2>&1
tail -3 gpurun_out/c26_dense_tests.log
timeout 600 python scripts/r02_chain_perf.py 128 512 def:dbg=1 nodp:dp_min_rows=-1 > gpurun_out/c26_perf128_T512.jsonl 2> gpurun_out/c26_perf128_T512.err
python - <<'PY'
import json
for l in open("gpurun_out/c26_perf128_T512.jsonl"):
    d = json.loads(l)
    print(d["variant"], d.get("error"), d.get("iterations"), d.get("ms_per_iter"), d.get("analysis_ms"), d.get("device_gb"), d.get("x_vs_first_variant"))
    if "levels" in d:
        print("   ", {k: v["ms"] for k, v in d["levels"].items()})
        for k, v in d["levels"].items():
            if "dp" in v and k not in ("fwd0", "bwd9"): print("      ", k, v["blocks"], v["max_rows"], v["dp"])
PY
