#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_dense.py -x -q --timeout 300 ) > gpurun_out/c32_dense_tests.log 2>&1
grep -E "passed|failed" gpurun_out/c32_dense_tests.log
timeout 600 python scripts/r02_dp_trace.py 128 512 0 9 > gpurun_out/c32_trace_fwd9.log 2>&1
tail -19 gpurun_out/c32_trace_fwd9.log
timeout 600 python scripts/r02_chain_perf.py 128 512 def:dbg=1 > gpurun_out/c32_perf128_T512.jsonl 2> gpurun_out/c32_perf128_T512.err
python - <<'PY'
import json
for l in open("gpurun_out/c32_perf128_T512.jsonl"):
    d = json.loads(l)
    print(d["variant"], d.get("error"), d.get("iterations"), d.get("ms_per_iter"), d.get("analysis_ms"))
    if "levels" in d:
        print("   ", {k: v["ms"] for k, v in d["levels"].items()})
PY
