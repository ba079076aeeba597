#!/bin/bash
# dense-panel levels: parity first, then per-level times with and without them
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_dense.py -x -q --timeout 300 ) > gpurun_out/c18_dense_tests.log 2>&1
tail -5 gpurun_out/c18_dense_tests.log
timeout 600 python scripts/r02_chain_perf.py 128 512 def: nodp:dp_min_rows=-1 c256:dp_panel=256 c512:dp_panel=512 c1024:dp_panel=1024 > gpurun_out/c18_perf128_T512.jsonl 2> gpurun_out/c18_perf128_T512.err
python - <<'PY'
import json
for l in open("gpurun_out/c18_perf128_T512.jsonl"):
    d = json.loads(l)
    print(d["variant"], d.get("error"), d.get("iterations"), d.get("ms_per_iter"), d.get("analysis_ms"), d.get("device_gb"), d.get("x_vs_first_variant"))
    if "levels" in d: print("   ", {k: v["ms"] for k, v in d["levels"].items()})
PY
timeout 1500 python scripts/r02_chain_perf.py 256 4096 def: nodp:dp_min_rows=-1 c512:dp_panel=512 > gpurun_out/c18_perf256_T4096.jsonl 2> gpurun_out/c18_perf256_T4096.err
python - <<'PY'
import json
for l in open("gpurun_out/c18_perf256_T4096.jsonl"):
    d = json.loads(l)
    print(d["variant"], d.get("error"), d.get("iterations"), d.get("ms_per_iter"), d.get("analysis_ms"), d.get("device_gb"), d.get("x_vs_first_variant"))
    if "levels" in d: print("   ", {k: v["ms"] for k, v in d["levels"].items()})
PY
