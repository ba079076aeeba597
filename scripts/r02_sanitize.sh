#!/bin/bash
# compute-sanitizer runs of the triangular-solve kernels on small cases (SURVEY section 5 asks for memcheck / racecheck
# evidence of the ready-flag protocols).  Usage (GPU box): bash scripts/r02_sanitize.sh > gpurun_out/r02_sanitize.txt 2>&1
# racecheck only understands barriers: the flag hand-offs of the chain kernel (volatile / release-acquire words in shared
# memory) are reported as hazards by construction; the summary lists them by kernel so that NEW ones stand out.
cd "$(dirname "$0")/.."
cat > /tmp/san_target.py <<'PY'
import sys, numpy as np
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from conftest import load_golden, make_problem, relerr
from rchol_b200 import capi
from oracle import oracle
case, wb = sys.argv[1], int(sys.argv[2])
dp = dict(dp_min_rows=32, dp_panel=64) if len(sys.argv) > 3 and sys.argv[3] == "c64" else {}   # dense-panel levels: default / many short hops
if case == "golden":
    g = load_golden("lap3d_12_t4"); A, G, b, part = g["A"], g["G"], g["b"], g["part"]
else:
    A, b, G, part, f = make_problem("lap3d", 40, 8)
zo = oracle.precond(*G, b)
with capi.Solver(0, use_graph=False, wb_min=wb, **dp) as s:
    s.set_matrix(*A); s.set_factor(*G, part)
    z = s.precond(b)
    print(case, "wb_min", wb, "dp", dp, "precond relerr", relerr(z, zo))
    x, relres, itr = s.pcg(b, 1e-8, 50)
    print("pcg", itr, relres)
PY
for tool in memcheck racecheck; do
  for case in golden lap40; do
    for wb in "-1 def" "2 def" "2 c64"; do
      echo "=== compute-sanitizer --tool $tool  case=$case wb_min / dense-panel variant = $wb ==="
      timeout 900 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san_target.py $case $wb 2>&1 | grep -E "relerr|pcg |ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|=========     at|Error" | sort | uniq -c | sort -rn | head -25
    done
  done
done
