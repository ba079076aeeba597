"""Post-processing of an ncu launch list (csv, --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum) of
scripts/r02_ncu_target.py: sums the DRAM bytes and the durations of the triangular-solve level launches of ONE
preconditioner application (the last one in the log) and writes profiles/r02_trsv_traffic.json, which bench.py reads for
`roofline.traffic`.  Usage: python scripts/r02_traffic.py launches.csv n T > profiles/r02_trsv_traffic.json"""
import csv, json, sys
path, n, T = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
launches = {}
for r in rows[1:]:
    try:
        lid, name, metric, val = int(r[ix["ID"]]), r[ix["Kernel Name"]], r[ix["Metric Name"]], float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
    except (ValueError, KeyError):
        continue
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1)
    launches.setdefault(lid, dict(name=name))[metric] = val * scale
seq = [launches[k] for k in sorted(launches)]
solve = [l for l in seq if any(t in l["name"] for t in ("k_wb_solve", "k_wb_pre", "k_bc_solve", "k_fc_solve", "k_dp_solve", "k_dp_pre"))]
half = len(solve) // 2                       # the target applies the preconditioner twice: keep the second application
solve = solve[half:]
by = {}
for l in solve:
    k = l["name"].split("(")[0].split("::")[-1]
    e = by.setdefault(k, dict(launches=0, ms=0.0, dram_bytes=0.0))
    e["launches"] += 1
    e["ms"] += l.get("gpu__time_duration.sum", 0.0)
    e["dram_bytes"] += l.get("dram__bytes_read.sum", 0.0) + l.get("dram__bytes_write.sum", 0.0)
out = dict(workload=f"lap3d_{n}^3_rchol_T{T}_pcg_tol1e-8", source=f"ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
           f"--clock-control none over scripts/r02_ncu_target.py {n} {T} 2 (second preconditioner application)",
           launches_per_solve_pair=len(solve), dram_bytes_per_solve_pair=sum(e["dram_bytes"] for e in by.values()),
           ncu_ms_per_solve_pair=sum(e["ms"] for e in by.values()), by_kernel=by)
print(json.dumps(out, indent=1))
