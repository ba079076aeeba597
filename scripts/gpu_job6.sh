#!/bin/bash
# 2-GPU functional check: sharded parity test + the N=2 leg of bench.py on a small workload
mkdir -p gpurun_out
nvidia-smi -L
( time timeout 600 python -m pytest tests/test_gpu_multi.py -x -q ) > gpurun_out/pytest_multi.log 2>&1
tail -5 gpurun_out/pytest_multi.log
export RCHOL_B200_BENCH_N=128
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_n128_g2.json 2> gpurun_out/bench_n128_g2.log
tail -5 gpurun_out/bench_n128_g2.log; cut -c1-1500 gpurun_out/bench_n128_g2.json
timeout 600 python bench.py --gpus 1 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n128_g1.json 2> gpurun_out/bench_n128_g1.log
python - <<'PY'
import json
for f in ("gpurun_out/bench_n128_g1.json", "gpurun_out/bench_n128_g2.json"):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, {k: d.get(k) for k in ("n_gpus", "value", "ms_per_iter", "iterations", "relres", "assembled_true_relres", "clocks")})
PY
