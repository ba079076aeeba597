#!/bin/bash
# two GPUs: sharded solve against the oracle (tests/test_gpu_multi.py, world 2) and a short two-rank bench at 128^3
mkdir -p gpurun_out
nvidia-smi -L
( time timeout 900 python -m pytest tests/test_gpu_multi.py -x -q --timeout 800 ) > gpurun_out/c38_multi_tests.log 2>&1
tail -6 gpurun_out/c38_multi_tests.log | cut -c1-600
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 2 --warmup 1 --n 128 --threads 512 > gpurun_out/c38_bench_n2.json 2> gpurun_out/c38_bench_n2.err
tail -2 gpurun_out/c38_bench_n2.err | cut -c1-300
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/c38_bench_n2.json").read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("value", "n_gpus", "ms_per_iter", "iterations", "relres", "e2e", "assembled_true_relres", "nccl_comm_init_ms")})
except Exception as e:
    print("no bench line", e)
PY
timeout 600 python bench.py --steps 2 --warmup 1 --n 128 --threads 512 --no-configs1 --no-cpu-baseline > gpurun_out/c38_bench_n1.json 2> gpurun_out/c38_bench_n1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/c38_bench_n1.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "n_gpus", "ms_per_iter", "iterations", "relres", "e2e", "time_to_solution_ms")})
PY
