#!/bin/bash
# 2-GPU call: the skipped group tests, bench at N=2 (ours + reference arm) as the driver launches it
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 900 python -m pytest tests/test_gpu_group.py tests/test_gpu_strikes.py tests/test_gpu_checked_build.py -m gpu -q > gpurun_out/r02_test_2gpu.log 2>&1; echo "pytest 2gpu rc=$?"; tail -5 gpurun_out/r02_test_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err; echo "bench N=2 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02_bench_2gpu_ref.json 2> gpurun_out/r02_bench_2gpu_ref.err; echo "ref arm N=2 rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_2gpu.json").read().strip().splitlines()[-1])
print("FE N=2", d["ms_per_step"], d["value"], "em", d["em"]["ms_per_step"], d["em"]["value"], "c5", d["c5_strong"]["fe"]["ms_per_step"], d["c5_strong"]["em"]["ms_per_step"])
print("group_check", d.get("group_check"))
r = json.loads(open("gpurun_out/r02_bench_2gpu_ref.json").read().strip().splitlines()[-1])
print("ref arm", r["value"], r["cpu_baseline"]["cores"], r["cpu_baseline"]["init_s"], r["n_gpus"])
PY
tail -3 gpurun_out/r02_bench_2gpu.err
