#!/bin/bash
# bench.py at N GPUs exactly as the driver launches it (ours, then the reference arm)
N=$1
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err; echo "bench N=$N rc=$?"
python - $N <<'PY'
import json, sys
N = sys.argv[1]
d = json.loads(open(f"gpurun_out/r02_bench_{N}gpu.json").read().strip().splitlines()[-1])
print("FE", d["n_gpus"], d["ms_per_step"], d["value"], "e2e", d["e2e"]["value"])
print("EM", d["em"]["ms_per_step"], d["em"]["value"])
print("C5", d["c5_strong"]["fe"]["ms_per_step"], d["c5_strong"]["em"]["ms_per_step"]); print("C4", d["c4_sweep"]["fe"]["launch_ms"], d["c4_sweep"]["em"]["launch_ms"], d["c4_sweep"]["fe"]["gpu_launches"], d["c4_sweep"]["em"]["nearest_to_README_point"]["E[X]"])
print("group_check", d.get("group_check"))
PY
tail -2 gpurun_out/r02_bench_${N}gpu.err
