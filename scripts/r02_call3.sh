#!/bin/bash
# round-2 GPU call 3: full GPU test suite on the new code, bench line, configs[0] size, microbench refresh
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( cd profiles/microbench && ./pipe_rates2 8 > ../../gpurun_out/r02_pipe_rates2_w8.txt 2>&1 )
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_test_gpu.log 2>&1; echo "pytest gpu rc=$?"
tail -15 gpurun_out/r02_test_gpu.log
timeout 600 python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; echo "bench rc=$?"
timeout 300 python bench.py --log2-paths 18 --steps 50 --no-sub-records --no-cpu-baseline --no-reference-cuda > gpurun_out/r02_bench_c1.json 2> gpurun_out/r02_bench_c1.err; echo "bench c1 rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_1gpu.json"))
print("FE", d["ms_per_step"], d["value"], d["roofline"]["frac"], d["roofline"]["mix_bound_frac"], "e2e", d["e2e"]["value"])
print("EM", d["em"]["ms_per_step"], d["em"]["value"], d["em"]["roofline"]["frac"], d["em"]["roofline"]["mix_bound_frac"])
print("C5", d["c5_strong"]["fe"]["ms_per_step"], d["c5_strong"]["em"]["ms_per_step"])
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], d["cpu_baseline"]["init_s"], d["cpu_baseline"]["sample_s"])
x = d["reference_cuda"]["xorwow"]; print("ref xorwow", x["exec_ms"], x["own_spread"], x["ours_same_draws"], x["ours_same_stream_fast"])
c = json.load(open("gpurun_out/r02_bench_c1.json")); print("C1 2^18", c["ms_per_step"], c["value"], c["roofline"]["frac"], c["kernel"])
PY
