#!/bin/bash
# round-2 final 1-GPU run: full GPU suite, smoke, bench (ours + reference arm), BASELINE configs C1-C5
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r02_test_gpu_final.log 2>&1; echo "pytest gpu rc=$?"; tail -4 gpurun_out/r02_test_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()"; echo "smoke rc=$?"
timeout 900 python bench.py --impl reference > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err; echo "ref arm rc=$?"
timeout 900 python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; echo "bench rc=$?"
timeout 900 python bench.py --method em --no-sub-records > gpurun_out/r02_bench_em_1gpu.json 2> gpurun_out/r02_bench_em_1gpu.err; echo "bench em rc=$?"
timeout 1500 python tests/run_configs.py --out gpurun_out/r02_configs_1gpu.json > gpurun_out/r02_configs_1gpu.log 2>&1; echo "configs rc=$?"; tail -3 gpurun_out/r02_configs_1gpu.log
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_1gpu.json"))
print("FE", d["ms_per_step"], d["value"], d["roofline"]["frac"], d["roofline"]["mix_bound_frac"], "e2e", d["e2e"]["value"])
print("EM", d["em"]["ms_per_step"], d["em"]["value"], d["em"]["roofline"]["frac"], d["em"]["roofline"]["mix_bound_frac"])
print("C5", d["c5_strong"]["fe"]["ms_per_step"], d["c5_strong"]["em"]["ms_per_step"])
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
x = d["reference_cuda"]["xorwow"]; print("ref xorwow", x["exec_ms"], x["own_spread"]["max_rel_spread_var"], x["ours_same_draws"]["max_rel_diff_var"], x["ours_same_stream_fast"]["max_rel_diff_var"], x.get("fast_vs_compat"))
r = json.load(open("gpurun_out/r02_bench_reference_arm.json")); print("ref arm", r["value"], r["cpu_baseline"]["cores"], r["ms_per_step"])
PY
