#!/bin/bash
# after the uniform-datapath change of the FE kernel: full GPU suite, bench line, FE ncu capture
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r02_test_gpu_final.log 2>&1; echo "pytest gpu rc=$?"; tail -3 gpurun_out/r02_test_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()"; echo "smoke rc=$?"
timeout 900 python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; echo "bench rc=$?"
FE="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-reference-cuda --no-sub-records"
$FE > gpurun_out/ncu_plain_fe.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_fe_launches.csv $FE > gpurun_out/ncu_fe_list.log 2>&1
echo "fe list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:fe_philox -s 3 -c 1 -f -o gpurun_out/r02_prof_fe $FE > gpurun_out/ncu_fe_full.log 2>&1
echo "fe full rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_1gpu.json"))
print("FE", d["ms_per_step"], d["value"], d["roofline"]["frac"], d["roofline"]["mix_bound_frac"], "e2e", d["e2e"]["value"])
print("EM", d["em"]["ms_per_step"], d["em"]["value"])
print("C5", d["c5_strong"]["fe"]["ms_per_step"], d["c5_strong"]["em"]["ms_per_step"])
x = d["reference_cuda"]; print("ref", x["xorwow"]["exec_ms"], x["philox"]["exec_ms"], x["philox"]["ours_native_same_words"])
print(d["dense_mode"]["ms_per_step"], d["xorwow_fast_mode"]["ms_per_step"])
PY
