#!/bin/bash
# round-2 closing rehearsal of what the driver runs on one GPU: full GPU suite, smoke, reference arm, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
if [ "$1" != bench ]; then
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02_test_gpu_final.log 2>&1; echo "pytest gpu rc=$?"; tail -4 gpurun_out/r02_test_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()"; echo "smoke rc=$?"
fi
t0=$SECONDS; timeout 900 python bench.py --impl reference > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err; echo "ref arm rc=$? wall $((SECONDS-t0)) s"
t0=$SECONDS; timeout 900 python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; echo "bench rc=$? wall $((SECONDS-t0)) s"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_1gpu.json").read().strip().splitlines()[-1])
print("FE", d["ms_per_step"], d["value"], d["roofline"]["frac"], "e2e", d["e2e"]["value"], d["clocks"])
print("EM", d["em"]["ms_per_step"], d["em"]["value"])
print("C5", d["c5_strong"]["fe"]["ms_per_step"], d["c5_strong"]["em"]["ms_per_step"])
print("C4", d["c4_sweep"]["fe"]["launch_ms"], d["c4_sweep"]["em"]["launch_ms"])
r = json.loads(open("gpurun_out/r02_bench_reference_arm.json").read().strip().splitlines()[-1]); print("ref arm", r["value"], r["cpu_baseline"]["cores"], r["ms_per_step"])
PY
