#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( cd profiles/microbench && ./pipe_rates2 8 > ../../gpurun_out/r02_pipe_rates2_w8.txt 2>&1 )
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r02_test_gpu.log 2>&1; echo "pytest gpu rc=$?"
tail -25 gpurun_out/r02_test_gpu.log
for cfg in "1 128" "1 256" "2 128" "2 256" "4 128" "4 256"; do
  set -- $cfg
  python bench.py --log2-paths 18 --steps 50 --paths-per-thread $1 --block-threads $2 --no-sub-records --no-cpu-baseline --no-reference-cuda 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('2^18 P=$1 T=$2', round(d['ms_per_step'],4), '%.4g' % d['value'], round(d['roofline']['frac'],3), d['kernel']['grid_x'], d['kernel']['regs_per_thread'])"
done
