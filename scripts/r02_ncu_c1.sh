#!/bin/bash
# ncu --set full of the native FE kernel at BASELINE configs[0] size (2^18 paths): where does the 0.65 come from?
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
C1="python bench.py --log2-paths 18 --steps 20 --warmup 3 --no-cpu-baseline --no-reference-cuda --no-sub-records"
$C1 > gpurun_out/r02_c1_plain.json 2> gpurun_out/r02_c1_plain.err; echo "plain rc=$?"
for cfg in "1 256" "1 128" "2 128"; do
  set -- $cfg
  $C1 --paths-per-thread $1 --block-threads $2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('P=$1 T=$2', d['ms_per_step'], d['roofline']['frac'], d['kernel'])"
done
ncu --set full --clock-control none --import-source on -k regex:fe_philox -s 3 -c 1 -f -o gpurun_out/r02_prof_fe_c1 $C1 > gpurun_out/ncu_fe_c1.log 2>&1
echo "ncu rc=$?"
true
python scripts/ncu_select.py gpurun_out/r02_prof_fe_c1.ncu-rep > gpurun_out/r02_fe_c1_ncu_raw_selected.csv 2>/dev/null
ls -la gpurun_out/r02_prof_fe_c1.ncu-rep
