#!/bin/bash
# round-2 profiles: launch lists of the bench commands and one --set full capture of each dominant kernel.
# Every command runs plain first (must exit 0) and only then under ncu; numbers printed under ncu are not bench values.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
FE="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-reference-cuda --no-sub-records"
EM="python bench.py --method em --steps 3 --warmup 3 --no-cpu-baseline --no-reference-cuda --no-sub-records"
$FE > gpurun_out/ncu_plain_fe.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_fe_launches.csv $FE > gpurun_out/ncu_fe_list.log 2>&1
echo "fe list rc=$?"
$EM > gpurun_out/ncu_plain_em.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_em_launches.csv $EM > gpurun_out/ncu_em_list.log 2>&1
echo "em list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:fe_philox -s 3 -c 1 -f -o gpurun_out/r02_prof_fe $FE > gpurun_out/ncu_fe_full.log 2>&1
echo "fe full rc=$?"
ncu --set full --clock-control none --import-source on -k regex:em_native -s 3 -c 1 -f -o gpurun_out/r02_prof_em $EM > gpurun_out/ncu_em_full.log 2>&1
echo "em full rc=$?"
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_*launches.csv
