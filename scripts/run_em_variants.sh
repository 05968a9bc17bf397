for v in "" minb8 minb5 t128b12 t128b16; do
  if [ -z "$v" ]; then unset NMCH_B200_LIB; else export NMCH_B200_LIB=$PWD/nmch_b200/variants/libnmch_b200_$v.so; fi
  python bench.py --method em --steps 5 --warmup 3 --no-cpu-baseline --no-reference-cuda 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$v', d['ms_per_step'], d['value'], d['kernel'], d['result']['E[X]'])"
done
unset NMCH_B200_LIB
ncu --set full --clock-control none --import-source on -k regex:em_native -c 1 -o gpurun_out/prof_em3 python bench.py --method em --steps 1 --warmup 1 --no-cpu-baseline --no-reference-cuda > gpurun_out/ncu_em3.log 2>&1; tail -2 gpurun_out/ncu_em3.log
