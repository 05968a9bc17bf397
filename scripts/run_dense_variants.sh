for v in "" nopin minb8; do
  if [ -z "$v" ]; then unset NMCH_B200_LIB; else export NMCH_B200_LIB=$PWD/nmch_b200/variants/libnmch_b200_$v.so; fi
  for P in 2 4; do
  python bench.py --rng dense --paths-per-thread $P --steps 5 --warmup 3 --no-cpu-baseline --no-reference-cuda 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('dense', '$v', d['ms_per_step'], d['value'], d['kernel']['paths_per_thread'], d['kernel']['regs_per_thread'], d['result']['E[X]'])"
  done
done
unset NMCH_B200_LIB
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-reference-cuda 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('native', d['ms_per_step'], d['value'], d['kernel'], d.get('dense_mode'))"
