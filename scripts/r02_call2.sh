#!/bin/bash
# round-2 GPU call: microbench v2, EM tests on the new sampler, EM / FE tuning variants
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( cd profiles/microbench && for w in 4 8; do ./pipe_rates2 $w > ../../gpurun_out/r02_pipe_rates2_w$w.txt 2>&1; done )
timeout 600 python -m pytest tests/test_gpu_em.py -x -q > gpurun_out/r02_test_em.log 2>&1; echo "pytest em rc=$?"
tail -5 gpurun_out/r02_test_em.log
one() {  # tag, lib, method
  if [ -n "$2" ]; then export NMCH_B200_LIB=$PWD/nmch_b200/variants/$2; else unset NMCH_B200_LIB; fi
  python bench.py --method $3 --steps 10 --warmup 3 --no-cpu-baseline --no-reference-cuda 2>gpurun_out/r02_v_$1.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],3), '%.4g' % d['value'], d['kernel'], d['result']['E[X]'], d['result']['std_error'], d['clocks']['sm_mhz'])"
  unset NMCH_B200_LIB
}
one em_base "" em
one em_minb5 libnmch_b200_minb5.so em
one em_minb4 libnmch_b200_minb4.so em
one fe_base "" fe
one fe_i2fp libnmch_b200_i2fp.so fe
