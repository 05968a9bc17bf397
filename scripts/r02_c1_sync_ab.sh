#!/bin/bash
# same-box A/B at BASELINE configs[0] size (2^18 paths): block-wide barrier every k loop iterations in the one-wave kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
C1="python bench.py --log2-paths 18 --steps 50 --warmup 5 --no-cpu-baseline --no-reference-cuda --no-sub-records"
out=gpurun_out/r02_c1_sync_ab.txt
echo "# bench.py --log2-paths 18 --steps 50: lib, P, T, ms per step, roofline fraction, E[X]" > $out
for lib in base sync8 sync32 sync128; do
  for cfg in "1 256" "1 128" "2 128" "2 256"; do
    set -- $cfg
    if [ $lib = base ]; then unset NMCH_B200_LIB; else export NMCH_B200_LIB=$PWD/nmch_b200/variants/libnmch_b200_$lib.so; fi
    $C1 --paths-per-thread $1 --block-threads $2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib P=$1 T=$2', round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['result']['E[X]'])" >> $out
  done
done
cat $out
