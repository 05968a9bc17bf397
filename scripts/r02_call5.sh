#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_gpu_fe.py tests/test_gpu_xorwow_fast.py tests/test_ref_cuda_golden.py tests/test_gpu_k1_legacy.py tests/test_gpu_checked_build.py tests/test_cli.py tests/test_bench_contract.py -m gpu -q > gpurun_out/r02_test_gpu5.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/r02_test_gpu5.log
# the reference's own sweep (exploration.cu:24-25: 5120 paths x 200 points x 1000 steps), default CLI run
for rng in xorwow xorwow-fast philox; do
  for i in 1 2; do ./bin/exploration --rng $rng > gpurun_out/explore_default_$rng.csv 2> gpurun_out/explore_default_$rng.err; done
  python - "$rng" <<'PY'
import sys, collections
rng = sys.argv[1]
tot = collections.Counter(); cnt = collections.Counter()
for line in open(f"gpurun_out/explore_default_{rng}.csv").read().splitlines()[1:]:
    f = [x.strip() for x in line.split(",")]
    if len(f) >= 6:
        tot[f[0]] += float(f[4]); cnt[f[0]] += 1
print("explore", rng, {m: (cnt[m], round(tot[m], 3)) for m in tot}, "(points, total ms)")
PY
done
./oracle/_ref/nmch_ref_harness --method fe --rng xorwow --kernel k3 --NTPB 512 --NB 10 --N 1000 --repeat 3 | tail -1
