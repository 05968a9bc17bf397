# Round-end evidence on ONE B200: tests, bench lines, launch list, ncu captures, BASELINE configs.  Writes gpurun_out/.
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/final_pytest_gpu.txt; cat gpurun_out/final_pytest_gpu.txt
python bench.py > gpurun_out/bench_fe.json 2> gpurun_out/bench_fe.err; tail -c 300 gpurun_out/bench_fe.json
python bench.py --method em > gpurun_out/bench_em.json 2> gpurun_out/bench_em.err; tail -c 300 gpurun_out/bench_em.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-reference-cuda > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fe_philox -c 1 -f -o gpurun_out/prof_fe3 \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-reference-cuda > gpurun_out/ncu_fe3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fe_dense -c 1 -f -o gpurun_out/prof_dense \
    python bench.py --rng dense --steps 1 --warmup 1 --no-cpu-baseline --no-reference-cuda > gpurun_out/ncu_dense.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:em_native -c 1 -f -o gpurun_out/prof_em3 \
    python bench.py --method em --steps 1 --warmup 1 --no-cpu-baseline --no-reference-cuda > gpurun_out/ncu_em3.log 2>&1
python tests/run_configs.py --out gpurun_out/configs_1gpu.json > gpurun_out/configs_1gpu.log 2>&1; tail -5 gpurun_out/configs_1gpu.log
