set -x
ncu --set full --clock-control none --import-source on -k regex:fe_philox -c 1 -f -o gpurun_out/prof_fe3 \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-reference-cuda > gpurun_out/ncu_fe3.log 2>&1
python -m pytest tests -m gpu -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()"
