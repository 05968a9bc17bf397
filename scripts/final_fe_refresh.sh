set -x
python bench.py > gpurun_out/bench_fe.json 2> gpurun_out/bench_fe.err; tail -c 200 gpurun_out/bench_fe.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-reference-cuda > gpurun_out/ncu_launches.log 2>&1
cat > /tmp/xf.py <<'PY'
import sys
sys.path.insert(0, ".")
from nmch_b200 import engine as E
n = 1 << 24
with E.Engine(NTPB=512, NB=n // 512, N=1000, rng=E.RNG_XORWOW_FAST) as e:
    e.init(1234)
    for _ in range(2):
        m = e.compute()
    print(m.exec_ms, m.mean, e.init_ms)
PY
python /tmp/xf.py
ncu --set full --clock-control none --import-source on -k regex:fe_xorwow_fast -c 1 -f -o gpurun_out/prof_xfast python /tmp/xf.py > gpurun_out/ncu_xfast.log 2>&1; tail -2 gpurun_out/ncu_xfast.log
python -m pytest tests -m gpu -q 2>&1 | tail -3
