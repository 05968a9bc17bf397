import sys
sys.path.insert(0, ".")
from nmch_b200 import engine as E
n = 1 << 24
for rep in range(2):
    for mode in (E.RNG_XORWOW_FAST, E.RNG_XORWOW_COMPAT):
        with E.Engine(NTPB=512, NB=n // 512, N=1000, rng=mode) as e:
            e.init(1234)
            m = e.compute()
            print(rep, mode, "init_ms", e.init_ms, "exec_ms", m.exec_ms, m.mean)
