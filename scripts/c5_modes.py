"""BASELINE configs[4] size (2^30 paths x 1000 steps, one GPU) for the opt-in FE streams: 64-bit indexing at full size."""
import sys
sys.path.insert(0, ".")
from nmch_b200 import engine as E
from oracle import oracle as o

n = 1 << 30
for name, mode in (("philox_dense", E.RNG_PHILOX_DENSE), ("xorwow_fast", E.RNG_XORWOW_FAST)):
    with E.Engine(NTPB=512, NB=1, n_paths=n, N=1000, rng=mode) as e:
        e.init(1234)
        m = e.compute()
        print(name, "init_ms", round(e.init_ms, 1), "exec_ms", round(m.exec_ms, 1), "E", m.mean, "SE", m.std_error,
              "z_vs_heston", (m.mean - o.heston_call()) / m.std_error, "path_steps_per_s", n * 1000 / (m.exec_ms * 1e-3), flush=True)
