"""Tiny invocation of every kernel (for compute-sanitizer memcheck / racecheck runs)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nmch_b200 import engine as E

k = np.array([0.5, 2.08, 0.1], np.float32)
th = np.array([0.1, 0.108, 0.5], np.float32)
sg = np.array([0.3, 1.0, 1.0], np.float32)
for method in (E.METHOD_FE, E.METHOD_EM):
    for rng in (E.RNG_PHILOX, E.RNG_XORWOW_COMPAT, E.RNG_PHILOX_COMPAT):
        for P in ((1, 2, 4, 8) if (method == E.METHOD_FE and rng == E.RNG_PHILOX) else (0,)):
            for floor in (0, 1):
                with E.Engine(NTPB=1, NB=1, n_paths=1500, N=15, method=method, rng=rng, floor=floor, paths_per_thread=P) as e:
                    e.init(7)
                    a = e.compute()
                    b = e.explore(k, th, sg)
                    S, V, c = e.compute_paths()
                    assert np.isfinite(S).all() and a.n_paths == 1500
print("sanitize_small ok")
