import sys
sys.path.insert(0, "/root/repo")
import numpy as np
from nmch_b200 import engine as E
from oracle import oracle as o
for N in (100, 1000):
    n = 1 << 14
    with E.Engine(NTPB=512, NB=n // 512, N=N, rng=E.RNG_PHILOX) as e:
        e.init(1234)
        S, V, m = e.compute_paths()
    ref = o.fe_run(o.Params(N=N), rng=o.RNG_PHILOX, n_paths=n, want_paths=True)
    rs = np.abs(S - ref["S"]) / ref["S"]
    rv = np.abs(V - ref["V"]) / np.maximum(ref["V"], 1e-3)
    print(N, "S rel: median %.2e p99 %.2e max %.2e | V rel: median %.2e p99 %.2e max %.2e" % (np.median(rs), np.quantile(rs, 0.99), rs.max(), np.median(rv), np.quantile(rv, 0.99), rv.max()))
