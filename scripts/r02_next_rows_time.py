"""Timings of the SURVEY §8f rows on one B200 at 2^24 paths: strike vector + delta (FE, EM), greeks (FE), QE-M."""
import json
import sys

import numpy as np

sys.path.insert(0, ".")
from nmch_b200 import engine as E  # noqa: E402
from oracle import oracle as o  # noqa: E402  (checker values only, after the timed calls)

out = {"paths": 1 << 24, "strikes": [0.9, 1.0, 1.1]}
n = 1 << 24
K = np.array(out["strikes"], np.float32)


def best(fn, reps=3):
    fn()
    return min(fn() for _ in range(reps))


with E.Engine(NTPB=512, NB=n // 512, N=1000) as e:
    e.init(1234)
    out["fe_compute_ms"] = best(lambda: e.compute().exec_ms)
    out["fe_compute_strikes_ms"] = best(lambda: e.compute_strikes(K)[0]["moments"].exec_ms)
    out["fe_compute_greeks_ms"] = best(lambda: e.compute_greeks(K)[0]["moments"].exec_ms)
    r = e.compute_greeks(K)
    out["fe_atm"] = {"price": r[1]["moments"].mean, "se": r[1]["moments"].std_error, "delta": r[1]["delta"],
                     "vega_v0": r[1]["vega_v0"], "vega_v0_se": r[1]["vega_v0_se"]}
with E.Engine(NTPB=512, NB=(1 << 22) // 512, N=1000, method=E.METHOD_EM) as e:
    e.init(1234)
    out["em_2p22_compute_ms"] = best(lambda: e.compute().exec_ms)
    out["em_2p22_compute_strikes_ms"] = best(lambda: e.compute_strikes(K)[0]["moments"].exec_ms)
for N in (50, 100):
    with E.Engine(NTPB=512, NB=n // 512, N=N, method=E.METHOD_QE) as e:
        e.init(1234)
        ms = best(lambda: e.compute().exec_ms)
        m = e.compute()
        out[f"qe_N{N}"] = {"ms": ms, "paths_per_s": n / (ms * 1e-3), "price": m.mean, "se": m.std_error}
h = 1e-4
out["semi_analytic"] = {"price": o.heston_call(), "delta": (o.heston_call(S0=1 + h) - o.heston_call(S0=1 - h)) / (2 * h),
                        "vega_v0": (o.heston_call(v0=0.1 + h) - o.heston_call(v0=0.1 - h)) / (2 * h)}
print(json.dumps(out))
