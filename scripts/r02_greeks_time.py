"""Timing of the tangent pass (compute_greeks) next to compute() and compute_strikes() at BASELINE configs[1] size."""
import json
import sys

import numpy as np

sys.path.insert(0, ".")
from nmch_b200 import engine as E  # noqa: E402

out = {}
n = 1 << 24
K = np.array([0.9, 1.0, 1.1], np.float32)
with E.Engine(NTPB=512, NB=n // 512, N=1000) as e:
    e.init(1234)
    e.compute()
    out["compute_ms"] = min(e.compute().exec_ms for _ in range(3))
    e.compute_strikes(K)
    out["compute_strikes_ms"] = min(e.compute_strikes(K)[0]["moments"].exec_ms for _ in range(3))
    e.compute_greeks(K)
    rows = [e.compute_greeks(K) for _ in range(3)]
    out["compute_greeks_ms"] = min(r[0]["moments"].exec_ms for r in rows)
    out["launch"] = e.launch_info()
    out["atm"] = {k: rows[-1][1][k] for k in ("delta", "itm", "vega_v0", "vega_v0_se")}
    out["atm"]["price"] = rows[-1][1]["moments"].mean
print(json.dumps(out))
