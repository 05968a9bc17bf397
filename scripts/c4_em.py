"""BASELINE configs[3], EM leg only (7778-point grid x 2^20 paths, one GPU): the whole sweep as ONE launch, each sampler
kind's points on their own (what the one launch is made of), and a few prices against the semi-analytic pricer."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from nmch_b200 import engine as E
from oracle import oracle as o

P = 20
ks = (0.1 + np.arange(P) * 9.9 / (P - 1)).astype(np.float32)
ths = (0.01 + np.arange(P) * 0.49 / (P - 1)).astype(np.float32)
sgs = (0.1 + np.arange(P) * 0.9 / (P - 1)).astype(np.float32)
pts = [(k, t, s) for s in sgs for t in ths for k in ks if not (20 * k * t < s * s)]
k, th, sg = (np.array(x, np.float32) for x in zip(*pts))
n = 1 << 20
d = 2.0 * k.astype(np.float64) * th / (sg.astype(np.float64) ** 2)
kinds = {"mixture (d <= 1/2)": d - 0.5 <= 1e-3, "split, boosted gamma (1/2 < d < 3/2)": (d - 0.5 > 1e-3) & (d - 0.5 < 1.0),
         "split, no boost (d >= 3/2)": d - 0.5 >= 1.0}
out = {"points": len(k), "paths_per_point": n, "N": 1000}
with E.Engine(NTPB=512, NB=n // 512, N=1000, method=E.METHOD_EM) as e:
    e.init(1234)
    e.explore(k[:8], th[:8], sg[:8])                       # warm-up
    l0 = e.launch_info()["kernel_launches"]
    t0 = time.perf_counter()
    res = e.explore(k, th, sg)
    out["one_launch"] = {"launch_ms": res[0].exec_ms, "wall_s": time.perf_counter() - t0,
                         "kernel_launches": e.launch_info()["kernel_launches"] - l0, "grid": e.launch_info()["grid_y"]}
    out["per_kind"] = {}
    for name, m in kinds.items():
        if m.sum():
            r = e.explore(k[m], th[m], sg[m])
            out["per_kind"][name] = {"points": int(m.sum()), "launch_ms": r[0].exec_ms,
                                     "ms_per_point": r[0].exec_ms / int(m.sum())}
out["sum_of_kinds_ms"] = sum(v["launch_ms"] for v in out["per_kind"].values())
zs = []
for i in np.linspace(0, len(k) - 1, 40).astype(int):
    want = o.heston_call(kappa=float(k[i]), theta=float(th[i]), sigma=float(sg[i]))
    zs.append((res[i].mean - want) / res[i].std_error)
out["z_vs_semi_analytic_40_points"] = {"mean": float(np.mean(zs)), "rms": float(np.sqrt(np.mean(np.square(zs)))), "max_abs": float(np.max(np.abs(zs)))}
print(json.dumps(out))
