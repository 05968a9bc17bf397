"""BASELINE configs[3], EM leg only (7778-point grid x 2^20 paths, one GPU): time and a few prices against the pricer."""
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
with E.Engine(NTPB=512, NB=n // 512, N=1000, method=E.METHOD_EM) as e:
    e.init(1234)
    t0 = time.perf_counter()
    res = e.explore(k, th, sg)
    wall = time.perf_counter() - t0
d = 2 * k * th / (sg * sg)
print("points", len(k), "mixture", int((d - 0.5 <= 1e-3).sum()), "split+boost", int(((d - 0.5 > 1e-3) & (d < 1.5)).sum()),
      "split packed", int((d >= 1.5).sum()), "launch_ms", res[0].exec_ms, "wall_s", wall)
zs = []
for i in np.linspace(0, len(k) - 1, 40).astype(int):
    want = o.heston_call(kappa=float(k[i]), theta=float(th[i]), sigma=float(sg[i]))
    zs.append((res[i].mean - want) / res[i].std_error)
print("z vs semi-analytic at 40 points: mean %.2f rms %.2f max|z| %.2f" % (np.mean(zs), np.sqrt(np.mean(np.square(zs))), np.max(np.abs(zs))))
