"""round-2 GPU call 6: default-sweep timing in process, EM tweak A/B on three kinds of points."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

code_sweep = r'''
import sys, json
sys.path.insert(0, %r)
import numpy as np
from nmch_b200 import engine as E
from oracle import oracle as o
k, th, sg = o.exploration_grid(5, True)
for rng, name in ((1, "xorwow_compat"), (5, "xorwow_fast"), (0, "philox")):
    with E.Engine(NTPB=512, NB=10, N=1000, rng=rng) as e:
        e.init(1234)
        ms = [e.explore(k, th, sg)[0].exec_ms for _ in range(4)]
        info = e.launch_info()
    print(json.dumps({"sweep": name, "points": len(k), "paths": 5120, "N": 1000, "ms": [round(x, 3) for x in ms], "grid": [info["grid_x"], info["grid_y"]]}))
''' % ROOT

code_em = r'''
import sys, json, os
sys.path.insert(0, %r)
from nmch_b200 import engine as E
n = 1 << 22
for name, k, th, sg in (("boost d=1.11", 0.5, 0.1, 0.3), ("packed d=5.7", 2.08, 0.108, 0.28), ("packed d=10", 10.0, 0.5, 1.0), ("mixture d=0.449", 2.08, 0.108, 1.0)):
    with E.Engine(NTPB=512, NB=n // 512, N=1000, method=E.METHOD_EM, k=k, theta=th, sigma=sg) as e:
        e.init(1234)
        e.compute()
        r = [e.compute() for _ in range(4)]
    print(json.dumps({"lib": os.path.basename(os.environ.get("NMCH_B200_LIB", "base")), "em": name, "ms": round(min(x.exec_ms for x in r), 3), "E": r[-1].mean, "se": r[-1].std_error}))
''' % ROOT

subprocess.run([sys.executable, "-c", code_sweep])
for lib in (None, os.path.join(ROOT, "nmch_b200", "variants", "libnmch_b200_tweak.so")):
    env = dict(os.environ)
    if lib:
        env["NMCH_B200_LIB"] = lib
    subprocess.run([sys.executable, "-c", code_em], env=env)
