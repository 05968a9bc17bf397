"""A/B of EM tuning variants on ONE box: alternate the libraries, three rounds, best exec_ms of 4 calls each."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, json, os
sys.path.insert(0, %r)
from nmch_b200 import engine as E
n = 1 << 22
out = {}
for name, k, th, sg in (("boost", 0.5, 0.1, 0.3), ("packed", 2.08, 0.108, 0.28), ("mixture", 2.08, 0.108, 1.0)):
    with E.Engine(NTPB=512, NB=n // 512, N=1000, method=E.METHOD_EM, k=k, theta=th, sigma=sg) as e:
        e.init(1234)
        e.compute()
        out[name] = round(min(e.compute().exec_ms for _ in range(4)), 3)
print(json.dumps({"lib": os.path.basename(os.environ.get("NMCH_B200_LIB", "base")), **out}))
''' % ROOT
libs = [None] + [os.path.join(ROOT, "nmch_b200", "variants", f"libnmch_b200_{t}.so") for t in sys.argv[1:]]
for rnd in range(3):
    for lib in libs:
        env = dict(os.environ)
        if lib:
            env["NMCH_B200_LIB"] = lib
        subprocess.run([sys.executable, "-c", code], env=env)
