"""A/B of FE tuning variants on ONE box: block size and resident-warp target of the native kernel."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, json, os
sys.path.insert(0, %r)
from nmch_b200 import engine as E
n = 1 << 24
out = {}
for P, bt in ((4, 128), (4, 256), (2, 128), (8, 128)):
    with E.Engine(NTPB=512, NB=n // 512, N=1000, paths_per_thread=P, block_threads=bt) as e:
        e.init(1234)
        e.compute()
        out[f"P{P}_T{bt}"] = round(min(e.compute().exec_ms for _ in range(4)), 3)
        out[f"regs_P{P}_T{bt}"] = e.launch_info()["regs_per_thread"]
print(json.dumps({"lib": os.path.basename(os.environ.get("NMCH_B200_LIB", "base")), **out}))
''' % ROOT
libs = [None] + [os.path.join(ROOT, "nmch_b200", "variants", f"libnmch_b200_{t}.so") for t in sys.argv[1:]]
for rnd in range(2):
    for lib in libs:
        env = dict(os.environ)
        if lib:
            env["NMCH_B200_LIB"] = lib
        subprocess.run([sys.executable, "-c", code], env=env)
