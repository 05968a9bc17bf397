"""Time the native FE kernel for several tuning builds of the library (NMCH_B200_LIB override)."""
import glob, json, os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, os, json
sys.path.insert(0, %r)
from nmch_b200 import engine as E
n = 1 << 24
for P, bt in ((4, 128), (4, 256)):
    with E.Engine(NTPB=512, NB=n // 512, N=1000, rng=0, paths_per_thread=P, block_threads=bt) as e:
        e.init(1234)
        ms = min(e.compute().exec_ms for _ in range(5))
        info = e.launch_info()
    print(json.dumps({"lib": os.path.basename(os.environ["NMCH_B200_LIB"]), "P": P, "threads": bt, "regs": info["regs_per_thread"],
                      "ms": round(ms, 3), "path_steps_per_s": n * 1000 / (ms * 1e-3)}))
''' % root
for lib in sorted(glob.glob(os.path.join(root, "nmch_b200", "build", "variants", "*.so"))):
    subprocess.run([sys.executable, "-c", code], env=dict(os.environ, NMCH_B200_LIB=lib))
