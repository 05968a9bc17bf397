"""Sweep the native FE kernel's launch shape on one GPU (paths per thread x block size)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nmch_b200 import engine as E

log2 = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n = 1 << log2
for floor in (0, 1):
    for P in (1, 2, 4, 8):
        for bt in (128, 256):
            with E.Engine(NTPB=512, NB=n // 512, N=1000, rng=0, floor=floor, paths_per_thread=P, block_threads=bt) as e:
                e.init(1234)
                ms = min(e.compute().exec_ms for _ in range(4))
                info = e.launch_info()
            print(json.dumps({"floor": floor, "P": P, "threads": bt, "regs": info["regs_per_thread"], "ms": round(ms, 3),
                              "path_steps_per_s": n * 1000 / (ms * 1e-3)}))
