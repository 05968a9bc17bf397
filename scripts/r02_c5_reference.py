"""BASELINE configs[4] size (2^30 paths, N = 1000) on ONE B200: the reference's own CUDA build against this engine on the
same seed.  The reference keeps one cuRAND state per path (48 B XORWOW / 64 B Philox: 51.5 / 68.7 GB -- it fits in 180 GB)
and sums per-thread payoff/n in FP32 through 2^21 float atomics; ours sums raw payoffs in FP64."""
import json
import os
import subprocess
import sys

sys.path.insert(0, ".")
from nmch_b200 import engine as E  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "nmch_ref_harness")
n, N = 1 << 30, 1000
out = {"paths": n, "N": N}


def ref(rng):
    r = subprocess.run([EXE, "--method", "fe", "--rng", rng, "--kernel", "k3", "--NTPB", "512", "--NB", str(n // 512), "--N", str(N),
                        "--repeat", "2"], capture_output=True, text=True, timeout=1200)
    rows = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and rows and rows[0]["cuda"] == "cudaSuccess", (r.returncode, r.stderr[-300:], rows[:1])
    return rows


def ours(mode):
    with E.Engine(NTPB=512, NB=n // 512, N=N, rng=mode) as e:
        e.init(1234)
        ms = [e.compute() for _ in range(2)]
        return [{"E": m.mean, "var": m.variance, "exec_ms": m.exec_ms, "std_error": m.std_error} for m in ms], e.init_ms


for rng, modes in (("xorwow", (("xorwow_compat", E.RNG_XORWOW_COMPAT), ("xorwow_fast", E.RNG_XORWOW_FAST))),
                   ("philox", (("philox_native", E.RNG_PHILOX),))):
    rows = ref(rng)
    rec = {"reference": [{"E": r["E"], "var": r["E2"] - r["E"] ** 2, "exec_ms": r["exec_ms"], "init_ms": r["init_ms"]} for r in rows]}
    for tag, mode in modes:
        mine, init_ms = ours(mode)
        rec[tag] = {"calls": mine, "init_ms": init_ms,
                    "max_rel_diff_E": max(abs(a["E"] - b["E"]) / b["E"] for a, b in zip(mine, rows)),
                    "max_rel_diff_var": max(abs(a["var"] - (b["E2"] - b["E"] ** 2)) / (b["E2"] - b["E"] ** 2) for a, b in zip(mine, rows))}
    out[rng] = rec
c, f = out["xorwow"]["xorwow_compat"]["calls"], out["xorwow"]["xorwow_fast"]["calls"]
out["xorwow"]["fast_vs_compat"] = {"max_rel_diff_E": max(abs(a["E"] - b["E"]) / b["E"] for a, b in zip(f, c)),
                                   "max_rel_diff_var": max(abs(a["var"] - b["var"]) / b["var"] for a, b in zip(f, c))}
print(json.dumps(out))
