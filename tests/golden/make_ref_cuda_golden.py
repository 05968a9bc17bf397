#!/usr/bin/env python
"""Regenerates tests/golden/ref_cuda_b200.json: outputs of the UNMODIFIED reference CUDA build
(oracle/_ref/nmch_ref_harness, built by `make -C oracle ref` from /root/reference) run on a B200.

Run on the GPU box:   python tests/golden/make_ref_cuda_golden.py
                      python tests/golden/make_ref_cuda_golden.py --plus     -> ref_cuda_b200_plus.json: the FE cases on
                      oracle/_ref/nmch_ref_harness_plus, the same build with the floor token `Vt = abs(Vt);` changed to
                      `Vt = fmaxf(Vt, 0.0f);` while compiling (oracle/Makefile) -- the (.)+ floor of BASELINE configs[0]/[1]
Each entry records the harness flags and, per compute() call, the float E[X], E[X^2] and err the
reference's getters returned.  The reference accumulates with float atomics (order-dependent), so its
own run-to-run spread is recorded as well (3 repeats of every case).
"""
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
EXE = os.path.join(ROOT, "oracle", "_ref", "nmch_ref_harness")
EXE_PLUS = EXE + "_plus"

CASES = [
    dict(method="fe", rng="xorwow", kernel="k3", NTPB=512, NB=512, N=1000, repeat=3),
    dict(method="fe", rng="philox", kernel="k3", NTPB=512, NB=512, N=1000, repeat=3),
    dict(method="fe", rng="xorwow", kernel="k2", NTPB=256, NB=100, N=250, repeat=2),
    dict(method="fe", rng="xorwow", kernel="k3", NTPB=32, NB=4, N=100, repeat=2),
    dict(method="fe", rng="xorwow", kernel="k3", NTPB=512, NB=512, N=1000, k=2.08, theta=0.108, sigma=1.0, repeat=2),
    dict(method="fe", rng="xorwow", kernel="k3", NTPB=512, NB=64, N=365, T=0.5, S_0=2.0, v_0=0.04, r=0.03, k=1.5, rho=0.3,
         theta=0.09, sigma=0.5, repeat=2),
    dict(method="fe", rng="mrg", kernel="k3", NTPB=512, NB=512, N=1000, repeat=2),
    dict(method="fe", rng="mrg", kernel="k3", NTPB=128, NB=64, N=333, k=2.08, theta=0.108, sigma=1.0, repeat=2),
    dict(method="em", rng="mrg", kernel="k3", NTPB=512, NB=64, N=500, repeat=2),
    dict(method="em", rng="xorwow", kernel="k3", NTPB=512, NB=64, N=1000, repeat=2),
    dict(method="em", rng="philox", kernel="k3", NTPB=512, NB=64, N=1000, repeat=2),
    dict(method="em", rng="xorwow", kernel="k3", NTPB=32, NB=4, N=100, repeat=2),
    dict(method="em", rng="xorwow", kernel="k3", NTPB=512, NB=32, N=500, k=2.08, theta=0.108, sigma=1.0, repeat=2),
    dict(method="em", rng="xorwow", kernel="k3", NTPB=512, NB=32, N=200, k=10.0, theta=0.5, sigma=1.0, repeat=2),
]
# (.)+ floor (--plus): FE only -- the floor does not exist in EM.  The Feller-violating point is where the floors differ.
PLUS_CASES = [
    dict(method="fe", rng="xorwow", kernel="k3", NTPB=512, NB=512, N=1000, repeat=3),
    dict(method="fe", rng="philox", kernel="k3", NTPB=512, NB=512, N=1000, repeat=3),
    dict(method="fe", rng="xorwow", kernel="k3", NTPB=512, NB=512, N=1000, k=2.08, theta=0.108, sigma=1.0, repeat=2),
    dict(method="fe", rng="philox", kernel="k3", NTPB=512, NB=512, N=1000, k=2.08, theta=0.108, sigma=1.0, repeat=2),
    dict(method="fe", rng="mrg", kernel="k3", NTPB=128, NB=64, N=333, k=2.08, theta=0.108, sigma=1.0, repeat=2),
    dict(method="fe", rng="xorwow", kernel="k2", NTPB=256, NB=100, N=250, k=6.04, theta=0.01, sigma=0.82, repeat=2),
    dict(method="fe", rng="xorwow", kernel="k3", NTPB=512, NB=64, N=365, T=0.5, S_0=2.0, v_0=0.04, r=0.03, k=1.5, rho=0.3,
         theta=0.09, sigma=0.9, repeat=2),
]
# the reference's exploration sweep (exploration.cu): warm-up compute + the first points, continued streams
SWEEP = [(0.5, 0.1, 0.3), (0.1, 0.01, 0.1), (2.08, 0.01, 0.1), (4.06, 0.01, 0.1), (6.04, 0.108, 0.28), (9.999999, 0.5, 1.0)]


def run(flags):
    cmd = [EXE_PLUS if "--plus" in sys.argv else EXE]
    for k, v in flags.items():
        cmd += [f"--{k}", str(v)]
    out = subprocess.run(cmd, check=True, capture_output=True, text=True, timeout=1800).stdout
    return [json.loads(l) for l in out.splitlines() if l.startswith("{")]


def main():
    plus = "--plus" in sys.argv
    name = "ref_cuda_b200_plus.json" if plus else "ref_cuda_b200.json"
    gold = {"generator": ("oracle/_ref/nmch_ref_harness_plus (reference sources with the FE floor token `Vt = abs(Vt);` changed to "
                          "`Vt = fmaxf(Vt, 0.0f);` while compiling, nvcc 12.9 -O3 -arch=sm_100) on NVIDIA B200" if plus else
                          "oracle/_ref/nmch_ref_harness (reference sources, nvcc 12.9 -O3 -arch=sm_100) on NVIDIA B200"),
            "cases": [], "sweeps": []}
    for c in (PLUS_CASES if plus else CASES):
        runs = [run(c) for _ in range(3)]
        calls = []
        for i in range(len(runs[0])):
            Es = [r[i]["E"] for r in runs]
            E2s = [r[i]["E2"] for r in runs]
            calls.append({"E": Es[0], "E2": E2s[0], "err": runs[0][i]["err"], "E_spread": max(Es) - min(Es),
                          "E2_spread": max(E2s) - min(E2s)})
        gold["cases"].append({"flags": c, "calls": calls})
    with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as f:
        for k, t, s in SWEEP:
            f.write(f"{k:.9g} {t:.9g} {s:.9g}\n")
        pts = f.name
    for method in (("fe",) if plus else ("fe", "em")):
        flags = dict(method=method, rng="xorwow", kernel="k3", NTPB=512, NB=10, N=1000, points=pts)
        rows = run(flags)
        gold["sweeps"].append({"flags": {k: v for k, v in flags.items() if k != "points"}, "points": SWEEP,
                               "calls": [{"E": r["E"], "E2": r["E2"], "err": r["err"]} for r in rows]})
    with open(os.path.join(HERE, name), "w") as f:
        json.dump(gold, f, indent=1)
    print("wrote", os.path.join(HERE, name))


if __name__ == "__main__":
    main()
