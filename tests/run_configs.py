"""Runs the BASELINE.json configurations on the visible GPU(s) and writes one JSON report.

  C1  FE (.)+ floor, README params, N=1000, 2^18 paths                      (+ the CPU oracle's time for it)
  C2  FE |.| and (.)+ floors, N=1000, 2^24 paths, 1 GPU  vs the reference CUDA build
  C3  EM, N=1000, 2^22 paths                              vs the reference CUDA build
  C4  exploration grid 20^3 over (k, theta, sigma), 2^20 paths per point, ONE launch per method
  C5  FE + EM at 2^30 paths, N=1000, on all visible GPUs (single-process group, one NCCL allreduce)

usage: python tests/run_configs.py [--out profiles/configs_r01.json] [--skip c4,c5] [--c4-log2-paths 20]
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from nmch_b200 import engine as E  # noqa: E402
from oracle import oracle as o  # noqa: E402

HESTON = 0.1197325094


def stats(m):
    return {"E": m.mean, "E2": m.mean_sq, "std_error": m.std_error, "exec_ms": m.exec_ms,
            "z_vs_heston": (m.mean - HESTON) / m.std_error if m.std_error > 0 else None}


def best_of(e, reps=3):
    ms = [e.compute() for _ in range(reps)]
    return min(ms, key=lambda m: m.exec_ms)


def ref_cuda(method, rng, nb, N=1000, repeat=3):
    exe = o.REF_HARNESS_PATH
    if not os.path.exists(exe):
        return None
    r = subprocess.run([exe, "--method", method, "--rng", rng, "--kernel", "k3", "--NTPB", "512", "--NB", str(nb), "--N",
                        str(N), "--repeat", str(repeat + 1)], capture_output=True, text=True, timeout=1800)
    rows = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
    if not rows:
        return {"error": r.stderr[-300:]}
    first = rows[0]
    best = min(rows[1:], key=lambda x: x["exec_ms"])
    return {"first_call": {k: first[k] for k in ("E", "E2", "err", "exec_ms", "init_ms")}, "best_exec_ms": best["exec_ms"]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "configs.json"))
    ap.add_argument("--skip", default="")
    ap.add_argument("--c4-log2-paths", type=int, default=20)
    ap.add_argument("--c4-points", type=int, default=20)
    ap.add_argument("--c5-log2-paths", type=int, default=30)
    args = ap.parse_args()
    skip = set(args.skip.split(",")) if args.skip else set()
    import torch
    ngpu = torch.cuda.device_count()
    rep = {"gpus_visible": ngpu, "gpu": torch.cuda.get_device_name(0)}

    if "c1" not in skip:
        n = 1 << 18
        with E.Engine(NTPB=512, NB=512, N=1000, floor=E.FLOOR_PLUS) as e:
            e.init(1234)
            e.compute()
            m = best_of(e)
        t0 = time.perf_counter()
        ref = o.fe_run(o.Params(), rng=o.RNG_XORWOW, floor=o.FLOOR_PLUS, n_paths=n)
        cpu_s = time.perf_counter() - t0
        rep["c1"] = {"ours": stats(m), "path_steps_per_s": n * 1000 / (m.exec_ms * 1e-3),
                     "cpu_oracle": {"E": ref["mean"], "seconds": cpu_s, "threads": o.max_threads(),
                                    "path_steps_per_s": n * 1000 / cpu_s}}
        print("c1", json.dumps(rep["c1"]), flush=True)

    if "c2" not in skip:
        n = 1 << 24
        out = {}
        for name, floor in (("abs", E.FLOOR_ABS), ("plus", E.FLOOR_PLUS)):
            with E.Engine(NTPB=512, NB=n // 512, N=1000, floor=floor) as e:
                e.init(1234)
                first = e.compute()
                m = best_of(e)
            out[name] = {"first_call": stats(first), "best_exec_ms": m.exec_ms, "path_steps_per_s": n * 1000 / (m.exec_ms * 1e-3)}
        with E.Engine(NTPB=512, NB=n // 512, N=1000, rng=E.RNG_XORWOW_COMPAT) as e:
            e.init(1234)
            first = e.compute()
            out["xorwow_compat"] = {"first_call": stats(first), "init_ms": e.init_ms}
        for name, mode in (("philox_dense", E.RNG_PHILOX_DENSE), ("xorwow_fast", E.RNG_XORWOW_FAST)):   # opt-in FE streams
            with E.Engine(NTPB=512, NB=n // 512, N=1000, rng=mode) as e:
                e.init(1234)
                first = e.compute()
                m = best_of(e)
                init_ms = e.init_ms
            out[name] = {"first_call": stats(first), "best_exec_ms": m.exec_ms, "init_ms": init_ms,
                         "path_steps_per_s": n * 1000 / (m.exec_ms * 1e-3)}
        out["reference_cuda_xorwow"] = ref_cuda("fe", "xorwow", n // 512)
        out["reference_cuda_philox"] = ref_cuda("fe", "philox", n // 512)
        r = out["reference_cuda_xorwow"]
        if r and "first_call" in r:
            f = r["first_call"]
            for key, mine in (("compat_vs_reference", "xorwow_compat"), ("xorwow_fast_vs_reference", "xorwow_fast")):
                c = out[mine]["first_call"]
                out[key] = {"rel_E": abs(c["E"] - f["E"]) / f["E"],
                            "rel_var": abs((c["E2"] - c["E"] ** 2) - (f["E2"] - f["E"] ** 2)) / (f["E2"] - f["E"] ** 2)}
        rep["c2"] = out
        print("c2", json.dumps(out), flush=True)

    if "c3" not in skip:
        n = 1 << 22
        out = {}
        with E.Engine(NTPB=512, NB=n // 512, N=1000, method=E.METHOD_EM) as e:
            e.init(1234)
            first = e.compute()
            m = best_of(e)
        out["native"] = {"first_call": stats(first), "best_exec_ms": m.exec_ms, "paths_per_s": n / (m.exec_ms * 1e-3)}
        with E.Engine(NTPB=512, NB=n // 512, N=1000, method=E.METHOD_EM, rng=E.RNG_XORWOW_COMPAT) as e:
            e.init(1234)
            first = e.compute()
            out["xorwow_compat"] = {"first_call": stats(first), "paths_per_s": n / (first.exec_ms * 1e-3)}
        out["reference_cuda_xorwow"] = ref_cuda("em", "xorwow", n // 512, repeat=1)
        r = out["reference_cuda_xorwow"]
        if r and "first_call" in r:
            c, f = out["xorwow_compat"]["first_call"], r["first_call"]
            out["compat_vs_reference"] = {"rel_E": abs(c["E"] - f["E"]) / f["E"],
                                          "rel_var": abs((c["E2"] - c["E"] ** 2) - (f["E2"] - f["E"] ** 2)) / (f["E2"] - f["E"] ** 2)}
        rep["c3"] = out
        print("c3", json.dumps(out), flush=True)

    if "c4" not in skip:
        P = args.c4_points
        ks = (0.1 + np.arange(P) * 9.9 / (P - 1)).astype(np.float32)
        ths = (0.01 + np.arange(P) * 0.49 / (P - 1)).astype(np.float32)
        sgs = (0.1 + np.arange(P) * 0.9 / (P - 1)).astype(np.float32)
        pts = [(k, t, s) for s in sgs for t in ths for k in ks if not (20 * k * t < s * s)]
        k, th, sg = (np.array(x, np.float32) for x in zip(*pts))
        n = 1 << args.c4_log2_paths
        out = {"points": len(k), "grid": f"{P}^3 with the reference's 20*k*theta<sigma^2 skip", "paths_per_point": n}
        for name, method, mode in (("fe", E.METHOD_FE, E.RNG_PHILOX), ("em", E.METHOD_EM, E.RNG_PHILOX),
                                   ("fe_philox_dense", E.METHOD_FE, E.RNG_PHILOX_DENSE),
                                   ("fe_xorwow_fast", E.METHOD_FE, E.RNG_XORWOW_FAST)):
            with E.Group(ngpu, NTPB=512, NB=n // 512, N=1000, method=method, rng=mode) as g:
                g.init(1234)
                t0 = time.perf_counter()
                res = g.explore(k, th, sg)
                wall = time.perf_counter() - t0
            ms = res[0].exec_ms
            units = len(k) * n * (1000 if name.startswith("fe") else 1)
            z = []
            for i in np.linspace(0, len(k) - 1, 12).astype(int):
                want = o.heston_call(kappa=float(k[i]), theta=float(th[i]), sigma=float(sg[i]))
                z.append({"k": float(k[i]), "theta": float(th[i]), "sigma": float(sg[i]), "E": res[i].mean, "heston": want,
                          "z": (res[i].mean - want) / res[i].std_error})
            out[name] = {"launch_ms": ms, "wall_s": wall, "gpus": ngpu,
                         ("path_steps_per_s" if name.startswith("fe") else "paths_per_s"): units / (ms * 1e-3), "sample_points": z}
            print("c4", name, json.dumps({k2: v for k2, v in out[name].items() if k2 != "sample_points"}), flush=True)
        rep["c4"] = out

    if "c5" not in skip:
        n = 1 << args.c5_log2_paths
        out = {"paths": n, "gpus": ngpu}
        for name, method in (("fe", E.METHOD_FE), ("em", E.METHOD_EM)):
            with E.Group(ngpu, NTPB=512, NB=1, n_paths=n, N=1000, method=method) as g:
                g.init(1234)
                g.compute()
                m = g.compute()
            units = n * (1000 if name == "fe" else 1)
            out[name] = dict(stats(m), **{("path_steps_per_s" if name == "fe" else "paths_per_s"): units / (m.exec_ms * 1e-3)})
            print("c5", name, json.dumps(out[name]), flush=True)
        rep["c5"] = out

    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(rep, f, indent=1)
    print("wrote", args.out)


if __name__ == "__main__":
    main()
