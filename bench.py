#!/usr/bin/env python
"""bench.py -- throughput of the Heston path-simulation hot path on N B200s (one process per GPU).

Contract: `python bench.py --gpus N --steps K --warmup W` (N>1: launched by torch.distributed.run) prints ONE
JSON line on rank 0.  A "step" is one compute() pass of the hot path over one batch of synthetic paths:
BASELINE.json configs[1] (FE Euler, N=1000 time steps, 2^24 paths per GPU, README parameters, seed 1234).
Paths shard over ranks by disjoint Philox subsequences (weak scaling: 2^24 paths per GPU); one NCCL
allreduce of the two FP64 moments per step when N>1.

  value     FE: path-steps/s (EM: paths/s), whole job, inputs are 11 scalars (nothing to stage in HBM)
  e2e       the same metric through the public C-ABI call nmch_engine_compute() with host buffers:
            kernel-parameter upload + launch + sync + the result landing in host memory, every step
  roofline  the FE kernel against the FP32/SFU ISSUE roofline of SURVEY.md §8d (this path moves no HBM
            bytes and has no GEMM: neither "hbm" nor "tensor" bounds it)
  cpu_baseline  the oracle port of the reference loop on this box's host cores (bounded sample)

`--impl reference`: the reference arm.  The reference is a CUDA program with no CPU implementation; its hot
loop restated in C (oracle/, kind "port") is timed on the host cores as the tier asks, and when the
reference's own CUDA build (oracle/_ref/nmch_ref_harness) travelled to the box its measured throughput on
the same GPU is attached as "reference_cuda" -- that is the like-for-like number.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

README = dict(T=1.0, S_0=1.0, v_0=0.1, r=0.0, k=0.5, rho=-0.7, theta=0.1, sigma=0.3)
ISSUE_PER_CLK_PER_SM = min(128.0 / 42.0, 16.0 / 5.0)      # 42 thread-instr, 5 MUFU per path-step
# profiles/r01_fe_mix_bound.txt: the FE kernel's own instruction mix, issued from independent chains on this GPU,
# needs 55.04 cycles per warp-step per SM sub-partition (the Philox IMAD.WIDE.U32 costs ~5.2 issue cycles)
MIX_BOUND_CYCLES_PER_WARP_STEP = 55.04
# SURVEY.md §8d for EM: "estimate 90-100 instr/step => 3.9e8 paths/s/GPU ceiling at N=1000" = 95 thread-instr per
# path-step at one instruction per scheduler per clock
EM_INSTR_PER_PATH_STEP = 95.0
# profiles/r01_fe_mix_bound_dense_em.txt: the EM trial's instruction mix (14 IMAD.WIDE, 18 LOP3, 2 SHF, 24 FP32,
# 9 MUFU) from independent chains needs 96.83 cycles per warp-trial per SM sub-partition; 1.045 trials per step
EM_MIX_BOUND_CYCLES_PER_WARP_STEP = 96.83 * 1.045


# ------------------------------------------------------------------------------------------------
# clocks (NVML) sampled DURING the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz, self.power = [], set(), None, []
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(
                    self.nv, "nvmlDeviceGetCurrentClocksEventReasons") else self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
                self.power.append(self.nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                pass
            self._stop.wait(0.02)

    def __enter__(self):
        if self.nv:
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._t:
            self._t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml_unavailable"]}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s), "power_w_max": max(self.power) if self.power else None}


# ------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port of the reference's loop, all host threads, bounded sample
# ------------------------------------------------------------------------------------------------
def cpu_baseline(method: str, N: int, budget_s: float = 12.0):
    from oracle import oracle as o
    p = o.Params(N=N, **README)
    threads = o.max_threads()
    n = 1 << 13
    run = (lambda n: o.fe_run(p, rng=o.RNG_XORWOW, n_paths=n)) if method == "fe" else (
        lambda n: o.em_run(p, rng=o.RNG_XORWOW, n_paths=n))
    run(256)                                     # builds the skip-ahead tables outside the timed region
    t0 = time.perf_counter()
    run(n)
    dt = time.perf_counter() - t0
    while dt < min(1.0, budget_s) and n < (1 << 24):   # find the rate, then size one sample to the budget
        n *= 4
        t0 = time.perf_counter()
        run(n)
        dt = time.perf_counter() - t0
    n_big = int(min(max(n, n * budget_s / max(dt, 1e-9)), 1 << 26))
    n_big = max(1 << 13, 1 << (n_big.bit_length() - 1))
    t0 = time.perf_counter()
    run(n_big)
    dt = time.perf_counter() - t0
    units = n_big * N if method == "fe" else n_big
    return {"value": units / dt, "unit": "path-steps/s" if method == "fe" else "paths/s", "cores": threads,
            "kind": "port", "sample": f"{method.upper()} oracle (XORWOW stream incl. per-path curand_init), "
                                      f"{n_big} paths x {N} steps, {dt:.2f} s, {threads} OpenMP threads"}


def reference_cuda(method: str, log2_paths: int, N: int, repeat: int = 3, with_ours: bool = True):
    """The reference's own CUDA build (unmodified sources, nvcc -arch=sm_100) on this GPU, if shipped.
    with_ours=False (the reference arm): only the reference binary runs -- nothing of this engine is loaded."""
    exe = os.path.join(ROOT, "oracle", "_ref", "nmch_ref_harness")
    if not os.path.exists(exe):
        return None
    try:
        import torch
        if not torch.cuda.is_available():
            return None
    except Exception:
        return None
    out = {}
    n = 1 << log2_paths
    for rng in ("xorwow", "philox"):
        try:
            r = subprocess.run([exe, "--method", method, "--rng", rng, "--kernel", "k3", "--NTPB", "512", "--NB",
                                str(n // 512), "--N", str(N), "--repeat", str(repeat + 1)],
                               capture_output=True, text=True, timeout=900, check=True)
            rows = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")][1:]   # drop the warm-up call
            ms = min(x["exec_ms"] for x in rows)
            units = n * N if method == "fe" else n
            out[rng] = {"value": units / (ms * 1e-3), "exec_ms": ms, "init_ms": rows[0]["init_ms"],
                        "E": rows[-1]["E"], "E2": rows[-1]["E2"]}
            if not with_ours:
                continue
            # same seed, same calls through OUR draw-compatible stream mode: identical results, our timing
            try:
                from nmch_b200 import engine as E
                mode = E.RNG_XORWOW_COMPAT if rng == "xorwow" else E.RNG_PHILOX_COMPAT
                with E.Engine(NTPB=512, NB=n // 512, N=N, method=E.METHOD_FE if method == "fe" else E.METHOD_EM, rng=mode,
                              **README) as eng:
                    eng.init(1234)
                    eng.compute()
                    ours = [eng.compute() for _ in range(repeat)]
                rel = max(abs(o_.mean - r_["E"]) / abs(r_["E"]) for o_, r_ in zip(ours, rows))
                best = min(o_.exec_ms for o_ in ours)
                out[rng]["ours_same_draws"] = {"value": units / (best * 1e-3), "exec_ms": best, "max_rel_diff_E": rel}
            except Exception as ex:  # noqa: BLE001
                out[rng]["ours_same_draws"] = {"error": str(ex)[:200]}
            if rng == "xorwow" and method == "fe":
                # the same integer draws through the native fast-math step (opt-in NMCH_RNG_XORWOW_FAST)
                try:
                    with E.Engine(NTPB=512, NB=n // 512, N=N, rng=E.RNG_XORWOW_FAST, **README) as eng:
                        eng.init(1234)
                        eng.compute()
                        ours = [eng.compute() for _ in range(repeat)]
                    rel = max(abs(o_.mean - r_["E"]) / abs(r_["E"]) for o_, r_ in zip(ours, rows))
                    relv = max(abs(o_.variance - (r_["E2"] - r_["E"] ** 2)) / (r_["E2"] - r_["E"] ** 2)
                               for o_, r_ in zip(ours, rows))
                    best = min(o_.exec_ms for o_ in ours)
                    out[rng]["ours_same_stream_fast"] = {"value": units / (best * 1e-3), "exec_ms": best,
                                                         "max_rel_diff_E": rel, "max_rel_diff_var": relv}
                except Exception as ex:  # noqa: BLE001
                    out[rng]["ours_same_stream_fast"] = {"error": str(ex)[:200]}
        except Exception as ex:  # noqa: BLE001
            out[rng] = {"error": str(ex)[:200]}
    out["what"] = (f"reference NMCH_{method.upper()}_K3_MM<rng> (unmodified sources, -O3 -arch=sm_100), 512 x {n // 512} "
                   f"paths, N={N}, best Tim_exec of {repeat} after one warm-up compute(); unit as `unit`")
    if with_ours:
        out["what"] += ("; ours_same_draws = this engine in the draw-compatible mode for that tag, same seed and calls "
                        "(relative difference of E[X]); ours_same_stream_fast = the same XORWOW integer draws through "
                        "the native fast-math step (NMCH_RNG_XORWOW_FAST)")
    return out


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--method", default="fe", choices=["fe", "em"])
    ap.add_argument("--log2-paths", type=int, default=None, help="paths per GPU (default 24 for FE, 22 for EM)")
    ap.add_argument("--N", type=int, default=1000)
    ap.add_argument("--floor", default="abs", choices=["abs", "plus"])
    ap.add_argument("--rng", default="philox", choices=["philox", "dense"],
                    help="philox: word-compatible native stream (default); dense: opt-in 3-steps-per-block FE stream")
    ap.add_argument("--paths-per-thread", type=int, default=0)
    ap.add_argument("--block-threads", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reference-cuda", action="store_true")
    ap.add_argument("--cpu-budget-s", type=float, default=None, help="seconds of host work per CPU-baseline sample")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    log2_paths = args.log2_paths if args.log2_paths is not None else (24 if args.method == "fe" else 22)
    n_per_gpu = 1 << log2_paths
    N = args.N
    metric = "fe_path_steps_per_s" if args.method == "fe" else "em_paths_per_s"
    unit = "path-steps/s" if args.method == "fe" else "paths/s"
    units_per_gpu_step = n_per_gpu * N if args.method == "fe" else n_per_gpu
    workload = (f"BASELINE configs[1]: FE Euler |.| floor, README params, N={N}, 2^{log2_paths} paths per GPU, seed 1234"
                if args.method == "fe" else
                f"BASELINE configs[2]: EM exact scheme, README params, N={N}, 2^{log2_paths} paths per GPU, seed 1234")
    config = {"workload": workload, "method": args.method, "floor": args.floor, "rng": args.rng, "n_steps": N,
              "paths_per_gpu": n_per_gpu, "global_paths": n_per_gpu * world, "parallelism": f"paths sharded x{world}",
              "l2": "n/a: the kernel has no HBM-resident inputs (11 scalars by value), state lives in registers"}

    # ---------------------------------------------------------------- reference arm (host CPU port)
    if args.impl == "reference":
        if rank != 0:
            return 0
        times, walls = [], []
        base = None
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            base = cpu_baseline(args.method, N, budget_s=args.cpu_budget_s if args.cpu_budget_s is not None
                                else max(2.0, 60.0 / max(1, args.steps + args.warmup)))
            if i >= args.warmup:
                times.append(base["value"])
                walls.append(time.perf_counter() - t0)
        v = sum(times) / len(times)
        base["value"] = v
        line = {"impl": "reference", "metric": metric, "value": v, "unit": unit, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * sum(walls) / len(walls), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config, "cpu_baseline": base,
                "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "note": "edo01/NMCH is CUDA-only; its hot loop restated in C (oracle/) is what runs on the host cores"}
        if not args.no_reference_cuda:
            line["reference_cuda"] = reference_cuda(args.method, log2_paths, N, with_ours=False)
        print(json.dumps(line))
        return 0

    # ---------------------------------------------------------------- our arm
    import torch
    import torch.distributed as dist

    from nmch_b200 import engine as E

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    saved_stdout = None
    if world > 1:
        # NCCL prints its version banner on the C-level stdout when the communicator comes up: park fd 1 on stderr
        # until the JSON line is due, so that stdout carries exactly one line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)

    from nmch_b200.distributed import ShardedEngine
    stream = torch.cuda.Stream(dev)             # one stream carries the kernel and the allreduce
    torch.cuda.set_stream(stream)
    sh = ShardedEngine(rank=rank, world=world, device=local_rank, NTPB=512, NB=(n_per_gpu * world) // 512, N=N,
                       method=E.METHOD_FE if args.method == "fe" else E.METHOD_EM,
                       floor=E.FLOOR_ABS if args.floor == "abs" else E.FLOOR_PLUS,
                       rng=E.RNG_PHILOX_DENSE if (args.rng == "dense" and args.method == "fe") else E.RNG_PHILOX,
                       paths_per_thread=args.paths_per_thread, block_threads=args.block_threads, **README)
    sh.init(1234)
    eng = sh.engine
    assert eng.n_local == n_per_gpu
    moments = None

    def step():
        nonlocal moments
        moments = sh.compute_async()             # kernel -> (N>1: one NCCL allreduce of 16 bytes of FP64 moments)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = eng.launch_info()["kernel_launches"]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        ev0.record(stream)
        for _ in range(args.steps):
            step()
        ev1.record(stream)
        barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    launches = eng.launch_info()["kernel_launches"] - launches0
    result = moments.cpu().numpy()
    ms_per_step = total_ms / args.steps
    value = units_per_gpu_step * world / (ms_per_step * 1e-3)

    # ---- end to end through the public C-ABI call (host in, host out), every step
    barrier()
    t0 = time.perf_counter()
    e2e_last = None
    for _ in range(args.steps):
        if world == 1:
            e2e_last = eng.compute()             # nmch_engine_compute: param upload + launch + sync + host result
        else:
            step()
            e2e_last = moments.cpu()             # D2H read of the reduced moments
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = units_per_gpu_step * world * args.steps / float(e2e_s.item())

    if saved_stdout is not None:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    if rank == 0:
        info = eng.launch_info()
        ck = clocks.summary()
        n_total = n_per_gpu * world
        mean = float(result[0]) / n_total
        var = float(result[1]) / n_total - mean * mean
        f_hz = (ck["sm_mhz"] or 1965) * 1e6
        per_gpu = value / world
        if args.method == "fe":
            peak = info["sm_count"] * f_hz * ISSUE_PER_CLK_PER_SM
        else:
            peak = info["sm_count"] * f_hz * 128.0 / (EM_INSTR_PER_PATH_STEP * N)
        # measured bound of the kernel's own instruction mix on this GPU (profiles/microbench/fe_mix_bound.cu)
        mix_peak = (info["sm_count"] * f_hz * 4 * 32 / MIX_BOUND_CYCLES_PER_WARP_STEP if args.method == "fe" else
                    info["sm_count"] * f_hz * 4 * 32 / (EM_MIX_BOUND_CYCLES_PER_WARP_STEP * N))
        traffic = None
        tj = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tj):
            try:
                traffic = json.load(open(tj)).get("fe_dense" if (args.method == "fe" and args.rng == "dense") else args.method)
            except Exception:
                traffic = None
        line = {
            "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config,
            "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": info["kernel_param_bytes"], "d2h_bytes_per_step": 16,
                    "api": "nmch_engine_compute (C ABI)" if world == 1 else "nmch_engine_compute_async + NCCL allreduce + D2H"},
            "gpu_launches": int(launches),
            "clocks": ck,
            "roofline": {"bound": "issue", "achieved": per_gpu, "peak": peak, "unit": unit, "frac": per_gpu / peak,
                         "traffic": traffic,
                         "model": "SURVEY.md §8d: SMs x f x min(128/42 issue, 16/5 MUFU) path-steps/s; f = median SM clock "
                                  "sampled during the timed region; per-GPU achieved" if args.method == "fe" else
                                  "SURVEY.md §8d estimate for EM: SMs x f x 128 / (95 thread-instr per path-step x N) paths/s "
                                  "(data-dependent scheme: reported, not targeted); f as for FE; per-GPU achieved",
                         "peak_at_max_clock": (info["sm_count"] * (ck["sm_max_mhz"] or 1965) * 1e6 * ISSUE_PER_CLK_PER_SM
                                               if args.method == "fe" else None),
                         "mix_bound_peak": mix_peak, "mix_bound_frac": per_gpu / mix_peak},
            "kernel": {k: info[k] for k in ("grid_x", "grid_y", "block_threads", "paths_per_thread", "regs_per_thread", "sm_count")},
            "result": {"E[X]": mean, "var": var, "std_error": (var / n_total) ** 0.5,
                       "heston_semi_analytic": 0.1197325094 if args.method in ("fe", "em") else None},
        }
        if world == 1 and args.method == "fe" and args.rng == "philox":
            try:                                  # the opt-in dense-draw stream, same workload, for the record
                with E.Engine(NTPB=512, NB=n_per_gpu // 512, N=N, rng=E.RNG_PHILOX_DENSE, device=local_rank,
                              floor=E.FLOOR_ABS if args.floor == "abs" else E.FLOOR_PLUS, **README) as de:
                    de.init(1234)
                    de.compute()
                    dms = min(de.compute().exec_ms for _ in range(3))
                dval = units_per_gpu_step / (dms * 1e-3)
                line["dense_mode"] = {"value": dval, "unit": unit, "ms_per_step": dms, "roofline_frac": dval / peak,
                                      "note": "NMCH_RNG_PHILOX_DENSE: three (23-bit, 19-bit) draws per Philox block; "
                                              "statistically equivalent, not cuRAND-word-compatible; bench.py --rng dense"}
            except Exception as ex:  # noqa: BLE001
                line["dense_mode"] = {"error": str(ex)[:200]}
            try:                                  # the reference's default stream (XORWOW) through the native step
                with E.Engine(NTPB=512, NB=n_per_gpu // 512, N=N, rng=E.RNG_XORWOW_FAST, device=local_rank,
                              floor=E.FLOOR_ABS if args.floor == "abs" else E.FLOOR_PLUS, **README) as xe:
                    xe.init(1234)
                    xe.compute()
                    xms = min(xe.compute().exec_ms for _ in range(3))
                    xinit = xe.init_ms
                xval = units_per_gpu_step / (xms * 1e-3)
                line["xorwow_fast_mode"] = {"value": xval, "unit": unit, "ms_per_step": xms, "init_ms": xinit,
                                            "note": "NMCH_RNG_XORWOW_FAST: cuRAND-XORWOW integer draws (the reference's "
                                                    "default generator, 24 B of state per path) through the native "
                                                    "fast-math step; same-seed agreement with the reference CUDA build "
                                                    "under reference_cuda.xorwow.ours_same_stream_fast"}
            except Exception as ex:  # noqa: BLE001
                line["xorwow_fast_mode"] = {"error": str(ex)[:200]}
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(args.method, N, budget_s=args.cpu_budget_s or 12.0)
        if not args.no_reference_cuda and world == 1:
            line["reference_cuda"] = reference_cuda(args.method, log2_paths, N)
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
