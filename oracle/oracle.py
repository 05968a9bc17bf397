"""ctypes front-end for the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py -- never by the product
package (nmch_b200/), which fails loudly without its CUDA library instead.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
CURAND_HOST_PATH = os.path.join(HERE, "_ref", "libcurand_host.so")
REF_HARNESS_PATH = os.path.join(HERE, "_ref", "nmch_ref_harness")
# the same build with the FE floor token changed to (.)+ (oracle/Makefile: _ref/nmch_ref_harness_plus)
REF_HARNESS_PLUS_PATH = os.path.join(HERE, "_ref", "nmch_ref_harness_plus")

RNG_XORWOW, RNG_PHILOX, RNG_MRG32K3A, RNG_PHILOX_DENSE = 0, 1, 2, 3
FLOOR_ABS, FLOOR_PLUS = 0, 1


class OrcParams(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("T", "S_0", "v_0", "r", "k", "rho", "theta", "sigma")] + [("N", C.c_int)]


class OrcRng(C.Structure):
    _fields_ = [
        ("kind", C.c_int), ("d", C.c_uint32), ("v", C.c_uint32 * 5),
        ("ctr", C.c_uint32 * 4), ("key", C.c_uint32 * 2), ("out", C.c_uint32 * 4), ("pos", C.c_int),
        ("s1", C.c_uint32 * 3), ("s2", C.c_uint32 * 3), ("dense_step", C.c_uint64),
        ("bm_flag", C.c_int), ("bm_extra", C.c_float), ("bm_flag_d", C.c_int), ("bm_extra_d", C.c_double),
    ]


@dataclass
class Params:
    """The 11 constructor values of nmch::methods::NMCH (NMCH.hpp:42), minus NTPB/NB."""
    T: float = 1.0
    S_0: float = 1.0
    v_0: float = 0.1
    r: float = 0.0
    k: float = 0.5
    rho: float = -0.7
    theta: float = 0.1
    sigma: float = 0.3
    N: int = 1000

    def c(self) -> OrcParams:
        return OrcParams(self.T, self.S_0, self.v_0, self.r, self.k, self.rho, self.theta, self.sigma, self.N)


def build(force: bool = False) -> None:
    """Compile liboracle.so (and, when the toolkit headers / reference are present, _ref/)."""
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(
            os.path.join(HERE, "nmch_oracle.c")):
        subprocess.run(["make", "-C", HERE, "liboracle.so"], check=True, capture_output=True)


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        u32p, f32p, f64p = C.POINTER(C.c_uint32), C.POINTER(C.c_float), C.POINTER(C.c_double)
        L.orc_philox4x32_10.argtypes = [u32p, u32p, u32p]
        L.orc_rng_init.argtypes = [C.POINTER(OrcRng), C.c_int, C.c_uint64, C.c_uint64, C.c_uint64]
        L.orc_rng_next.argtypes = [C.POINTER(OrcRng)]
        L.orc_rng_next.restype = C.c_uint32
        L.orc_uniform.argtypes = [C.POINTER(OrcRng)]
        L.orc_uniform.restype = C.c_float
        L.orc_normal2.argtypes = [C.POINTER(OrcRng), f32p, f32p]
        L.orc_normal.argtypes = [C.POINTER(OrcRng)]
        L.orc_normal.restype = C.c_float
        L.orc_normal_double.argtypes = [C.POINTER(OrcRng)]
        L.orc_normal_double.restype = C.c_double
        L.orc_poisson.argtypes = [C.POINTER(OrcRng), C.c_double]
        L.orc_poisson.restype = C.c_uint
        L.orc_gamma.argtypes = [C.POINTER(OrcRng), C.c_float]
        L.orc_gamma.restype = C.c_float
        L.orc_fe_run.argtypes = [C.POINTER(OrcParams), C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64,
                                 C.c_int, f32p, f32p, f64p, f64p, C.c_int]
        L.orc_fe_run_at.argtypes = [C.POINTER(OrcParams), C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64,
                                    C.c_int, f32p, f32p, f64p, f64p, C.c_int]
        L.orc_fe_sweep.argtypes = [C.POINTER(OrcParams), C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64,
                                   C.c_int, f32p, f32p, f32p, f64p, C.c_int]
        L.orc_em_run.argtypes = [C.POINTER(OrcParams), C.c_int, C.c_uint64, C.c_uint64, C.c_uint64,
                                 C.c_int, f32p, f32p, f64p, f64p, C.c_int]
        L.orc_qe_run.argtypes = [C.POINTER(OrcParams), C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, f32p, f32p, f64p,
                                 f64p, C.c_int]
        L.orc_em_native_run.argtypes = [C.POINTER(OrcParams), C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, f32p, f32p, f64p,
                                        f64p, C.c_int]
        L.orc_em_exact_run.argtypes = [C.POINTER(OrcParams), C.c_uint64, C.c_uint64, f64p, f64p, f64p, C.c_int]
        L.orc_get_err.argtypes = [C.c_int, C.c_float, C.c_float]
        L.orc_get_err.restype = C.c_float
        L.orc_NP.argtypes = [C.c_double]
        L.orc_NP.restype = C.c_double
        L.orc_print_true_price.argtypes = [C.c_float] * 4
        L.orc_print_true_price.restype = C.c_float
        L.orc_heston_call.argtypes = [C.c_double] * 9
        L.orc_heston_call.restype = C.c_double
        L.orc_exploration_grid.argtypes = [C.c_int, C.c_int, f32p, f32p, f32p, C.c_int]
        L.orc_exploration_grid.restype = C.c_int
        L.orc_max_threads.restype = C.c_int
        L.orc_rng_init_only.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int]
        L.orc_rng_init_only.restype = C.c_uint64
        _lib = L
    return _lib


def _fp(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def philox4x32_10(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, o)
    return list(o)


class Rng:
    """One cuRAND-compatible stream: curand_init(seed, subsequence, offset)."""

    def __init__(self, kind: int, seed: int, subsequence: int = 0, offset: int = 0):
        self.s = OrcRng()
        lib().orc_rng_init(C.byref(self.s), kind, seed, subsequence, offset)

    def next(self) -> int:
        return lib().orc_rng_next(C.byref(self.s))

    def uniform(self) -> float:
        return lib().orc_uniform(C.byref(self.s))

    def normal2(self):
        a, b = C.c_float(), C.c_float()
        lib().orc_normal2(C.byref(self.s), C.byref(a), C.byref(b))
        return a.value, b.value

    def normal(self) -> float:
        return lib().orc_normal(C.byref(self.s))

    def normal_double(self) -> float:
        return lib().orc_normal_double(C.byref(self.s))

    def poisson(self, lam: float) -> int:
        return lib().orc_poisson(C.byref(self.s), lam)

    def gamma(self, alpha: float) -> float:
        return lib().orc_gamma(C.byref(self.s), alpha)

    @property
    def xorwow_state(self):
        return self.s.d, list(self.s.v)

    @property
    def mrg_state(self):
        return list(self.s.s1) + list(self.s.s2)


def _run(fn, p: Params, args_mid, first_path, n_paths, calls, want_paths, threads):
    S = np.empty(n_paths, np.float32) if want_paths else None
    V = np.empty(n_paths, np.float32) if want_paths else None
    s, s2 = C.c_double(), C.c_double()
    cp = p.c()
    fn(C.byref(cp), *args_mid, first_path, n_paths, calls, _fp(S, C.c_float), _fp(V, C.c_float),
       C.byref(s), C.byref(s2), threads)
    return {"sum": s.value, "sumsq": s2.value, "n": n_paths, "mean": s.value / n_paths,
            "mean_sq": s2.value / n_paths, "S": S, "V": V}


def fe_run(p: Params, rng=RNG_XORWOW, floor=FLOOR_ABS, seed=1234, first_path=0, n_paths=1024, calls=1,
           want_paths=False, threads=0):
    """Reference FE kernel semantics (NMCH_FE.cu:145-175) for paths [first_path, first_path+n_paths)."""
    return _run(lib().orc_fe_run, p, (rng, floor, seed), first_path, n_paths, calls, want_paths, threads)


def fe_run_at(p: Params, offset, rng=RNG_PHILOX, floor=FLOOR_ABS, seed=1234, first_path=0, n_paths=1024, calls=1,
              want_paths=False, threads=0):
    """fe_run with the streams started `offset` 32-bit draws in (curand_init's offset argument)."""
    return _run(lib().orc_fe_run_at, p, (rng, floor, seed, offset), first_path, n_paths, calls, want_paths, threads)


def fe_tangent_run(p: Params, rng=RNG_PHILOX, floor=FLOOR_ABS, seed=1234, first_path=0, n_paths=1024, threads=0):
    """FE paths with the pathwise tangent dS_T/dv_0 (checker of compute_greeks): returns S, V (float32), B (float64)."""
    S = np.empty(n_paths, np.float32)
    V = np.empty(n_paths, np.float32)
    B = np.empty(n_paths, np.float64)
    cp = p.c()
    fn = lib().orc_fe_tangent_run
    fn.restype = None
    fn.argtypes = [C.POINTER(OrcParams), C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64,
                   C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_double), C.c_int]
    fn(C.byref(cp), rng, floor, seed, first_path, n_paths, _fp(S, C.c_float), _fp(V, C.c_float), _fp(B, C.c_double), threads)
    return {"S": S, "V": V, "B": B}


def fe_sweep(p: Params, k, theta, sigma, rng=RNG_XORWOW, floor=FLOOR_ABS, seed=1234, first_path=0, n_paths=1024,
             threads=0):
    """exploration.cu:71-88 for FE: returns an (n_points, 2) array of raw payoff sums."""
    k = np.ascontiguousarray(k, np.float32)
    theta = np.ascontiguousarray(theta, np.float32)
    sigma = np.ascontiguousarray(sigma, np.float32)
    sums = np.zeros(2 * len(k), np.float64)
    cp = p.c()
    lib().orc_fe_sweep(C.byref(cp), rng, floor, seed, first_path, n_paths, len(k), _fp(k, C.c_float),
                       _fp(theta, C.c_float), _fp(sigma, C.c_float), _fp(sums, C.c_double), threads)
    return sums.reshape(-1, 2)


def em_run(p: Params, rng=RNG_XORWOW, seed=1234, first_path=0, n_paths=1024, calls=1, want_paths=False,
           threads=0):
    """Reference EM kernel semantics (NMCH_EM.cu:213-260)."""
    return _run(lib().orc_em_run, p, (rng, seed), first_path, n_paths, calls, want_paths, threads)


def qe_run(p: Params, seed=1234, first_path=0, n_paths=1024, call=0, want_paths=False, threads=0):
    """QE-M restated on the kernel's own draw mapping (checker for nmch_b200/csrc/qe_kernels.cu)."""
    S = np.empty(n_paths, np.float32) if want_paths else None
    V = np.empty(n_paths, np.float32) if want_paths else None
    s, s2 = C.c_double(), C.c_double()
    cp = p.c()
    lib().orc_qe_run(C.byref(cp), seed, first_path, n_paths, call, _fp(S, C.c_float), _fp(V, C.c_float), C.byref(s),
                     C.byref(s2), threads)
    return {"sum": s.value, "sumsq": s2.value, "n": n_paths, "mean": s.value / n_paths, "mean_sq": s2.value / n_paths,
            "S": S, "V": V}


def em_native_run(p: Params, seed=1234, first_path=0, n_paths=1024, call=0, want_paths=False, threads=0):
    """The product's native EM sampler restated on the kernel's own Philox words (checker for em_native_kernel)."""
    S = np.empty(n_paths, np.float32) if want_paths else None
    V = np.empty(n_paths, np.float32) if want_paths else None
    s, s2 = C.c_double(), C.c_double()
    cp = p.c()
    lib().orc_em_native_run(C.byref(cp), seed, first_path, n_paths, call, _fp(S, C.c_float), _fp(V, C.c_float), C.byref(s),
                            C.byref(s2), threads)
    return {"sum": s.value, "sumsq": s2.value, "n": n_paths, "mean": s.value / n_paths, "mean_sq": s2.value / n_paths,
            "S": S, "V": V}


def em_exact_run(p: Params, seed=1, n_paths=1024, threads=0):
    s, s2, sS = C.c_double(), C.c_double(), C.c_double()
    cp = p.c()
    lib().orc_em_exact_run(C.byref(cp), seed, n_paths, C.byref(s), C.byref(s2), C.byref(sS), threads)
    return {"sum": s.value, "sumsq": s2.value, "n": n_paths, "mean": s.value / n_paths,
            "mean_sq": s2.value / n_paths, "mean_ST": sS.value / n_paths}


def get_err(n: int, strike_price: float, price_squared: float) -> float:
    return lib().orc_get_err(n, strike_price, price_squared)


def heston_call(S0=1.0, K=1.0, v0=0.1, r=0.0, kappa=0.5, theta=0.1, sigma=0.3, rho=-0.7, T=1.0) -> float:
    return lib().orc_heston_call(S0, K, v0, r, kappa, theta, sigma, rho, T)


def exploration_grid(steps=5, apply_filter=True):
    cap = (steps + 2) ** 3
    k = np.empty(cap, np.float32)
    th = np.empty(cap, np.float32)
    sg = np.empty(cap, np.float32)
    n = lib().orc_exploration_grid(steps, int(apply_filter), _fp(k, C.c_float), _fp(th, C.c_float),
                                   _fp(sg, C.c_float), cap)
    return k[:n].copy(), th[:n].copy(), sg[:n].copy()


def std_error(mean: float, mean_sq: float, n: int) -> float:
    """Plain standard error of the mean (what the reference's get_err over-states)."""
    return float(np.sqrt(max(mean_sq - mean * mean, 0.0) / n))


def max_threads() -> int:
    return lib().orc_max_threads()


def host_threads() -> int:
    """Host threads this process may run on (its CPU affinity), whatever OMP_NUM_THREADS says: torch.distributed.run
    exports OMP_NUM_THREADS=1, which would make omp_get_max_threads() -- and a CPU baseline sized by it -- one thread."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def rng_init_only(rng=RNG_XORWOW, seed=1234, first_path=0, n_paths=1024, threads=0) -> int:
    """curand_init for every path and nothing else (the reference's Tim_init span); returns a checksum."""
    return lib().orc_rng_init_only(rng, seed, first_path, n_paths, threads)
