"""Python mirror of the reference's method API (nmch::methods, include/NMCH/methods/*.hpp).

Same class names, constructor argument order, lifecycle (init(seed) / compute() / print_stats() /
finalize()), getters and setters as the C++ templates, so tests read like the reference's own usage
(README.md:60-93, src/NMCH/test/nmch.cu:115-134).  The `rnd_state` template tag becomes a keyword:
"curandStateXORWOW_t" -> cuRAND-XORWOW-compatible stream, "curandStateMRG32k3a_t" -> cuRAND-MRG32k3a-compatible
stream, "curandStatePhilox4_32_10_t" -> native fused Philox (pass compat=True for the cuRAND-Philox
draw-and-arithmetic compatible validation mode).
All K1/K2/K3/MM/PgM/PiM variants map to the one engine (their differences were memory-space and
reduction experiments, NMCH_FE.hpp:76-189).  No arithmetic happens here.
"""
from __future__ import annotations

import math

import numpy as np

from . import engine as _eng

XORWOW = "curandStateXORWOW_t"
PHILOX = "curandStatePhilox4_32_10_t"
MRG32K3A = "curandStateMRG32k3a_t"


def _NP(x: float) -> float:
    # Abramowitz-Stegun 26.2.17 as the reference prints it (src/NMCH/utils/utils.cu:5-25)
    p, b1, b2, b3, b4, b5, c = 0.2316419, 0.319381530, -0.356563782, 1.781477937, -1.821255978, 1.330274429, 0.39894228
    if x >= 0.0:
        t = 1.0 / (1.0 + p * x)
        return 1.0 - c * math.exp(-x * x / 2.0) * t * (t * (t * (t * (t * b5 + b4) + b3) + b2) + b1)
    t = 1.0 / (1.0 - p * x)
    return c * math.exp(-x * x / 2.0) * t * (t * (t * (t * (t * b5 + b4) + b3) + b2) + b1)


class NMCH:
    """Abstract base (NMCH.hpp:28-115): parameter bag + result fields."""
    _method = _eng.METHOD_FE
    _title = ""
    _k1_kernel = False      # True for the classes that launch FE_k1 / EM_k1 in the reference (K1_MM, K1_PgM, K1_PiM)

    def __init__(self, NTPB, NB, T, S_0, v_0, r, k, rho, theta, sigma, N, rnd_state=PHILOX, *, compat=False,
                 dense=False, fast=False, floor="abs", device=-1, first_path=0, n_local=0, n_paths=0,
                 paths_per_thread=0):
        self.NTPB, self.NB, self.T, self.S_0, self.v_0, self.r = NTPB, NB, T, S_0, v_0, r
        self.k, self.rho, self.theta, self.sigma, self.N = k, rho, theta, sigma, N
        self.K = S_0                                   # NMCH.cu:7
        self.dt = np.float32(T) / np.float32(N)        # NMCH.cu:9
        self.state_numbers = n_paths or NTPB * NB      # NMCH_FE.cu:317
        self.strike_price = 0.0
        self.price_squared = 0.0
        self.Tim_exec = 0.0
        self.Tim_init = 0.0
        fe = self._method == _eng.METHOD_FE
        if rnd_state == XORWOW:                        # fast: cuRAND's integer draws, native fast-math step (FE only)
            rng = _eng.RNG_XORWOW_FAST if (fast and fe) else _eng.RNG_XORWOW_COMPAT
        elif rnd_state == PHILOX:                      # dense: three steps per Philox block (FE only)
            rng = _eng.RNG_PHILOX_COMPAT if compat else (_eng.RNG_PHILOX_DENSE if (dense and fe) else _eng.RNG_PHILOX)
        elif rnd_state == MRG32K3A:
            rng = _eng.RNG_MRG32K3A_COMPAT
        else:
            raise ValueError(f"unknown rnd_state tag {rnd_state!r}")
        self._engine = _eng.Engine(NTPB, NB, T, S_0, v_0, r, k, rho, theta, sigma, N, method=self._method,
                                   floor=_eng.FLOOR_ABS if floor == "abs" else _eng.FLOOR_PLUS, rng=rng,
                                   device=device, n_paths=n_paths, first_path=first_path, n_local=n_local,
                                   paths_per_thread=paths_per_thread)
        self.last_moments = None
        self.legacy_k1_moment = False

    def set_legacy_k1_moment(self, on: bool) -> None:
        """Exact legacy output of the reference's K1 kernels: price_squared = E[X^2]/n^2 (NMCH_FE.cu:56-58,
        NMCH_EM.cu:129-131: they reduce (payoff/n)^2/n).  K1 classes only; K2 / K3 are unaffected."""
        self.legacy_k1_moment = bool(on)

    # lifecycle -------------------------------------------------------------------------------
    def init(self, seed: int) -> None:
        self._engine.init(seed)
        self.Tim_init = self._engine.init_ms

    def compute(self) -> None:
        self._engine.set_params(self.k, self.theta, self.sigma)
        m = self._engine.compute()
        self.last_moments = m
        self.strike_price = float(np.float32(m.mean))
        self.price_squared = float(np.float32(m.mean_sq))
        if self.legacy_k1_moment and self._k1_kernel:
            n = float(self.state_numbers)
            self.price_squared = float(np.float32(m.sum_payoff_sq / n / n / n))
        self.Tim_exec = m.exec_ms

    def finalize(self) -> None:
        self._engine.finalize()

    # getters / setters (NMCH.hpp:65-80, NMCH_FE.hpp:43-55) -------------------------------------
    def get_strike_price(self) -> float:
        return self.strike_price

    def get_price_squared(self) -> float:
        return self.price_squared

    def get_execution_time(self) -> float:
        return self.Tim_exec

    def set_k(self, k): self.k = k
    def set_theta(self, theta): self.theta = theta
    def set_sigma(self, sigma): self.sigma = sigma

    def get_err(self) -> float:
        """The reference's 95% half-width formula, float/double mix included (NMCH_FE.hpp:50-55)."""
        n = self.state_numbers
        inv = float(np.float32(1.0) / np.float32(n - 1))
        inner = float(np.float32(n) * np.float32(self.price_squared)
                      - np.float32(self.strike_price) * np.float32(self.strike_price))
        return float(np.float32(1.96 * math.sqrt(inv * inner) / math.sqrt(float(n))))

    def true_price_line(self) -> float:
        # Black-Scholes with vol := sigma, T := 1 (NMCH_FE.cu:336-338) -- not the Heston price
        s = float(self.sigma)
        return float(np.float32(self.S_0 * _NP((self.r + 0.5 * s * s) / s)
                                - self.K * float(np.exp(np.float32(-self.r))) * _NP((self.r - 0.5 * s * s) / s)))

    def stats_text(self) -> str:
        lines = ["Base parameters:", "NTPB    = %d" % self.NTPB, "NB      = %d" % self.NB, "T       = %f" % self.T,
                 "S_0,K   = %f" % self.S_0, "v_0     = %f" % self.v_0, "r       = %f" % self.r, "k       = %f" % self.k,
                 "theta   = %f" % self.theta, "sigma   = %f" % self.sigma, "N       = %d" % self.N,
                 "dt      = %f" % float(self.dt), "METHOD: %s" % self._title,
                 "The estimated price E[X] is equal to %f" % self.strike_price,
                 "The estimated E[X^2] is equal to %f" % self.price_squared,
                 "The true price %f" % self.true_price_line(),
                 "error associated to a confidence interval of 95%% = %f" % self.get_err(),
                 "Execution time %f ms" % self.Tim_exec, "Initialization time %f ms" % self.Tim_init]
        return "\n".join(lines) + "\n"

    def print_stats(self) -> None:
        print(self.stats_text(), end="")


class NMCH_FE_K1(NMCH):
    _method = _eng.METHOD_FE
    _title = "FORWARD-EULER"


class NMCH_FE_K1_MM(NMCH_FE_K1): _k1_kernel = True
class NMCH_FE_K2_MM(NMCH_FE_K1_MM): _k1_kernel = False
class NMCH_FE_K3_MM(NMCH_FE_K2_MM): pass
class NMCH_FE_K1_PgM(NMCH_FE_K1): _k1_kernel = True
class NMCH_FE_K1_PiM(NMCH_FE_K1): _k1_kernel = True


class NMCH_FE_K2_PHILOX_MM(NMCH_FE_K1_MM):
    """Non-template in the reference (NMCH_FE.hpp:142): always the Philox tag."""

    _k1_kernel = False

    def __init__(self, NTPB, NB, T, S_0, v_0, r, k, rho, theta, sigma, N, **kw):
        super().__init__(NTPB, NB, T, S_0, v_0, r, k, rho, theta, sigma, N, PHILOX, **kw)


class NMCH_EM_K1(NMCH):
    _method = _eng.METHOD_EM
    _title = "EXACT-METHOD"


class NMCH_EM_K1_MM(NMCH_EM_K1): _k1_kernel = True
class NMCH_EM_K2_MM(NMCH_EM_K1_MM): _k1_kernel = False
class NMCH_EM_K3_MM(NMCH_EM_K2_MM): pass


class NMCH_QE_K1_MM(NMCH):
    """Third method (no reference counterpart): Andersen's QE-M large-step scheme; native Philox tag only."""
    _method = _eng.METHOD_QE
    _title = "QUADRATIC-EXPONENTIAL"
