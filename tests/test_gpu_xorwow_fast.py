"""Opt-in NMCH_RNG_XORWOW_FAST FE mode through the C ABI: the reference's default stream (cuRAND XORWOW: same integer
draws per path as its CUDA build on the same seed) pushed through the native fast-math step.  Checkers: the oracle's
XORWOW paths (IEEE transforms) per path, and the UNMODIFIED reference CUDA build's results on a B200
(tests/golden/ref_cuda_b200.json) to the north-star tolerance of the XORWOW-compatible mode (1e-5 relative on price
and variance)."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as o

pytestmark = pytest.mark.gpu
FAST = 5
GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_cuda_b200.json")))
PKEYS = ("T", "S_0", "v_0", "r", "k", "rho", "theta", "sigma")
FE_XORWOW = [c for c in GOLD["cases"] if c["flags"]["method"] == "fe" and c["flags"]["rng"] == "xorwow"]


def _rel(a, b):
    return abs(a - b) / abs(b)


@pytest.mark.parametrize("floor", [0, 1])
@pytest.mark.parametrize("N", [1, 4, 5, 7, 100])
def test_paths_track_the_oracle_xorwow_stream(N, floor):
    from nmch_b200 import engine as E
    n = 4096 + 77
    with E.Engine(NTPB=1, NB=1, n_paths=n, N=N, rng=FAST, floor=floor) as e:
        e.init(1234)
        S, V, m = e.compute_paths()
        S2, V2, _ = e.compute_paths()                       # the stream continues (state written back)
    one = o.fe_run(o.Params(N=N), rng=o.RNG_XORWOW, floor=floor, n_paths=n, want_paths=True)
    two = o.fe_run(o.Params(N=N), rng=o.RNG_XORWOW, floor=floor, n_paths=n, calls=2, want_paths=True)
    # same draws, approximate transforms (23-bit uniforms, MUFU): 1e-6 per step, random walk over N steps
    for got, want in ((S, one["S"]), (S2, two["S"])):
        err = np.abs(got - want) / np.abs(want)
        assert np.median(err) < 2e-6 * max(1.0, np.sqrt(N)), np.median(err)
        assert (err < 2e-3).mean() > 0.999 and err.max() < 5e-2, (err.max(), (err < 2e-3).mean())
    np.testing.assert_allclose(V, one["V"], rtol=5e-2, atol=2e-4)
    assert abs(m.mean - one["mean"]) < 1e-5 * max(one["mean"], 1e-3) + 2e-7


@pytest.mark.parametrize("case", FE_XORWOW, ids=lambda c: "{NTPB}x{NB}-N{N}".format(**c["flags"]))
def test_prices_match_the_reference_cuda_build_on_identical_seeds(case):
    from nmch_b200 import engine as E
    f = case["flags"]
    kw = {k: f[k] for k in PKEYS if k in f}
    with E.Engine(NTPB=f["NTPB"], NB=f["NB"], N=f["N"], rng=FAST, **kw) as e:
        e.init(1234)
        for want in case["calls"]:                           # consecutive compute() calls on continued streams
            m = e.compute()
            var_ref = want["E2"] - want["E"] ** 2
            print("rel diff vs reference CUDA build:", f["NTPB"], f["NB"], f["N"], _rel(m.mean, want["E"]), _rel(m.variance, var_ref))
            assert _rel(m.mean, want["E"]) < 1e-5 + 4 * want["E_spread"] / want["E"], (m.mean, want)
            assert _rel(m.variance, var_ref) < 1e-5 + 4 * want["E2_spread"] / var_ref, (m.variance, var_ref)


def test_sweep_matches_the_reference_and_equals_sequential_computes():
    from nmch_b200 import engine as E
    sw = [s for s in GOLD["sweeps"] if s["flags"]["method"] == "fe"][0]
    f = sw["flags"]
    k, th, sg = (np.array(x, np.float32) for x in zip(*sw["points"]))
    with E.Engine(NTPB=f["NTPB"], NB=f["NB"], N=f["N"], rng=FAST) as e:
        e.init(1234)
        got = e.explore(k, th, sg)                           # one launch, points walked in the reference's order
    for m, want in zip(got, sw["calls"]):
        assert _rel(m.mean, want["E"]) < 2e-5, (m.mean, want["E"])
    with E.Engine(NTPB=f["NTPB"], NB=f["NB"], N=f["N"], rng=FAST) as e:
        e.init(1234)
        for i, m in enumerate(got):
            e.set_params(float(k[i]), float(th[i]), float(sg[i]))
            s = e.compute()
            assert s.sum_payoff == m.sum_payoff and s.sum_payoff_sq == m.sum_payoff_sq


def test_shards_strikes_and_argument_checks():
    from nmch_b200 import capi
    from nmch_b200 import engine as E
    n, N = 1 << 15, 60
    with E.Engine(NTPB=512, NB=n // 512, N=N, rng=FAST) as e:
        e.init(9)
        S, V, whole = e.compute_paths()
    parts = []
    for g in range(2):                                        # path index = subsequence: shards reproduce the paths
        with E.Engine(NTPB=512, NB=n // 512, N=N, rng=FAST, first_path=g * n // 2 + (13 if g else 0),
                      n_local=n // 2 - (13 if g else 0)) as e:
            e.init(9)
            parts.append(e.compute_paths())
    np.testing.assert_array_equal(parts[0][0], S[: n // 2])
    np.testing.assert_array_equal(parts[1][0], S[n // 2 + 13:])
    with E.Engine(NTPB=512, NB=n // 512, N=N, rng=FAST) as e:
        e.init(9)
        res = e.compute_strikes(np.array([0.9, 1.0, 1.1], np.float32))
    assert abs(res[1]["moments"].sum_payoff - whole.sum_payoff) < 1e-9 * n
    with pytest.raises(capi.NmchError):
        E.Engine(NTPB=32, NB=4, N=10, rng=FAST, method=E.METHOD_EM)


def test_full_size_price_and_speed_class():
    from nmch_b200 import engine as E
    n = 1 << 22
    with E.Engine(NTPB=512, NB=n // 512, N=1000, rng=FAST) as e:
        e.init(1234)
        m = e.compute()
        m = e.compute()
    assert abs(m.mean - o.heston_call()) < 3.5 * m.std_error + 1e-4     # Euler bias is O(dt) = +1e-4 here
    with E.Engine(NTPB=512, NB=n // 512, N=1000, rng=E.RNG_XORWOW_COMPAT) as e:
        e.init(1234)
        e.compute()
        c = e.compute()
    assert _rel(m.mean, c.mean) < 1e-5 and _rel(m.variance, c.variance) < 1e-5
    assert m.exec_ms < 0.7 * c.exec_ms                                   # same draws, well under the IEEE path's time
