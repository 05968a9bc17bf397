"""Strike vector + pathwise delta in the same pass, through the C ABI.

Checker: the oracle's terminal prices for the SAME paths (compat stream), folded on the host in FP64, and the
semi-analytic Heston price / delta (central difference of the oracle's pricer in S_0) for the native streams."""
import numpy as np
import pytest

from oracle import oracle as o

pytestmark = pytest.mark.gpu

STRIKES = np.array([0.7, 0.9, 1.0, 1.05, 1.3], np.float32)


def _host_fold(S, K, S0):
    S = S.astype(np.float64)
    pay = np.maximum(S.astype(np.float32) - np.float32(K), 0).astype(np.float64)
    itm = S.astype(np.float32) > np.float32(K)
    return pay.sum(), (pay * pay).sum(), (S[itm].astype(np.float32) * np.float32(1.0 / S0)).astype(np.float64).sum(), itm.sum()


@pytest.mark.parametrize("method", [0, 1])
@pytest.mark.parametrize("rng", [0, 1])
def test_strike_sums_equal_host_fold_of_the_same_paths(method, rng):
    from nmch_b200 import engine as E
    n, N = 4096 + 37, 60
    with E.Engine(NTPB=1, NB=1, n_paths=n, N=N, method=method, rng=rng) as e:
        e.init(1234)
        S, V, m0 = e.compute_paths()
    with E.Engine(NTPB=1, NB=1, n_paths=n, N=N, method=method, rng=rng) as e:
        e.init(1234)
        res = e.compute_strikes(STRIKES)
    for r, K in zip(res, STRIKES):
        pay, pay2, dl, itm = _host_fold(S, K, 1.0)
        assert abs(r["moments"].sum_payoff - pay) < 1e-9 * n
        assert abs(r["moments"].sum_payoff_sq - pay2) < 1e-9 * n
        assert abs(r["delta"] * n - dl) < 1e-6 * n and round(r["itm"] * n) == itm
    # the at-the-money entry is the plain compute() result
    atm = res[2]["moments"]
    assert abs(atm.sum_payoff - m0.sum_payoff) < 1e-9 * n


def test_compat_strike_sums_match_oracle_paths():
    from nmch_b200 import engine as E
    n, N = 8192, 100
    with E.Engine(NTPB=512, NB=16, N=N, rng=E.RNG_XORWOW_COMPAT) as e:
        e.init(1234)
        res = e.compute_strikes(STRIKES)
    ref = o.fe_run(o.Params(N=N), rng=o.RNG_XORWOW, n_paths=n, want_paths=True)
    for r, K in zip(res, STRIKES):
        pay, pay2, dl, itm = _host_fold(ref["S"], K, 1.0)
        assert abs(r["moments"].sum_payoff - pay) < 3e-5 * n
        assert abs(r["delta"] * n - dl) < 2e-3 * n          # a path within 1e-6 of the strike may flip sides


@pytest.mark.parametrize("method", [0, 1])
def test_native_prices_and_deltas_match_semi_analytic(method):
    from nmch_b200 import engine as E
    n = 1 << 21
    with E.Engine(NTPB=512, NB=n // 512, N=500, method=method) as e:
        e.init(11)
        res = e.compute_strikes(STRIKES)
        again = e.compute()                                   # streams advanced like a compute() call
    h = 1e-4
    for r, K in zip(res, STRIKES):
        m = r["moments"]
        price = o.heston_call(K=float(K))
        delta = (o.heston_call(S0=1 + h, K=float(K)) - o.heston_call(S0=1 - h, K=float(K))) / (2 * h)
        bias = 3e-4 if method == 0 else 1e-4                   # Euler scheme bias at N = 500
        assert abs(m.mean - price) < 3.5 * m.std_error + bias, (K, m.mean, price)
        assert abs(r["delta"] - delta) < 3.5 * 0.6 / np.sqrt(n) + 2e-3, (K, r["delta"], delta)
    assert again.sum_payoff != res[2]["moments"].sum_payoff


def test_strike_argument_validation():
    from nmch_b200 import capi
    from nmch_b200 import engine as E
    with E.Engine(NTPB=32, NB=4, N=10) as e:
        e.init(1)
        with pytest.raises(capi.NmchError):
            e.compute_strikes(np.zeros(65, np.float32))
        assert len(e.compute_strikes([1.0])) == 1


def test_group_and_cli_strikes():
    import os
    import subprocess
    from nmch_b200 import Engine, Group
    kw = dict(NTPB=512, NB=64, N=100)
    with Engine(**kw) as e:
        e.init(1234)
        a = e.compute_strikes(STRIKES)
    with Group(1, **kw) as g:
        g.init(1234)
        b = g.compute_strikes(STRIKES)
    for x, y in zip(a, b):
        assert x["moments"].sum_payoff == y["moments"].sum_payoff and x["delta"] == y["delta"]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([os.path.join(root, "bin", "NMCH"), "--NB", "64", "--N", "100", "--strikes", "0.9,1.0,1.1"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    rows = [l.split(", ") for l in r.stdout.splitlines() if l[:1].isdigit() and l.count(",") == 3]
    assert len(rows) == 3 and abs(float(rows[1][0]) - 1.0) < 1e-6
    assert float(rows[0][1]) > float(rows[1][1]) > float(rows[2][1]) > 0          # call price falls with the strike
    assert 1 > float(rows[0][3]) > float(rows[1][3]) > float(rows[2][3]) > 0      # so does the delta


def test_group_strikes_two_gpus():
    import torch
    from nmch_b200 import Engine, Group
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    kw = dict(NTPB=512, NB=256, N=100)
    with Engine(**kw) as e:
        e.init(1234)
        a = e.compute_strikes(STRIKES)
    with Group(2, **kw) as g:
        g.init(1234)
        b = g.compute_strikes(STRIKES)
    for x, y in zip(a, b):
        np.testing.assert_allclose(x["moments"].sum_payoff, y["moments"].sum_payoff, rtol=1e-12)
        np.testing.assert_allclose(x["delta"], y["delta"], rtol=1e-12)
