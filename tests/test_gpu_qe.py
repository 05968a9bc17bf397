"""QE-M large-step scheme (no reference counterpart), through the C ABI.

Checkers: the semi-analytic Heston price (the scheme's accuracy at LARGE steps is the point) and the oracle's
restatement of the same scheme on the same Philox draws (per-path)."""
import numpy as np
import pytest

from oracle import oracle as o

pytestmark = pytest.mark.gpu


def qe_engine(n, N, **kw):
    from nmch_b200 import engine as E
    return E.Engine(NTPB=512, NB=n // 512, N=N, method=E.METHOD_QE, **kw)


@pytest.mark.parametrize("kw", [dict(), dict(k=2.08, theta=0.108, sigma=1.0), dict(k=10.0, theta=0.5, sigma=1.0),
                                dict(T=0.5, S_0=2.0, v_0=0.04, r=0.03, k=1.5, rho=-0.3, theta=0.09, sigma=0.5)])
def test_per_path_matches_oracle_restatement(kw):
    n, N = 4096, 40
    with qe_engine(n, N, **kw) as e:
        e.init(77)
        S, V, m = e.compute_paths()
    ref = o.qe_run(o.Params(N=N, **kw), seed=77, n_paths=n, want_paths=True)
    np.testing.assert_allclose(S, ref["S"], rtol=3e-3, atol=3e-4)
    np.testing.assert_allclose(V, ref["V"], rtol=2e-2, atol=2e-4)
    assert abs(m.mean - ref["mean"]) < 0.05 * m.std_error + 1e-5


@pytest.mark.parametrize("N", [10, 50, 100])
@pytest.mark.parametrize("k,theta,sigma", [(0.5, 0.1, 0.3), (2.08, 0.108, 1.0), (0.1, 0.5, 1.0), (10.0, 0.01, 0.1)])
def test_large_steps_match_semi_analytic_price(N, k, theta, sigma):
    n = 1 << 20
    with qe_engine(n, N, k=k, theta=theta, sigma=sigma) as e:
        e.init(5)
        S, V, m = e.compute_paths()
    want = o.heston_call(kappa=k, theta=theta, sigma=sigma)
    bias = 6e-4 if N == 10 else 2e-4                       # QE's discretisation bias at dt = 0.1 / <= 0.02
    assert abs(m.mean - want) < 3.5 * m.std_error + bias, (N, m.mean, want, m.std_error)
    S64 = S.astype(np.float64)
    assert abs(S64.mean() - 1.0) < 4 * S64.std() / np.sqrt(n) + 1e-4      # martingale correction: E[S_T] = S_0


def test_streams_shards_explore_and_rng_restriction():
    from nmch_b200 import capi
    from nmch_b200 import engine as E
    n, N = 1 << 15, 50
    with qe_engine(n, N) as e:
        e.init(9)
        a, b = e.compute(), e.compute()
    assert a.sum_payoff != b.sum_payoff
    halves = []
    for g in range(2):
        with qe_engine(n, N, first_path=g * n // 2, n_local=n // 2) as e:
            e.init(9)
            halves.append(e.compute())
    assert abs(halves[0].sum_payoff + halves[1].sum_payoff - a.sum_payoff) < 1e-9 * n
    k, th, sg = o.exploration_grid(5, True)
    with qe_engine(n, N) as e:
        e.init(9)
        ex = e.explore(k[:4], th[:4], sg[:4])
    with qe_engine(n, N) as e:
        e.init(9)
        for i in range(4):
            e.set_params(float(k[i]), float(th[i]), float(sg[i]))
            assert e.compute().sum_payoff == ex[i].sum_payoff
    with pytest.raises(capi.NmchError):
        E.Engine(NTPB=512, NB=8, N=50, method=E.METHOD_QE, rng=E.RNG_XORWOW_COMPAT)


def test_cli_method_qe():
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([os.path.join(root, "bin", "NMCH"), "--method", "qe", "--N", "50", "--NB", "2048", "--json"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "METHOD: QUADRATIC-EXPONENTIAL" in r.stdout and "The true price 0.119733" in r.stdout
    import json
    j = json.loads(r.stdout.splitlines()[-1])
    assert abs(j["E"] - 0.1197325) < 4 * j["std_error"] + 2e-4
