"""Strike vector with pathwise delta AND vega ("pathwise delta/vega"), through the C ABI.

Checkers: the oracle's tangent restatement on the same Philox draws (oracle.fe_tangent_run: the reference's Euler
step, NMCH_FE.cu:156-163, differentiated in v_0 and carried in double), bump-and-revalue with the product's own
compute() on the same streams, and the semi-analytic d price / d v_0."""
import numpy as np
import pytest

from oracle import oracle as o

pytestmark = pytest.mark.gpu

STRIKES = np.array([0.7, 0.9, 1.0, 1.05, 1.3], np.float32)


def test_greeks_pass_reproduces_the_strike_pass_bit_for_bit():
    """Same words, same instructions for (S, V): payoff / delta / in-the-money sums equal compute_strikes() exactly."""
    from nmch_b200 import engine as E
    n, N = 8192 + 37, 60
    for floor in (0, 1):
        with E.Engine(NTPB=1, NB=1, n_paths=n, N=N, floor=floor) as e:
            e.init(1234)
            a = e.compute_strikes(STRIKES)
        with E.Engine(NTPB=1, NB=1, n_paths=n, N=N, floor=floor) as e:
            e.init(1234)
            b = e.compute_greeks(STRIKES)
            after = e.compute()
        with E.Engine(NTPB=1, NB=1, n_paths=n, N=N, floor=floor) as e:
            e.init(1234)
            e.compute()
            second = e.compute()
        for x, y in zip(a, b):
            assert x["moments"].sum_payoff == y["moments"].sum_payoff
            assert x["moments"].sum_payoff_sq == y["moments"].sum_payoff_sq
            assert x["delta"] == y["delta"] and x["itm"] == y["itm"]
        assert after.sum_payoff == second.sum_payoff          # the stream advanced like one compute()


@pytest.mark.parametrize("floor,kw", [(0, {}), (1, dict(k=2.08, theta=0.108, sigma=1.0))])
def test_vega_sums_match_the_oracle_tangent_on_the_same_draws(floor, kw):
    from nmch_b200 import engine as E
    n, N = 16384 + 5, 60
    with E.Engine(NTPB=1, NB=1, n_paths=n, N=N, floor=floor, **kw) as e:
        e.init(1234)
        res = e.compute_greeks(STRIKES)
    ref = o.fe_tangent_run(o.Params(N=N, **kw), floor=floor, n_paths=n)
    for r, K in zip(res, STRIKES):
        est = np.where(ref["S"] > np.float32(K), ref["B"], 0.0)
        # fast-math transforms against IEEE ones on identical uniforms: per-path differences of 1e-5 relative, a few
        # paths within that of the strike change sides (each moves the mean by |B| / n)
        tol = 2e-4 * np.abs(est).mean() + 4 * np.abs(ref["B"]).max() / n + 1e-6
        assert abs(r["vega_v0"] - est.mean()) < tol, (K, r["vega_v0"], est.mean(), tol)
        se = est.std() / np.sqrt(n)
        assert abs(r["vega_v0_se"] - se) < 0.02 * se + 1e-9


@pytest.mark.parametrize("floor,kw", [(0, {}), (1, dict(k=2.08, theta=0.108, sigma=1.0))])
def test_vega_matches_bump_and_revalue_and_the_semi_analytic_sensitivity(floor, kw):
    from nmch_b200 import engine as E
    n, N, h = 1 << 20, 250, 1e-3
    common = dict(NTPB=512, NB=n // 512, N=N, floor=floor, **kw)
    with E.Engine(**common) as e:
        e.init(7)
        res = e.compute_greeks(STRIKES)
    prices = {}
    for sgn in (+1, -1):
        with E.Engine(v_0=0.1 + sgn * h, **common) as e:
            e.init(7)
            prices[sgn] = e.compute_strikes(STRIKES)
    akw = {("kappa" if a == "k" else a): b for a, b in kw.items()}
    for j, (r, K) in enumerate(zip(res, STRIKES)):
        bump = (prices[+1][j]["moments"].mean - prices[-1][j]["moments"].mean) / (2 * h)
        se = r["vega_v0_se"]
        assert se > 0
        # same streams: the two estimators differ by the O(h^2) curvature and the paths that cross the floor's kink
        assert abs(r["vega_v0"] - bump) < 2 * se + 2e-3, (K, r["vega_v0"], bump, se)
        analytic = (o.heston_call(K=float(K), v0=0.1 + h, **akw) - o.heston_call(K=float(K), v0=0.1 - h, **akw)) / (2 * h)
        bias = 0.01 if floor == 0 else 0.04                   # Euler scheme bias at N = 250 (larger beyond Feller)
        assert abs(r["vega_v0"] - analytic) < 4 * se + bias, (K, r["vega_v0"], analytic, se)


def test_greeks_are_refused_outside_the_native_fe_stream():
    from nmch_b200 import capi
    from nmch_b200 import engine as E
    for kw in (dict(method=1), dict(rng=E.RNG_XORWOW_COMPAT), dict(rng=E.RNG_PHILOX_DENSE)):
        with E.Engine(NTPB=32, NB=4, N=10, **kw) as e:
            e.init(1)
            with pytest.raises(capi.NmchError):
                e.compute_greeks([1.0])
            e.compute()                                        # the engine stays usable
    with E.Engine(NTPB=32, NB=4, N=10) as e:
        e.init(1)
        with pytest.raises(capi.NmchError):
            e.compute_greeks(np.zeros(65, np.float32))
        assert len(e.compute_greeks([1.0])) == 1


def test_greeks_resume_in_the_middle_of_a_philox_block():
    """An odd N leaves the stream half way through a block: the next pass starts on the block's second word pair."""
    from nmch_b200 import engine as E
    n, N = 4096, 31
    with E.Engine(NTPB=1, NB=1, n_paths=n, N=N) as e:
        e.init(5)
        e.compute()
        a = e.compute_strikes(STRIKES)
    with E.Engine(NTPB=1, NB=1, n_paths=n, N=N) as e:
        e.init(5)
        e.compute()
        b = e.compute_greeks(STRIKES)
    for x, y in zip(a, b):
        assert x["moments"].sum_payoff == y["moments"].sum_payoff and x["delta"] == y["delta"]


def test_group_shard_and_cli_greeks():
    import os
    import subprocess
    from nmch_b200 import Engine, Group
    kw = dict(NTPB=512, NB=64, N=100)
    with Engine(**kw) as e:
        e.init(1234)
        a = e.compute_greeks(STRIKES)
    with Group(1, **kw) as g:
        g.init(1234)
        b = g.compute_greeks(STRIKES)
    for x, y in zip(a, b):
        assert x["vega_v0"] == y["vega_v0"] and x["delta"] == y["delta"]
    # two shards of the same global path range add up to the whole (what the group's allreduce sums)
    n = 512 * 64
    parts = []
    for first in (0, n // 2):
        with Engine(n_paths=n, first_path=first, n_local=n // 2, **kw) as e:
            e.init(1234)
            parts.append(e.compute_greeks(STRIKES))
    for x, p0, p1 in zip(a, parts[0], parts[1]):
        np.testing.assert_allclose(x["vega_v0"], 0.5 * (p0["vega_v0"] + p1["vega_v0"]), rtol=1e-12)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([os.path.join(root, "bin", "NMCH"), "--NB", "64", "--N", "100", "--strikes", "0.9,1.0,1.1", "--vega"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    rows = [l.split(", ") for l in r.stdout.splitlines() if l[:1].isdigit() and l.count(",") == 5]
    assert len(rows) == 3
    with Engine(**kw) as e:                                   # the CLI reports compute() first, then this pass
        e.init(1234)
        e.compute()
        second = e.compute_greeks(STRIKES)
    assert abs(float(rows[1][4]) - second[2]["vega_v0"]) < 1e-5 and float(rows[1][5]) > 0
    r = subprocess.run([os.path.join(root, "bin", "NMCH"), "--method", "em", "--NB", "8", "--N", "10", "--strikes", "1.0", "--vega"],
                       capture_output=True, text=True)
    assert r.returncode != 0                                  # the engine error convention: message + exit(EXIT_FAILURE)


def test_group_greeks_two_gpus():
    import torch
    from nmch_b200 import Engine, Group
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    kw = dict(NTPB=512, NB=256, N=100)
    with Engine(**kw) as e:
        e.init(1234)
        a = e.compute_greeks(STRIKES)
    with Group(2, **kw) as g:
        g.init(1234)
        b = g.compute_greeks(STRIKES)
    for x, y in zip(a, b):
        np.testing.assert_allclose(x["moments"].sum_payoff, y["moments"].sum_payoff, rtol=1e-12)
        np.testing.assert_allclose(x["vega_v0"], y["vega_v0"], rtol=1e-12)
