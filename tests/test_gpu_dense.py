"""Opt-in dense-draw FE stream mode (NMCH_RNG_PHILOX_DENSE: three (23-bit radius, 19-bit angle) draws per Philox
block), through the C ABI.  It is NOT word-compatible with cuRAND's per-step layout, so the checkers are: a
restatement of its own mapping in the oracle (per path), exact moment identities of the Euler scheme (they isolate
the quality of the normals), and the semi-analytic price."""
import numpy as np
import pytest

from oracle import oracle as o

pytestmark = pytest.mark.gpu
DENSE = 4


def engine(n, N, **kw):
    from nmch_b200 import engine as E
    ntpb = min(n, 512)
    return E.Engine(NTPB=ntpb, NB=max(1, n // ntpb), N=N, rng=DENSE, n_paths=n, **kw)


@pytest.mark.parametrize("N", [1, 2, 3, 100, 101])
@pytest.mark.parametrize("floor", [0, 1])
def test_per_path_matches_restatement(N, floor):
    n = 4096 + 5
    with engine(n, N, floor=floor) as e:
        e.init(1234)
        S, V, m = e.compute_paths()
        S2, V2, _ = e.compute_paths()                 # second call resumes inside a block when N % 3 != 0
    one = o.fe_run(o.Params(N=N), rng=o.RNG_PHILOX_DENSE, floor=floor, n_paths=n, want_paths=True)
    two = o.fe_run(o.Params(N=N), rng=o.RNG_PHILOX_DENSE, floor=floor, n_paths=n, calls=2, want_paths=True)
    np.testing.assert_allclose(S, one["S"], rtol=2e-3, atol=2e-4)
    np.testing.assert_allclose(V, one["V"], rtol=1e-2, atol=5e-4)
    np.testing.assert_allclose(S2, two["S"], rtol=2e-3, atol=2e-4)
    assert abs(m.mean - one["mean"]) < 0.1 * m.std_error + 1e-5


def test_exact_moment_identities_of_the_euler_scheme():
    # sigma = 0, k = 0: V stays v0, S_T = prod(1 + sqrt(v0 dt) Z_i) with Z_i iid N(0,1):
    #   E[S_T] = 1 and E[S_T^2] = (1 + v0 dt)^N exactly -- mean, variance and independence of the generated normals
    n, N, v0 = 1 << 22, 60, 0.1
    with engine(n, N, k=0.0, theta=0.1, sigma=0.0, v_0=v0) as e:
        e.init(99)
        S, V, m = e.compute_paths()
    S = S.astype(np.float64)
    assert np.allclose(V, v0)
    assert abs(S.mean() - 1.0) < 4 * S.std() / np.sqrt(n)
    want2 = (1.0 + v0 / N) ** N
    s2 = S * S
    assert abs(s2.mean() - want2) < 4 * s2.std() / np.sqrt(n), (s2.mean(), want2)
    # fourth moment of the one-step return pins the kurtosis of the normals: E[(1 + aZ)^4] = 1 + 6 a^2 + 3 a^4
    with engine(n, 1, k=0.0, theta=0.1, sigma=0.0, v_0=v0, T=1.0) as e:
        e.init(5)
        S1 = e.compute_paths()[0].astype(np.float64)
    a2 = v0
    want4 = 1 + 6 * a2 + 3 * a2 * a2
    s4 = S1 ** 4
    assert abs(s4.mean() - want4) < 4 * s4.std() / np.sqrt(n), (s4.mean(), want4)


def test_price_matches_semi_analytic_and_other_modes():
    from nmch_b200 import engine as E
    n = 1 << 22
    with engine(n, 1000) as e:
        e.init(1234)
        m = e.compute()
    assert abs(m.mean - o.heston_call()) < 3 * m.std_error + 1e-4
    with E.Engine(NTPB=512, NB=n // 512, N=1000, rng=E.RNG_PHILOX) as e:
        e.init(1234)
        w = e.compute()
    assert abs(m.mean - w.mean) < 3 * np.hypot(m.std_error, w.std_error)          # independent draws: 3 SE
    assert abs(m.variance - w.variance) < 5e-3 * w.variance


def test_explore_seek_shards_and_layout_independence():
    from nmch_b200 import capi
    from nmch_b200 import engine as E
    n, N = 1 << 15, 50
    k, th, sg = o.exploration_grid(5, True)
    with engine(n, N) as e:
        e.init(7)
        ex = e.explore(k[:4], th[:4], sg[:4])
    with engine(n, N) as e:
        e.init(7)
        for i in range(4):
            e.set_params(float(k[i]), float(th[i]), float(sg[i]))
            s = e.compute()
            assert s.sum_payoff == ex[i].sum_payoff and s.sum_payoff_sq == ex[i].sum_payoff_sq
    with engine(n, N) as e:
        e.init(7)
        whole = e.compute_paths()
    halves = []
    for g in range(2):
        with engine(n, N, first_path=g * n // 2, n_local=n // 2) as e:
            e.init(7)
            halves.append(e.compute())
    assert abs(halves[0].sum_payoff + halves[1].sum_payoff - whole[2].sum_payoff) < 1e-9 * n
    for P in (1, 2, 4):
        with engine(n, N, paths_per_thread=P) as e:
            e.init(7)
            np.testing.assert_array_equal(e.compute_paths()[0], whole[0])
    with engine(4096, 10) as e:                                   # seek: position in logical draws, two per step
        e.init(7)
        e.seek(2 * 7)
        S = e.compute_paths()[0]
    ref = o.fe_run_at(o.Params(N=10), 14, rng=o.RNG_PHILOX_DENSE, seed=7, n_paths=4096, want_paths=True)
    np.testing.assert_allclose(S, ref["S"], rtol=2e-3, atol=2e-4)
    with pytest.raises(capi.NmchError):
        E.Engine(NTPB=32, NB=4, N=10, rng=DENSE, method=E.METHOD_EM)


def test_explore_beyond_49152_points_with_N_mod_3_equal_2():
    """the block index of point p is (r0 + p * (N % 3)) / 3; formed with a 32-bit multiply it wrapped from
    t0 = 98304 on, i.e. for N % 3 == 2 and p >= 49152.  Compare late points of a 60000-point sweep with compute()
    calls positioned at the same stream offset."""
    n, N, n_points = 4096, 5, 60000
    rng = np.random.default_rng(3)
    k = rng.uniform(0.1, 5.0, n_points).astype(np.float32)
    th = rng.uniform(0.01, 0.5, n_points).astype(np.float32)
    sg = rng.uniform(0.1, 1.0, n_points).astype(np.float32)
    with engine(n, N) as e:
        e.init(11)
        ex = e.explore(k, th, sg)
    for p in (0, 49151, 49152, 49153, 55001, 59999):
        with engine(n, N) as e:
            e.init(11)
            e.seek(2 * N * p)
            e.set_params(float(k[p]), float(th[p]), float(sg[p]))
            s = e.compute()
        assert abs(s.sum_payoff - ex[p].sum_payoff) <= 1e-12 * abs(s.sum_payoff) + 1e-300, p
        assert abs(s.sum_payoff_sq - ex[p].sum_payoff_sq) <= 1e-12 * abs(s.sum_payoff_sq) + 1e-300, p
