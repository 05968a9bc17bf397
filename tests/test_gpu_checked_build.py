"""Memory-safety evidence in place of compute-sanitizer (closed on the B200 pool): the CHECKED build of the library
(nmch_b200/libnmch_b200_checked.so = the same sources with -DNMCHB_CHECKS: device-side asserts on every index the
kernels form, guard bands around every device buffer swept after each blocking call) runs the cases where an
indexing mistake would show: ragged path counts, 64-bit path indices, shards, several tiles per block, many-point
sweeps, every stream mode and method, the strike and greeks passes, the single-process group.  A failed assert or an overwritten
guard band comes back as NMCH_ERR_CUDA and fails the case.  The sweep itself is proved by a planted overrun."""
import json
import os
import subprocess
import sys
import textwrap

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASES = textwrap.dedent(r"""
    import json, sys
    import numpy as np
    sys.path.insert(0, %r)
    from nmch_b200 import capi, engine as E
    L = capi.load()
    assert L.nmch_checked_build() == 1, "the checked library is not the one loaded"
    done = []

    def case(name, fn):
        fn()
        done.append(name)

    def selftest():
        with E.Engine(NTPB=128, NB=4, N=5) as e:
            e.init(1)
            e.compute()
            capi.check(L.nmch_checked_selftest(e._h))
    case("guard sweep catches a planted overrun", selftest)

    grid_k = np.linspace(0.1, 10.0, 37, dtype=np.float32)
    grid_t = np.linspace(0.01, 0.5, 37, dtype=np.float32)
    grid_s = np.linspace(0.1, 1.0, 37, dtype=np.float32)

    # FE, every stream mode: ragged sizes (not a multiple of the tile, of the block, of the warp), one path, 2 calls
    for rng in (E.RNG_PHILOX, E.RNG_XORWOW_COMPAT, E.RNG_PHILOX_COMPAT, E.RNG_MRG32K3A_COMPAT, E.RNG_PHILOX_DENSE, E.RNG_XORWOW_FAST):
        for n in (1, 31, 4096 + 5, 3 * 4096 + 1023):
            for ppt in ((0, 1, 2, 4, 8) if rng == E.RNG_PHILOX else (0,)):
                def run(rng=rng, n=n, ppt=ppt):
                    with E.Engine(NTPB=1, NB=1, N=7, rng=rng, n_paths=n, paths_per_thread=ppt) as e:
                        e.init(3)
                        e.compute(); e.compute()
                        S, V, m = e.compute_paths()
                        assert np.isfinite(S).all() and m.n_paths == n
                        e.check()
                case(f"fe rng={rng} n={n} ppt={ppt}", run)

    # 64-bit path indices: a shard that starts beyond 2^32 and one that straddles a multiple of 2^32
    for first in ((1 << 32) + 4096 * 3, (1 << 32) - 4096):
        for rng, method in ((E.RNG_PHILOX, E.METHOD_FE), (E.RNG_PHILOX_DENSE, E.METHOD_FE), (E.RNG_XORWOW_COMPAT, E.METHOD_FE),
                            (E.RNG_PHILOX, E.METHOD_EM), (E.RNG_PHILOX, E.METHOD_QE)):
            def run(first=first, rng=rng, method=method):
                with E.Engine(NTPB=1, NB=1, N=6, rng=rng, method=method, n_paths=1 << 33, first_path=first, n_local=8192 + 17) as e:
                    e.init(5)
                    e.compute()
                    e.compute_paths()
            case(f"64-bit first_path={first} rng={rng} method={method}", run)

    # sweeps: several tiles per block (tiles > 256 with more than one point), ragged, many points
    def sweep_tiles():
        with E.Engine(NTPB=1, NB=1, N=3, n_paths=300 * 1024 + 77) as e:     # 301 tiles of 1024 paths at P = 4
            e.init(7)
            out = e.explore(grid_k[:3], grid_t[:3], grid_s[:3])
            assert len(out) == 3 and all(np.isfinite(m.sum_payoff) for m in out)
    case("fe sweep, tiles_per_block > 1", sweep_tiles)
    for rng in (E.RNG_PHILOX, E.RNG_PHILOX_DENSE, E.RNG_XORWOW_COMPAT, E.RNG_XORWOW_FAST, E.RNG_PHILOX_COMPAT):
        def run(rng=rng):
            with E.Engine(NTPB=1, NB=1, N=5, rng=rng, n_paths=5120 + 3) as e:
                e.init(9)
                e.explore(grid_k, grid_t, grid_s)
                e.explore(grid_k[:2], grid_t[:2], grid_s[:2])              # the buffers shrink back to a smaller sweep
        case(f"fe sweep 37 points rng={rng}", run)
    def many_points():
        rng_ = np.random.default_rng(1)
        n_pts = 20000
        with E.Engine(NTPB=1, NB=1, N=2, n_paths=4096) as e:
            e.init(11)
            e.explore(rng_.uniform(0.1, 5, n_pts).astype(np.float32), rng_.uniform(0.01, 0.5, n_pts).astype(np.float32),
                      rng_.uniform(0.1, 1, n_pts).astype(np.float32))
    case("fe sweep 20000 points", many_points)

    # EM: the three samplers alone and mixed in one launch (kEmAny), compat streams, ragged
    em_pts = [(0.5, 0.1, 0.3), (10.0, 0.5, 1.0), (2.08, 0.108, 1.0), (0.1, 0.5, 1.0), (0.5, 0.1, 0.42)]
    for rng in (E.RNG_PHILOX, E.RNG_XORWOW_COMPAT, E.RNG_PHILOX_COMPAT, E.RNG_MRG32K3A_COMPAT):
        def run(rng=rng):
            with E.Engine(NTPB=1, NB=1, N=9, rng=rng, method=E.METHOD_EM, n_paths=2 * 256 + 3) as e:
                e.init(13)
                for k, t, s in em_pts:
                    e.set_params(k, t, s)
                    e.compute()
                a = np.array(em_pts, np.float32)
                e.explore(a[:, 0], a[:, 1], a[:, 2])
                e.compute_paths()
        case(f"em rng={rng}", run)

    # QE, strikes (ragged n: float4 body + tail), and a sharded pair that must add up
    def qe():
        with E.Engine(NTPB=1, NB=1, N=10, method=E.METHOD_QE, n_paths=4096 + 9) as e:
            e.init(2)
            e.compute()
            e.explore(grid_k[:5], grid_t[:5], grid_s[:5])
    case("qe", qe)
    for method in (E.METHOD_FE, E.METHOD_EM):
        def run(method=method):
            with E.Engine(NTPB=1, NB=1, N=8, method=method, n_paths=4096 * 2 + 3) as e:
                e.init(4)
                r = e.compute_strikes(np.linspace(0.7, 1.3, 64))
                assert len(r) == 64
        case(f"strikes method={method}", run)
    def greeks():
        for n, first in ((4096 * 2 + 3, 0), (1, 0), (8192 + 17, (1 << 32) - 4096)):
            for floor in (0, 1):
                with E.Engine(NTPB=1, NB=1, N=9, floor=floor, n_paths=(1 << 33) if first else n, first_path=first, n_local=n) as e:
                    e.init(4)
                    r = e.compute_greeks(np.linspace(0.7, 1.3, 64))
                    assert len(r) == 64 and all(np.isfinite(x["vega_v0"]) for x in r)
                    e.compute_greeks([1.0])
                    e.compute()
    case("greeks (tangent pass + fold), ragged and 64-bit shard", greeks)
    def shards():
        n = 3 * 4096 + 100
        whole = None
        with E.Engine(NTPB=1, NB=1, N=11, n_paths=n) as e:
            e.init(6)
            whole = e.compute()
        parts = []
        for first, cnt in ((0, 4096), (4096, 8192), (3 * 4096, 100)):
            with E.Engine(NTPB=1, NB=1, N=11, n_paths=n, first_path=first, n_local=cnt) as e:
                e.init(6)
                parts.append(e.compute())
        assert abs(sum(p.sum_payoff for p in parts) - whole.sum_payoff) < 1e-9 * n
    case("shards add up", shards)
    def group():
        with E.Group(1, NTPB=1, NB=1, N=5, n_paths=4096 * 2 + 11) as g:
            g.init(8)
            g.compute()
            g.explore(grid_k[:4], grid_t[:4], grid_s[:4])
            g.compute_strikes([0.9, 1.0, 1.1])
            g.compute_greeks([0.9, 1.0, 1.1])
        if L.nmch_device_count() > 1:
            with E.Group(2, NTPB=1, NB=1, N=5, n_paths=4096 * 4 + 11) as g:
                g.init(8)
                g.compute()
    case("group", group)
    print(json.dumps({"checked": 1, "cases": len(done)}))
""")


def test_checked_build_runs_the_indexing_cases_clean():
    from nmch_b200 import _build
    lib = _build.build_checked(only_if_missing=True)
    env = dict(os.environ, NMCH_B200_LIB=lib)
    r = subprocess.run([sys.executable, "-c", CASES % ROOT], capture_output=True, text=True, timeout=1500, env=env)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-3000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["checked"] == 1 and out["cases"] >= 60, out


def test_normal_build_reports_unchecked_and_check_is_a_sync():
    from nmch_b200 import capi, engine as E
    L = capi.load()
    if os.environ.get("NMCH_B200_LIB"):
        pytest.skip("a library override is active")
    assert L.nmch_checked_build() == 0
    with E.Engine(NTPB=128, NB=4, N=5) as e:
        e.init(1)
        e.compute()
        e.check()
        assert L.nmch_checked_selftest(e._h) == capi.ERR_STATE
