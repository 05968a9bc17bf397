"""Single-process multi-GPU group (nmch_group_*): sharding + one NCCL allreduce == one engine on all paths."""
import numpy as np
import pytest

from oracle import oracle as o

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("method", [0, 1])
def test_group_of_one_equals_engine(method):
    from nmch_b200 import Engine, Group
    kw = dict(NTPB=512, NB=64, N=100, method=method)
    with Engine(**kw) as e:
        e.init(1234)
        a = e.compute()
    with Group(1, **kw) as g:
        g.init(1234)
        b = g.compute()
        assert g.size == 1
    assert a.sum_payoff == b.sum_payoff and a.sum_payoff_sq == b.sum_payoff_sq and b.n_paths == 512 * 64


@pytest.mark.parametrize("method,rng", [(0, 0), (0, 1), (1, 0), (0, 4), (0, 5)])
def test_group_shards_add_up(method, rng):
    from nmch_b200 import Engine, Group
    n = _ngpu()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    kw = dict(NTPB=512, NB=256, N=100, method=method, rng=rng)
    with Engine(**kw) as e:
        e.init(1234)
        whole = e.compute()
        k, th, sg = o.exploration_grid(5, True)
        ex1 = e.explore(k[:5], th[:5], sg[:5])
    for G in sorted({2, n}):
        with Group(G, **kw) as g:
            g.init(1234)
            m = g.compute()
            ex = g.explore(k[:5], th[:5], sg[:5])
        assert m.n_paths == whole.n_paths
        np.testing.assert_allclose([m.sum_payoff, m.sum_payoff_sq], [whole.sum_payoff, whole.sum_payoff_sq], rtol=1e-12)
        for a, b in zip(ex, ex1):
            np.testing.assert_allclose([a.sum_payoff, a.sum_payoff_sq], [b.sum_payoff, b.sum_payoff_sq], rtol=1e-12)


def test_cpp_cli_gpus_flag():
    import json
    import os
    import subprocess
    n = _ngpu()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "bin", "NMCH")
    one = json.loads(subprocess.run([exe, "--NB", "256", "--N", "100", "--json"], capture_output=True, text=True).stdout.splitlines()[-1])
    two = json.loads(subprocess.run([exe, "--NB", "256", "--N", "100", "--json", "--gpus", "2"], capture_output=True, text=True).stdout.splitlines()[-1])
    assert abs(one["sum_payoff"] - two["sum_payoff"]) < 1e-9 * one["n_paths"]
