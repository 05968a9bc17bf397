"""World-size-2 gloo tests (CPU) of the N>1 host logic: shard bounds + the single allreduce of the moments.

The per-shard compute is played by the oracle here (no GPU in this container); the logic under test is
nmch_b200.distributed.shard_bounds / allreduce_moments -- the same functions the GPU path uses with NCCL."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from nmch_b200.distributed import allreduce_moments, shard_bounds


def test_shard_bounds_cover_exactly_once():
    for n in (8192, 1 << 20, (1 << 20) + 12345, 3 * 4096 + 1):
        for world in (1, 2, 3, 4, 8):
            if n // world < 4096 and world > 1:
                with pytest.raises(ValueError):
                    shard_bounds(n, 0, world)
                continue
            covered = 0
            for r in range(world):
                first, cnt = shard_bounds(n, r, world)
                assert first == covered and cnt > 0
                assert world == 1 or first % 4096 == 0
                covered += cnt
            assert covered == n
    with pytest.raises(ValueError):
        shard_bounds(100, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, N, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as o
    first, cnt = shard_bounds(n, rank, world)
    r = o.fe_run(o.Params(N=N), rng=o.RNG_PHILOX, first_path=first, n_paths=cnt, threads=2)
    local = np.array([r["sum"], r["sumsq"]], np.float64)
    glob = allreduce_moments(local.copy())
    # exploration payload: 2 * n_points doubles in one collective
    k = np.array([0.5, 2.0, 4.0], np.float32)
    sw = o.fe_sweep(o.Params(N=N), k, np.full(3, 0.1, np.float32), np.full(3, 0.3, np.float32), rng=o.RNG_PHILOX,
                    first_path=first, n_paths=cnt, threads=2)
    gsw = allreduce_moments(sw.reshape(-1).copy())
    np.save(os.path.join(out_dir, f"r{rank}.npy"), np.concatenate([glob, gsw]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_moment_allreduce_equals_single_rank(tmp_path):
    from oracle import oracle as o
    n, N, world = 8192 + 777, 20, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, N, str(tmp_path)), nprocs=world, join=True)
    whole = o.fe_run(o.Params(N=N), rng=o.RNG_PHILOX, n_paths=n)
    k = np.array([0.5, 2.0, 4.0], np.float32)
    sw = o.fe_sweep(o.Params(N=N), k, np.full(3, 0.1, np.float32), np.full(3, 0.3, np.float32), rng=o.RNG_PHILOX, n_paths=n)
    for r in range(world):
        got = np.load(tmp_path / f"r{r}.npy")
        np.testing.assert_allclose(got[:2], [whole["sum"], whole["sumsq"]], rtol=1e-12)
        np.testing.assert_allclose(got[2:], sw.reshape(-1), rtol=1e-12)
