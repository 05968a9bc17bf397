"""One process per GPU: shard the path axis over ranks, one allreduce of the FP64 partial moments.

The hot path shards naturally: global path index == generator subsequence, so rank g of G
simulates paths [first, first + n_local) with no data-path collective; the only exchange is the sum of the
2 (or 2 * n_points) partial moments, done with torch.distributed (NCCL over NVLink on GPUs; gloo in the CPU
tests of this host logic).  The reference has no multi-GPU path at all.
"""
from __future__ import annotations

import numpy as np

ALIGN = 4096      # tile alignment of first_path in the Philox FE modes (include/nmch_b200.h)


def shard_bounds(n_paths: int, rank: int, world: int, align: int = ALIGN):
    """Paths of `rank`: equal multiples of `align`, the last rank takes the remainder."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    if world == 1:
        return 0, n_paths
    per = (n_paths // world) // align * align
    if per == 0:
        raise ValueError(f"need at least {align} paths per rank to shard {n_paths} paths over {world} ranks")
    first = per * rank
    return first, (n_paths - first) if rank == world - 1 else per


def allreduce_moments(local, group=None):
    """Sum raw moment arrays over ranks (in place for tensors); identity when not distributed."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    if isinstance(local, torch.Tensor):
        dist.all_reduce(local, op=dist.ReduceOp.SUM, group=group)
        return local
    t = torch.from_numpy(np.ascontiguousarray(local, np.float64))
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.numpy()


class ShardedEngine:
    """The engine of this rank's shard plus the allreduce; API mirrors Engine (compute / explore)."""

    def __init__(self, rank: int | None = None, world: int | None = None, device: int | None = None, group=None, **kw):
        import torch
        import torch.distributed as dist

        from .engine import Engine
        if rank is None:
            rank = dist.get_rank(group) if dist.is_initialized() else 0
        if world is None:
            world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank, self.world, self.group = rank, world, group
        n_paths = kw.pop("n_paths", 0) or kw.get("NTPB", 512) * kw.get("NB", 512)
        first, n_local = shard_bounds(n_paths, rank, world)
        self.n_paths = n_paths
        self.device = torch.cuda.current_device() if device is None else device
        self.engine = Engine(n_paths=n_paths, first_path=first, n_local=n_local, device=self.device, **kw)
        self._torch = torch
        self._moments = None

    def init(self, seed: int = 1234):
        self.engine.init(seed)
        return self

    def _buffer(self, n_points: int):
        torch = self._torch
        if self._moments is None or self._moments.numel() < 2 * n_points:
            self._moments = torch.zeros(2 * n_points, dtype=torch.float64, device=torch.device("cuda", self.device))
        return self._moments[: 2 * n_points]

    def compute_async(self):
        """Enqueue kernel + allreduce on the current torch stream; returns the device tensor of GLOBAL sums."""
        torch = self._torch
        buf = self._buffer(1)
        self.engine.compute_async(torch.cuda.current_stream(self.device).cuda_stream, buf.data_ptr())
        return allreduce_moments(buf, self.group)

    def compute(self):
        from .engine import Moments
        s = self.compute_async().cpu().numpy()
        return Moments(float(s[0]), float(s[1]), self.n_paths, float("nan"))

    def explore(self, k, theta, sigma):
        from .engine import Moments
        torch = self._torch
        buf = self._buffer(len(k))
        self.engine.explore_async(torch.cuda.current_stream(self.device).cuda_stream, k, theta, sigma, buf.data_ptr())
        s = allreduce_moments(buf, self.group).cpu().numpy().reshape(-1, 2)
        return [Moments(float(a), float(b), self.n_paths, float("nan")) for a, b in s]

    def close(self):
        self.engine.close()
