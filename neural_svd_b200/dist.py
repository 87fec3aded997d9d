"""Data parallelism over collocation points: one process per GPU, parameters
replicated, two all-reduces per step.  The reference has no executed multi-GPU path; the
equivalence target is "G ranks == 1 rank on the concatenated batch".

Rank r treats the first half of its local points as part of f1 and the second half as part of
f2, so the union over ranks reproduces torch.chunk(f, 2) (methods/nestedlora.py:263) on the
global batch [all first halves | all second halves].
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class PointParallel:
    def __init__(self, group=None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.group = group
        self.world_size = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self._counts_cache = {}

    def reset_counts(self):
        """Forget the cached global counts (call on every rank if ANY rank changes its local batch size)."""
        self._counts_cache = {}

    def global_counts(self, n_local: int, b1_local: int, device):
        """(B, B1, B2) summed over ranks.  Cached per (n_local, b1_local) of this rank — batch sizes are static in
        the reference's training loop — so the steady state has no host synchronisation; see reset_counts()."""
        key = (n_local, b1_local)
        if key not in self._counts_cache:
            c = torch.tensor([n_local, b1_local], dtype=torch.int64, device=device)
            dist.all_reduce(c, op=dist.ReduceOp.SUM, group=self.group)
            Bg, B1g = int(c[0]), int(c[1])
            self._counts_cache[key] = (Bg, B1g, Bg - B1g)
        return self._counts_cache[key]

    def allreduce_terms(self, terms: torch.Tensor, n_local: int, b1_local: int):
        """all-reduce #1: the un-normalised [G1 | G2 | operator sum] buffer (2 L^2 + 1 floats)."""
        dist.all_reduce(terms, op=dist.ReduceOp.SUM, group=self.group)
        return self.global_counts(n_local, b1_local, terms.device)

    def allreduce_grads(self, flat: torch.Tensor):
        """all-reduce #2: the flat parameter-gradient buffer. SUM, no averaging: dF already carries 1/B_global."""
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        return flat


def shard_points(x_global: torch.Tensor, rank: int, world_size: int) -> torch.Tensor:
    """Slice a global batch so that rank halves tile the global halves (parity checks)."""
    B = x_global.shape[0]
    b1 = (B + 1) // 2
    h1, h2 = x_global[:b1], x_global[b1:]
    if h1.shape[0] % world_size or h2.shape[0] % world_size:
        raise ValueError("global halves must divide evenly across ranks")
    n1, n2 = h1.shape[0] // world_size, h2.shape[0] // world_size
    return torch.cat([h1[rank * n1:(rank + 1) * n1], h2[rank * n2:(rank + 1) * n2]], 0)
