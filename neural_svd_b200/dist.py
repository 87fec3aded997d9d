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
        self._count_rows = {}          # (n_local, b1_local, device) -> the four count floats, on the device
        self._synced = set()

    def sync_parameters(self, module):
        """Broadcast every parameter / buffer of `module` from rank 0, once per module object: replicas that were
        constructed from different seeds would otherwise diverge silently (the gradients are summed, not checked)."""
        if id(module) in self._synced:
            return
        with torch.no_grad():
            for t in list(module.parameters()) + list(module.buffers()):
                dist.broadcast(t.data, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0,
                               group=self.group)
        self._synced.add(id(module))

    def count_row(self, n_local: int, b1_local: int, device):
        """[n mod 2^16, n / 2^16, b1 mod 2^16, b1 / 2^16] of THIS rank as four floats on the device (cached: it is the
        rank's own constant).  Appended to the `terms` buffer, the counts travel with all-reduce #1 on every step, so a
        rank whose batch size changes cannot desynchronise the collectives, and nothing is copied to the host."""
        key = (n_local, b1_local, str(device))
        if key not in self._count_rows:
            self._count_rows[key] = torch.tensor([n_local & 0xFFFF, n_local >> 16, b1_local & 0xFFFF, b1_local >> 16],
                                                 dtype=torch.float32, device=device)
        return self._count_rows[key]

    def allreduce_terms(self, terms: torch.Tensor, n_local: int, b1_local: int):
        """all-reduce #1: the un-normalised [G1 | G2 | operator sum | counts] buffer (2 L^2 + 5 floats).  `terms` must
        have room for the four count floats at its end."""
        terms[-4:] = self.count_row(n_local, b1_local, terms.device)
        dist.all_reduce(terms, op=dist.ReduceOp.SUM, group=self.group)
        return terms

    def global_counts(self, n_local: int, b1_local: int, device):
        """(B, B1, B2) summed over ranks, on the HOST (one small all-reduce + copy; for callers that need python ints,
        e.g. the CDK loss).  Always runs the collective: no cache that could leave ranks with mismatched calls."""
        c = torch.tensor([n_local, b1_local], dtype=torch.int64, device=device)
        dist.all_reduce(c, op=dist.ReduceOp.SUM, group=self.group)
        Bg, B1g = int(c[0]), int(c[1])
        return Bg, B1g, Bg - B1g

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
