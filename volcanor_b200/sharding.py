"""Target sharding across ranks (one process per GPU) -- the multi-GPU plumbing of the hot path.

Every target's velocity depends on the full, read-only source set of the stage (the reference's OpenMP loop over
targets, src/libCommon.f90:132-139, has no cross-iteration dependence), so the target list is cut into contiguous
slices, one per rank; sources are replicated.  The only exchange step is an all-gather of the convected node
positions after each convection stage (predictor and corrector: 2 per step for fdScheme 1/3).

Works with any torch.distributed backend: NCCL on the GPUs (bench.py), gloo on CPU (tests/test_sharding_gloo.py).
"""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class TargetShard:
    m: int          # total targets
    world: int
    rank: int

    @property
    def per(self) -> int:
        """Slice length every rank owns (the last slices may be partly or wholly padding)."""
        return (self.m + self.world - 1) // self.world if self.m > 0 else 0

    @property
    def lo(self) -> int:
        return min(self.rank * self.per, self.m)

    @property
    def hi(self) -> int:
        return min((self.rank + 1) * self.per, self.m)

    @property
    def count(self) -> int:
        return self.hi - self.lo

    @property
    def padded(self) -> int:
        return self.world * self.per


def allgather_slices(P_all, shard: TargetShard, group=None):
    """In place: every rank contributes rows [rank*per, (rank+1)*per) of P_all (padded, (world*per, 3)) and
    receives everybody else's.  No-op for world == 1."""
    if shard.world == 1:
        return P_all
    import torch.distributed as dist
    mine = P_all[shard.rank * shard.per:(shard.rank + 1) * shard.per].clone()
    dist.all_gather_into_tensor(P_all, mine, group=group)
    return P_all
