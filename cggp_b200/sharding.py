"""Row sharding of the training inputs over the ranks of one box: rank r owns the contiguous rows
``[start, stop)``; Z, Kuu and all CG vectors are replicated.  The only exchange on the path is the all-reduce of the
partial ``[B, M]`` product of every operator application (``cggp_allreduce_sum``)."""
from __future__ import annotations

from typing import Tuple


def shard_rows(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split: the first ``n % world`` ranks get one extra row."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(int(n), int(world))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)
