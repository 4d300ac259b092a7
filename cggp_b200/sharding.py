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


def shard_rows_weighted(n: int, rank: int, speeds, block: int = 24) -> Tuple[int, int]:
    """Contiguous split in proportion to the ranks' measured speeds (e.g. ``1 / product_ms`` of a calibration solve,
    all-gathered): the devices of one box differ by a constant 1 - 2 %, and with an even split every iteration waits for
    the slowest one.  Boundaries fall on multiples of ``block`` rows (the row block of the pipelined
    kernel), every rank keeps at least one block while rows last, and the spans partition ``[0, n)``."""
    world = len(speeds)
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    if any(not (s > 0.0) for s in speeds):
        raise ValueError("speeds must be positive")
    n, block = int(n), max(1, int(block))
    total = float(sum(speeds))
    nblocks = (n + block - 1) // block
    bounds, acc = [0], 0.0
    for r in range(world):
        acc += float(speeds[r])
        b = int(round(nblocks * acc / total)) if r < world - 1 else nblocks
        b = max(b, min(bounds[-1] + 1, nblocks))   # at least one block per rank while blocks last
        b = min(b, nblocks)
        bounds.append(b)
    start, stop = min(bounds[rank] * block, n), min(bounds[rank + 1] * block, n)
    return start, stop
