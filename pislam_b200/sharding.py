"""Frame sharding for the multi-GPU path: pyramids are independent units, so a batch is
split into contiguous shards, one per rank, each rank runs the single-GPU pipeline, and the fixed-capacity frame
records meet in ONE all-gather (rank-major).  Host-side bookkeeping only -- no data-path code here."""
from typing import List, Tuple


def shard_frames(nframes: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous shard of `nframes` for `rank`: (first frame, count); every rank gets ceil(n/world) slots,
    trailing ranks may own fewer (or zero) real frames."""
    per = (nframes + world - 1) // world
    start = min(rank * per, nframes)
    return start, max(0, min(per, nframes - start))


def frames_per_rank(nframes: int, world: int) -> int:
    """Slots per rank in the gathered buffer (static size => one fixed-size all-gather)."""
    return (nframes + world - 1) // world


def gathered_index(nframes: int, world: int) -> List[Tuple[int, int]]:
    """For every global frame f: (rank, slot) of its record in the rank-major gathered buffer."""
    per = frames_per_rank(nframes, world)
    return [(f // per, f % per) for f in range(nframes)]
