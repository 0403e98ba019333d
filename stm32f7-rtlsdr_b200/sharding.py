"""Multi-GPU host logic: captures are independent, so the batch is partitioned by capture index
with no data-path collective.  One process per GPU; rank r of W owns a
contiguous range of captures.  torch.distributed is used only for the barrier / max-over-ranks
timing in bench.py and for gathering tiny result summaries."""


def shard_range(n_captures, rank, world):
    """Contiguous, balanced partition: the first n % W ranks get one extra capture."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    base, extra = divmod(n_captures, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def all_shards(n_captures, world):
    return [shard_range(n_captures, r, world) for r in range(world)]
