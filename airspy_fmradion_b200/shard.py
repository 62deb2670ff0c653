"""Channel sharding across ranks (one process per GPU). Channels are independent streams, so
the path partitions with no data-path collective: rank r of W owns a contiguous range."""


def channel_range(rank, world, total_channels):
    """Contiguous, balanced partition: the first (total % world) ranks get one extra channel."""
    if not (0 <= rank < world) or total_channels < 0:
        raise ValueError("bad rank/world/total")
    base, extra = divmod(total_channels, world)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi
