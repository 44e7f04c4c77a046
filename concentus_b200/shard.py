"""Host-side stream sharding: streams are independent, so rank r of W simply owns a contiguous block.
No collective is needed on the data path; the only cross-rank communication in bench.py is the timing barrier/max."""


def shard_range(n_streams, world, rank):
    """[begin, end) of the streams owned by `rank`; blocks differ in size by at most one."""
    base, rem = divmod(n_streams, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)
