"""Query sharding across ranks: contiguous blocks of reads per rank, both strands of a read
stay together, the index is replicated; hit lists are gathered on the host in rank order, which restores
query order.  No data-path collective: the only communication is the gather of the results."""
import numpy as np


def shard_range(n_reads, rank, world):
    """reads [lo, hi) owned by `rank`"""
    per = (n_reads + world - 1) // world
    lo = min(n_reads, per * rank)
    hi = min(n_reads, per * (rank + 1))
    return lo, hi


def globalize(hits, first_query):
    """shard-local query ids -> global query ids"""
    hits = np.asarray(hits, dtype=np.uint64).reshape(-1, 4).copy()
    hits[:, 0] += np.uint64(first_query)
    return hits


def gather_hits(local_hits, dist=None):
    """all ranks -> rank 0: concatenation of the per-rank hit arrays in rank order (None on other ranks)"""
    local_hits = np.asarray(local_hits, dtype=np.uint64).reshape(-1, 4)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local_hits
    out = [None] * dist.get_world_size() if dist.get_rank() == 0 else None
    dist.gather_object(local_hits, out, dst=0)
    if dist.get_rank() != 0:
        return None
    return np.concatenate(out, axis=0)
