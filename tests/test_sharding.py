"""world_size-2 gloo test of the query sharding (the only multi-rank logic of the path):
contiguous read shards per rank, replicated index, host-side gather in rank order == unsharded result."""
import os
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle as O
import workloads as W
import sahara_b200 as sb
from sahara_b200 import sharding


def _worker(rank, world, port, q, seqs, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ix = O.OracleIndex.build(seqs, 6, 16)  # index replicated on every rank
    sch = sb.SearchScheme.generate("h2-k2", 0, 2, q.shape[1])
    lo, hi = sharding.shard_range(q.shape[0] // 2, rank, world)
    local = ix.locate(ix.search(q[2 * lo: 2 * hi], sch, True)) if hi > lo else np.zeros((0, 4), np.uint64)
    allhits = sharding.gather_hits(sharding.globalize(local, 2 * lo), dist)
    if rank == 0:
        np.save(out_path, allhits)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process(tmp_path):
    rng = np.random.default_rng(9)
    seqs = [W.repetitive_genome(rng, 8000)]
    q = W.sample_reads(rng, seqs, 51, 30, 2, True)  # odd number of reads: uneven shards
    out = os.path.join(tmp_path, "hits.npy")
    mp.spawn(_worker, args=(2, 29517, q, seqs, out), nprocs=2, join=True)
    got = np.load(out)
    ix = O.OracleIndex.build(seqs, 6, 16)
    sch = sb.SearchScheme.generate("h2-k2", 0, 2, 30)
    want = ix.locate(ix.search(q, sch, True))
    assert np.array_equal(got, want)  # rank-order concatenation restores the reference's query order


def test_shard_ranges_cover_everything():
    for n in (0, 1, 7, 100, 1001):
        for world in (1, 2, 3, 8):
            r = [sharding.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
