"""world_size-2 gloo test (CPU) of the host side of the multi-rank similarity stage: the row
partition, the variable-length pair gather and the grouping give the single-process answer."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from facet_b200.ops import balanced_row_blocks
    from facet_b200.synth import synth_embeddings, synth_hashes
    from facet_b200.utils.duplicate import all_gather_embeddings, gather_pairs, group_duplicates
    from oracle import grouping as og
    n = 600
    e = synth_embeddings(n, seed=1, cluster_fraction=0.4)
    shard = torch.from_numpy(e[rank * (n // world):(rank + 1) * (n // world)])
    full = all_gather_embeddings(shard)
    assert np.array_equal(full.numpy(), e)
    # the kernel's job (pairs of this rank's row block) is done by the oracle here: CPU box
    b = balanced_row_blocks(n, world)
    allp = og.cosine_pairs(e, 0.9)
    mine = allp[(allp[:, 0] >= b[rank]) & (allp[:, 0] < b[rank + 1])]
    pairs = gather_pairs(torch.from_numpy(mine.astype(np.int32)))
    assert sorted(map(tuple, pairs.tolist())) == sorted(map(tuple, allp.tolist()))
    aggs = np.random.default_rng(0).uniform(0, 10, n).round(2).tolist()
    gid, lead = group_duplicates(n, pairs, aggs)
    wg, wl = og.cosine_groups(e, aggs, 0.9)
    assert gid.tolist() == wg.tolist() and lead.tolist() == wl.tolist()
    # any disjoint dealing of the pair triangle over the ranks must union to the full set (the kernel deals upper-triangle
    # tiles round-robin; here rows are dealt in blocks of 2048)
    h = synth_hashes(5000, seed=2, dup_fraction=0.3)
    hp = og.hamming_pairs(h, 6)
    mine = hp[((hp[:, 0] // 2048) % world) == rank]
    pairs = gather_pairs(torch.from_numpy(mine.astype(np.int32)))
    g2, l2 = group_duplicates(5000, pairs, [1.0] * 5000)
    w2, wl2 = og.duplicate_groups(h, [1.0] * 5000, 90)
    assert g2.tolist() == w2.tolist() and l2.tolist() == wl2.tolist()
    if rank == 0:
        ret["ok"] = True
    dist.destroy_process_group()


def test_two_rank_gather_and_grouping():
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret.get("ok")


def test_balanced_row_blocks_cover_and_balance():
    from facet_b200.ops import balanced_row_blocks
    for n, p in [(1000, 8), (100000, 8), (7, 2), (128, 3)]:
        b = balanced_row_blocks(n, p)
        assert b[0] == 0 and b[-1] == n and all(x <= y for x, y in zip(b, b[1:]))
        if n >= 1000:
            work = [sum(n - 1 - i for i in range(b[k], b[k + 1])) for k in range(p)]
            assert max(work) / (sum(work) / p) < 1.05
