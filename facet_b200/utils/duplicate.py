"""Global duplicate detection on pHash values — drop-in for utils/duplicate.py:44-169.

The O(N^2) XOR/popcount/threshold loop (duplicate.py:94-119) runs on the GPU
(csrc/hamming.cu, optionally sharded over ranks); Union-Find, group numbering and lead
selection stay on the host and follow the reference step by step so the database ends up
byte-identical: groups are numbered in ascending order of their Union-Find root
(duplicate.py:150), the lead is the first member with the highest aggregate
(duplicate.py:152, Python ``max`` keeps the first maximum).
"""
from __future__ import annotations

import sqlite3

import numpy as np

from .. import ops


class _UnionFind:
    """Path-halving / union-by-rank, same tie-breaking as duplicate.py:15-36 (roots matter:
    group ids follow sorted roots)."""

    def __init__(self, n):
        self.parent = list(range(n))
        self.rank = [0] * n

    def find(self, x):
        p = self.parent
        while p[x] != x:
            p[x] = p[p[x]]
            x = p[x]
        return x

    def union(self, a, b):
        ra, rb = self.find(a), self.find(b)
        if ra == rb:
            return
        if self.rank[ra] < self.rank[rb]:
            ra, rb = rb, ra
        self.parent[rb] = ra
        if self.rank[ra] == self.rank[rb]:
            self.rank[ra] += 1


def max_hamming_distance(similarity_pct) -> int:
    """duplicate.py:63 / scorer.py:1891: int(64 * (1 - pct/100))."""
    return int(64 * (1 - similarity_pct / 100))


def gather_pairs(pairs_local, group=None) -> np.ndarray:
    """Concatenate the pair lists of all ranks (variable length) on every rank.

    Single process: just moves the list to the host.  Multi-process: all_gather of the counts,
    then all_gather of lists padded to the maximum.
    """
    import torch
    import torch.distributed as dist
    if group is None and not (dist.is_available() and dist.is_initialized()):
        return pairs_local.cpu().numpy().astype(np.int64).reshape(-1, 2)
    world = dist.get_world_size(group)
    if world == 1:
        return pairs_local.cpu().numpy().astype(np.int64).reshape(-1, 2)
    dev = pairs_local.device
    count = torch.tensor([pairs_local.shape[0]], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(count) for _ in range(world)]
    dist.all_gather(counts, count, group=group)
    counts = [int(c.item()) for c in counts]
    mx = max(max(counts), 1)
    padded = torch.zeros((mx, 2), dtype=torch.int32, device=dev)
    padded[: pairs_local.shape[0]] = pairs_local
    bufs = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(bufs, padded, group=group)
    parts = [b[:c].cpu().numpy() for b, c in zip(bufs, counts)]
    return np.concatenate(parts, axis=0).astype(np.int64).reshape(-1, 2)


def group_duplicates(n: int, pairs: np.ndarray, aggregates):
    """Union-Find + numbering + lead selection.  Returns (group_id int64[n] with 0 = none,
    is_lead uint8[n]).

    The reference unions matches in lexicographic (i, j) order (duplicate.py:97-119); the shape
    of the forest — and so the sorted roots that number the groups — depends on that order, so
    the GPU's unordered pair list is sorted first.
    """
    gid = np.zeros(n, np.int64)
    lead = np.zeros(n, np.uint8)
    if len(pairs) == 0:
        return gid, lead
    pairs = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)
    order = np.lexsort((pairs[:, 1], pairs[:, 0]))
    uf = _UnionFind(n)
    for i, j in pairs[order].tolist():
        uf.union(i, j)
    groups: dict[int, list[int]] = {}
    for idx in range(n):
        groups.setdefault(uf.find(idx), []).append(idx)
    next_id = 1
    for _root, members in sorted(groups.items()):
        if len(members) < 2:
            continue
        best = max(members, key=lambda k: aggregates[k])
        for k in members:
            gid[k] = next_id
        lead[best] = 1
        next_id += 1
    return gid, lead


def find_duplicate_groups(hashes: np.ndarray, aggregates, max_distance: int, group=None):
    """hashes uint64[n] (row order = ORDER BY path) -> (group_id, is_lead).  With an initialised
    torch.distributed group every rank scans its share of row tiles and the pair lists are
    all-gathered; all ranks return the same answer."""
    import torch.distributed as dist
    rank, world = 0, 1
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    local = ops.hamming_pairs(hashes, max_distance, part=rank, nparts=world)
    pairs = gather_pairs(local, group)
    return group_duplicates(len(hashes), pairs, aggregates)


def detect_duplicates(db_path, config_path=None):
    """Same contract as the reference's detect_duplicates(db_path, config_path): reads
    (path, phash, aggregate) ordered by path, writes duplicate_group_id / is_duplicate_lead."""
    from ..config import ScoringConfig
    config = ScoringConfig(config_path, validate=False)
    pct = config.get_duplicate_detection_settings().get("similarity_threshold_percent", 90)
    max_distance = max_hamming_distance(pct)
    print(f"Duplicate detection: similarity >= {pct}% (Hamming distance <= {max_distance})")
    with sqlite3.connect(db_path) as conn:
        rows = conn.execute("SELECT path, phash, aggregate FROM photos WHERE phash IS NOT NULL ORDER BY path").fetchall()
    if not rows:
        print("No photos with pHash found.")
        return
    paths = [r[0] for r in rows]
    aggregates = [r[2] or 0.0 for r in rows]
    hashes = np.array([int(r[1], 16) for r in rows], dtype=np.uint64)
    print(f"Comparing {len(paths)} photos...")
    gid, lead = find_duplicate_groups(hashes, aggregates, max_distance)
    with sqlite3.connect(db_path) as conn:
        conn.execute("UPDATE photos SET duplicate_group_id = NULL, is_duplicate_lead = 0")
        sel = np.flatnonzero(gid)
        if len(sel) == 0:
            print("No duplicates found.")
        else:
            conn.executemany("UPDATE photos SET duplicate_group_id = ?, is_duplicate_lead = ? WHERE path = ?",
                             [(int(gid[k]), int(lead[k]), paths[k]) for k in sel.tolist()])
            print(f"Marked {int(gid.max())} groups: {len(sel)} photos")
        conn.commit()


# ---------------------------------------------------------------------------------------------
# Cosine mode over CLIP embeddings (north_star kernel 3; formula sites models/tagger.py:99-101,
# api/routers/gallery.py:465-471).  Ranks hold disjoint shards of the [N,768] embedding matrix;
# one NCCL all-gather makes the matrix resident everywhere, then every rank scans its balanced
# row block against all columns and the surviving pair lists are gathered.
# ---------------------------------------------------------------------------------------------

def all_gather_embeddings(local_emb, group=None):
    """[n_local,d] float32 shards (equal n_local on every rank) -> [world*n_local, d] on every rank."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local_emb
    world = dist.get_world_size(group)
    out = torch.empty((world * local_emb.shape[0], local_emb.shape[1]), dtype=local_emb.dtype, device=local_emb.device)
    dist.all_gather_into_tensor(out, local_emb.contiguous(), group=group)
    return out


def shard_blocks(rank: int, world: int, n_local: int):
    """The shard-against-shard blocks rank `rank` scans: [(a_shard, a_lo, a_hi, b_shard, b_lo, b_hi, triangle)] with row ranges
    inside the shards.  Every unordered pair of rows is covered by exactly one block of exactly one rank, every rank scans
    the same number of row pairs: its own shard against itself (upper triangle; needs no other rank's data), the next
    (world - 1) // 2 shards in ring order in full, and for an even world half of the opposite shard's block (the lower ranks
    take the first half of their own rows against the whole opposite shard, the upper ranks their whole shard against the second
    half of the opposite one)."""
    blocks = [(rank, 0, n_local, rank, 0, n_local, True)]
    for k in range(1, (world - 1) // 2 + 1):
        blocks.append((rank, 0, n_local, (rank + k) % world, 0, n_local, False))
    if world % 2 == 0 and world > 1:
        half = n_local // 2
        if rank < world // 2:
            blocks.append((rank, 0, half, rank + world // 2, 0, n_local, False))
        else:
            blocks.append((rank, 0, n_local, rank - world // 2, half, n_local, False))
    return blocks


def cosine_pairs_sharded(local_emb, tau: float, group=None):
    """This rank's share of the all-pairs cosine scan over the shards of every rank.  [n_local, d] float32 (equal n_local
    everywhere) -> (pairs int32 [m,2] with global row indices i < j, sims float32 [m]).

    The scan is cut into shard-against-shard blocks (`shard_blocks`).  The block of the rank's own shard against itself starts at
    once, on the local bf16 copy, WHILE NCCL gathers the bf16 copies of the other shards (half the bytes of the float32 rows) on
    its own stream; the other blocks wait for that gather; the float32 rows, which only the recheck of the few candidates
    needs, are gathered behind it and are waited for last."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return ops.cosine_pairs(local_emb, tau)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    local = local_emb.contiguous()
    n_local, d = local.shape
    eb_local = ops.to_bf16(local)
    eb = torch.empty((world * n_local, d), dtype=torch.bfloat16, device=local.device)
    work_b = dist.all_gather_into_tensor(eb, eb_local, group=group, async_op=True)
    e32 = torch.empty((world * n_local, d), dtype=torch.float32, device=local.device)
    work_f = dist.all_gather_into_tensor(e32, local, group=group, async_op=True)
    blocks = []
    for (sa, a_lo, a_hi, sb, b_lo, b_hi, tri) in shard_blocks(rank, world, n_local):
        a = eb_local[a_lo:a_hi] if sa == rank else eb[sa * n_local + a_lo: sa * n_local + a_hi]
        b = eb_local[b_lo:b_hi] if sb == rank else eb[sb * n_local + b_lo: sb * n_local + b_hi]
        blocks.append((a, sa * n_local + a_lo, b, sb * n_local + b_lo, tri))

    def before_block(k):
        if k == 1:
            work_b.wait()           # the compute stream waits for the NCCL stream; the host does not block

    def wait_f32():
        work_f.wait()
        return e32

    return ops.cosine_blocks(blocks, wait_f32, d, tau, before_block=before_block)


def find_similar_groups(local_emb, aggregates, tau: float, group=None):
    """Cosine-threshold grouping: returns (group_id, is_lead) for the gathered rows on every rank."""
    import torch.distributed as dist
    rank, world = 0, 1
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    local_pairs, _ = cosine_pairs_sharded(local_emb, tau, group)
    pairs = gather_pairs(local_pairs, group)
    return group_duplicates(world * local_emb.shape[0], pairs, aggregates)
