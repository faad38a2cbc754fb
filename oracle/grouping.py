"""ORACLE (test infrastructure): plain-Python restatement of the reference's grouping logic.

* ``duplicate_groups``  follows utils/duplicate.py:44-169 — all pairs i<j with
  popcount(h_i ^ h_j) <= int(64*(1-pct/100)), Union-Find in (i, j) order, groups numbered by
  ascending root, lead = first max aggregate.
* ``burst_leads``       follows processing/scorer.py:1880-1986 — sequential chain over photos
  ordered by date_taken with the rapid / slow rules.
* ``cosine_groups``     the north_star's cosine mode: sim_ij = <e_i, e_j> on stored L2-normalised
  float32 embeddings (formula sites: models/tagger.py:99-101, api/routers/gallery.py:465-471),
  pair kept iff sim >= tau, grouped exactly like duplicate_groups.

Pinned by tests/golden/grouping_golden.json (outputs of the reference functions run on a
temporary SQLite database, see tests/golden/make_golden_grouping.py).
"""
from __future__ import annotations

from datetime import datetime

import numpy as np


class UnionFind:
    def __init__(self, n):
        self.parent = list(range(n))
        self.rank = [0] * n

    def find(self, x):
        while self.parent[x] != x:
            self.parent[x] = self.parent[self.parent[x]]
            x = self.parent[x]
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


def popcount64(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint64)
    out = np.zeros(x.shape, np.int32)
    table = np.array([bin(i).count("1") for i in range(256)], np.int32)
    for b in range(8):
        out += table[((x >> np.uint64(8 * b)) & np.uint64(0xFF)).astype(np.int64)]
    return out


def hamming_pairs(hashes: np.ndarray, max_distance: int) -> np.ndarray:
    """All (i<j) with distance <= max_distance, in lexicographic order."""
    n = len(hashes)
    out = []
    for i in range(n - 1):
        d = popcount64(np.bitwise_xor(hashes[i], hashes[i + 1:]))
        for m in np.where(d <= max_distance)[0]:
            out.append((i, i + 1 + int(m)))
    return np.array(out, dtype=np.int64).reshape(-1, 2)


def groups_from_pairs(n: int, pairs: np.ndarray, aggregates):
    uf = UnionFind(n)
    for i, j in pairs.tolist():
        uf.union(i, j)
    groups = {}
    for idx in range(n):
        groups.setdefault(uf.find(idx), []).append(idx)
    gid = np.zeros(n, np.int64)
    lead = np.zeros(n, np.uint8)
    k = 1
    for _root, members in sorted(groups.items()):
        if len(members) < 2:
            continue
        best = max(members, key=lambda idx: aggregates[idx])
        for m in members:
            gid[m] = k
        lead[best] = 1
        k += 1
    return gid, lead


def duplicate_groups(hashes: np.ndarray, aggregates, similarity_pct=90):
    max_distance = int(64 * (1 - similarity_pct / 100))
    return groups_from_pairs(len(hashes), hamming_pairs(hashes, max_distance), aggregates)


def cosine_pairs(emb: np.ndarray, tau: float, block: int = 2048) -> np.ndarray:
    """All (i<j) with float32 dot >= tau (embeddings are stored L2-normalised)."""
    n = emb.shape[0]
    e = np.ascontiguousarray(emb, dtype=np.float32)
    out = []
    for i0 in range(0, n, block):
        sims = e[i0:i0 + block] @ e.T
        ii, jj = np.nonzero(sims >= np.float32(tau))
        ii = ii + i0
        keep = jj > ii
        out.append(np.stack([ii[keep], jj[keep]], axis=1))
    p = np.concatenate(out, axis=0) if out else np.zeros((0, 2), np.int64)
    return p[np.lexsort((p[:, 1], p[:, 0]))].astype(np.int64)


def cosine_pairs_exact(emb: np.ndarray, tau: float, block: int = 1024) -> np.ndarray:
    """All (i<j) with float64 dot of the stored float32 rows >= float64(float32(tau)): the order-independent
    criterion (to ~1e-13) that the device recheck implements; cosine_pairs above is the reference's float32 formula,
    whose result near tau depends on the BLAS summation order."""
    n = emb.shape[0]
    e = np.ascontiguousarray(emb, dtype=np.float32).astype(np.float64)
    t = np.float64(np.float32(tau))
    out = []
    for i0 in range(0, n, block):
        sims = e[i0:i0 + block] @ e.T
        ii, jj = np.nonzero(sims >= t)
        ii = ii + i0
        keep = jj > ii
        out.append(np.stack([ii[keep], jj[keep]], axis=1))
    p = np.concatenate(out, axis=0) if out else np.zeros((0, 2), np.int64)
    return p[np.lexsort((p[:, 1], p[:, 0]))].astype(np.int64)


def cosine_groups(emb: np.ndarray, aggregates, tau: float):
    return groups_from_pairs(emb.shape[0], cosine_pairs(emb, tau), aggregates)


def _parse(date_str):
    if not date_str:
        return None
    try:
        return datetime.strptime(date_str[:19], "%Y:%m:%d %H:%M:%S")
    except (ValueError, TypeError):
        return None


def burst_leads(dates, hashes_hex, aggregates, paths=None, photo_persons=None, similarity_percent=88,
                time_window_minutes=60, rapid_burst_seconds=5):
    n = len(dates)
    lead = np.zeros(n, np.uint8)
    if n == 0:
        return lead
    thr = int(64 * (1 - similarity_percent / 100))
    persons = photo_persons or {}

    def dist(a, b):
        if not a or not b:
            return 999
        return bin(int(a, 16) ^ int(b, 16)).count("1")

    def shares(i, b):
        if paths is None:
            return True
        p1, p2 = persons.get(paths[i], set()), persons.get(paths[b], set())
        if not p1 or not p2:
            return True
        return bool(p1 & p2)

    def similar(i, burst):
        di = _parse(dates[i])
        if di is None:
            return False
        for b in burst:
            db = _parse(dates[b])
            if db is None:
                continue
            td = abs((di - db).total_seconds())
            if td <= rapid_burst_seconds and shares(i, b) and dist(hashes_hex[i], hashes_hex[b]) <= thr * 2:
                return True
            if td <= time_window_minutes * 60 and dist(hashes_hex[i], hashes_hex[b]) <= thr:
                return True
        return False

    def close(burst):
        w = max(burst, key=lambda k: aggregates[k] or 0)
        lead[w] = 1

    cur = [0]
    for i in range(1, n):
        if similar(i, cur):
            cur.append(i)
        else:
            close(cur)
            cur = [i]
    close(cur)
    return lead
