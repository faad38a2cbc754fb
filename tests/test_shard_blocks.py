"""Block decomposition of the multi-GPU all-pairs scan (utils/duplicate.py shard_blocks): every unordered pair of rows is covered
exactly once, and every rank scans the same number of row pairs."""
import itertools

import pytest

from facet_b200.utils.duplicate import shard_blocks


@pytest.mark.parametrize("world", [1, 2, 3, 4, 5, 8])
@pytest.mark.parametrize("n_local", [4, 7, 10])
def test_every_pair_once_and_balanced(world, n_local):
    seen = {}
    loads = []
    for rank in range(world):
        load = 0
        for sa, a_lo, a_hi, sb, b_lo, b_hi, tri in shard_blocks(rank, world, n_local):
            assert sa == rank or sb == rank
            for i in range(sa * n_local + a_lo, sa * n_local + a_hi):
                for j in range(sb * n_local + b_lo, sb * n_local + b_hi):
                    if tri and not j > i:
                        continue
                    key = (min(i, j), max(i, j))
                    assert i != j
                    seen[key] = seen.get(key, 0) + 1
                    load += 1
        loads.append(load)
    n = world * n_local
    assert set(seen) == set(itertools.combinations(range(n), 2))
    assert all(v == 1 for v in seen.values())
    assert max(loads) - min(loads) <= n_local * ((n_local + 1) // 2)        # odd shard sizes split the opposite block unevenly
