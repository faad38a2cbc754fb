"""GPU parity for duplicate / burst grouping (csrc/hamming.cu + host chain) through the
reference-shaped entry points, against reference goldens and the oracle."""
import os

import numpy as np
import pytest

from grouping_util import load_cases, make_db, read_result, sqlite_order, write_config
from facet_b200.synth import synth_hashes, synth_timestamps

pytestmark = pytest.mark.gpu


def test_detect_duplicates_and_bursts_match_reference_golden(tmp_path):
    from facet_b200.processing.bursts import process_bursts
    from facet_b200.utils.duplicate import detect_duplicates
    cfg = str(tmp_path / "cfg.json")
    write_config(cfg)
    for k, case in enumerate(load_cases()):
        db = str(tmp_path / f"t{k}.db")
        make_db(db, case["rows"], case["persons"])
        detect_duplicates(db, cfg)
        process_bursts(db, cfg)
        got = read_result(db)
        for p, want in case["result"].items():
            assert got[p] == want, (case["n"], p, got[p], want)


@pytest.mark.parametrize("n", [1, 2, 2047, 2048, 2049, 5000, 20000])
def test_hamming_pairs_bit_exact(n):
    from facet_b200 import ops
    from oracle import grouping as og
    h = synth_hashes(n, seed=n, dup_fraction=0.25, max_flip=9)
    for thr in (0, 6, 19):
        got = ops.hamming_pairs(h, thr).cpu().numpy().astype(np.int64)
        got = got[np.lexsort((got[:, 1], got[:, 0]))] if len(got) else got.reshape(0, 2)
        if n <= 5000:
            want = og.hamming_pairs(h, thr)
        else:   # vectorised oracle for the larger case
            want = []
            for i0 in range(0, n, 1024):
                d = og.popcount64(h[i0:i0 + 1024, None] ^ h[None, :])
                ii, jj = np.nonzero(d <= thr)
                ii += i0
                want.append(np.stack([ii[jj > ii], jj[jj > ii]], 1))
            want = np.concatenate(want)
            want = want[np.lexsort((want[:, 1], want[:, 0]))]
        assert got.tolist() == want.tolist()


def test_hamming_parts_partition_the_pair_set():
    from facet_b200 import ops
    h = synth_hashes(9000, seed=5, dup_fraction=0.3, max_flip=8)
    full = ops.hamming_pairs(h, 6).cpu().numpy()
    full = {tuple(p) for p in full.tolist()}
    for nparts in (2, 3, 8):
        parts = [ops.hamming_pairs(h, 6, part=r, nparts=nparts).cpu().numpy() for r in range(nparts)]
        sets = [{tuple(p) for p in x.tolist()} for x in parts]
        assert sum(len(s) for s in sets) == len(full), "parts overlap"
        assert set().union(*sets) == full
    # truncated buffer is reported and retried
    small = ops.hamming_pairs(h, 6, cap=3).cpu().numpy()
    assert {tuple(p) for p in small.tolist()} == full


def test_burst_chain_vs_oracle_random():
    from facet_b200.processing.bursts import burst_leads
    from oracle import grouping as og
    for seed in range(4):
        n = 3000
        rng = np.random.default_rng(seed)
        h = synth_hashes(n, seed=seed, dup_fraction=0.5, max_flip=40)
        ts = synth_timestamps(n, seed=seed)
        if seed == 3:
            ts = ts[rng.permutation(n)]          # non-monotonic order must still be exact
        dates = ["2024:05:%02d %02d:%02d:%02d" % (1 + s // 86400, (s // 3600) % 24, (s // 60) % 60, s % 60) for s in ts.tolist()]
        if seed == 2:
            for k in range(0, n, 41):
                dates[k] = None
        hexes = ["%016x" % int(x) for x in h]
        aggs = np.round(rng.uniform(0, 10, n), 1).tolist()
        kw = dict(similarity_percent=70, time_window_minutes=0.8, rapid_burst_seconds=0.4)
        got = burst_leads(dates, hexes, aggs, **kw)
        want = og.burst_leads(dates, hexes, aggs, **kw)
        assert got.tolist() == want.tolist()
        kw = dict(similarity_percent=88, time_window_minutes=60, rapid_burst_seconds=5)   # reference defaults
        assert burst_leads(dates, hexes, aggs, **kw).tolist() == og.burst_leads(dates, hexes, aggs, **kw).tolist()
