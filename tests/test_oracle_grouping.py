"""Pin the grouping oracle against outputs of the reference's detect_duplicates / process_bursts."""
import numpy as np

from grouping_util import load_cases, sqlite_order
from oracle import grouping as og


def test_duplicate_groups_match_reference_golden():
    for case in load_cases():
        rows = case["rows"]
        order = sqlite_order(rows, "path")
        hashes = np.array([int(rows[i]["phash"], 16) for i in order], dtype=np.uint64)
        aggs = [rows[i]["aggregate"] or 0.0 for i in order]
        gid, lead = og.duplicate_groups(hashes, aggs, 90)
        for k, i in enumerate(order):
            g, l, _ = case["result"][rows[i]["path"]]
            assert (g or 0) == int(gid[k]), (case["n"], rows[i]["path"])
            assert int(l) == int(lead[k])


def test_burst_leads_match_reference_golden():
    for case in load_cases():
        rows = case["rows"]
        order = sqlite_order(rows, "date_taken")
        persons = {p: set(v) for p, v in case["persons"].items()}
        lead = og.burst_leads([rows[i]["date_taken"] for i in order], [rows[i]["phash"] for i in order],
                              [rows[i]["aggregate"] for i in order], [rows[i]["path"] for i in order], persons,
                              similarity_percent=70, time_window_minutes=0.8, rapid_burst_seconds=0.4)
        for k, i in enumerate(order):
            assert int(case["result"][rows[i]["path"]][2]) == int(lead[k]), (case["n"], k)


def test_popcount_and_pairs_small():
    rng = np.random.default_rng(0)
    h = rng.integers(0, 2**63, size=50, dtype=np.uint64)
    p = og.hamming_pairs(h, 30)
    brute = [(i, j) for i in range(50) for j in range(i + 1, 50) if bin(int(h[i]) ^ int(h[j])).count("1") <= 30]
    assert p.tolist() == [list(x) for x in brute]
