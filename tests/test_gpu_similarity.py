"""GPU parity for the cosine similarity stage: tcgen05 GEMM + threshold epilogue + fp32 re-score."""
import numpy as np
import pytest

from facet_b200.synth import synth_embeddings

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [2, 130, 1000, 5000, 12000])
def test_cosine_pair_set_matches_oracle(n):
    import torch
    from facet_b200 import ops
    from oracle import grouping as og
    e = synth_embeddings(n, seed=n, cluster_fraction=0.3)
    et = torch.from_numpy(e).cuda()
    for tau in (0.90, 0.97):
        want = og.cosine_pairs(e, tau)
        pairs, sims = ops.cosine_pairs(et, tau)
        got = pairs.cpu().numpy().astype(np.int64)
        got = got[np.lexsort((got[:, 1], got[:, 0]))] if len(got) else got.reshape(0, 2)
        # bit-exact against the order-independent criterion (float64 dot of the float32 rows >= tau) ...
        assert got.tolist() == og.cosine_pairs_exact(e, tau).tolist()
        # ... and equal to the reference's float32 formula wherever that formula is itself well defined: a float32 dot
        # within 2e-6 of tau flips with the BLAS summation order
        full = e @ e.T
        def strict(p, margin):
            return {tuple(x) for x in p.tolist() if abs(float(full[x[0], x[1]]) - tau) > margin}
        assert strict(got, 2e-6) == strict(want, 2e-6)
        if len(got):
            ref = np.einsum("ij,ij->i", e[got[:, 0]], e[got[:, 1]])
            srt = sims.cpu().numpy()[np.lexsort((pairs.cpu().numpy()[:, 1], pairs.cpu().numpy()[:, 0]))]
            np.testing.assert_allclose(srt, ref, rtol=0, atol=2e-6)


def test_cosine_parts_and_grouping():
    import torch
    from facet_b200 import ops
    from facet_b200.utils.duplicate import group_duplicates
    from oracle import grouping as og
    n = 6000
    e = synth_embeddings(n, seed=3, cluster_fraction=0.4)
    et = torch.from_numpy(e).cuda()
    full = {tuple(p) for p in ops.cosine_pairs(et, 0.9)[0].cpu().numpy().tolist()}
    for nparts in (2, 8):
        sets = [{tuple(p) for p in ops.cosine_pairs(et, 0.9, part=r, nparts=nparts)[0].cpu().numpy().tolist()} for r in range(nparts)]
        assert sum(len(s) for s in sets) == len(full)
        assert set().union(*sets) == full
    aggs = np.random.default_rng(0).uniform(0, 10, n).round(2).tolist()
    gid, lead = group_duplicates(n, np.array(sorted(full), dtype=np.int64), aggs)
    wg, wl = og.cosine_groups(e, aggs, 0.9)
    assert gid.tolist() == wg.tolist() and lead.tolist() == wl.tolist()
    # truncated candidate buffer -> retried
    small = {tuple(p) for p in ops.cosine_pairs(et, 0.9, cap=5)[0].cpu().numpy().tolist()}
    assert small == full


@pytest.mark.parametrize("world", [2, 3, 8])
def test_block_decomposition_equals_the_single_scan(world):
    """The shard-against-shard blocks of the multi-GPU flow (fb_cosine_block: diagonal blocks with the triangle test,
    rectangular blocks, wrapped blocks whose columns lie before their rows), scanned one rank after the other on one GPU,
    find exactly the pairs of the single-matrix scan."""
    import torch
    from facet_b200 import ops
    from facet_b200.synth import synth_embeddings
    from facet_b200.utils.duplicate import shard_blocks
    n_local = 1000 if world != 8 else 520
    n = world * n_local
    emb = torch.from_numpy(synth_embeddings(n, seed=23 + world)).cuda()
    emb = emb[torch.randperm(n, generator=torch.Generator().manual_seed(1)).cuda()].contiguous()     # near-duplicates across shards
    want, want_s = ops.cosine_pairs(emb, 0.9)
    eb = ops.to_bf16(emb)
    got = []
    for rank in range(world):
        blocks = []
        for sa, a_lo, a_hi, sb, b_lo, b_hi, tri in shard_blocks(rank, world, n_local):
            blocks.append((eb[sa * n_local + a_lo: sa * n_local + a_hi], sa * n_local + a_lo,
                           eb[sb * n_local + b_lo: sb * n_local + b_hi], sb * n_local + b_lo, tri))
        p, s = ops.cosine_blocks(blocks, lambda: emb, emb.shape[1], 0.9)
        got.append(torch.cat([p.long(), s.view(torch.int32).long()[:, None]], dim=1))
    got = torch.cat(got).cpu().numpy()
    ref = torch.cat([want.long(), want_s.view(torch.int32).long()[:, None]], dim=1).cpu().numpy()
    assert len(ref) > 50
    assert sorted(map(tuple, got)) == sorted(map(tuple, ref))
