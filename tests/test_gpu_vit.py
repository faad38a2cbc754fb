"""GPU numerics for the tensor-core path: GEMM / LayerNorm / attention vs plain PyTorch fp32, and
the whole ViT-L/14 tower + heads vs the fp32 oracle (north_star: cosine >= 0.999, aesthetic +-0.01)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rand_bf16(shape, seed, scale=1.0, dtype=None):
    import torch
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, device="cuda", generator=g) * scale).to(dtype or torch.bfloat16)


@pytest.mark.parametrize("m,n,k", [(128, 256, 64), (300, 256, 128), (1000, 768, 1024), (257 * 8, 3072, 1024),
                                   (2056, 1024, 4096), (129, 32, 640), (4096, 1024, 640),
                                   # CTA-pair form (M >= 512, N % 256 == 0): exactly two pairs, an odd number of row
                                   # tiles with a ragged last one, many column tiles, a single k-block, a long K
                                   (512, 256, 64), (513, 512, 128), (640, 4096, 1024), (1283, 256, 64), (768, 256, 8192)])
@pytest.mark.parametrize("dt", ["bf16", "fp16"])
def test_gemm_all_epilogues(m, n, k, dt):
    import torch
    from facet_b200 import ops
    dtype = torch.float16 if dt == "fp16" else torch.bfloat16
    a, b = _rand_bf16((m, k), 1, dtype=dtype), _rand_bf16((n, k), 2, scale=k ** -0.5, dtype=dtype)
    bias = torch.randn(n, device="cuda")
    res = torch.randn(m, n, device="cuda")
    ref = a.float() @ b.float().T
    tol = dict(rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(ops.gemm_bf16(a, b, ops.GEMM_F32), ref, rtol=1e-4, atol=1e-3)
    torch.testing.assert_close(ops.gemm_bf16(a, b, ops.GEMM_F32, bias=bias), ref + bias, rtol=1e-4, atol=1e-3)
    torch.testing.assert_close(ops.gemm_bf16(a, b, ops.GEMM_BIAS_BF16, bias=bias).float(), ref + bias, **tol)
    torch.testing.assert_close(ops.gemm_bf16(a, b, ops.GEMM_BIAS_GELU_BF16, bias=bias).float(),
                               torch.nn.functional.gelu(ref + bias), **tol)
    out = res.clone()
    ops.gemm_bf16(a, b, ops.GEMM_BIAS_RESIDUAL_F32, bias=bias, residual=out, out=out)   # in place
    torch.testing.assert_close(out, ref + bias + res, rtol=1e-4, atol=1e-3)


def test_layernorm_and_assembly():
    import torch
    from facet_b200 import ops
    x = torch.randn(1000, 1024, device="cuda") * 3 + 0.5
    g, b = torch.randn(1024, device="cuda"), torch.randn(1024, device="cuda")
    ref = torch.nn.functional.layer_norm(x, (1024,), g, b, 1e-5)
    torch.testing.assert_close(ops.vit_layernorm(x, g, b, out_bf16=False), ref, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(ops.vit_layernorm(x, g, b, out_bf16=True).float(), ref, rtol=1e-2, atol=1e-2)
    pe = torch.randn(3 * 256, 1024, device="cuda")
    cls, pos = torch.randn(1024, device="cuda"), torch.randn(257, 1024, device="cuda")
    tok = torch.cat([cls.expand(3, 1, 1024), pe.reshape(3, 256, 1024)], 1) + pos
    ref = torch.nn.functional.layer_norm(tok, (1024,), g, b, 1e-5).reshape(-1, 1024)
    got = ops.vit_layernorm(pe, g, b, out_bf16=False, class_emb=cls, pos_emb=pos, rows=3 * 257)
    torch.testing.assert_close(got, ref, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("variant", ["tcgen05-bf16", "tcgen05-fp16"])
def test_attention_vs_torch(variant):
    import torch
    from facet_b200 import ops
    bsz = 3
    qkv = _rand_bf16((bsz * 257, 3072), 5, scale=1.5, dtype=torch.float16 if variant.endswith("fp16") else None)
    got = ops.vit_attention(qkv, bsz).float().reshape(bsz, 257, 16, 64)
    q, k, v = qkv.float().reshape(bsz, 257, 3, 16, 64).unbind(2)
    att = torch.softmax(torch.einsum("bqhd,bkhd->bhqk", q, k) * 0.125, dim=-1)
    ref = torch.einsum("bhqk,bkhd->bqhd", att, v)
    torch.testing.assert_close(got, ref, rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize("dt,aest_tol", [
    ("fp16", 0.01),
    # bf16 operands: cosine 0.999999 but 0.004 mean / 0.013 max aesthetic error on this random-init head (measured,
    # DESIGN.md §7) — outside the north_star's 0.01, so the bound is NOT loosened and the case is an expected failure.
    # The default (and everything the bench measures) is fp16, the precision the reference itself runs on CUDA
    # (`self.model.half()`, processing/scorer.py:515).
    pytest.param("bf16", 0.01, marks=pytest.mark.xfail(reason="bf16 activations: aesthetic max error 0.013 > 0.01", strict=False)),
])
def test_vit_tower_vs_fp32_oracle(dt, aest_tol):
    """north_star tolerances: cosine >= 0.999 and aesthetic within 0.01 of the fp32 oracle."""
    import torch
    from facet_b200.models.clip_vit import ClipVitL14, random_state_dict
    from oracle import vit_torch
    sd = random_state_dict(0)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(8, 3, 224, 224, generator=g)
    tags = torch.nn.functional.normalize(torch.randn(240, 768, generator=torch.Generator().manual_seed(7)), dim=-1)
    sd_gpu = {k: v.cuda() for k, v in sd.items()}
    ref = vit_torch.score_batch(sd_gpu, x.cuda(), tags.cuda())       # fp32 oracle (TF32 off by default)
    model = ClipVitL14(sd, tag_embeddings=tags.numpy(), dtype=dt)
    out = model.encode(x.cuda())
    cos = torch.nn.functional.cosine_similarity(out["embedding"], ref["embedding"], dim=-1)
    assert float(cos.min()) >= 0.999, cos
    aest = ((out["aesthetic_raw"] + 1) * 5).clamp(0, 10)
    assert float((aest - ref["aesthetic"]).abs().max()) <= aest_tol, (aest, ref["aesthetic"])
    assert float((out["tag_sims"] - ref["tag_sims"]).abs().max()) <= 5e-3
    nrm = out["embedding"].norm(dim=-1)
    torch.testing.assert_close(nrm, torch.ones_like(nrm), rtol=1e-5, atol=1e-5)


def test_embedding_heads_match_torch():
    """fb_embedding_heads (score_from_embedding / tagger similarities on stored embeddings) vs plain fp32 PyTorch."""
    import torch
    from facet_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(4)
    v = torch.nn.functional.normalize(torch.randn(5, 768, device="cuda", generator=g), dim=-1)
    w1, b1 = torch.randn(256, 768, device="cuda", generator=g) * 768 ** -0.5, torch.randn(256, device="cuda", generator=g) * 0.02
    w2, b2 = torch.randn(256, device="cuda", generator=g) * 256 ** -0.5, torch.randn(1, device="cuda", generator=g) * 0.02
    tags = torch.nn.functional.normalize(torch.randn(240, 768, device="cuda", generator=g), dim=-1)
    raw, sims = ops.embedding_heads(v, head=(w1, b1, w2, b2), tag_embeddings=tags)
    ref_raw = torch.relu(v @ w1.T + b1) @ w2 + b2
    torch.testing.assert_close(raw, ref_raw, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(sims, v @ tags.T, rtol=1e-5, atol=1e-5)
    assert ops.embedding_heads(v, tag_embeddings=tags)[0] is None and ops.embedding_heads(v, head=(w1, b1, w2, b2))[1] is None
