"""GPU: repeated launches must give bit-identical results.  The technical pass only uses integer atomics, the
tensor-core kernels (CTA-pair GEMM, attention with two threads per row) no atomics at all, so any run-to-run
difference would be a race in their barrier / shared-memory hand-offs."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_technical_pass_is_deterministic():
    import torch
    from facet_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(1)
    fr = torch.randint(0, 256, (6, 1000, 1544, 3), dtype=torch.uint8, device="cuda", generator=g)
    luma0 = torch.empty(fr.shape[:3], dtype=torch.uint8, device="cuda")
    first = [t.clone() for t in ops.tech_stats_raw(fr, luma_out=luma0)]
    for _ in range(20):
        luma = torch.empty_like(luma0)
        got = ops.tech_stats_raw(fr, luma_out=luma)
        for a, b in zip(first, got):
            assert torch.equal(a, b)
        assert torch.equal(luma, luma0)


def test_vit_tower_is_deterministic():
    import torch
    from facet_b200.models.clip_vit import ClipVitL14, random_state_dict
    model = ClipVitL14(random_state_dict(0, layers=4))
    x = torch.randn(24, 3, 224, 224, device="cuda", generator=torch.Generator(device="cuda").manual_seed(2))
    first = {k: v.clone() for k, v in model.encode(x).items() if v is not None}
    for _ in range(8):
        got = model.encode(x)
        for k, v in first.items():
            assert torch.equal(v, got[k]), k


def test_gemm_pair_kernel_is_deterministic():
    import torch
    from facet_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    a = torch.randn((2056, 1024), device="cuda", generator=g).to(torch.float16)
    b = (torch.randn((3072, 1024), device="cuda", generator=g) / 32).to(torch.float16)
    bias = torch.randn(3072, device="cuda", generator=g)
    first = ops.gemm_bf16(a, b, ops.GEMM_BIAS_GELU_BF16, bias=bias).clone()
    for _ in range(10):
        assert torch.equal(ops.gemm_bf16(a, b, ops.GEMM_BIAS_GELU_BF16, bias=bias), first)


def test_cosine_pair_set_is_deterministic():
    import torch
    from facet_b200 import ops
    from facet_b200.synth import synth_embeddings
    e = torch.from_numpy(synth_embeddings(6000, seed=5)).cuda()
    def run():
        p, _ = ops.cosine_pairs(e, 0.90)
        return set(map(tuple, p.cpu().numpy().tolist()))
    first = run()
    for _ in range(5):
        assert run() == first
