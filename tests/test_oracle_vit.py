"""Check the fp32 ViT restatement (oracle/vit_torch.py) against torch's own MultiheadAttention /
functional primitives on a 2-layer slice (CPU, seconds)."""
import torch

from facet_b200.models.clip_vit import random_state_dict
from oracle import vit_torch


def test_block_matches_nn_multihead_attention():
    sd = random_state_dict(1, layers=2)
    x = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        got = vit_torch.encode_image(sd, x)
        # independent path: nn.MultiheadAttention modules loaded with the same tensors
        t = torch.nn.functional.conv2d(x, sd["conv1.weight"], stride=14).flatten(2).transpose(1, 2)
        t = torch.cat([sd["class_embedding"].expand(2, 1, 1024), t], 1) + sd["positional_embedding"]
        t = torch.nn.functional.layer_norm(t, (1024,), sd["ln_pre.weight"], sd["ln_pre.bias"])
        for l in range(2):
            p = f"transformer.resblocks.{l}."
            mha = torch.nn.MultiheadAttention(1024, 16, batch_first=True)
            mha.in_proj_weight.copy_(sd[p + "attn.in_proj_weight"]); mha.in_proj_bias.copy_(sd[p + "attn.in_proj_bias"])
            mha.out_proj.weight.copy_(sd[p + "attn.out_proj.weight"]); mha.out_proj.bias.copy_(sd[p + "attn.out_proj.bias"])
            y = torch.nn.functional.layer_norm(t, (1024,), sd[p + "ln_1.weight"], sd[p + "ln_1.bias"])
            t = t + mha(y, y, y, need_weights=False)[0]
            y = torch.nn.functional.layer_norm(t, (1024,), sd[p + "ln_2.weight"], sd[p + "ln_2.bias"])
            y = torch.nn.functional.gelu(torch.nn.functional.linear(y, sd[p + "mlp.c_fc.weight"], sd[p + "mlp.c_fc.bias"]))
            t = t + torch.nn.functional.linear(y, sd[p + "mlp.c_proj.weight"], sd[p + "mlp.c_proj.bias"])
        want = torch.nn.functional.layer_norm(t[:, 0], (1024,), sd["ln_post.weight"], sd["ln_post.bias"]) @ sd["proj"]
    torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-4)
    out = vit_torch.score_batch(sd, x, torch.nn.functional.normalize(torch.randn(5, 768), dim=-1))
    assert out["embedding"].shape == (2, 768) and out["tag_sims"].shape == (2, 5)
    assert float(out["aesthetic"].min()) >= 0 and float(out["aesthetic"].max()) <= 10
