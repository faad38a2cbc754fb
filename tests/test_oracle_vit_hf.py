"""Pin the ViT-L/14 oracle (oracle/vit_torch.py) against an INDEPENDENT implementation of the same
tower: HuggingFace `transformers.CLIPVisionModelWithProjection` (installed in this image; its CLIP
vision tower is the architecture open_clip's `ViT-L-14` weights are published for, and HF ships the
converter between the two layouts).  The same random tensors are loaded into both — open_clip's
fused `in_proj_weight` / `in_proj_bias` split into HF's q / k / v projections — and all 24 layers
must agree.  CPU, fp32, about half a minute.

Reference call sites being pinned: processing/scorer.py:508-516 (model), :661-664 (encode_image).
"""
import pytest
import torch

from facet_b200.models.clip_vit import random_state_dict, WIDTH, LAYERS
from oracle import vit_torch

transformers = pytest.importorskip("transformers")


def hf_state_dict(sd, layers):
    """open_clip visual-tower keys -> HF CLIPVisionModelWithProjection keys (the mapping of HF's own
    convert_open_clip script: conv1 -> patch_embedding, in_proj rows [0:W|W:2W|2W:3W] -> q|k|v,
    ln_1/ln_2 -> layer_norm1/2, c_fc/c_proj -> fc1/fc2, proj [W,768] -> visual_projection.weight^T)."""
    out = {
        "vision_model.embeddings.class_embedding": sd["class_embedding"],
        "vision_model.embeddings.patch_embedding.weight": sd["conv1.weight"],
        "vision_model.embeddings.position_embedding.weight": sd["positional_embedding"],
        "vision_model.pre_layrnorm.weight": sd["ln_pre.weight"],
        "vision_model.pre_layrnorm.bias": sd["ln_pre.bias"],
        "vision_model.post_layernorm.weight": sd["ln_post.weight"],
        "vision_model.post_layernorm.bias": sd["ln_post.bias"],
        "visual_projection.weight": sd["proj"].T.contiguous(),
    }
    for l in range(layers):
        p, h = f"transformer.resblocks.{l}.", f"vision_model.encoder.layers.{l}."
        w, b = sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"]
        for i, name in enumerate(("q_proj", "k_proj", "v_proj")):
            out[h + f"self_attn.{name}.weight"] = w[i * WIDTH:(i + 1) * WIDTH]
            out[h + f"self_attn.{name}.bias"] = b[i * WIDTH:(i + 1) * WIDTH]
        out[h + "self_attn.out_proj.weight"] = sd[p + "attn.out_proj.weight"]
        out[h + "self_attn.out_proj.bias"] = sd[p + "attn.out_proj.bias"]
        out[h + "layer_norm1.weight"], out[h + "layer_norm1.bias"] = sd[p + "ln_1.weight"], sd[p + "ln_1.bias"]
        out[h + "layer_norm2.weight"], out[h + "layer_norm2.bias"] = sd[p + "ln_2.weight"], sd[p + "ln_2.bias"]
        out[h + "mlp.fc1.weight"], out[h + "mlp.fc1.bias"] = sd[p + "mlp.c_fc.weight"], sd[p + "mlp.c_fc.bias"]
        out[h + "mlp.fc2.weight"], out[h + "mlp.fc2.bias"] = sd[p + "mlp.c_proj.weight"], sd[p + "mlp.c_proj.bias"]
    return out


def test_oracle_tower_matches_hf_clip_vision_model_all_layers():
    from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection
    torch.manual_seed(0)
    sd = random_state_dict(5)
    cfg = CLIPVisionConfig(hidden_size=WIDTH, intermediate_size=4096, num_hidden_layers=LAYERS, num_attention_heads=16,
                           patch_size=14, image_size=224, projection_dim=768, hidden_act="gelu", layer_norm_eps=1e-5)
    model = CLIPVisionModelWithProjection(cfg).eval()
    missing, unexpected = model.load_state_dict(hf_state_dict(sd, LAYERS), strict=False)
    # every parameter of the HF tower is covered by the mapping (position_ids is a buffer HF rebuilds)
    assert not unexpected and all("position_ids" in k for k in missing), (missing, unexpected)
    x = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        want = model(pixel_values=x).image_embeds
        got = vit_torch.encode_image(sd, x)
    assert want.shape == got.shape == (2, 768)
    torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-4)
    cos = torch.nn.functional.cosine_similarity(got, want, dim=-1)
    assert float(cos.min()) > 0.999999
    # the per-layer hidden states agree too (not only the pooled output): layer outputs of the oracle
    # truncated to l layers vs HF's hidden_states[l]
    with torch.no_grad():
        hs = model.vision_model(pixel_values=x[:1], output_hidden_states=True).hidden_states
        for l in (1, 12, 24):
            pooled = model.visual_projection(model.vision_model.post_layernorm(hs[l][:, 0]))
            torch.testing.assert_close(vit_torch.encode_image(sd, x[:1], layers=l), pooled, rtol=1e-4, atol=1e-4)
