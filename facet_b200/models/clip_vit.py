"""CLIP ViT-L/14 image tower + MLP aesthetic head + tag similarities on the B200 kernels.

Stands in for the objects `Facet.__init__` builds at processing/scorer.py:508-516 (`self.model` =
open_clip 'ViT-L-14') and :571-585 (`self.aesthetic_head`), for the calls made by
`get_aesthetic_and_quality_batch` (scorer.py:640-673).  open_clip is a third-party package that
is neither vendored in the reference nor installed here; the architecture is restated from its
published `VisionTransformer` (conv 14x14/14 patch embedding without bias, class token, learned
positional embedding, ln_pre, 24 pre-LN residual blocks with fused in_proj / out_proj and an
erf-GELU MLP 1024-4096-1024, ln_post on the class token, bias-free projection to 768).

Weights come as a plain ``state_dict`` with open_clip's visual-tower key names (``conv1.weight``,
``class_embedding``, ``positional_embedding``, ``ln_pre.*``, ``transformer.resblocks.N.*``,
``ln_post.*``, ``proj``) so a real checkpoint can be loaded as is; there is no network here, so
tests and the bench use ``random_state_dict(seed)``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _lib

WIDTH, LAYERS, HEADS, MLP, TOKENS, OUT_DIM, PATCH_K, PATCH_K_PAD = 1024, 24, 16, 4096, 257, 768, 588, 640


def random_state_dict(seed: int = 0, layers: int = LAYERS, dtype=None):
    """Random-init visual tower + aesthetic head with open_clip's init scales (CPU float32).

    Biases get a small random value (open_clip zero-inits them) so the bias paths are exercised.
    The five GEMM weight families (conv1, in_proj, out_proj, c_fc, c_proj) are drawn and then
    rounded to values representable in bf16 and fp16: the tower stores them in 16 bits, so the oracle
    and the CUDA path hold bit-identical weights and the comparison measures arithmetic, not storage.
    """
    import torch
    g = torch.Generator().manual_seed(seed)

    def rn(*shape, std=1.0):
        return torch.randn(*shape, generator=g, dtype=torch.float32) * std

    def rw(*shape, std=1.0):
        # representable in both 16-bit storage formats the tower supports (bf16 and fp16)
        return rn(*shape, std=std).to(torch.bfloat16).to(torch.float16).to(torch.float32)

    sd = {}
    scale = WIDTH ** -0.5
    sd["conv1.weight"] = rw(WIDTH, 3, 14, 14, std=(3 * 14 * 14) ** -0.5)
    sd["class_embedding"] = rn(WIDTH, std=scale)
    sd["positional_embedding"] = rn(TOKENS, WIDTH, std=scale)
    for name in ("ln_pre", "ln_post"):
        sd[f"{name}.weight"] = 1.0 + rn(WIDTH, std=0.02)
        sd[f"{name}.bias"] = rn(WIDTH, std=0.02)
    attn_std = WIDTH ** -0.5
    proj_std = (WIDTH ** -0.5) * ((2 * LAYERS) ** -0.5)
    fc_std = (2 * WIDTH) ** -0.5
    for l in range(layers):
        p = f"transformer.resblocks.{l}."
        sd[p + "ln_1.weight"] = 1.0 + rn(WIDTH, std=0.02)
        sd[p + "ln_1.bias"] = rn(WIDTH, std=0.02)
        sd[p + "ln_2.weight"] = 1.0 + rn(WIDTH, std=0.02)
        sd[p + "ln_2.bias"] = rn(WIDTH, std=0.02)
        sd[p + "attn.in_proj_weight"] = rw(3 * WIDTH, WIDTH, std=attn_std)
        sd[p + "attn.in_proj_bias"] = rn(3 * WIDTH, std=0.02)
        sd[p + "attn.out_proj.weight"] = rw(WIDTH, WIDTH, std=proj_std)
        sd[p + "attn.out_proj.bias"] = rn(WIDTH, std=0.02)
        sd[p + "mlp.c_fc.weight"] = rw(MLP, WIDTH, std=fc_std)
        sd[p + "mlp.c_fc.bias"] = rn(MLP, std=0.02)
        sd[p + "mlp.c_proj.weight"] = rw(WIDTH, MLP, std=proj_std)
        sd[p + "mlp.c_proj.bias"] = rn(WIDTH, std=0.02)
    sd["proj"] = rn(WIDTH, OUT_DIM, std=scale)
    # aesthetic head: Sequential(Linear(768,256), ReLU, Linear(256,1)) left at its random init in the
    # reference (scorer.py:578-583, strict=False load of mismatching keys)
    sd["aesthetic_head.0.weight"] = rn(256, OUT_DIM, std=OUT_DIM ** -0.5)
    sd["aesthetic_head.0.bias"] = rn(256, std=0.02)
    sd["aesthetic_head.2.weight"] = rn(1, 256, std=256 ** -0.5)
    sd["aesthetic_head.2.bias"] = rn(1, std=0.02)
    return sd


class ClipVitL14:
    """Device-resident packed weights + workspace; ``encode`` runs the whole tower in one C call."""

    def __init__(self, state_dict, tag_embeddings=None, device=None, dtype="fp16", fold_layernorm=True):
        """dtype: 16-bit format of GEMM weights and activations.  "fp16" (default) is the precision the
        reference itself runs on CUDA (`self.model.half()`, processing/scorer.py:515) and keeps the
        aesthetic score within 0.01 of the fp32 oracle with a wide margin; "bf16" has the same
        tensor-core rate and 8x coarser rounding (cosine stays >= 0.9999, aesthetic within ~0.015)."""
        torch = _lib.require_cuda()
        if dtype not in ("fp16", "bf16"):
            raise ValueError("dtype must be 'fp16' or 'bf16'")
        self.dtype = dtype
        t16 = torch.float16 if dtype == "fp16" else torch.bfloat16
        self._lib = _lib.load()
        self.device = torch.device(device or f"cuda:{torch.cuda.current_device()}")
        sd = state_dict
        self.n_layers = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("transformer.resblocks."))
        keep = []

        def f32(t):
            x = t.detach().to(self.device, torch.float32).contiguous()
            keep.append(x)
            return x

        def bf16(t):
            x = t.detach().to(self.device, torch.float32).to(t16).contiguous()
            keep.append(x)
            return x

        wp = torch.zeros((WIDTH, PATCH_K_PAD), dtype=torch.float32)
        wp[:, :PATCH_K] = sd["conv1.weight"].reshape(WIDTH, PATCH_K).float()
        self.layers = (_lib.VitLayer * self.n_layers)()
        for l in range(self.n_layers):
            p = f"transformer.resblocks.{l}."
            L = self.layers[l]
            L.ln1_g, L.ln1_b = f32(sd[p + "ln_1.weight"]).data_ptr(), f32(sd[p + "ln_1.bias"]).data_ptr()
            L.ln2_g, L.ln2_b = f32(sd[p + "ln_2.weight"]).data_ptr(), f32(sd[p + "ln_2.bias"]).data_ptr()
            L.w_qkv, L.b_qkv = bf16(sd[p + "attn.in_proj_weight"]).data_ptr(), f32(sd[p + "attn.in_proj_bias"]).data_ptr()
            L.w_out, L.b_out = bf16(sd[p + "attn.out_proj.weight"]).data_ptr(), f32(sd[p + "attn.out_proj.bias"]).data_ptr()
            L.w_fc, L.b_fc = bf16(sd[p + "mlp.c_fc.weight"]).data_ptr(), f32(sd[p + "mlp.c_fc.bias"]).data_ptr()
            L.w_proj, L.b_proj = bf16(sd[p + "mlp.c_proj.weight"]).data_ptr(), f32(sd[p + "mlp.c_proj.bias"]).data_ptr()
            if fold_layernorm:
                # LayerNorm(x) W^T + b = rstd (x (W * gamma)^T - mean * s) + c, s = row sums of the 16-bit (W * gamma), c = W beta + b
                for ln, wk, bk, names in (("ln_1", "attn.in_proj_weight", "attn.in_proj_bias", ("w_qkv_ln", "s_qkv", "c_qkv")),
                                          ("ln_2", "mlp.c_fc.weight", "mlp.c_fc.bias", ("w_fc_ln", "s_fc", "c_fc"))):
                    wt = sd[p + wk].detach().to(self.device, torch.float32)
                    gam = sd[p + ln + ".weight"].detach().to(self.device, torch.float32)
                    bet = sd[p + ln + ".bias"].detach().to(self.device, torch.float32)
                    w_ln = bf16(wt * gam[None, :])
                    s_n = f32(w_ln.to(torch.float64).sum(dim=1).to(torch.float32))
                    c_n = f32((wt.to(torch.float64) @ bet.to(torch.float64) + sd[p + bk].detach().to(self.device, torch.float64)).to(torch.float32))
                    setattr(L, names[0], w_ln.data_ptr())
                    setattr(L, names[1], s_n.data_ptr())
                    setattr(L, names[2], c_n.data_ptr())
        w = _lib.VitWeights()
        w.w_patch = bf16(wp).data_ptr()
        w.class_emb = f32(sd["class_embedding"]).data_ptr()
        w.pos_emb = f32(sd["positional_embedding"]).data_ptr()
        w.ln_pre_g, w.ln_pre_b = f32(sd["ln_pre.weight"]).data_ptr(), f32(sd["ln_pre.bias"]).data_ptr()
        w.ln_post_g, w.ln_post_b = f32(sd["ln_post.weight"]).data_ptr(), f32(sd["ln_post.bias"]).data_ptr()
        w.proj = f32(sd["proj"]).data_ptr()
        w.head_w1, w.head_b1 = f32(sd["aesthetic_head.0.weight"]).data_ptr(), f32(sd["aesthetic_head.0.bias"]).data_ptr()
        w.head_w2, w.head_b2 = f32(sd["aesthetic_head.2.weight"].reshape(-1)).data_ptr(), f32(sd["aesthetic_head.2.bias"]).data_ptr()
        self.n_tags = 0
        if tag_embeddings is not None:
            te = f32(torch.as_tensor(np.asarray(tag_embeddings, dtype=np.float32)))
            w.tag_emb = te.data_ptr()
            self.n_tags = int(te.shape[0])
        w.n_tags = self.n_tags
        w.f16 = 1 if dtype == "fp16" else 0
        w.fused_ln = 1 if fold_layernorm else 0
        w.n_layers = self.n_layers
        w.layers = C.cast(self.layers, C.POINTER(_lib.VitLayer))
        self.weights = w
        self._keep = keep
        self._ws = None
        self._ws_batch = 0

    def _workspace(self, batch):
        torch = _lib.require_cuda()
        if self._ws is None or batch > self._ws_batch:
            nbytes = int(self._lib.fb_vit_workspace_bytes(batch))
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self._ws_batch = batch
        return self._ws

    def encode(self, clip_in):
        """clip_in: CUDA float32 [B,3,224,224] -> dict of CUDA float32 tensors:
        features [B,768], embedding [B,768] (L2-normalised), aesthetic_raw [B], tag_sims [B,n_tags]."""
        torch = _lib.require_cuda()
        x = clip_in.to(self.device, torch.float32).contiguous()
        b = int(x.shape[0])
        ws = self._workspace(b)
        feats = torch.empty((b, OUT_DIM), dtype=torch.float32, device=self.device)
        emb = torch.empty_like(feats)
        raw = torch.empty((b,), dtype=torch.float32, device=self.device)
        sims = torch.empty((b, max(self.n_tags, 1)), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.fb_vit_forward(C.byref(self.weights), C.c_void_p(x.data_ptr()), b,
                                                C.c_void_p(ws.data_ptr()), ws.numel(), C.c_void_p(feats.data_ptr()),
                                                C.c_void_p(emb.data_ptr()), C.c_void_p(raw.data_ptr()),
                                                C.c_void_p(sims.data_ptr()) if self.n_tags else None,
                                                _lib.stream_ptr()), "fb_vit_forward")
        return {"features": feats, "embedding": emb, "aesthetic_raw": raw,
                "tag_sims": sims[:, : self.n_tags] if self.n_tags else None}
