// Whole-tower driver: enqueues every kernel of one ViT-L/14 forward pass on a stream.
// See include/facet_b200.h (fb_vit_forward) for the contract.
#include "../../include/facet_b200.h"
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace fb {

namespace {
constexpr int kTokens = 257, kWidth = 1024, kMlp = 4096, kPatchKPad = 640;
inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct Workspace {
    float* x;        // [M][1024] fp32 residual stream
    void* xn;        // [M][1024] bf16 LayerNorm output / attention output
    void* qkv;       // [M][3072] bf16 (aliases the im2col matrix [B*256][640])
    void* attn;      // [M][1024] bf16
    void* h;         // [M][4096] bf16 (aliases the patch-embedding output [B*256][1024] fp32)
    float* stats;    // [M][8][2] fp32: per-row (sum, sum of squares) of the residual stream per 128-column slice (LayerNorm fold)
    size_t bytes;
};

Workspace carve(void* base, int batch) {
    const size_t M = (size_t)batch * kTokens;
    Workspace w;
    uint8_t* p = reinterpret_cast<uint8_t*>(base);
    size_t off = 0;
    auto take = [&](size_t n) { uint8_t* r = p ? p + off : nullptr; off += align256(n); return r; };
    w.x = reinterpret_cast<float*>(take(M * kWidth * 4));
    w.xn = take(M * kWidth * 2);
    w.qkv = take(M * 3 * kWidth * 2);
    w.attn = take(M * kWidth * 2);
    w.h = take(M * kMlp * 2);
    w.stats = reinterpret_cast<float*>(take(M * (kWidth / 128) * 2 * 4));
    w.bytes = off;
    return w;
}
}  // namespace

size_t vit_workspace_bytes(int batch) { return carve(nullptr, batch).bytes; }

int vit_forward(const fb_vit_weights* w, const float* d_clip_in, int batch, void* d_workspace, size_t workspace_bytes,
                float* d_features, float* d_embedding, float* d_aesthetic_raw, float* d_tag_sims, cudaStream_t st) {
    FB_REQUIRE(w && d_clip_in && d_workspace && d_features && d_embedding && d_aesthetic_raw, "fb_vit_forward: null pointer");
    FB_REQUIRE(batch >= 1 && w->n_layers >= 1 && w->layers, "fb_vit_forward: bad batch / layers");
    FB_REQUIRE(workspace_bytes >= vit_workspace_bytes(batch), "fb_vit_forward: workspace too small (%zu < %zu)",
               workspace_bytes, vit_workspace_bytes(batch));
    FB_REQUIRE((reinterpret_cast<uintptr_t>(d_workspace) & 255) == 0, "fb_vit_forward: workspace must be 256-byte aligned");
    Workspace ws = carve(d_workspace, batch);
    const int M = batch * kTokens;
    const int f16 = w->f16 ? 1 : 0;
    const int gflag = f16 ? FB_GEMM_F16_FLAG : 0;
    const int ln_out = f16 ? 2 : 1;
    int rc;
#define STEP_CAT(cat, call, n)          \
    do {                                \
        {                               \
            ProfScope ps_(cat, st);     \
            rc = (call);                \
        }                               \
        if (rc) return rc;              \
        count_launch(n);                \
    } while (0)
    // patch embedding: im2col (aliases qkv) -> GEMM -> fp32 [B*256][1024] (aliases h)
    STEP_CAT(PROF_IM2COL, launch_im2col_patch14(d_clip_in, batch, ws.qkv, f16, st), 1);
    STEP_CAT(PROF_GEMM, launch_gemm_bf16(ws.qkv, kPatchKPad, w->w_patch, kPatchKPad, batch * 256, kWidth, kPatchKPad, FB_GEMM_F32 | gflag, nullptr,
                          ws.h, kWidth, nullptr, 0, st), 1);
    bool fold = w->fused_ln != 0 && !getenv("FB_VIT_NO_LN_FOLD");
    for (int l = 0; l < w->n_layers && fold; ++l) {
        const fb_vit_layer& L = w->layers[l];
        fold = L.w_qkv_ln && L.s_qkv && L.c_qkv && L.w_fc_ln && L.s_fc && L.c_fc;
    }
    if (fold) {
        // LayerNorm folded into the GEMMs: the residual epilogues (and ln_pre) write a 16-bit copy of the residual stream (ws.xn)
        // and per-row sums (ws.stats); the QKV / fc GEMMs multiply that copy with gain-scaled weights and finish the
        // normalisation in their epilogues.  No pass over the residual stream between the GEMMs.
        constexpr int kSlots = kWidth / 128;
        GemmLnFold prod{ws.xn, kWidth, ws.stats, kSlots, nullptr, 0};
        STEP_CAT(PROF_LAYERNORM, launch_layernorm_pre_fold(reinterpret_cast<const float*>(ws.h), kWidth, M, w->ln_pre_g, w->ln_pre_b, w->class_emb,
                              w->pos_emb, ws.x, kWidth, ws.xn, f16, ws.stats, kSlots, st), 1);
        for (int l = 0; l < w->n_layers; ++l) {
            const fb_vit_layer& L = w->layers[l];
            GemmLnFold cq{nullptr, 0, ws.stats, kSlots, L.s_qkv, kWidth};
            STEP_CAT(PROF_GEMM, launch_gemm_bf16_ln(ws.xn, kWidth, L.w_qkv_ln, kWidth, M, 3 * kWidth, kWidth, FB_GEMM_BIAS_BF16 | gflag, L.c_qkv, ws.qkv,
                                  3 * kWidth, nullptr, 0, &cq, st), 1);
            STEP_CAT(PROF_ATTENTION, launch_attention_tc(ws.qkv, batch, ws.attn, f16, st), 1);
            STEP_CAT(PROF_GEMM, launch_gemm_bf16_ln(ws.attn, kWidth, L.w_out, kWidth, M, kWidth, kWidth, FB_GEMM_BIAS_RESIDUAL_F32 | gflag, L.b_out, ws.x,
                                  kWidth, ws.x, kWidth, &prod, st), 1);
            GemmLnFold cf{nullptr, 0, ws.stats, kSlots, L.s_fc, kWidth};
            STEP_CAT(PROF_GEMM, launch_gemm_bf16_ln(ws.xn, kWidth, L.w_fc_ln, kWidth, M, kMlp, kWidth, FB_GEMM_BIAS_GELU_BF16 | gflag, L.c_fc, ws.h, kMlp,
                                  nullptr, 0, &cf, st), 1);
            // the last block's output feeds only ln_post on the class token (fp32, vit_tail_kernel): no copy needed, but harmless
            STEP_CAT(PROF_GEMM, launch_gemm_bf16_ln(ws.h, kMlp, L.w_proj, kMlp, M, kWidth, kMlp, FB_GEMM_BIAS_RESIDUAL_F32 | gflag, L.b_proj, ws.x, kWidth,
                                  ws.x, kWidth, &prod, st), 1);
        }
    } else {
    // class token + positional embedding + ln_pre -> residual stream
    STEP_CAT(PROF_LAYERNORM, launch_layernorm(reinterpret_cast<const float*>(ws.h), kWidth, M, w->ln_pre_g, w->ln_pre_b, w->class_emb,
                          w->pos_emb, ws.x, kWidth, 0, st), 1);
    for (int l = 0; l < w->n_layers; ++l) {
        const fb_vit_layer& L = w->layers[l];
        STEP_CAT(PROF_LAYERNORM, launch_layernorm(ws.x, kWidth, M, L.ln1_g, L.ln1_b, nullptr, nullptr, ws.xn, kWidth, ln_out, st), 1);
        STEP_CAT(PROF_GEMM, launch_gemm_bf16(ws.xn, kWidth, L.w_qkv, kWidth, M, 3 * kWidth, kWidth, FB_GEMM_BIAS_BF16 | gflag, L.b_qkv, ws.qkv,
                              3 * kWidth, nullptr, 0, st), 1);
        STEP_CAT(PROF_ATTENTION, launch_attention_tc(ws.qkv, batch, ws.attn, f16, st), 1);
        STEP_CAT(PROF_GEMM, launch_gemm_bf16(ws.attn, kWidth, L.w_out, kWidth, M, kWidth, kWidth, FB_GEMM_BIAS_RESIDUAL_F32 | gflag, L.b_out, ws.x,
                              kWidth, ws.x, kWidth, st), 1);
        STEP_CAT(PROF_LAYERNORM, launch_layernorm(ws.x, kWidth, M, L.ln2_g, L.ln2_b, nullptr, nullptr, ws.xn, kWidth, ln_out, st), 1);
        STEP_CAT(PROF_GEMM, launch_gemm_bf16(ws.xn, kWidth, L.w_fc, kWidth, M, kMlp, kWidth, FB_GEMM_BIAS_GELU_BF16 | gflag, L.b_fc, ws.h, kMlp,
                              nullptr, 0, st), 1);
        STEP_CAT(PROF_GEMM, launch_gemm_bf16(ws.h, kMlp, L.w_proj, kMlp, M, kWidth, kMlp, FB_GEMM_BIAS_RESIDUAL_F32 | gflag, L.b_proj, ws.x, kWidth,
                              ws.x, kWidth, st), 1);
    }
    }
    STEP_CAT(PROF_TAIL, launch_vit_tail(ws.x, batch, w->ln_post_g, w->ln_post_b, w->proj, w->head_w1, w->head_b1, w->head_w2, w->head_b2,
                         w->tag_emb, w->n_tags, d_features, d_embedding, d_aesthetic_raw, d_tag_sims, st), 1);
#undef STEP_CAT
    return 0;
}

}  // namespace fb
