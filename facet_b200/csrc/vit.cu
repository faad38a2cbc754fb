// Non-GEMM kernels of the CLIP ViT-L/14 image tower and its heads.
//
// Replaces what `self.model.encode_image(inputs)` + `F.normalize` + `self.aesthetic_head`
// (processing/scorer.py:661-664) and `image_features @ self.text_embeddings.T`
// (models/tagger.py:99-101) execute through open_clip / PyTorch in the reference:
//   im2col_patch14_kernel   conv 14x14/14 patch embedding as the A operand of a GEMM
//   layernorm_kernel        ln_pre (with class-token / positional-embedding assembly), ln_1, ln_2
//   (attention lives in csrc/attention_tc.cu)
//   vit_tail_kernel         ln_post(CLS) @ proj -> features; L2-normalised embedding;
//                           MLP aesthetic head on the un-normalised features; tag similarities
// The residual stream stays in fp32; GEMM inputs are bf16 (csrc/gemm.cu).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "kernels.h"

namespace fb {

namespace {

constexpr int kWidth = 1024;
constexpr int kTokens = 257;
constexpr int kHeads = 16;
constexpr int kHeadDim = 64;
constexpr int kPatchK = 588;      // 3 * 14 * 14
constexpr int kPatchKPad = 640;

// ---------------------------------------------------------------------------------------------
template <bool F16>
__global__ void __launch_bounds__(256) im2col_patch14_kernel(const float* __restrict__ x, int batch,
                                                             uint32_t* __restrict__ out) {
    // out[(b*256 + py*16 + px)][k], k = c*196 + ky*14 + kx (conv weight [1024,3,14,14] flattened), zero padded to 640
    const long long total = (long long)batch * 256 * (kPatchKPad / 2);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int kp = (int)(i % (kPatchKPad / 2)) * 2;
        const long long row = i / (kPatchKPad / 2);
        const int p = (int)(row & 255);
        const int b = (int)(row >> 8);
        float v[2] = {0.f, 0.f};
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int k = kp + e;
            if (k < kPatchK) {
                const int c = k / 196, r = k - c * 196;
                const int ky = r / 14, kx = r - ky * 14;
                const int y = (p >> 4) * 14 + ky, xx = (p & 15) * 14 + kx;
                v[e] = __ldg(x + (((size_t)b * 3 + c) * 224 + y) * 224 + xx);
            }
        }
        if (F16) {
            __half2 h2 = __floats2half2_rn(v[0], v[1]);
            out[i] = *reinterpret_cast<uint32_t*>(&h2);
        } else {
            __nv_bfloat162 b2 = __floats2bfloat162_rn(v[0], v[1]);
            out[i] = *reinterpret_cast<uint32_t*>(&b2);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm over 1024 columns, one warp per row, eps = 1e-5 (two-pass, fp32).
//   ASSEMBLE: the input row is built on the fly as (token 0 ? class_emb : patch_out[b, t-1]) + pos[t]
//   OUT: 0 = fp32 output (residual stream), 1 = bf16, 2 = fp16 (GEMM operand)
template <bool ASSEMBLE, int OUT>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ in, long long ld_in, int rows,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        const float* __restrict__ cls, const float* __restrict__ pos,
                                                        void* __restrict__ out, long long ld_out) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    float v[32];
    if (ASSEMBLE) {
        const int b = row / kTokens, t = row - b * kTokens;
        const float* src = (t == 0) ? cls : in + ((size_t)b * 256 + (t - 1)) * ld_in;
        const float* pp = pos + (size_t)t * kWidth;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 a = *reinterpret_cast<const float4*>(src + j * 128 + lane * 4);
            const float4 p4 = __ldg(reinterpret_cast<const float4*>(pp + j * 128 + lane * 4));
            v[4 * j] = a.x + p4.x; v[4 * j + 1] = a.y + p4.y; v[4 * j + 2] = a.z + p4.z; v[4 * j + 3] = a.w + p4.w;
        }
    } else {
        const float* src = in + (size_t)row * ld_in;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 a = *reinterpret_cast<const float4*>(src + j * 128 + lane * 4);
            v[4 * j] = a.x; v[4 * j + 1] = a.y; v[4 * j + 2] = a.z; v[4 * j + 3] = a.w;
        }
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) s += v[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.0f / kWidth);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const float d = v[j] - mean;
        q += d * d;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * (1.0f / kWidth) + 1e-5f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma + j * 128 + lane * 4));
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(beta + j * 128 + lane * 4));
        const float o0 = (v[4 * j] - mean) * rstd * g4.x + b4.x;
        const float o1 = (v[4 * j + 1] - mean) * rstd * g4.y + b4.y;
        const float o2 = (v[4 * j + 2] - mean) * rstd * g4.z + b4.z;
        const float o3 = (v[4 * j + 3] - mean) * rstd * g4.w + b4.w;
        if (OUT == 1) {
            __nv_bfloat162 lo = __floats2bfloat162_rn(o0, o1), hi = __floats2bfloat162_rn(o2, o3);
            uint2 pk = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
            *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out) + (size_t)row * ld_out + j * 128 + lane * 4) = pk;
        } else if (OUT == 2) {
            __half2 lo = __floats2half2_rn(o0, o1), hi = __floats2half2_rn(o2, o3);
            uint2 pk = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
            *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(out) + (size_t)row * ld_out + j * 128 + lane * 4) = pk;
        } else {
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + (size_t)row * ld_out + j * 128 + lane * 4) =
                make_float4(o0, o1, o2, o3);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Tail: one CTA (256 threads) per image.
//   y = ln_post(x[b, 0, :]);  feat = y @ proj[1024,768]  (fp32, weights bf16-rounded like the GEMMs? no: fp32)
//   emb = feat / max(||feat||, 1e-12)                      F.normalize, scorer.py:663
//   raw = W2 . relu(W1 feat + b1) + b2                     aesthetic head on un-normalised features, scorer.py:664
//   sims = emb @ T^T                                       tagger.py:101
__global__ void __launch_bounds__(768) vit_tail_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                       const float* __restrict__ be, const float* __restrict__ proj /*[1024][768]*/,
                                                       const float* __restrict__ w1 /*[256][768]*/, const float* __restrict__ b1,
                                                       const float* __restrict__ w2 /*[256]*/, const float* __restrict__ b2,
                                                       const float* __restrict__ tags /*[ntags][768]*/, int ntags,
                                                       float* __restrict__ feat_out, float* __restrict__ emb_out,
                                                       float* __restrict__ raw_out, float* __restrict__ sims_out) {
    __shared__ float y[kWidth];
    __shared__ float feat[768];
    __shared__ float red[24];
    __shared__ float hid[256];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;   // 24 warps
    const float* row = x + (size_t)b * kTokens * kWidth;
    // ln_post on the class token (two-pass, fp32)
    float v0 = row[tid], v1 = (tid < 256) ? row[768 + tid] : 0.f;
    float s = v0 + v1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    float mean = 0.f;
    for (int i = 0; i < 24; ++i) mean += red[i];
    mean *= (1.0f / kWidth);
    __syncthreads();
    float d0 = v0 - mean, d1 = (tid < 256) ? (v1 - mean) : 0.f;
    float q = d0 * d0 + d1 * d1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    if (lane == 0) red[warp] = q;
    __syncthreads();
    float var = 0.f;
    for (int i = 0; i < 24; ++i) var += red[i];
    const float rstd = rsqrtf(var * (1.0f / kWidth) + 1e-5f);
    y[tid] = d0 * rstd * g[tid] + be[tid];
    if (tid < 256) y[768 + tid] = d1 * rstd * g[768 + tid] + be[768 + tid];
    __syncthreads();
    // projection: one output column per thread, coalesced rows of proj, 8 independent accumulators
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < kWidth; k += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(y[k + j], __ldg(proj + (size_t)(k + j) * 768 + tid), acc[j]);
    }
    const float f = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
    feat[tid] = f;
    float nrm = f * f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
    __syncthreads();            // red[] reads above are done, feat[] is complete after the next barrier
    if (lane == 0) red[warp] = nrm;
    __syncthreads();
    float tot = 0.f;
    for (int i = 0; i < 24; ++i) tot += red[i];
    const float inv = 1.0f / fmaxf(sqrtf(tot), 1e-12f);
    feat_out[(size_t)b * 768 + tid] = f;
    emb_out[(size_t)b * 768 + tid] = f * inv;
    // MLP head: warp per hidden unit, coalesced rows of w1
    for (int u = warp; u < 256; u += 24) {
        const float* wr = w1 + (size_t)u * 768;
        float h = 0.f;
#pragma unroll 8
        for (int k = lane; k < 768; k += 32) h = fmaf(__ldg(wr + k), feat[k], h);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
        if (lane == 0) hid[u] = fmaxf(h + b1[u], 0.f) * w2[u];
    }
    // tag similarities on the normalised embedding: warp per prompt
    for (int tg = warp; tg < ntags; tg += 24) {
        const float* tr = tags + (size_t)tg * 768;
        float d = 0.f;
#pragma unroll 8
        for (int k = lane; k < 768; k += 32) d = fmaf(feat[k] * inv, __ldg(tr + k), d);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
        if (lane == 0) sims_out[(size_t)b * ntags + tg] = d;
    }
    __syncthreads();
    if (warp == 0) {
        float r = 0.f;
        for (int k = lane; k < 256; k += 32) r += hid[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
        if (lane == 0) raw_out[b] = r + b2[0];
    }
}

// ---------------------------------------------------------------------------------------------
// Heads on stored embeddings (no tower): the per-image calls of the reference that start from the 3072-byte
// `clip_embedding` BLOB — `Facet.score_from_embedding` (processing/scorer.py:620-629: the MLP head applied to
// the vector as stored) and `CLIPTagger.get_tags_from_embedding` (models/tagger.py:99-101: emb @ T^T).
// One CTA (8 warps) per vector; fp32 throughout, warp per hidden unit / per prompt like the tail kernel.
__global__ void __launch_bounds__(256) embedding_heads_kernel(const float* __restrict__ x /*[n][768]*/,
                                                              const float* __restrict__ w1, const float* __restrict__ b1,
                                                              const float* __restrict__ w2, const float* __restrict__ b2,
                                                              const float* __restrict__ tags, int ntags,
                                                              float* __restrict__ raw_out, float* __restrict__ sims_out) {
    __shared__ float v[768];
    __shared__ float hid[256];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int k = tid; k < 768; k += 256) v[k] = x[(size_t)b * 768 + k];
    __syncthreads();
    if (raw_out) {
        for (int u = warp; u < 256; u += 8) {
            const float* wr = w1 + (size_t)u * 768;
            float h = 0.f;
#pragma unroll 8
            for (int k = lane; k < 768; k += 32) h = fmaf(__ldg(wr + k), v[k], h);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
            if (lane == 0) hid[u] = fmaxf(h + b1[u], 0.f) * w2[u];
        }
    }
    for (int tg = warp; tg < ntags; tg += 8) {
        const float* tr = tags + (size_t)tg * 768;
        float d = 0.f;
#pragma unroll 8
        for (int k = lane; k < 768; k += 32) d = fmaf(v[k], __ldg(tr + k), d);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
        if (lane == 0) sims_out[(size_t)b * ntags + tg] = d;
    }
    __syncthreads();
    if (raw_out && warp == 0) {
        float r = 0.f;
        for (int k = lane; k < 256; k += 32) r += hid[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
        if (lane == 0) raw_out[b] = r + b2[0];
    }
}

}  // namespace

int launch_im2col_patch14(const float* d_x, int batch, void* d_out, int f16, cudaStream_t stream) {
    FB_REQUIRE(d_x && d_out && batch >= 1, "fb_vit_im2col: bad arguments");
    const long long total = (long long)batch * 256 * (kPatchKPad / 2);
    int blocks = (int)((total + 255) / 256);
    if (blocks > sm_count() * 32) blocks = sm_count() * 32;
    if (f16) im2col_patch14_kernel<true><<<blocks, 256, 0, stream>>>(d_x, batch, reinterpret_cast<uint32_t*>(d_out));
    else im2col_patch14_kernel<false><<<blocks, 256, 0, stream>>>(d_x, batch, reinterpret_cast<uint32_t*>(d_out));
    FB_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_layernorm(const float* d_in, long long ld_in, int rows, const float* gamma, const float* beta,
                     const float* cls, const float* pos, void* d_out, long long ld_out, int out_bf16,
                     cudaStream_t stream) {
    FB_REQUIRE(d_in && gamma && beta && d_out && rows >= 1, "fb_vit_layernorm: bad arguments");
    const int blocks = (rows + 7) / 8;
    if (cls) {
        FB_REQUIRE(pos && !out_bf16, "fb_vit_layernorm: assembly mode writes the fp32 residual stream");
        layernorm_kernel<true, 0><<<blocks, 256, 0, stream>>>(d_in, ld_in, rows, gamma, beta, cls, pos, d_out, ld_out);
    } else if (out_bf16 == 1) {
        layernorm_kernel<false, 1><<<blocks, 256, 0, stream>>>(d_in, ld_in, rows, gamma, beta, nullptr, nullptr, d_out, ld_out);
    } else if (out_bf16 == 2) {
        layernorm_kernel<false, 2><<<blocks, 256, 0, stream>>>(d_in, ld_in, rows, gamma, beta, nullptr, nullptr, d_out, ld_out);
    } else {
        layernorm_kernel<false, 0><<<blocks, 256, 0, stream>>>(d_in, ld_in, rows, gamma, beta, nullptr, nullptr, d_out, ld_out);
    }
    FB_CUDA_OK(cudaGetLastError());
    return 0;
}

// ln_pre (class token + positional embedding assembly) for the LayerNorm-fold path: fp32 residual stream, its 16-bit copy and
// the (sum, sum of squares) of every output row in slot 0 of the row-sum array (the other slots are zeroed).
template <bool F16>
__global__ void __launch_bounds__(256) layernorm_pre_fold_kernel(const float* __restrict__ in, long long ld_in, int rows,
                                                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                 const float* __restrict__ cls, const float* __restrict__ pos,
                                                                 float* __restrict__ out, long long ld_out, uint16_t* __restrict__ out16,
                                                                 float* __restrict__ row_stats, int ln_slots) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    float v[32];
    const int b = row / kTokens, t = row - b * kTokens;
    const float* src = (t == 0) ? cls : in + ((size_t)b * 256 + (t - 1)) * ld_in;
    const float* pp = pos + (size_t)t * kWidth;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 a = *reinterpret_cast<const float4*>(src + j * 128 + lane * 4);
        const float4 p4 = __ldg(reinterpret_cast<const float4*>(pp + j * 128 + lane * 4));
        v[4 * j] = a.x + p4.x; v[4 * j + 1] = a.y + p4.y; v[4 * j + 2] = a.z + p4.z; v[4 * j + 3] = a.w + p4.w;
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) s += v[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.0f / kWidth);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const float d = v[j] - mean;
        q += d * d;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * (1.0f / kWidth) + 1e-5f);
    float os = 0.f, oq = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma + j * 128 + lane * 4));
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(beta + j * 128 + lane * 4));
        const float o0 = (v[4 * j] - mean) * rstd * g4.x + b4.x;
        const float o1 = (v[4 * j + 1] - mean) * rstd * g4.y + b4.y;
        const float o2 = (v[4 * j + 2] - mean) * rstd * g4.z + b4.z;
        const float o3 = (v[4 * j + 3] - mean) * rstd * g4.w + b4.w;
        *reinterpret_cast<float4*>(out + (size_t)row * ld_out + j * 128 + lane * 4) = make_float4(o0, o1, o2, o3);
        uint2 pk;
        if (F16) {
            __half2 lo = __floats2half2_rn(o0, o1), hi = __floats2half2_rn(o2, o3);
            pk = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
        } else {
            __nv_bfloat162 lo = __floats2bfloat162_rn(o0, o1), hi = __floats2bfloat162_rn(o2, o3);
            pk = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
        }
        *reinterpret_cast<uint2*>(out16 + (size_t)row * kWidth + j * 128 + lane * 4) = pk;
        os += (o0 + o1) + (o2 + o3);
        oq += fmaf(o0, o0, o1 * o1) + fmaf(o2, o2, o3 * o3);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        os += __shfl_xor_sync(0xffffffffu, os, o);
        oq += __shfl_xor_sync(0xffffffffu, oq, o);
    }
    if (lane < ln_slots) {
        const float2 v2 = lane == 0 ? make_float2(os, oq) : make_float2(0.f, 0.f);
        *reinterpret_cast<float2*>(row_stats + ((size_t)row * ln_slots + lane) * 2) = v2;
    }
}

int launch_layernorm_pre_fold(const float* d_in, long long ld_in, int rows, const float* gamma, const float* beta, const float* cls,
                              const float* pos, float* d_out, long long ld_out, void* d_out16, int f16, float* d_row_stats, int ln_slots,
                              cudaStream_t stream) {
    FB_REQUIRE(d_in && gamma && beta && cls && pos && d_out && d_out16 && d_row_stats && rows >= 1 && ln_slots >= 1 && ln_slots <= 32,
               "fb_vit_layernorm (fold): bad arguments");
    const int blocks = (rows + 7) / 8;
    if (f16) layernorm_pre_fold_kernel<true><<<blocks, 256, 0, stream>>>(d_in, ld_in, rows, gamma, beta, cls, pos, d_out, ld_out,
                                                                         reinterpret_cast<uint16_t*>(d_out16), d_row_stats, ln_slots);
    else layernorm_pre_fold_kernel<false><<<blocks, 256, 0, stream>>>(d_in, ld_in, rows, gamma, beta, cls, pos, d_out, ld_out,
                                                                      reinterpret_cast<uint16_t*>(d_out16), d_row_stats, ln_slots);
    FB_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_vit_tail(const float* d_x, int batch, const float* g, const float* be, const float* proj, const float* w1,
                    const float* b1, const float* w2, const float* b2, const float* tags, int ntags, float* feat,
                    float* emb, float* raw, float* sims, cudaStream_t stream) {
    FB_REQUIRE(d_x && g && be && proj && w1 && b1 && w2 && b2 && feat && emb && raw && batch >= 1, "fb_vit_tail: bad arguments");
    FB_REQUIRE(ntags == 0 || (tags && sims), "fb_vit_tail: tag matrix / output missing");
    vit_tail_kernel<<<batch, 768, 0, stream>>>(d_x, g, be, proj, w1, b1, w2, b2, tags, ntags, feat, emb, raw, sims);
    FB_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_embedding_heads(const float* d_x, int n, const float* w1, const float* b1, const float* w2, const float* b2,
                           const float* tags, int ntags, float* raw, float* sims, cudaStream_t stream) {
    FB_REQUIRE(d_x && n >= 1, "fb_embedding_heads: bad arguments");
    FB_REQUIRE(!raw || (w1 && b1 && w2 && b2), "fb_embedding_heads: head weights missing");
    FB_REQUIRE(ntags == 0 || (tags && sims), "fb_embedding_heads: tag matrix / output missing");
    FB_REQUIRE(raw || ntags > 0, "fb_embedding_heads: nothing to compute");
    embedding_heads_kernel<<<n, 256, 0, stream>>>(d_x, w1, b1, w2, b2, tags, ntags, raw, sims);
    FB_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace fb
