// Attention of the ViT-L/14 tower on tcgen05: softmax(Q K^T / 8) V per (image, head), 257 tokens, d = 64.
// (what open_clip's nn.MultiheadAttention does inside `encode_image`, processing/scorer.py:662.)
//
// Persistent kernel, one CTA per SM, 512 threads = two independent groups of 8 warps; each group walks
// its own queue of (image, head) pairs with its own shared-memory tiles and its own 256 TMEM
// columns, so the tensor-core / TMA phases of one group overlap the softmax of the other.  Inside a
// group two threads share a query row (warps w and w+4 address the same 32 TMEM lanes): one takes the
// scores of keys 0..127, the other those of keys 128..255; row maximum and sum are exchanged through
// shared memory.
// Per (image, head), for the two 128-row query tiles t = 0, 1:
//   TMA      Q (2 x 128 rows), K and V (256 rows each) from the fused qkv buffer, 128-byte swizzle;
//            the next pair's Q/K (V) loads are issued as soon as the last MMA reading them retires
//   tcgen05  S = Q_t K^T           M=128, N=256, K=64, fp32 accumulator in TMEM columns [0,256)
//   softmax  two threads per query row (tcgen05.ld), exact fp32 max / exp2 / sum; P is written back
//            to TMEM as packed 16-bit pairs over S columns its writer has already consumed
//            (tcgen05.st): keys 0..127 -> columns [0,64), keys 128..255 -> columns [128,192)
//   tcgen05  O = P V               A operand from TMEM, V as an MN-major B operand straight from
//            the TMA tile; M=128, N=64, K=256; accumulator in TMEM columns [64,128)
//   the 257th token is handled on the CUDA cores: as a key (one extra score per row folded into the
//   softmax, one rank-1 update in the epilogue) and as a query (one row against all 257 keys).  The query
//   part runs while the group's P V products are in flight (scores and softmax of the row behind P V of
//   tile 0, its output row behind P V of tile 1), with mixed-precision FMAs straight on the 16-bit
//   operands (FHFMA: fp16 x fp16 + fp32, no conversions); its q / k / v rows are fetched one pair ahead
#include <cuda_bf16.h>

#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace fb {

namespace {

constexpr int kTok = 257, kW = 1024, kD = 64;
constexpr int kThreadsAttn = 512;      // two groups of 256 threads
constexpr float kScaleLog2 = 0.125f * 1.4426950408889634f;

// per-warpgroup shared memory map (bytes)
constexpr int kOffQ = 0;                 // 2 x [128][64] bf16, 128B swizzle
constexpr int kOffK = 32768;             // [256][64]
constexpr int kOffV = 65536;             // [256][64]
constexpr int kOffX = 98304;             // scalars, exchange arrays, barriers
constexpr int kWgBytes = 98304 + 8192;
constexpr int kSmemAttn = 2 * kWgBytes + 1024;

struct XArea {
    uint32_t q256h[32], k256h[32];   // q / k rows of the 257th token as 16-bit pairs
    float v256[64];
    float cls_p[264];
    float cls_red[16];
    float cls_o[64];
    float s256[2][128];      // score of query row r of tile t against key 256
    float xmax[2][128];      // row maximum / sum of each column half
    float xsum[2][128];
    uint64_t bar_qk, bar_v, bar_s, bar_pv;
};


// 16-byte chunk c of row r in a [rows][64] bf16 tile stored with the 128-byte swizzle
__device__ __forceinline__ const uint4* sw_chunk(const uint8_t* tile, int r, int c) {
    return reinterpret_cast<const uint4*>(tile + r * 128 + ((c ^ (r & 7)) << 4));
}

// acc + lo(a) lo(b) + hi(a) hi(b) on 16-bit pairs: fp16 products are exact in fp32, so the mixed-precision FMA
// (FHFMA with half selectors) gives what convert-then-fmaf gives, without the conversions
template <bool F16>
__device__ __forceinline__ float mac2(uint32_t a, uint32_t b, float acc) {
    if (F16) {
        asm("{.reg .b16 al, ah, bl, bh;\n\t"
            "mov.b32 {al, ah}, %1;\n\t"
            "mov.b32 {bl, bh}, %2;\n\t"
            "fma.rn.f32.f16 %0, al, bl, %0;\n\t"
            "fma.rn.f32.f16 %0, ah, bh, %0;}"
            : "+f"(acc) : "r"(a), "r"(b));
        return acc;
    }
    acc = fmaf(tc::lo16<false>(a), tc::lo16<false>(b), acc);
    return fmaf(tc::hi16<false>(a), tc::hi16<false>(b), acc);
}
// acc[0], acc[1] += p * (lo(v), hi(v)); p16 = p as an fp16 pair (F16) — the precision P has in the tensor-core product
template <bool F16>
__device__ __forceinline__ void axpy2(float p, uint32_t p16, uint32_t v, float& a0, float& a1) {
    if (F16) {
        asm("{.reg .b16 pl, ph, vl, vh;\n\t"
            "mov.b32 {pl, ph}, %2;\n\t"
            "mov.b32 {vl, vh}, %3;\n\t"
            "fma.rn.f32.f16 %0, pl, vl, %0;\n\t"
            "fma.rn.f32.f16 %1, pl, vh, %1;}"
            : "+f"(a0), "+f"(a1) : "r"(p16), "r"(v));
        return;
    }
    a0 = fmaf(p, tc::lo16<false>(v), a0);
    a1 = fmaf(p, tc::hi16<false>(v), a1);
}

// dot product of row r of a swizzled [rows][64] tile with a 64-element vector held as 32 words of 16-bit pairs
template <bool F16>
__device__ __forceinline__ float dot64_row(const uint8_t* tile, int r, const uint32_t* vec) {
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const uint4 w = *sw_chunk(tile, r, c);
        const uint4 v = *reinterpret_cast<const uint4*>(vec + 4 * c);
        acc = mac2<F16>(w.x, v.x, acc);
        acc = mac2<F16>(w.y, v.y, acc);
        acc = mac2<F16>(w.z, v.z, acc);
        acc = mac2<F16>(w.w, v.w, acc);
    }
    return acc;
}

// MN-major B operand (V: [keys][64 d], d contiguous, 128-byte swizzle): groups of 8 keys are 1024 B apart.
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t smem_addr) {
    const uint32_t lo = ((smem_addr >> 4) & 0x3FFFu) | (64u << 16);
    const uint32_t hi = 64u | (1u << 14) | (2u << 29);
    return ((uint64_t)hi << 32) | lo;
}

// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
          "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2_fast(float x) {      // MUFU.EX2, flush-to-zero: arguments here are <= 0
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void wg_sync(int wg) { asm volatile("bar.sync %0, 256;" ::"r"(wg + 1) : "memory"); }

template <bool F16>
__global__ void __launch_bounds__(kThreadsAttn, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const uint16_t* __restrict__ qkv,
                    uint16_t* __restrict__ out, int n_pairs) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint32_t tmem_slot;
    // 1024-byte alignment by pointer arithmetic on the shared pointer (a round trip through an integer would
    // make every later access a generic LD/ST instead of LDS/STS)
    uint8_t* smem0 = smem_raw + ((1024u - ((uint32_t)__cvta_generic_to_shared(smem_raw) & 1023u)) & 1023u);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wg = warp >> 3;                       // group 0 / 1
    const int wt = tid & 255;                       // thread within the group
    const int row = wt & 127;                       // query row within a tile
    const int half = wt >> 7;                       // which 128 keys of the row this thread owns
    uint8_t* smem = smem0 + wg * kWgBytes;
    XArea* X = reinterpret_cast<XArea*>(smem + kOffX);

    if (wt == 0) {
        tc::mbar_init(&X->bar_qk, 1);
        tc::mbar_init(&X->bar_v, 1);
        tc::mbar_init(&X->bar_s, 1);
        tc::mbar_init(&X->bar_pv, 1);
        tc::mbar_fence_init();
        tc::fence_proxy_async();
        if (tid == 0) tc::tma_prefetch_desc(&tmap_qkv);
    }
    if (warp == 0) tc::tmem_alloc(&tmem_slot, 512);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_slot + wg * 256;                               // this group's 256 columns
    const uint32_t lane_base = (uint32_t)(32 * (warp & 3)) << 16;             // TMEM lanes this warp may touch
    const uint32_t t_s = tmem + lane_base + half * 128;                       // this thread's 128 S columns
    const uint32_t t_p = tmem + lane_base + half * 128;                       // P of keys [128 half, +128): 64 columns
    const uint32_t t_o = tmem + lane_base + 64 + half * 32;                   // O: columns [64,128), 32 per thread
    const uint32_t idesc_s = tc::make_idesc16<F16>(128, 256);
    const uint32_t idesc_o = tc::make_idesc16<F16>(128, 64, 0, 1);

    const int first = blockIdx.x * 2 + wg, stride = gridDim.x * 2;
    auto issue_qk = [&](int pair) {
        const int b = pair >> 4, h = pair & 15, r0 = b * kTok;
        tc::mbar_expect_tx(&X->bar_qk, 4 * 16384);
        tc::tma_load_2d(&tmap_qkv, &X->bar_qk, smem + kOffQ, h * kD, r0);
        tc::tma_load_2d(&tmap_qkv, &X->bar_qk, smem + kOffQ + 16384, h * kD, r0 + 128);
        tc::tma_load_2d(&tmap_qkv, &X->bar_qk, smem + kOffK, kW + h * kD, r0);
        tc::tma_load_2d(&tmap_qkv, &X->bar_qk, smem + kOffK + 16384, kW + h * kD, r0 + 128);
    };
    auto issue_v = [&](int pair) {
        const int b = pair >> 4, h = pair & 15, r0 = b * kTok;
        tc::mbar_expect_tx(&X->bar_v, 2 * 16384);
        tc::tma_load_2d(&tmap_qkv, &X->bar_v, smem + kOffV, 2 * kW + h * kD, r0);
        tc::tma_load_2d(&tmap_qkv, &X->bar_v, smem + kOffV + 16384, 2 * kW + h * kD, r0 + 128);
    };
    if (wt == 0 && first < n_pairs) {
        issue_qk(first);
        issue_v(first);
    }

    // this thread's word (two values) of the 257th token's q / k / v rows, fetched one pair ahead
    auto load256 = [&](int pair) -> uint32_t {
        const int b = pair >> 4, h = pair & 15;
        return __ldg(reinterpret_cast<const uint32_t*>(qkv + ((size_t)b * kTok + 256) * (3 * kW) + (wt >> 5) * kW + h * kD + (wt & 31) * 2));
    };
    uint32_t pre256 = 0;
    if (wt < 96 && first < n_pairs) pre256 = load256(first);
    const int w8 = warp & 7;

    uint32_t ph_load = 0, ph_s = 0, ph_pv = 0;
    for (int pair = first; pair < n_pairs; pair += stride) {
        const int b = pair >> 4, h = pair & 15;
        const size_t tok0 = (size_t)b * kTok;
        const int next = pair + stride;
        if (wt < 96) {
            const int which = wt >> 5, i2 = wt & 31;
            if (which == 0) X->q256h[i2] = pre256;
            else if (which == 1) X->k256h[i2] = pre256;
            else {
                X->v256[2 * i2] = tc::lo16<F16>(pre256);
                X->v256[2 * i2 + 1] = tc::hi16<F16>(pre256);
            }
        }
        wg_sync(wg);
        tc::mbar_wait(&X->bar_qk, ph_load);
        if (wt == 0) {
            // S = Q_0 K^T
            const uint64_t dk = tc::make_desc_k_sw128(tc::smem_u32(smem + kOffK));
            const uint64_t dq = tc::make_desc_k_sw128(tc::smem_u32(smem + kOffQ));
#pragma unroll
            for (int k = 0; k < 4; ++k) tc::umma_bf16(tmem, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
            tc::umma_commit(&X->bar_s);
        }
        // score of query row `row` of tile `half` against key 256; CUDA cores, overlaps the MMA
        X->s256[half][row] = dot64_row<F16>(smem + kOffQ + half * 16384, row, X->k256h);

#pragma unroll 1
        for (int t = 0; t < 2; ++t) {
            tc::mbar_wait(&X->bar_s, ph_s);
            ph_s ^= 1;
            tc::tc_fence_after();
            if (t == 1 && wt == 0 && next < n_pairs) issue_qk(next);     // Q and K are dead once S_1 is complete
            // pass 1: maximum over this thread's 128 keys (four independent chains), key 256 folded into half 0
            // (X->s256[t][row] of tile 0 was written by this very thread; tile 1 is several barriers later)
            float mx;
            {
                uint32_t va[32], vb[32];
                float m0 = half == 0 ? X->s256[t][row] : -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
                tc::tmem_ld_32x32(t_s, va);
#pragma unroll 1
                for (int c = 0; c < 4; c += 2) {
                    tc::tmem_ld_wait();
                    tc::tmem_ld_32x32(t_s + (c + 1) * 32, vb);
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        m0 = fmaxf(m0, __uint_as_float(va[j]));
                        m1 = fmaxf(m1, __uint_as_float(va[j + 1]));
                        m2 = fmaxf(m2, __uint_as_float(va[j + 2]));
                        m3 = fmaxf(m3, __uint_as_float(va[j + 3]));
                    }
                    tc::tmem_ld_wait();
                    if (c + 2 < 4) tc::tmem_ld_32x32(t_s + (c + 2) * 32, va);
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        m0 = fmaxf(m0, __uint_as_float(vb[j]));
                        m1 = fmaxf(m1, __uint_as_float(vb[j + 1]));
                        m2 = fmaxf(m2, __uint_as_float(vb[j + 2]));
                        m3 = fmaxf(m3, __uint_as_float(vb[j + 3]));
                    }
                }
                mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
            }
            X->xmax[half][row] = mx;
            wg_sync(wg);
            mx = fmaxf(X->xmax[0][row], X->xmax[1][row]);
            const float s256 = X->s256[t][row];
            // pass 2: p = exp2((s - m) / 8 * log2 e); P (16-bit pairs) overwrites S columns this thread has consumed
            const float mxs = mx * kScaleLog2;
            const float p_last = ex2_fast(fmaf(s256, kScaleLog2, -mxs));
            float sum0 = half == 0 ? p_last : 0.f, sum1 = 0.f;
            {
                uint32_t va[32], vb[32];
                auto emit = [&](const uint32_t (&v)[32], int c) {
                    uint32_t pk[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float x0 = fmaf(__uint_as_float(v[2 * j]), kScaleLog2, -mxs);
                        const float x1 = fmaf(__uint_as_float(v[2 * j + 1]), kScaleLog2, -mxs);
                        const float p0 = ex2_fast(x0), p1 = ex2_fast(x1);
                        sum0 += p0;
                        sum1 += p1;
                        pk[j] = tc::pack16<F16>(p0, p1);
                    }
                    tmem_st_32x16(t_p + c * 16, pk);
                };
                tc::tmem_ld_32x32(t_s, va);
#pragma unroll 1
                for (int c = 0; c < 4; c += 2) {
                    tc::tmem_ld_wait();
                    tc::tmem_ld_32x32(t_s + (c + 1) * 32, vb);
                    emit(va, c);
                    tc::tmem_ld_wait();
                    if (c + 2 < 4) tc::tmem_ld_32x32(t_s + (c + 2) * 32, va);
                    emit(vb, c + 1);
                }
            }
            X->xsum[half][row] = sum0 + sum1;
            tmem_st_wait();
            tc::tc_fence_before();
            wg_sync(wg);
            const float inv_l = 1.0f / (X->xsum[0][row] + X->xsum[1][row]);
            if (t == 0) tc::mbar_wait(&X->bar_v, ph_load);
            if (wt == 0) {
                tc::tc_fence_after();
                // O = P V : 16 UMMAs of K = 16 keys (8 TMEM columns of P, 16 rows of V each)
                const uint32_t vbase = tc::smem_u32(smem + kOffV);
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    umma_bf16_ts(tmem + 64, tmem + (i < 8 ? 8 * i : 128 + 8 * (i - 8)), make_desc_mn_sw128(vbase + i * 2048), idesc_o, i != 0);
                tc::umma_commit(&X->bar_pv);
            }
            // ---- the 257th query row on the CUDA cores while P V runs ----
            if (t == 0) {
                // scores against all 257 keys (thread = key), maximum, exponentials, sum
                const float sa = dot64_row<F16>(smem + kOffK, wt, X->q256h);
                float s_last = -INFINITY;
                if (wt == 0) {
                    s_last = 0.f;
#pragma unroll
                    for (int d = 0; d < 32; ++d) s_last = mac2<F16>(X->q256h[d], X->k256h[d], s_last);
                }
                float cm = fmaxf(sa, s_last);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, o));
                if (lane == 0) X->cls_red[w8] = cm;
                wg_sync(wg);
                cm = fmaxf(fmaxf(fmaxf(X->cls_red[0], X->cls_red[1]), fmaxf(X->cls_red[2], X->cls_red[3])),
                           fmaxf(fmaxf(X->cls_red[4], X->cls_red[5]), fmaxf(X->cls_red[6], X->cls_red[7])));
                const float pa = ex2_fast((sa - cm) * kScaleLog2);
                X->cls_p[wt] = pa;
                float sum = pa;
                if (wt == 0) {
                    const float pl = ex2_fast((s_last - cm) * kScaleLog2);
                    X->cls_p[256] = pl;
                    sum += pl;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
                if (lane == 0) X->cls_red[8 + w8] = sum;      // read after the barriers of tile 1
            } else {
                if (wt < 96 && next < n_pairs) pre256 = load256(next);
                // o[d] = sum_j p_j V[j][d]: warp = 8 values of d, lane = keys lane + 32 i (conflict-free 16-byte reads of the
                // swizzled tile), then a transposing butterfly leaves one total per group of four lanes
                float a8[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) a8[k] = 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int key = lane + 32 * i;
                    const uint4 q = *sw_chunk(smem + kOffV, key, w8);
                    const float pj = X->cls_p[key];
                    const uint32_t p16 = F16 ? tc::pack_f16(pj, pj) : 0u;
                    axpy2<F16>(pj, p16, q.x, a8[0], a8[1]);
                    axpy2<F16>(pj, p16, q.y, a8[2], a8[3]);
                    axpy2<F16>(pj, p16, q.z, a8[4], a8[5]);
                    axpy2<F16>(pj, p16, q.w, a8[6], a8[7]);
                }
                float b4[4], b2[2];
                const bool u16 = (lane & 16) != 0, u8 = (lane & 8) != 0, u4 = (lane & 4) != 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float send = u16 ? a8[k] : a8[k + 4];
                    const float keep = u16 ? a8[k + 4] : a8[k];
                    b4[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                }
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const float send = u8 ? b4[k] : b4[k + 2];
                    const float keep = u8 ? b4[k + 2] : b4[k];
                    b2[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                }
                float tot1 = (u4 ? b2[1] : b2[0]) + __shfl_xor_sync(0xffffffffu, u4 ? b2[0] : b2[1], 4);
                tot1 += __shfl_xor_sync(0xffffffffu, tot1, 2);
                tot1 += __shfl_xor_sync(0xffffffffu, tot1, 1);
                if ((lane & 3) == 0) X->cls_o[8 * w8 + (u16 ? 4 : 0) + (u8 ? 2 : 0) + (u4 ? 1 : 0)] = tot1;
            }
            tc::mbar_wait(&X->bar_pv, ph_pv);
            ph_pv ^= 1;
            tc::tc_fence_after();
            // epilogue: this thread's 32 columns of the O row + rank-1 contribution of key 256, normalised, 16-bit
            {
                uint32_t v0[32];
                tc::tmem_ld_32x32(t_o, v0);
                tc::tmem_ld_wait();
                tc::tc_fence_before();
                wg_sync(wg);                       // every row of O is in registers: the accumulator columns may be reused
                if (t == 0 && wt == 0) {
                    tc::tc_fence_after();
                    const uint64_t dk = tc::make_desc_k_sw128(tc::smem_u32(smem + kOffK));
                    const uint64_t dq = tc::make_desc_k_sw128(tc::smem_u32(smem + kOffQ + 16384));
#pragma unroll
                    for (int k = 0; k < 4; ++k) tc::umma_bf16(tmem, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
                    tc::umma_commit(&X->bar_s);
                }
                if (t == 1) {
                    // V is dead: O_1 is complete and every thread is past its reads of the tile (barrier above)
                    if (wt == 0 && next < n_pairs) issue_v(next);
                    if (wt < 64) {
                        const float tot = ((X->cls_red[8] + X->cls_red[9]) + (X->cls_red[10] + X->cls_red[11])) +
                                          ((X->cls_red[12] + X->cls_red[13]) + (X->cls_red[14] + X->cls_red[15]));
                        const float o = X->cls_o[wt] + X->cls_p[256] * X->v256[wt];
                        out[(tok0 + 256) * kW + h * kD + wt] = (uint16_t)(tc::pack16<F16>(o / tot, 0.f) & 0xffffu);
                    }
                }
                uint4* dst = reinterpret_cast<uint4*>(out + (tok0 + t * 128 + row) * kW + h * kD + half * 32);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float o[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int dcol = 8 * q + j;
                        o[j] = (__uint_as_float(v0[dcol]) + p_last * X->v256[half * 32 + dcol]) * inv_l;
                    }
                    dst[q] = make_uint4(tc::pack16<F16>(o[0], o[1]), tc::pack16<F16>(o[2], o[3]), tc::pack16<F16>(o[4], o[5]),
                                        tc::pack16<F16>(o[6], o[7]));
                }
            }
        }
        ph_load ^= 1;
        wg_sync(wg);        // X arrays (q256h / k256h / v256 / cls_* / s256) are rewritten by the next pair
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    if (warp == 0) tc::tmem_dealloc(tmem_slot, 512);
}

}  // namespace

int launch_attention_tc(const void* d_qkv, int batch, void* d_out, int f16, cudaStream_t stream) {
    FB_REQUIRE(d_qkv && d_out && batch >= 1, "fb_vit_attention: bad arguments");
    CUtensorMap tm;
    int rc = make_tmap_bf16_2d(&tm, d_qkv, (uint64_t)batch * kTok, 3 * kW, 3 * kW, 128, 64);
    if (rc) return rc;
    static PerDeviceFlag attr_set;
    if (!attr_set.get()) {
        FB_CUDA_OK(cudaFuncSetAttribute(attention_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemAttn));
        FB_CUDA_OK(cudaFuncSetAttribute(attention_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemAttn));
        attr_set.set();
    }
    const int n_pairs = batch * 16;
    int grid = (n_pairs + 1) / 2;
    if (grid > sm_count()) grid = sm_count();
    if (f16) attention_tc_kernel<true><<<grid, kThreadsAttn, kSmemAttn, stream>>>(tm, reinterpret_cast<const uint16_t*>(d_qkv),
                                                                              reinterpret_cast<uint16_t*>(d_out), n_pairs);
    else attention_tc_kernel<false><<<grid, kThreadsAttn, kSmemAttn, stream>>>(tm, reinterpret_cast<const uint16_t*>(d_qkv),
                                                                            reinterpret_cast<uint16_t*>(d_out), n_pairs);
    FB_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace fb
