// Cosine similarity stage helpers (the north_star's kernel (3) in cosine mode).
//
// The reference only has cosine over stored CLIP embeddings at models/tagger.py:99-101 and
// api/routers/gallery.py:465-471; the grouping semantics are those of utils/duplicate.py
// (pair kept iff similarity >= threshold, Union-Find).  The all-pairs product runs as a bf16
// tcgen05 GEMM with a threshold epilogue (csrc/gemm.cu, FB_GEMM_THRESHOLD_PAIRS) that emits
// candidates with sim >= tau - band; bf16 input rounding moves a unit-vector dot product by at most
// 2^-8, so band = 0.01 cannot lose a pair.  Candidates are then re-scored here on the stored float32
// embeddings with float64 accumulation (every product of two floats is exact in a double, the 768-term sum is
// good to ~1e-13), which decides the final pair set: `double(dot) >= double(float(tau))`.  A float32 dot product
// (what NumPy / BLAS computes in the reference's formula sites) depends on the summation order by ~1e-7, so
// it cannot define a pair set bit-exactly; the float64 criterion can, on both sides (oracle: float64 matmul).
#include <cuda_bf16.h>

#include "common.cuh"
#include "kernels.h"

namespace fb {

namespace {

__global__ void __launch_bounds__(256) f32_to_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n) {
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 2; i < n; i += (long long)gridDim.x * blockDim.x * 2) {
        if (i + 1 < n) {
            *reinterpret_cast<__nv_bfloat162*>(out + i) = __floats2bfloat162_rn(in[i], in[i + 1]);
        } else {
            out[i] = __float2bfloat16(in[i]);
        }
    }
}

// one warp per candidate: float64 dot product of the float32 rows in a fixed order, keep iff >= tau
__global__ void __launch_bounds__(256) cosine_recheck_kernel(const float* __restrict__ emb, long long ld, int k,
                                                             const int* __restrict__ cand, const unsigned long long* __restrict__ ncand,
                                                             long long cand_cap, float tau, int* __restrict__ pairs,
                                                             float* __restrict__ sims, long long cap,
                                                             unsigned long long* __restrict__ count) {
    const long long total = (long long)min((unsigned long long)cand_cap, *ncand);
    const int lane = threadIdx.x & 31;
    for (long long c = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); c < total; c += (long long)gridDim.x * 8) {
        const int i = cand[2 * c], j = cand[2 * c + 1];
        const float* a = emb + (size_t)i * ld;
        const float* b = emb + (size_t)j * ld;
        double d = 0.0;
        for (int x = lane; x < k; x += 32) d = fma((double)a[x], (double)b[x], d);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
        if (lane == 0 && d >= (double)tau) {
            const unsigned long long pos = atomicAdd(count, 1ull);
            if ((long long)pos < cap) {
                pairs[2 * pos] = i;
                pairs[2 * pos + 1] = j;
                sims[pos] = (float)d;
            }
        }
    }
}

}  // namespace

int launch_f32_to_bf16(const float* d_in, void* d_out, long long n, cudaStream_t stream) {
    FB_REQUIRE(d_in && d_out && n >= 0, "fb_f32_to_bf16: bad arguments");
    if (n == 0) return 0;
    long long blocks = (n / 2 + 255) / 256;
    if (blocks > sm_count() * 16) blocks = sm_count() * 16;
    if (blocks < 1) blocks = 1;
    f32_to_bf16_kernel<<<(unsigned)blocks, 256, 0, stream>>>(d_in, reinterpret_cast<__nv_bfloat16*>(d_out), n);
    FB_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_cosine_recheck(const float* d_emb_f32, long long ld, int k, const int* d_cand, const unsigned long long* d_ncand,
                          long long cand_cap, float tau, int* d_pairs, float* d_sims, long long cap,
                          unsigned long long* d_count, cudaStream_t stream) {
    FB_REQUIRE(d_emb_f32 && d_cand && d_ncand && d_pairs && d_sims && d_count, "fb_cosine_pairs: null pointer");
    long long blocks = (cand_cap + 7) / 8;
    if (blocks > sm_count() * 8) blocks = sm_count() * 8;
    if (blocks < 1) blocks = 1;
    cosine_recheck_kernel<<<(unsigned)blocks, 256, 0, stream>>>(d_emb_f32, ld, k, d_cand, d_ncand, cand_cap, tau, d_pairs, d_sims,
                                                               cap, d_count);
    FB_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace fb
