// Perceptual hash (64-bit pHash) of a batch of frames.
//
// Replaces `imagehash.phash(pil_img)` at processing/batch_processor.py:216 (scorer.py:972,
// multi_pass.py:449).  imagehash (requirements.txt:25, >= 4.3.0) is third-party and not vendored; its
// published algorithm is: PIL convert('L') -> resize((32,32), LANCZOS) -> 2-D DCT-II
// (scipy.fftpack.dct on both axes) -> top-left 8x8 -> bit = coefficient > median, row-major, MSB first.
//   luma_hresample_kernel   ITU-R 601 luma exactly as Pillow (L = (19595 R + 38470 G + 7471 B + 2^15) >> 16)
//                           fused with the horizontal Lanczos pass of Pillow's 8-bit resampler
//   phash_finish_kernel     vertical pass -> 32x32 uint8, float64 DCT-II low block, median, 64 bits
// The uint8 32x32 image is bit-exact with Pillow; the DCT is evaluated directly in float64.
#include "common.cuh"
#include "kernels.h"

namespace fb {

namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;
constexpr int kRows = 8;          // input rows per CTA in the horizontal pass
constexpr int kOut = 32;          // hash image is 32 x 32

__device__ __forceinline__ int clip8(int v) {
    v >>= kPrecisionBits;
    return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// thread = (output column xo = tid / 8, tap residue tid % 8); 8 rows per CTA share every coefficient load
__global__ void __launch_bounds__(256) luma_hresample_kernel(const uint8_t* __restrict__ img, long long img_stride, int H, int W,
                                                             int rgb_order, const int* __restrict__ bounds,
                                                             const int* __restrict__ coef, int ksize,
                                                             uint8_t* __restrict__ tmp /*[n][H][32]*/) {
    extern __shared__ __align__(16) uint8_t s_luma[];     // [kRows][W]
    const int n_img = blockIdx.y;
    const uint8_t* base = img + (size_t)n_img * img_stride;
    const int r_first = blockIdx.x * kRows;
    const int wr = rgb_order ? 19595 : 7471, wb = rgb_order ? 7471 : 19595;    // weight of byte 0 / byte 2
    const size_t row_bytes = (size_t)W * 3;
    const bool vec = ((row_bytes & 3) == 0) && ((reinterpret_cast<uintptr_t>(base) & 3) == 0) && (W % 4 == 0);
    for (int rr = 0; rr < kRows; ++rr) {
        const int r = min(r_first + rr, H - 1);
        const uint8_t* src = base + (size_t)r * row_bytes;
        uint8_t* dst = s_luma + (size_t)rr * W;
        if (vec) {
            for (int g = threadIdx.x; g < W / 4; g += blockDim.x) {
                const uint32_t* wp = reinterpret_cast<const uint32_t*>(src) + 3 * g;
                const uint32_t w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = __ldg(wp + 2);
                const int l0 = (wr * (int)(w0 & 255) + 38470 * (int)((w0 >> 8) & 255) + wb * (int)((w0 >> 16) & 255) + 0x8000) >> 16;
                const int l1 = (wr * (int)(w0 >> 24) + 38470 * (int)(w1 & 255) + wb * (int)((w1 >> 8) & 255) + 0x8000) >> 16;
                const int l2 = (wr * (int)((w1 >> 16) & 255) + 38470 * (int)(w1 >> 24) + wb * (int)(w2 & 255) + 0x8000) >> 16;
                const int l3 = (wr * (int)((w2 >> 8) & 255) + 38470 * (int)((w2 >> 16) & 255) + wb * (int)(w2 >> 24) + 0x8000) >> 16;
                *reinterpret_cast<uint32_t*>(dst + 4 * g) = (uint32_t)l0 | ((uint32_t)l1 << 8) | ((uint32_t)l2 << 16) | ((uint32_t)l3 << 24);
            }
        } else {
            for (int x = threadIdx.x; x < W; x += blockDim.x) {
                const uint8_t* p = src + (size_t)x * 3;
                dst[x] = (uint8_t)((wr * (int)p[0] + 38470 * (int)p[1] + wb * (int)p[2] + 0x8000) >> 16);
            }
        }
    }
    __syncthreads();
    const int xo = threadIdx.x >> 3, part = threadIdx.x & 7;
    const int first = bounds[2 * xo], cnt = bounds[2 * xo + 1];
    const int* k = coef + (size_t)xo * ksize;
    int acc[kRows];
#pragma unroll
    for (int rr = 0; rr < kRows; ++rr) acc[rr] = 0;
    for (int j = part; j < cnt; j += 8) {
        const int c = __ldg(k + j);
        const uint8_t* p = s_luma + first + j;
#pragma unroll
        for (int rr = 0; rr < kRows; ++rr) acc[rr] += (int)p[(size_t)rr * W] * c;
    }
#pragma unroll
    for (int rr = 0; rr < kRows; ++rr) {
        int v = acc[rr];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        const int r = r_first + rr;
        if (part == 0 && r < H) tmp[((size_t)n_img * H + r) * kOut + xo] = (uint8_t)clip8(v + (1 << (kPrecisionBits - 1)));
    }
}

// Pillow luma plane [n][H][W] uint8 (input of the tensor-core horizontal pass): 16 pixels (48 bytes) per thread
__global__ void __launch_bounds__(256) luma_plane_kernel(const uint8_t* __restrict__ img, long long total_px, int rgb_order,
                                                         uint8_t* __restrict__ luma) {
    const int wr = rgb_order ? 19595 : 7471, wb = rgb_order ? 7471 : 19595;
    const long long groups = total_px / 16;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (long long)gridDim.x * blockDim.x) {
        const uint4* src = reinterpret_cast<const uint4*>(img) + 3 * g;
        const uint4 a = ldg_nc_v4(src), b = ldg_nc_v4(src + 1), c = ldg_nc_v4(src + 2);
        const uint32_t w[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
        uint32_t o[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t w0 = w[3 * q], w1 = w[3 * q + 1], w2 = w[3 * q + 2];
            const int l0 = (wr * (int)(w0 & 255) + 38470 * (int)((w0 >> 8) & 255) + wb * (int)((w0 >> 16) & 255) + 0x8000) >> 16;
            const int l1 = (wr * (int)(w0 >> 24) + 38470 * (int)(w1 & 255) + wb * (int)((w1 >> 8) & 255) + 0x8000) >> 16;
            const int l2 = (wr * (int)((w1 >> 16) & 255) + 38470 * (int)(w1 >> 24) + wb * (int)(w2 & 255) + 0x8000) >> 16;
            const int l3 = (wr * (int)((w2 >> 8) & 255) + 38470 * (int)((w2 >> 16) & 255) + wb * (int)(w2 >> 24) + 0x8000) >> 16;
            o[q] = (uint32_t)l0 | ((uint32_t)l1 << 8) | ((uint32_t)l2 << 16) | ((uint32_t)l3 << 24);
        }
        reinterpret_cast<uint4*>(luma)[g] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// one CTA (1024 threads = the 32 x 32 pixels) per image
__global__ void __launch_bounds__(1024) phash_finish_kernel(const uint8_t* __restrict__ tmp, int H, const int* __restrict__ bounds,
                                                            const int* __restrict__ coef, int ksize,
                                                            unsigned long long* __restrict__ hashes, uint8_t* __restrict__ small /*[n][32][32] or null*/,
                                                            double* __restrict__ dct_out /*[n][64] or null*/) {
    __shared__ double px[kOut][kOut];
    __shared__ double rowdct[kOut][8];     // DCT along x, first 8 frequencies
    __shared__ double low[64];
    const int n_img = blockIdx.x;
    const int yo = threadIdx.x >> 5, xo = threadIdx.x & 31;
    {
        const int first = bounds[2 * yo], cnt = bounds[2 * yo + 1];
        const int* k = coef + (size_t)yo * ksize;
        const uint8_t* col = tmp + ((size_t)n_img * H + first) * kOut + xo;
        int acc = 1 << (kPrecisionBits - 1);
        for (int y = 0; y < cnt; ++y) acc += (int)col[(size_t)y * kOut] * __ldg(k + y);
        const int u = clip8(acc);
        px[yo][xo] = (double)u;
        if (small) small[((size_t)n_img * kOut + yo) * kOut + xo] = (uint8_t)u;
    }
    __syncthreads();
    // scipy.fftpack.dct type 2, unnormalised: y[k] = 2 sum_n x[n] cos(pi k (2n+1) / (2N)), on axis 0 then axis 1
    if (threadIdx.x < kOut * 8) {
        const int y = threadIdx.x >> 3, v = threadIdx.x & 7;
        double s = 0.0;
        for (int x = 0; x < kOut; ++x) s += px[y][x] * cospi((double)(v * (2 * x + 1)) / 64.0);
        rowdct[y][v] = 2.0 * s;
    }
    __syncthreads();
    if (threadIdx.x < 64) {
        const int u = threadIdx.x >> 3, v = threadIdx.x & 7;
        double s = 0.0;
        for (int y = 0; y < kOut; ++y) s += rowdct[y][v] * cospi((double)(u * (2 * y + 1)) / 64.0);
        low[threadIdx.x] = 2.0 * s;
        if (dct_out) dct_out[(size_t)n_img * 64 + threadIdx.x] = 2.0 * s;
    }
    __syncthreads();
    // median of 64 = mean of the 32nd and 33rd smallest (numpy.median): order statistics by rank counting,
    // one warp, two values per lane
    if (threadIdx.x < 32) {
        double sorted_lo = 0.0, sorted_hi = 0.0;
        for (int half = 0; half < 2; ++half) {
            const double me = low[threadIdx.x + 32 * half];
            int less = 0, equal = 0;
            for (int i = 0; i < 64; ++i) {
                less += low[i] < me;
                equal += low[i] == me;
            }
            const bool is31 = (less <= 31 && 31 < less + equal), is32 = (less <= 32 && 32 < less + equal);
            const unsigned m31 = __ballot_sync(0xffffffffu, is31), m32 = __ballot_sync(0xffffffffu, is32);
            if (m31) sorted_lo = __shfl_sync(0xffffffffu, me, __ffs(m31) - 1);
            if (m32) sorted_hi = __shfl_sync(0xffffffffu, me, __ffs(m32) - 1);
        }
        const double med = (sorted_lo + sorted_hi) / 2.0;
        unsigned long long h = 0ull;
        for (int half = 0; half < 2; ++half) {
            const int i = threadIdx.x + 32 * half;
            const unsigned m = __ballot_sync(0xffffffffu, low[i] > med);
            if (threadIdx.x == 0) {
                for (int b = 0; b < 32; ++b)
                    if (m & (1u << b)) h |= 1ull << (63 - (b + 32 * half));
            }
        }
        if (threadIdx.x == 0) hashes[n_img] = h;
    }
}

}  // namespace

int launch_phash(const uint8_t* d_images, int n, int H, int W, long long image_stride, int rgb_order, const int* d_hbounds,
                 const int* d_hcoef, int hk, const int* d_vbounds, const int* d_vcoef, int vk, uint8_t* d_tmp,
                 unsigned long long* d_hashes, uint8_t* d_small, double* d_dct, uint8_t* d_luma, int luma_ready,
                 const int8_t* d_tc_coef, int tc_kw, int tc_limbs, const int* d_tc_kb0, cudaStream_t stream) {
    FB_REQUIRE(d_images && d_hbounds && d_hcoef && d_vbounds && d_vcoef && d_tmp && d_hashes, "fb_phash: null pointer");
    FB_REQUIRE(n >= 1 && H >= 1 && W >= 1, "fb_phash: bad shape");
    // tensor-core route: Pillow luma plane, then the horizontal Lanczos pass as an exact u8 x s8 product
    int tc = 1;
    const bool plane_ok = d_luma && d_tc_coef && d_tc_kb0 && (W % 16 == 0) && image_stride == (long long)H * W * 3 &&
                          (reinterpret_cast<uintptr_t>(d_images) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_luma) & 15) == 0;
    FB_REQUIRE(!luma_ready || plane_ok, "fb_phash: luma_ready needs the tensor-core route (width %% 16 == 0, tables, aligned plane)");
    if (plane_ok) {
        if (!luma_ready) {     // otherwise fb_tech_stats_luma already produced the plane in its pass over the frame
            const long long total_px = (long long)n * H * W;
            long long blocks = (total_px / 16 + 255) / 256;
            if (blocks > sm_count() * 16) blocks = sm_count() * 16;
            luma_plane_kernel<<<(unsigned)blocks, 256, 0, stream>>>(d_images, total_px, rgb_order, d_luma);
            FB_CUDA_OK(cudaGetLastError());
        }
        tc = launch_resample_h_tc(d_luma, n, H, W, (long long)H * W, kOut, d_tc_coef, tc_kw, tc_limbs, d_tc_kb0, 0, H, d_tmp, 1, stream);
        if (tc < 0 || tc > 1) return tc;
    }
    if (tc == 1) {
        const size_t smem = (size_t)kRows * W;
        FB_REQUIRE(smem <= 200 * 1024, "fb_phash: image width %d exceeds shared memory staging", W);
        static PerDeviceFlag attr_set;
        if (!attr_set.get()) {
            FB_CUDA_OK(cudaFuncSetAttribute(luma_hresample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            attr_set.set();
        }
        dim3 gh((H + kRows - 1) / kRows, n);
        luma_hresample_kernel<<<gh, 256, smem, stream>>>(d_images, image_stride, H, W, rgb_order, d_hbounds, d_hcoef, hk, d_tmp);
        FB_CUDA_OK(cudaGetLastError());
    }
    phash_finish_kernel<<<n, 1024, 0, stream>>>(d_tmp, H, d_vbounds, d_vcoef, vk, d_hashes, d_small, d_dct);
    FB_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace fb
