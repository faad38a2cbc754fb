// CLIP preprocess and crop-Laplacian kernels.
//
//   resample_h_kernel / resample_v_kernel  replace `scorer.preprocess` (processing/scorer.py:508-510,
//       called per image at processing/batch_processor.py:95): Pillow's two-pass 8-bit antialiased
//       bicubic resampler (horizontal pass first, uint8 intermediate, 22-bit fixed-point taps),
//       CenterCrop, ToTensor (/255) and Normalize, all in float32 like torchvision.
//   roi_laplacian_kernel                   replaces analyzers/face.py:272-279 `_get_crop_sharpness`.
// The coefficient tables are built on the host (facet_b200/utils/resample.py) exactly as Pillow's
// precompute_coeffs + normalize_coeffs_8bpc do; the kernels only do the integer dot products.
#include "common.cuh"
#include "kernels.h"

namespace fb {

namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;   // Pillow PRECISION_BITS

__device__ __forceinline__ int clip8(int v) {
    v >>= kPrecisionBits;
    return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// Horizontal pass.  One CTA = kRowsPerBlock input rows of one image, staged in shared memory with
// 16-byte loads; thread = output column, accumulating all staged rows at once so every coefficient is
// fetched once per kRowsPerBlock rows.  Taps are re-based per column to a multiple of 4 pixels
// (12 bytes = 3 aligned words: B0 G0 R0 B1 | G1 R1 B2 G2 | R2 B3 G3 R3); the host pads the
// coefficient table with zeros accordingly (facet_b200/utils/resample.py, ResamplePlan.hcpad).
constexpr int kRowsPerBlock = 4;

__global__ void __launch_bounds__(256) resample_h_kernel(
    const uint8_t* __restrict__ img, long long img_stride, int H, int W, int out, const int* __restrict__ p0tab,
    const int* __restrict__ cpad, int ngroups, int row0, int rows, int px_lo, int span_px,
    uint8_t* __restrict__ tmp) {
    extern __shared__ __align__(16) uint8_t s_rows[];     // [kRowsPerBlock][span_bytes], span_bytes % 48 == 0
    const int n_img = blockIdx.y;
    const uint8_t* base = img + (size_t)n_img * img_stride;
    uint8_t* tbase = tmp + (size_t)n_img * rows * out * 3;
    const int span_bytes = span_px * 3;
    const int r_first = blockIdx.x * kRowsPerBlock;
    const size_t row_bytes = (size_t)W * 3;
    const int valid_bytes = min(span_bytes, (W - px_lo) * 3);      // bytes of the span that exist in the row
    const bool vec_ok = ((row_bytes & 15) == 0) && ((reinterpret_cast<uintptr_t>(base) & 15) == 0);
    for (int rr = 0; rr < kRowsPerBlock; ++rr) {
        const int r = min(r_first + rr, rows - 1);
        const uint8_t* src = base + (size_t)(row0 + r) * row_bytes + (size_t)px_lo * 3;
        uint8_t* dst = s_rows + (size_t)rr * span_bytes;
        if (vec_ok) {
            for (int i = threadIdx.x * 16; i < span_bytes; i += blockDim.x * 16) {
                uint4 v = make_uint4(0, 0, 0, 0);
                if (i + 16 <= valid_bytes) v = __ldg(reinterpret_cast<const uint4*>(src + i));
                else {
                    uint8_t t[16];
                    for (int b = 0; b < 16; ++b) t[b] = (i + b < valid_bytes) ? __ldg(src + i + b) : (uint8_t)0;
                    v = *reinterpret_cast<uint4*>(t);
                }
                *reinterpret_cast<uint4*>(dst + i) = v;
            }
        } else {
            for (int i = threadIdx.x; i < span_bytes; i += blockDim.x) dst[i] = (i < valid_bytes) ? __ldg(src + i) : (uint8_t)0;
        }
    }
    __syncthreads();
    for (int xo = threadIdx.x; xo < out; xo += blockDim.x) {
        const int word0 = ((p0tab[xo] - px_lo) * 3) >> 2;
        int acc[kRowsPerBlock][3];
#pragma unroll
        for (int rr = 0; rr < kRowsPerBlock; ++rr) acc[rr][0] = acc[rr][1] = acc[rr][2] = 1 << (kPrecisionBits - 1);
        for (int g = 0; g < ngroups; ++g) {
            const int c0 = __ldg(cpad + (size_t)(4 * g) * out + xo), c1 = __ldg(cpad + (size_t)(4 * g + 1) * out + xo);
            const int c2 = __ldg(cpad + (size_t)(4 * g + 2) * out + xo), c3 = __ldg(cpad + (size_t)(4 * g + 3) * out + xo);
#pragma unroll
            for (int rr = 0; rr < kRowsPerBlock; ++rr) {
                const uint32_t* wp = reinterpret_cast<const uint32_t*>(s_rows + (size_t)rr * span_bytes) + word0 + 3 * g;
                const uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2];
                acc[rr][0] += (int)(w0 & 255) * c0 + (int)(w0 >> 24) * c1 + (int)((w1 >> 16) & 255) * c2 + (int)((w2 >> 8) & 255) * c3;
                acc[rr][1] += (int)((w0 >> 8) & 255) * c0 + (int)(w1 & 255) * c1 + (int)(w1 >> 24) * c2 + (int)((w2 >> 16) & 255) * c3;
                acc[rr][2] += (int)((w0 >> 16) & 255) * c0 + (int)((w1 >> 8) & 255) * c1 + (int)(w2 & 255) * c2 + (int)(w2 >> 24) * c3;
            }
        }
#pragma unroll
        for (int rr = 0; rr < kRowsPerBlock; ++rr) {
            const int r = r_first + rr;
            if (r < rows) {
                uint8_t* o = tbase + ((size_t)r * out + xo) * 3;
                o[0] = (uint8_t)clip8(acc[rr][0]);
                o[1] = (uint8_t)clip8(acc[rr][1]);
                o[2] = (uint8_t)clip8(acc[rr][2]);
            }
        }
    }
}

// Vertical pass.  One CTA = one output row of one image; thread = 4 consecutive (column, channel)
// bytes of the row (32-bit loads of the uint8 intermediate), then ToTensor + Normalize in float32.
__global__ void __launch_bounds__(256) resample_v_kernel(
    const uint8_t* __restrict__ tmp, int rows, int row0, int out, const int* __restrict__ bounds,
    const int* __restrict__ coef, int ksize, int rgb_order, float m0, float m1, float m2, float s0, float s1,
    float s2, float* __restrict__ dst) {
    const int n_img = blockIdx.y;
    const int yo = blockIdx.x;
    const uint8_t* tbase = tmp + (size_t)n_img * rows * out * 3;
    const int first = bounds[2 * yo] - row0, cnt = bounds[2 * yo + 1];
    const int* k = coef + (size_t)yo * ksize;
    const int rowb = out * 3;
    const bool vec = (rowb & 3) == 0 && ((reinterpret_cast<uintptr_t>(tbase) & 3) == 0);
    for (int i = threadIdx.x * 4; i < rowb; i += blockDim.x * 4) {
        int acc[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[b] = 1 << (kPrecisionBits - 1);
        if (vec) {
#pragma unroll 4
            for (int y = 0; y < cnt; ++y) {
                const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(tbase + (size_t)(first + y) * rowb + i));
                const int kv = __ldg(k + y);
                acc[0] += (int)(w & 255) * kv;
                acc[1] += (int)((w >> 8) & 255) * kv;
                acc[2] += (int)((w >> 16) & 255) * kv;
                acc[3] += (int)(w >> 24) * kv;
            }
        } else {
            for (int y = 0; y < cnt; ++y) {
                const int kv = __ldg(k + y);
#pragma unroll
                for (int b = 0; b < 4; ++b)
                    if (i + b < rowb) acc[b] += (int)tbase[(size_t)(first + y) * rowb + i + b] * kv;
            }
        }
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int ib = i + b;
            if (ib >= rowb) break;
            const int u = clip8(acc[b]);
            const int xo = ib / 3, c = ib - 3 * xo;
            const int co = rgb_order ? c : 2 - c;                 // output planes are R,G,B
            const float mean = co == 0 ? m0 : (co == 1 ? m1 : m2);
            const float sd = co == 0 ? s0 : (co == 1 ? s1 : s2);
            const float v = __fdiv_rn((float)u, 255.0f);          // ToTensor
            dst[(((size_t)n_img * 3 + co) * out + yo) * out + xo] = __fdiv_rn(__fsub_rn(v, mean), sd);   // Normalize
        }
    }
}

__device__ __forceinline__ int gray_at(const uint8_t* img, int W, int y, int x, int rgb) {
    const uint8_t* p = img + ((size_t)y * W + x) * 3;
    const int c0 = p[0], c1 = p[1], c2 = p[2];
    const int b = rgb ? c2 : c0, r = rgb ? c0 : c2;
    return (3735 * b + 19235 * c1 + 9798 * r + 16384) >> 15;
}

__global__ void __launch_bounds__(256) roi_laplacian_kernel(const uint8_t* __restrict__ img, int H, int W, int rgb,
                                                            const int* __restrict__ boxes, long long* __restrict__ out) {
    __shared__ unsigned long long s_l[8], s_q[8];
    const int k = blockIdx.x;
    const int x1 = boxes[4 * k], y1 = boxes[4 * k + 1], x2 = boxes[4 * k + 2], y2 = boxes[4 * k + 3];
    const int w = x2 - x1, h = y2 - y1;
    long long sl = 0;
    unsigned long long sq = 0;
    const bool ok = w >= 1 && h >= 1 && x1 >= 0 && y1 >= 0 && x2 <= W && y2 <= H;
    if (ok) {
        const long long npx = (long long)w * h;
        for (long long i = threadIdx.x; i < npx; i += blockDim.x) {
            const int yy = (int)(i / w), xx = (int)(i % w);
            // reflect-101 of the crop; an extent of one pixel reflects onto itself
            const int yu = yy == 0 ? (h > 1) : yy - 1, yd = yy == h - 1 ? (h > 1 ? h - 2 : 0) : yy + 1;
            const int xl = xx == 0 ? (w > 1) : xx - 1, xr = xx == w - 1 ? (w > 1 ? w - 2 : 0) : xx + 1;
            const int c = gray_at(img, W, y1 + yy, x1 + xx, rgb);
            const int L = gray_at(img, W, y1 + yu, x1 + xx, rgb) + gray_at(img, W, y1 + yd, x1 + xx, rgb) +
                          gray_at(img, W, y1 + yy, x1 + xl, rgb) + gray_at(img, W, y1 + yy, x1 + xr, rgb) - 4 * c;
            sl += L;
            sq += (unsigned long long)(L * L);
        }
    }
    unsigned long long l_bits = warp_sum_u64((unsigned long long)sl);
    sq = warp_sum_u64(sq);
    if ((threadIdx.x & 31) == 0) {
        s_l[threadIdx.x >> 5] = l_bits;
        s_q[threadIdx.x >> 5] = sq;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long L = 0, Q = 0;
        for (int i = 0; i < 8; ++i) {
            L += s_l[i];
            Q += s_q[i];
        }
        out[3 * k] = ok ? (long long)w * h : 0;
        out[3 * k + 1] = (long long)L;
        out[3 * k + 2] = (long long)Q;
    }
}

}  // namespace

int launch_clip_preprocess(const uint8_t* d_images, int n, int H, int W, long long image_stride, int rgb_order,
                           int out_size, const int* d_hp0, const int* d_hcpad, int hgroups, int h_px_lo,
                           int h_span_px, const int* d_vbounds, const int* d_vcoef, int vk, int row0, int rows,
                           const float* mean3, const float* std3, uint8_t* d_tmp, float* d_out,
                           const int8_t* d_tc_coef, int tc_kw, int tc_limbs, const int* d_tc_kb0, cudaStream_t stream) {
    FB_REQUIRE(d_images && d_hp0 && d_hcpad && d_vbounds && d_vcoef && d_tmp && d_out && mean3 && std3,
               "fb_clip_preprocess: null pointer");
    FB_REQUIRE(n >= 1 && out_size >= 1 && out_size <= 1024, "fb_clip_preprocess: bad n/out_size");
    FB_REQUIRE(row0 >= 0 && rows >= 1 && row0 + rows <= H, "fb_clip_preprocess: row range outside the image");
    FB_REQUIRE(h_px_lo >= 0 && h_px_lo % 16 == 0 && h_span_px % 16 == 0 && h_px_lo < W && hgroups >= 1,
               "fb_clip_preprocess: horizontal span must start and end on multiples of 16 pixels");
    const size_t smem = (size_t)kRowsPerBlock * h_span_px * 3;
    FB_REQUIRE(smem <= 200 * 1024, "fb_clip_preprocess: row span %d px exceeds shared memory staging", h_span_px);
    static PerDeviceFlag attr_set;
    if (!attr_set.get()) {
        FB_CUDA_OK(cudaFuncSetAttribute(resample_h_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set.set();
    }
    // horizontal pass: exact int8 tensor-core product when the layout allows it, CUDA cores otherwise
    int tc = launch_resample_h_tc(d_images, n, H, W, image_stride, out_size, d_tc_coef, tc_kw, tc_limbs, d_tc_kb0, row0, rows,
                                  d_tmp, 3, stream);
    if (tc < 0 || tc > 1) return tc;
    if (tc == 1) {
        dim3 gh((rows + kRowsPerBlock - 1) / kRowsPerBlock, n);
        resample_h_kernel<<<gh, 256, smem, stream>>>(d_images, image_stride, H, W, out_size, d_hp0, d_hcpad, hgroups, row0, rows,
                                                     h_px_lo, h_span_px, d_tmp);
        FB_CUDA_OK(cudaGetLastError());
    }
    dim3 gv(out_size, n);
    int threads = (out_size * 3 + 3) / 4;
    threads = threads > 256 ? 256 : ((threads + 31) / 32) * 32;
    resample_v_kernel<<<gv, threads, 0, stream>>>(d_tmp, rows, row0, out_size, d_vbounds, d_vcoef, vk, rgb_order,
                                                  mean3[0], mean3[1], mean3[2], std3[0], std3[1], std3[2], d_out);
    FB_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_roi_laplacian(const uint8_t* d_image, int H, int W, int rgb_order, const int* d_boxes, int k,
                         long long* d_out, cudaStream_t stream) {
    FB_REQUIRE(d_image && d_boxes && d_out && k >= 1, "fb_roi_laplacian: bad arguments");
    roi_laplacian_kernel<<<k, 256, 0, stream>>>(d_image, H, W, rgb_order, d_boxes, d_out);
    FB_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace fb
