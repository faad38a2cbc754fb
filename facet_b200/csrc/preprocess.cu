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

// One CTA = kRowsPerBlock input rows of one image; thread = output column.
constexpr int kRowsPerBlock = 8;

__global__ void __launch_bounds__(256) resample_h_kernel(
    const uint8_t* __restrict__ img, long long img_stride, int H, int W, int out, const int* __restrict__ bounds,
    const int* __restrict__ coef, int ksize, int row0, int rows, int byte_lo, int byte_hi,
    uint8_t* __restrict__ tmp) {
    extern __shared__ __align__(16) uint8_t s_row[];
    const int n_img = blockIdx.y;
    const uint8_t* base = img + (size_t)n_img * img_stride;
    uint8_t* tbase = tmp + (size_t)n_img * rows * out * 3;
    const int span = byte_hi - byte_lo;
    for (int rr = 0; rr < kRowsPerBlock; ++rr) {
        const int r = blockIdx.x * kRowsPerBlock + rr;
        if (r >= rows) break;
        const uint8_t* src = base + (size_t)(row0 + r) * W * 3 + byte_lo;
        __syncthreads();
        for (int i = threadIdx.x; i < span; i += blockDim.x) s_row[i] = __ldg(src + i);
        __syncthreads();
        for (int xo = threadIdx.x; xo < out; xo += blockDim.x) {
            const int first = bounds[2 * xo], cnt = bounds[2 * xo + 1];
            const int* k = coef + (size_t)xo * ksize;
            const uint8_t* p = s_row + first * 3 - byte_lo;
            int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
            for (int x = 0; x < cnt; ++x) {
                const int kv = __ldg(k + x);
                a0 += (int)p[3 * x] * kv;
                a1 += (int)p[3 * x + 1] * kv;
                a2 += (int)p[3 * x + 2] * kv;
            }
            uint8_t* o = tbase + ((size_t)r * out + xo) * 3;
            o[0] = (uint8_t)clip8(a0);
            o[1] = (uint8_t)clip8(a1);
            o[2] = (uint8_t)clip8(a2);
        }
    }
}

// One CTA = one output row of one image; thread = (column, channel) byte of the row.
__global__ void __launch_bounds__(1024) resample_v_kernel(
    const uint8_t* __restrict__ tmp, int rows, int row0, int out, const int* __restrict__ bounds,
    const int* __restrict__ coef, int ksize, int rgb_order, float m0, float m1, float m2, float s0, float s1,
    float s2, float* __restrict__ dst) {
    const int n_img = blockIdx.y;
    const int yo = blockIdx.x;
    const uint8_t* tbase = tmp + (size_t)n_img * rows * out * 3;
    const int first = bounds[2 * yo] - row0, cnt = bounds[2 * yo + 1];
    const int* k = coef + (size_t)yo * ksize;
    const int rowb = out * 3;
    for (int i = threadIdx.x; i < rowb; i += blockDim.x) {
        int acc = 1 << (kPrecisionBits - 1);
        for (int y = 0; y < cnt; ++y) acc += (int)tbase[(size_t)(first + y) * rowb + i] * __ldg(k + y);
        const int u = clip8(acc);
        const int xo = i / 3, c = i - 3 * xo;
        const int co = rgb_order ? c : 2 - c;                 // output planes are R,G,B
        const float mean = co == 0 ? m0 : (co == 1 ? m1 : m2);
        const float sd = co == 0 ? s0 : (co == 1 ? s1 : s2);
        const float v = __fdiv_rn((float)u, 255.0f);          // ToTensor
        dst[(((size_t)n_img * 3 + co) * out + yo) * out + xo] = __fdiv_rn(__fsub_rn(v, mean), sd);   // Normalize
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
    const bool ok = w >= 2 && h >= 2 && x1 >= 0 && y1 >= 0 && x2 <= W && y2 <= H;
    if (ok) {
        const long long npx = (long long)w * h;
        for (long long i = threadIdx.x; i < npx; i += blockDim.x) {
            const int yy = (int)(i / w), xx = (int)(i % w);
            const int yu = yy == 0 ? 1 : yy - 1, yd = yy == h - 1 ? h - 2 : yy + 1;
            const int xl = xx == 0 ? 1 : xx - 1, xr = xx == w - 1 ? w - 2 : xx + 1;
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
                           int out_size, const int* d_hbounds, const int* d_hcoef, int hk, int h_byte_lo,
                           int h_byte_hi, const int* d_vbounds, const int* d_vcoef, int vk, int row0, int rows,
                           const float* mean3, const float* std3, uint8_t* d_tmp, float* d_out,
                           cudaStream_t stream) {
    FB_REQUIRE(d_images && d_hbounds && d_hcoef && d_vbounds && d_vcoef && d_tmp && d_out && mean3 && std3,
               "fb_clip_preprocess: null pointer");
    FB_REQUIRE(n >= 1 && out_size >= 1 && out_size <= 1024, "fb_clip_preprocess: bad n/out_size");
    FB_REQUIRE(row0 >= 0 && rows >= 1 && row0 + rows <= H, "fb_clip_preprocess: row range outside the image");
    FB_REQUIRE(h_byte_lo >= 0 && h_byte_hi <= W * 3 && h_byte_lo < h_byte_hi, "fb_clip_preprocess: byte range outside the row");
    const int span = h_byte_hi - h_byte_lo;
    FB_REQUIRE(span <= 200 * 1024, "fb_clip_preprocess: row span %d B exceeds shared memory staging", span);
    FB_CUDA_OK(cudaFuncSetAttribute(resample_h_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    dim3 gh((rows + kRowsPerBlock - 1) / kRowsPerBlock, n);
    resample_h_kernel<<<gh, 256, span, stream>>>(d_images, image_stride, H, W, out_size, d_hbounds, d_hcoef, hk,
                                                row0, rows, h_byte_lo, h_byte_hi, d_tmp);
    FB_CUDA_OK(cudaGetLastError());
    dim3 gv(out_size, n);
    int threads = out_size * 3;
    threads = threads > 1024 ? 1024 : ((threads + 31) / 32) * 32;
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
