// Photo thumbnail: Pillow's Image.thumbnail((size, size), LANCZOS) on the GPU, bit-exact.
//
// Replaces the pixel work of `generate_photo_thumbnail` (utils/image_transforms.py:32-50 of the reference,
// called when a photo row is saved, processing/scorer.py:1681-1686):
//   1. box_reduce_kernel   ImagingReduce by (fx, fy) = int(scale / reducing_gap): ((sum + n/2) * multiplier) >> 24
//                          per channel, narrower boxes (with their own multiplier) on the right / bottom edge
//   2. resample_h_u8_kernel / resample_v_u8_kernel   the two-pass 8-bit Lanczos resampler on the reduced image
//                          (22-bit fixed-point taps built on the host, facet_b200/utils/thumbnail.py)
// The JPEG encoding of the few hundred KB result stays on the host (PIL, as in the reference).
#include "common.cuh"
#include "kernels.h"

namespace fb {

namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;   // Pillow PRECISION_BITS

struct ReduceArgs {
    const uint8_t* img;
    long long img_stride;
    int H, W, fx, fy, rh, rw;
    unsigned int mult[4];      // multiplier of (full, right edge, bottom edge, corner) boxes
    uint8_t* out;
};

// One thread per output pixel.  fx == 4 on 4-byte aligned rows reads each of the box's rows as three words
// (B0 G0 R0 B1 | G1 R1 B2 G2 | R2 B3 G3 R3); everything else takes the byte loop.
__global__ void __launch_bounds__(256) box_reduce_kernel(ReduceArgs a) {
    const int n_img = blockIdx.z;
    const int xo = blockIdx.x * blockDim.x + threadIdx.x;
    const int yo = blockIdx.y;
    if (xo >= a.rw) return;
    const uint8_t* base = a.img + (size_t)n_img * a.img_stride;
    const int x0 = xo * a.fx, y0 = yo * a.fy;
    const int xs = min(a.fx, a.W - x0), ys = min(a.fy, a.H - y0);
    const size_t row_bytes = (size_t)a.W * 3;
    unsigned int s0 = 0, s1 = 0, s2 = 0;
    const bool fast4 = (a.fx == 4) && (xs == 4) && ((row_bytes & 3) == 0) && ((reinterpret_cast<uintptr_t>(base) & 3) == 0);
    if (fast4) {
        for (int y = 0; y < ys; ++y) {
            const uint32_t* p = reinterpret_cast<const uint32_t*>(base + (size_t)(y0 + y) * row_bytes + (size_t)x0 * 3);
            const uint32_t w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2);
            s0 += (w0 & 255) + (w0 >> 24) + ((w1 >> 16) & 255) + ((w2 >> 8) & 255);
            s1 += ((w0 >> 8) & 255) + (w1 & 255) + (w1 >> 24) + ((w2 >> 16) & 255);
            s2 += ((w0 >> 16) & 255) + ((w1 >> 8) & 255) + (w2 & 255) + (w2 >> 24);
        }
    } else {
        for (int y = 0; y < ys; ++y) {
            const uint8_t* p = base + (size_t)(y0 + y) * row_bytes + (size_t)x0 * 3;
            for (int x = 0; x < xs; ++x) {
                s0 += __ldg(p + 3 * x);
                s1 += __ldg(p + 3 * x + 1);
                s2 += __ldg(p + 3 * x + 2);
            }
        }
    }
    const unsigned int n = (unsigned int)(xs * ys);
    const unsigned int m = a.mult[(xs < a.fx ? 1 : 0) + (ys < a.fy ? 2 : 0)];
    uint8_t* o = a.out + ((size_t)n_img * a.rh * a.rw + (size_t)yo * a.rw + xo) * 3;
    o[0] = (uint8_t)(((s0 + n / 2) * m) >> 24);
    o[1] = (uint8_t)(((s1 + n / 2) * m) >> 24);
    o[2] = (uint8_t)(((s2 + n / 2) * m) >> 24);
}

// fx == 4 on 16-byte aligned rows whose width is a multiple of 16: one thread = four output pixels = 48
// contiguous bytes of each of the box's rows (3 x LDG.128); channel sums with byte-selecting dot products.
__global__ void __launch_bounds__(256) box_reduce4_kernel(ReduceArgs a) {
    const int n_img = blockIdx.z;
    const int xq = blockIdx.x * blockDim.x + threadIdx.x;      // group of 4 output pixels
    const int yo = blockIdx.y;
    if (xq * 4 >= a.rw) return;
    const uint8_t* base = a.img + (size_t)n_img * a.img_stride;
    const int y0 = yo * a.fy;
    const int ys = min(a.fy, a.H - y0);
    const size_t row_bytes = (size_t)a.W * 3;
    unsigned int s[4][3] = {};
    for (int y = 0; y < ys; ++y) {
        const uint4* p = reinterpret_cast<const uint4*>(base + (size_t)(y0 + y) * row_bytes + (size_t)xq * 48);
        const uint4 q0 = ldg_nc_v4(p), q1 = ldg_nc_v4(p + 1), q2 = ldg_nc_v4(p + 2);
        const uint32_t w[12] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            // 12 bytes of output pixel o: B0 G0 R0 B1 | G1 R1 B2 G2 | R2 B3 G3 R3
            const uint32_t w0 = w[3 * o], w1 = w[3 * o + 1], w2 = w[3 * o + 2];
            s[o][0] = __dp4a(w0, 0x01000001u, __dp4a(w1, 0x00010000u, __dp4a(w2, 0x00000100u, s[o][0])));
            s[o][1] = __dp4a(w0, 0x00000100u, __dp4a(w1, 0x01000001u, __dp4a(w2, 0x00010000u, s[o][1])));
            s[o][2] = __dp4a(w0, 0x00010000u, __dp4a(w1, 0x00000100u, __dp4a(w2, 0x01000001u, s[o][2])));
        }
    }
    const unsigned int n = (unsigned int)(4 * ys);
    const unsigned int m = a.mult[ys < a.fy ? 2 : 0];
    uint8_t* o = a.out + ((size_t)n_img * a.rh * a.rw + (size_t)yo * a.rw + (size_t)xq * 4) * 3;
    uint32_t r[3] = {0, 0, 0};
#pragma unroll
    for (int i = 0; i < 12; ++i) r[i >> 2] |= (((s[i / 3][i % 3] + n / 2) * m) >> 24) << (8 * (i & 3));
    if ((a.rw & 3) == 0) {
        *reinterpret_cast<uint32_t*>(o) = r[0];
        *reinterpret_cast<uint32_t*>(o + 4) = r[1];
        *reinterpret_cast<uint32_t*>(o + 8) = r[2];
    } else {
        for (int i = 0; i < 12; ++i) o[i] = (uint8_t)(r[i >> 2] >> (8 * (i & 3)));
    }
}

__device__ __forceinline__ uint8_t clip8(int v) {
    v >>= kPrecisionBits;
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// Horizontal pass: in [n][rows][in_w][3] (row pitch in_pitch bytes, image stride in_stride) -> out [n][rows][out_w][3].
// One CTA per (row, image): the input row is staged in shared memory with word loads (byte loads when the
// row is not 4-byte aligned), then thread = output pixel.
__global__ void __launch_bounds__(256) resample_h_u8_kernel(const uint8_t* __restrict__ in, long long in_stride, long long in_pitch,
                                                            int in_w, int rows, int out_w, const int* __restrict__ bounds,
                                                            const int* __restrict__ coef, int ksize, uint8_t* __restrict__ out) {
    extern __shared__ __align__(16) uint8_t s_row[];
    const int n_img = blockIdx.y, y = blockIdx.x;
    const uint8_t* src = in + (size_t)n_img * in_stride + (size_t)y * in_pitch;
    const int nbytes = in_w * 3;
    if ((reinterpret_cast<uintptr_t>(src) & 3) == 0) {
        const int nwords = nbytes >> 2;
        for (int i = threadIdx.x; i < nwords; i += blockDim.x)
            reinterpret_cast<uint32_t*>(s_row)[i] = __ldg(reinterpret_cast<const uint32_t*>(src) + i);
        for (int i = (nwords << 2) + threadIdx.x; i < nbytes; i += blockDim.x) s_row[i] = __ldg(src + i);
    } else {
        for (int i = threadIdx.x; i < nbytes; i += blockDim.x) s_row[i] = __ldg(src + i);
    }
    __syncthreads();
    uint8_t* orow = out + ((size_t)n_img * rows + y) * out_w * 3;
    for (int xo = threadIdx.x; xo < out_w; xo += blockDim.x) {
        const int first = __ldg(bounds + 2 * xo), cnt = __ldg(bounds + 2 * xo + 1);
        const int* k = coef + (size_t)xo * ksize;
        const uint8_t* p = s_row + first * 3;
        int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
        for (int t = 0; t < cnt; ++t) {
            const int kv = __ldg(k + t);
            a0 += (int)p[3 * t] * kv;
            a1 += (int)p[3 * t + 1] * kv;
            a2 += (int)p[3 * t + 2] * kv;
        }
        orow[3 * xo] = clip8(a0);
        orow[3 * xo + 1] = clip8(a1);
        orow[3 * xo + 2] = clip8(a2);
    }
}

// Vertical pass: in [n][rows][w][3] -> out [n][out_h][w][3]; thread = 4 consecutive bytes of the row when the row
// length allows word loads (one byte otherwise); swap_rb reverses the channel order on the way out (BGR frames
// -> RGB thumbnails).
template <int VEC>
__global__ void __launch_bounds__(256) resample_v_u8_kernel(const uint8_t* __restrict__ in, int rows, int w, int out_h,
                                                            const int* __restrict__ bounds, const int* __restrict__ coef, int ksize,
                                                            int swap_rb, uint8_t* __restrict__ out) {
    const int n_img = blockIdx.z, yo = blockIdx.y;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) * VEC;       // first byte within the row
    const int rowb = w * 3;
    if (i >= rowb) return;
    const uint8_t* src = in + (size_t)n_img * rows * rowb;
    const int first = bounds[2 * yo], cnt = bounds[2 * yo + 1];
    const int* k = coef + (size_t)yo * ksize;
    int acc[VEC];
#pragma unroll
    for (int b = 0; b < VEC; ++b) acc[b] = 1 << (kPrecisionBits - 1);
    for (int t = 0; t < cnt; ++t) {
        const int kv = __ldg(k + t);
        if (VEC == 4) {
            const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(src + (size_t)(first + t) * rowb + i));
            acc[0] += (int)(v & 255) * kv;
            acc[1 % VEC] += (int)((v >> 8) & 255) * kv;
            acc[2 % VEC] += (int)((v >> 16) & 255) * kv;
            acc[3 % VEC] += (int)(v >> 24) * kv;
        } else {
            acc[0] += (int)__ldg(src + (size_t)(first + t) * rowb + i) * kv;
        }
    }
    uint8_t* orow = out + ((size_t)n_img * out_h + yo) * rowb;
#pragma unroll
    for (int b = 0; b < VEC; ++b) {
        int oi = i + b;
        if (swap_rb) {
            const int c = oi % 3;
            oi = oi - c + (2 - c);
        }
        orow[oi] = clip8(acc[b]);
    }
}

}  // namespace

int launch_box_reduce(const uint8_t* d_images, int n, int H, int W, long long image_stride, int fx, int fy, int red_h, int red_w,
                      const unsigned int* mult4, uint8_t* d_reduced, cudaStream_t stream) {
    FB_REQUIRE(d_images && d_reduced && mult4, "box reduction: null pointer");
    FB_REQUIRE(red_h == (H + fy - 1) / fy && red_w == (W + fx - 1) / fx, "box reduction: reduced size does not match the factors");
    FB_REQUIRE(n <= 65535 && red_h <= 65535, "box reduction: batch or height too large for one launch");
    ReduceArgs a;
    a.img = d_images; a.img_stride = image_stride; a.H = H; a.W = W; a.fx = fx; a.fy = fy; a.rh = red_h; a.rw = red_w;
    for (int i = 0; i < 4; ++i) a.mult[i] = mult4[i];
    a.out = d_reduced;
    const bool vec4 = fx == 4 && W % 16 == 0 && image_stride % 16 == 0 && (reinterpret_cast<uintptr_t>(d_images) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(d_reduced) & 3) == 0;
    if (vec4) {
        dim3 grid((red_w / 4 + 127) / 128, red_h, n);
        box_reduce4_kernel<<<grid, 128, 0, stream>>>(a);
    } else {
        dim3 grid((red_w + 255) / 256, red_h, n);
        box_reduce_kernel<<<grid, 256, 0, stream>>>(a);
    }
    FB_CUDA_OK(cudaGetLastError());
    count_launch(1);
    return 0;
}

int launch_thumbnail(const uint8_t* d_images, int n, int H, int W, long long image_stride, int fx, int fy, int red_h, int red_w,
                     const unsigned int* mult4, const int* d_hbounds, const int* d_hcoef, int hk, const int* d_vbounds,
                     const int* d_vcoef, int vk, int out_h, int out_w, int swap_rb, uint8_t* d_reduced, uint8_t* d_tmp,
                     uint8_t* d_out, int reduced_ready, cudaStream_t stream) {
    FB_REQUIRE((d_images || reduced_ready) && d_hbounds && d_hcoef && d_vbounds && d_vcoef && d_tmp && d_out, "fb_thumbnail: null pointer");
    FB_REQUIRE(n >= 1 && H >= 1 && W >= 1 && fx >= 1 && fy >= 1 && out_h >= 1 && out_w >= 1, "fb_thumbnail: bad sizes");
    FB_REQUIRE(red_h == (H + fy - 1) / fy && red_w == (W + fx - 1) / fx, "fb_thumbnail: reduced size does not match the factors");
    FB_REQUIRE(n <= 65535 && out_h <= 65535, "fb_thumbnail: batch or height too large for one launch");
    const uint8_t* src = d_images;
    long long src_stride = image_stride, src_pitch = (long long)W * 3;
    if (fx > 1 || fy > 1) {
        FB_REQUIRE(d_reduced, "fb_thumbnail: the reduction needs its buffer");
        if (!reduced_ready) {
            int rc = launch_box_reduce(d_images, n, H, W, image_stride, fx, fy, red_h, red_w, mult4, d_reduced, stream);
            if (rc) return rc;
        }
        src = d_reduced;
        src_stride = (long long)red_h * red_w * 3;
        src_pitch = (long long)red_w * 3;
    }
    {
        const size_t smem = ((size_t)red_w * 3 + 15) & ~(size_t)15;
        FB_REQUIRE(smem <= 200 * 1024, "fb_thumbnail: a row of %d pixels does not fit in shared memory", red_w);
        if (smem > 48 * 1024)
            FB_CUDA_OK(cudaFuncSetAttribute(resample_h_u8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid(red_h, n);
        resample_h_u8_kernel<<<grid, 256, smem, stream>>>(src, src_stride, src_pitch, red_w, red_h, out_w, d_hbounds, d_hcoef, hk, d_tmp);
        FB_CUDA_OK(cudaGetLastError());
    }
    {
        const int rowb = out_w * 3;
        if (rowb % 4 == 0 && (reinterpret_cast<uintptr_t>(d_tmp) & 3) == 0) {
            dim3 grid((rowb / 4 + 255) / 256, out_h, n);
            resample_v_u8_kernel<4><<<grid, 256, 0, stream>>>(d_tmp, red_h, out_w, out_h, d_vbounds, d_vcoef, vk, swap_rb, d_out);
        } else {
            dim3 grid((rowb + 255) / 256, out_h, n);
            resample_v_u8_kernel<1><<<grid, 256, 0, stream>>>(d_tmp, red_h, out_w, out_h, d_vbounds, d_vcoef, vk, swap_rb, d_out);
        }
        FB_CUDA_OK(cudaGetLastError());
    }
    count_launch(2);
    return 0;
}

}  // namespace fb
