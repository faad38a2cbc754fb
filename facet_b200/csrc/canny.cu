// Edge maps for the rule-based composition analyzer (analyzers/composition.py of the reference):
//   :215   cv2.GaussianBlur(gray, (5, 5), 0)
//   :218   cv2.Canny(blurred, 50, 150)                   (leading lines, before cv2.HoughLinesP)
//   :33-36 np.median(gray), cv2.Canny(gray, lower, upper) (subject search, before cv2.findContours)
// Bit-exact with OpenCV 4.x on uint8 input:
//   * GaussianBlur 5x5, sigma 0: the fixed-point kernel [1 4 6 4 1] / 16 in both directions,
//     BORDER_REFLECT_101, result (sum + 128) >> 8 (all intermediate values are exact integers)
//   * Canny, aperture 3, L1 gradient: Sobel with BORDER_REPLICATE, magnitude |dx| + |dy| (zero outside the
//     image), non-maximum suppression with the 15-bit fixed-point tangents tan 22.5 / tan 67.5, the
//     asymmetric comparisons (> on one side, >= on the other for the horizontal / vertical sectors),
//     hysteresis = every candidate (m > low, local maximum) that is 8-connected through candidates to one
//     with m > high
// The hysteresis is a connected-component problem: candidates are joined with a lock-free union-find
// (atomicMin links towards the smaller index), components that contain a strong pixel are flagged at their
// root, and the output pass keeps the candidates of flagged components.  Its result does not depend on the
// order of the unions, so the map equals OpenCV's stack-based propagation.
#include "common.cuh"
#include "kernels.h"

namespace fb {

namespace {

constexpr int kTW = 64, kTH = 32, kCThreads = 256;

__device__ __forceinline__ int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
    return i;
}
__device__ __forceinline__ int clampi(int i, int n) { return i < 0 ? 0 : (i >= n ? n - 1 : i); }

// gray = (3735 B + 19235 G + 9798 R + 2^14) >> 15 (cv2.cvtColor BGR2GRAY, image_cache.py:30), optional 256-bin histogram
template <bool RGB>
__global__ void __launch_bounds__(256) gray_plane_kernel(const uint8_t* __restrict__ img, long long npx, uint8_t* __restrict__ gray,
                                                         unsigned int* __restrict__ hist) {
    __shared__ unsigned int s_h[256];
    s_h[threadIdx.x] = 0u;
    __syncthreads();
    const bool words = ((reinterpret_cast<uintptr_t>(img) | reinterpret_cast<uintptr_t>(gray)) & 3) == 0;
    const long long groups = npx >> 2;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < groups; q += (long long)gridDim.x * blockDim.x) {
        uint32_t a, b, c;
        if (words) {
            const uint32_t* p = reinterpret_cast<const uint32_t*>(img) + 3 * q;
            a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
        } else {
            const uint8_t* p = img + 12 * q;
            a = p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24);
            b = p[4] | (p[5] << 8) | (p[6] << 16) | ((uint32_t)p[7] << 24);
            c = p[8] | (p[9] << 8) | (p[10] << 16) | ((uint32_t)p[11] << 24);
        }
        const uint32_t by[12] = {a & 255u, (a >> 8) & 255u, (a >> 16) & 255u, a >> 24, b & 255u, (b >> 8) & 255u,
                                 (b >> 16) & 255u, b >> 24, c & 255u, (c >> 8) & 255u, (c >> 16) & 255u, c >> 24};
        uint32_t out = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t c0 = by[3 * k], c1 = by[3 * k + 1], c2 = by[3 * k + 2];
            const uint32_t g = (3735u * (RGB ? c2 : c0) + 19235u * c1 + 9798u * (RGB ? c0 : c2) + 16384u) >> 15;
            out |= g << (8 * k);
            if (hist) atomicAdd(&s_h[g], 1u);
        }
        if (words) reinterpret_cast<uint32_t*>(gray)[q] = out;
        else {
            gray[4 * q] = (uint8_t)out, gray[4 * q + 1] = (uint8_t)(out >> 8);
            gray[4 * q + 2] = (uint8_t)(out >> 16), gray[4 * q + 3] = (uint8_t)(out >> 24);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(npx & 3)) {       // the last npx % 4 pixels
        const long long i = (groups << 2) + threadIdx.x;
        const uint32_t c0 = img[3 * i], c1 = img[3 * i + 1], c2 = img[3 * i + 2];
        const uint32_t g = (3735u * (RGB ? c2 : c0) + 19235u * c1 + 9798u * (RGB ? c0 : c2) + 16384u) >> 15;
        gray[i] = (uint8_t)g;
        if (hist) atomicAdd(&s_h[g], 1u);
    }
    __syncthreads();
    if (hist && s_h[threadIdx.x]) atomicAdd(hist + threadIdx.x, s_h[threadIdx.x]);
}

// 5x5 binomial blur, BORDER_REFLECT_101.  Tile of 64 x 32 outputs; rows of the tile + 2 above / below are
// filtered horizontally into 16-bit sums (<= 255 * 16), then vertically.
__global__ void __launch_bounds__(kCThreads) blur5_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int H, int W) {
    __shared__ uint8_t s_in[kTH + 4][kTW + 4 + 4];
    __shared__ uint16_t s_h[kTH + 4][kTW];
    const int x0 = blockIdx.x * kTW, y0 = blockIdx.y * kTH, tid = threadIdx.x;
    for (int i = tid; i < (kTH + 4) * (kTW + 4); i += kCThreads) {
        const int r = i / (kTW + 4), c = i % (kTW + 4);
        s_in[r][c] = src[(size_t)reflect101(y0 + r - 2, H) * W + reflect101(x0 + c - 2, W)];
    }
    __syncthreads();
    for (int i = tid; i < (kTH + 4) * kTW; i += kCThreads) {
        const int r = i / kTW, c = i % kTW;
        s_h[r][c] = (uint16_t)(s_in[r][c] + 4 * s_in[r][c + 1] + 6 * s_in[r][c + 2] + 4 * s_in[r][c + 3] + s_in[r][c + 4]);
    }
    __syncthreads();
    for (int i = tid; i < kTH * kTW; i += kCThreads) {
        const int r = i / kTW, c = i % kTW;
        const int y = y0 + r, x = x0 + c;
        if (y < H && x < W) {
            const int v = s_h[r][c] + 4 * s_h[r + 1][c] + 6 * s_h[r + 2][c] + 4 * s_h[r + 3][c] + s_h[r + 4][c];
            dst[(size_t)y * W + x] = (uint8_t)((v + 128) >> 8);
        }
    }
}

// Sobel + L1 magnitude + non-maximum suppression -> class (0 none, 1 candidate, 2 strong candidate) and the
// union-find label (own index for candidates, -1 otherwise).
__global__ void __launch_bounds__(kCThreads) sobel_nms_kernel(const uint8_t* __restrict__ src, int H, int W, int low, int high,
                                                              uint8_t* __restrict__ cls, int* __restrict__ label) {
    __shared__ uint8_t s_in[kTH + 4][kTW + 4 + 4];
    __shared__ uint16_t s_mag[kTH + 2][kTW + 2];
    const int x0 = blockIdx.x * kTW, y0 = blockIdx.y * kTH, tid = threadIdx.x;
    for (int i = tid; i < (kTH + 4) * (kTW + 4); i += kCThreads) {
        const int r = i / (kTW + 4), c = i % (kTW + 4);
        s_in[r][c] = src[(size_t)clampi(y0 + r - 2, H) * W + clampi(x0 + c - 2, W)];       // BORDER_REPLICATE
    }
    __syncthreads();
    auto grad = [&](int r, int c, int& dx, int& dy) {      // (r, c) in s_mag coordinates = s_in (r + 1, c + 1)
        const int a00 = s_in[r][c], a01 = s_in[r][c + 1], a02 = s_in[r][c + 2];
        const int a10 = s_in[r + 1][c], a12 = s_in[r + 1][c + 2];
        const int a20 = s_in[r + 2][c], a21 = s_in[r + 2][c + 1], a22 = s_in[r + 2][c + 2];
        dx = (a02 + 2 * a12 + a22) - (a00 + 2 * a10 + a20);
        dy = (a20 + 2 * a21 + a22) - (a00 + 2 * a01 + a02);
    };
    for (int i = tid; i < (kTH + 2) * (kTW + 2); i += kCThreads) {
        const int r = i / (kTW + 2), c = i % (kTW + 2);
        const int y = y0 + r - 1, x = x0 + c - 1;
        int m = 0;
        if (y >= 0 && y < H && x >= 0 && x < W) {
            int dx, dy;
            grad(r, c, dx, dy);
            m = abs(dx) + abs(dy);
        }
        s_mag[r][c] = (uint16_t)m;
    }
    __syncthreads();
    constexpr int kTg22 = 13573;      // (int)(0.4142135623730950488016887242097 * (1 << 15) + 0.5)
    for (int i = tid; i < kTH * kTW; i += kCThreads) {
        const int r = i / kTW, c = i % kTW;
        const int y = y0 + r, x = x0 + c;
        if (y >= H || x >= W) continue;
        const int m = s_mag[r + 1][c + 1];
        int k = 0;
        if (m > low) {
            int dx, dy;
            grad(r + 1, c + 1, dx, dy);
            const int ax = abs(dx), ay = abs(dy) << 15;
            const int tg22x = ax * kTg22;
            bool keep;
            if (ay < tg22x) keep = m > s_mag[r + 1][c] && m >= s_mag[r + 1][c + 2];
            else {
                const int tg67x = tg22x + (ax << 16);
                if (ay > tg67x) keep = m > s_mag[r][c + 1] && m >= s_mag[r + 2][c + 1];
                else {
                    const int s = (dx ^ dy) < 0 ? -1 : 1;
                    keep = m > s_mag[r][c + 1 - s] && m > s_mag[r + 2][c + 1 + s];
                }
            }
            if (keep) k = m > high ? 2 : 1;
        }
        const size_t idx = (size_t)y * W + x;
        cls[idx] = (uint8_t)k;
        label[idx] = k ? (int)idx : -1;
    }
}

__device__ __forceinline__ int uf_find(const int* label, int i) {
    int p = label[i];
    while (p != i) {
        i = p;
        p = label[i];
    }
    return i;
}
__device__ __forceinline__ void uf_union(int* label, int a, int b) {
    for (;;) {
        a = uf_find(label, a);
        b = uf_find(label, b);
        if (a == b) return;
        if (a < b) {
            const int t = a;
            a = b, b = t;
        }
        const int old = atomicMin(label + a, b);      // link the larger root below the smaller one
        if (old == a) return;
        a = old;                                      // someone else linked it first: merge with that target too
    }
}

// join every candidate with its candidate neighbours to the west, north-west, north and north-east
__global__ void __launch_bounds__(256) uf_merge_kernel(const uint8_t* __restrict__ cls, int* label, int H, int W) {
    const long long n = (long long)H * W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if (!cls[i]) continue;
        const int y = (int)(i / W), x = (int)(i % W);
        if (x > 0 && cls[i - 1]) uf_union(label, (int)i, (int)i - 1);
        if (y > 0) {
            const long long up = i - W;
            if (x > 0 && cls[up - 1]) uf_union(label, (int)i, (int)up - 1);
            if (cls[up]) uf_union(label, (int)i, (int)up);
            if (x + 1 < W && cls[up + 1]) uf_union(label, (int)i, (int)up + 1);
        }
    }
}

// strong candidates flag the root of their component (bit 2 of the root's class byte)
__global__ void __launch_bounds__(256) uf_flag_kernel(uint8_t* cls, const int* __restrict__ label, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if ((cls[i] & 3) != 2) continue;
        const int r = uf_find(label, (int)i);
        unsigned int* w = reinterpret_cast<unsigned int*>(cls) + (r >> 2);
        const unsigned int bit = 4u << (8 * (r & 3));
        if (!(*reinterpret_cast<volatile unsigned int*>(w) & bit)) atomicOr(w, bit);
    }
}

__global__ void __launch_bounds__(256) uf_out_kernel(const uint8_t* __restrict__ cls, const int* __restrict__ label, long long n,
                                                     uint8_t* __restrict__ edges, unsigned long long* __restrict__ count) {
    unsigned int c = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        uint8_t e = 0;
        if (cls[i] & 3) {
            const int r = uf_find(label, (int)i);
            if (cls[r] & 4) e = 255, ++c;
        }
        edges[i] = e;
    }
    if (count) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, (unsigned long long)c);
    }
}

inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace

int launch_gray_plane(const uint8_t* d_image, int H, int W, int rgb_order, uint8_t* d_gray, unsigned int* d_hist256,
                      cudaStream_t stream) {
    FB_REQUIRE(d_image && d_gray && H >= 1 && W >= 1, "fb_gray_plane: bad arguments");
    const long long npx = (long long)H * W;
    if (d_hist256) FB_CUDA_OK(cudaMemsetAsync(d_hist256, 0, 256 * sizeof(unsigned int), stream));
    long long want = ((npx >> 2) + 255) / 256;
    int blocks = (int)(want < 1 ? 1 : (want > (long long)sm_count() * 16 ? (long long)sm_count() * 16 : want));
    if (rgb_order) gray_plane_kernel<true><<<blocks, 256, 0, stream>>>(d_image, npx, d_gray, d_hist256);
    else gray_plane_kernel<false><<<blocks, 256, 0, stream>>>(d_image, npx, d_gray, d_hist256);
    FB_CUDA_OK(cudaGetLastError());
    return 0;
}

size_t canny_workspace_bytes(int H, int W) {
    const size_t n = (size_t)H * W;
    return align256(n) /* blurred plane */ + align256(n + 4) /* classes */ + align256(4 * n) /* labels */;
}

int launch_canny(const uint8_t* d_gray, int H, int W, int blur, int low, int high, void* d_ws, size_t ws_bytes,
                 uint8_t* d_edges, unsigned long long* d_count, cudaStream_t stream) {
    FB_REQUIRE(d_gray && d_edges && d_ws, "fb_canny: null pointer");
    FB_REQUIRE(H >= 1 && W >= 1 && (long long)H * W < (1ll << 31), "fb_canny: image of %d x %d is outside the supported range", W, H);
    FB_REQUIRE(ws_bytes >= canny_workspace_bytes(H, W), "fb_canny: workspace of %zu bytes, need %zu", ws_bytes, canny_workspace_bytes(H, W));
    FB_REQUIRE((reinterpret_cast<uintptr_t>(d_ws) & 255) == 0, "fb_canny: workspace must be 256-byte aligned");
    const size_t n = (size_t)H * W;
    uint8_t* ws = static_cast<uint8_t*>(d_ws);
    uint8_t* d_blur = ws;
    uint8_t* d_cls = d_blur + align256(n);
    int* d_label = reinterpret_cast<int*>(d_cls + align256(n + 4));
    const dim3 tiles((W + kTW - 1) / kTW, (H + kTH - 1) / kTH);
    FB_REQUIRE(tiles.y <= 65535, "fb_canny: image too tall");
    const uint8_t* src = d_gray;
    if (blur) {
        blur5_kernel<<<tiles, kCThreads, 0, stream>>>(d_gray, d_blur, H, W);
        src = d_blur;
    }
    if (n & 3) FB_CUDA_OK(cudaMemsetAsync(d_cls + n, 0, 4 - (n & 3), stream));      // padding bytes of the last class word
    if (d_count) FB_CUDA_OK(cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), stream));
    sobel_nms_kernel<<<tiles, kCThreads, 0, stream>>>(src, H, W, low, high, d_cls, d_label);
    long long want = ((long long)n + 255) / 256;
    const int blocks = (int)(want > (long long)sm_count() * 32 ? (long long)sm_count() * 32 : want);
    uf_merge_kernel<<<blocks, 256, 0, stream>>>(d_cls, d_label, H, W);
    uf_flag_kernel<<<blocks, 256, 0, stream>>>(d_cls, d_label, (long long)n);
    uf_out_kernel<<<blocks, 256, 0, stream>>>(d_cls, d_label, (long long)n, d_edges, d_count);
    FB_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace fb
