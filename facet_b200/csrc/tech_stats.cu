// Technical-metrics pass: one read of a [n,H,W,3] uint8 batch -> sufficient statistics.
//
// Replaces, per image (reference paths under /root/reference):
//   analyzers/image_cache.py:30-32   cvtColor(BGR2GRAY), cvtColor(BGR2HSV), Laplacian(gray).var()
//   analyzers/technical.py:94        calcHist(hsv,[0,1],[180,256])          -> hs_hist[180][256]
//   analyzers/technical.py:153       calcHist(gray,[0],[256])               -> hist256[256]
//   analyzers/technical.py:302       sum |filter2D(gray, Immerkaer 3x3)|    -> sums[2]
// Outputs are exact integers; every metric dict of technical.py is a closed form of them
// (facet_b200/analyzers/_closed_form.py).
//
// Fast kernel (W % 16 == 0, 16-byte aligned rows):
//   * persistent grid = #SMs, 512 threads, 1 CTA/SM; the CTA owns a contiguous range of
//     "units" (512-px x R-row tiles) so that it mostly stays inside one image
//   * a warp walks down its tile: lane = 16 consecutive pixels (3 x LDG.128 per row,
//     prefetched two rows ahead); gray of the two previous rows and their horizontal second
//     differences stay in registers as exact fp16 pairs, so the 4-neighbour Laplacian and the
//     Immerkaer response (outer product of [1,-2,1]) cost one pass of packed half2 adds
//   * HS histogram (46080 x u32 = 180 KB) and a lane-private luminance histogram
//     (256 x 32 lanes, conflict-free) live in shared memory; they are merged into the
//     per-image global histograms only when the CTA crosses an image boundary
// Generic kernel: any W/H >= 2, any alignment; one thread per pixel, global atomics.
#include <cuda_fp16.h>

#include "common.cuh"
#include "kernels.h"

namespace fb {

namespace {

constexpr int kHsBins = 180 * 256;
constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kLanePx = 16;
constexpr int kTileW = 32 * kLanePx;   // 512 px per warp row
constexpr int kH256Copies = 32;        // one luminance-histogram column per lane
constexpr size_t kSmemWords = kHsBins + 256 * kH256Copies + 512 + 16;   // + work counter
constexpr int kOffH256 = kHsBins;                       // word offsets inside the dynamic smem block
constexpr int kOffSdiv = kOffH256 + 256 * kH256Copies;
constexpr int kOffHdiv = kOffSdiv + 256;

extern __shared__ __align__(16) unsigned int fb_smem[];

struct TechArgs {
    const uint8_t* img;
    long long img_stride;
    int n, H, W;
    int rows_per_unit, tiles_x, units_y;
    unsigned int* hist256;
    unsigned int* hs;
    unsigned long long* sums;
    int gray_round;   // 1 << 14, passed through the constant bank so IMAD can take it as an addend
    uint8_t* luma;    // optional [n][H][W] Pillow luma plane (input of the pHash resampler), or nullptr
    int luma_round;   // 1 << 15
};

__device__ __forceinline__ int sdiv_entry(int i) {   // round-half-even(255*4096 / i)
    return i == 0 ? 0 : __double2int_rn(1044480.0 / (double)i);
}
__device__ __forceinline__ int hdiv_entry(int i) {   // round-half-even(180*4096 / (6 i))
    return i == 0 ? 0 : __double2int_rn(737280.0 / (6.0 * (double)i));
}

__device__ __forceinline__ int gray_of(int b, int g, int r) {
    return (3735 * b + 19235 * g + 9798 * r + 16384) >> 15;
}

// OpenCV RGB2HSV_b, hue range 180: returns h*256+s.
__device__ __forceinline__ int hs_bin_of(int b, int g, int r, const unsigned int* sdiv,
                                         const unsigned int* hdiv) {
    int v = max(max(b, g), r);
    int mn = min(min(b, g), r);
    int d = v - mn;
    int s = (d * (int)sdiv[v] + 2048) >> 12;
    int hr = (v == r) ? (g - b) : ((v == g) ? (b - r + 2 * d) : (r - g + 4 * d));
    int h = (hr * (int)hdiv[d] + 2048) >> 12;
    h += (h < 0) ? 180 : 0;
    return h * 256 + s;
}

__device__ __forceinline__ __half2 u16x2_to_half2(uint32_t packed) {
    // 0x6400 | x is the fp16 number 1024 + x for x < 1024
    uint32_t bits = packed | 0x64006400u;
    return __hsub2(*reinterpret_cast<__half2*>(&bits), __half2half2(__ushort_as_half(0x6400)));
}
__device__ __forceinline__ uint32_t h2_bits(__half2 v) { return *reinterpret_cast<uint32_t*>(&v); }
__device__ __forceinline__ __half2 bits_h2(uint32_t v) { return *reinterpret_cast<__half2*>(&v); }

__device__ __forceinline__ float fma_f32_f16(uint16_t a, float c) {   // c + a*a, a is fp16 bits
    float r;
    asm("fma.rn.f32.f16 %0, %1, %1, %2;" : "=f"(r) : "h"(a), "f"(c));
    return r;
}
__device__ __forceinline__ float add_f32_f16(uint16_t a, float c) {   // c + a
    float r;
    asm("add.f32.f16 %0, %1, %2;" : "=f"(r) : "h"(a), "f"(c));
    return r;
}

struct RowRegs {
    uint4 q0, q1, q2;   // 48 bytes = 16 pixels
    uint32_t h0, h1, h2;   // the 3 bytes of the one extra pixel lane 0 / lane 31 may need (kept raw:
                           // combining them here would stall on the load two rows early)
};

// histogram increment in shared memory (ptxas turns "+1" into the warp-aggregating
// ATOMS.POPC.INC, which must not be predicated: FULL tiles call it unconditionally)
__device__ __forceinline__ void smem_inc(uint32_t smem_base, int word) {
    asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(smem_base + 4u * (uint32_t)word), "r"(1u) : "memory");
}

template <bool RGB>
__device__ __forceinline__ void load_row(RowRegs& r, const uint8_t* row, int xl, bool need_halo, int hx) {
    const uint8_t* p = row + (size_t)xl * 3;
    r.q0 = ldg_nc_v4(p);
    r.q1 = ldg_nc_v4(p + 16);
    r.q2 = ldg_nc_v4(p + 32);
    r.h0 = r.h1 = r.h2 = 0;
    if (need_halo) {
        const uint8_t* h = row + (size_t)hx * 3;
        r.h0 = __ldg(h);
        r.h1 = __ldg(h + 1);
        r.h2 = __ldg(h + 2);
    }
}

template <bool RGB, bool FULL, bool LUMA>
__device__ __forceinline__ void process_unit(const TechArgs& a, const uint8_t* img, int tx, int uy,
                                             unsigned long long& out_l2, unsigned long long& out_n,
                                             long long& out_l) {
    const int lane = (int)lane_id();
    const int W = a.W, H = a.H;
    const int r0 = uy * a.rows_per_unit;
    const int r1 = min(H, r0 + a.rows_per_unit);
    const int rb = r1 - r0;
    const int x0 = tx * kTileW + lane * kLanePx;
    const bool active = x0 < W;
    const int xl = active ? x0 : (W - kLanePx);
    // left / right neighbour of the lane's 16-pixel span
    const bool left_own = (xl == 0);                 // reflect-101: x=-1 -> x=1
    const bool right_own = (xl + kLanePx == W);      // x=W -> x=W-2
    const bool halo_left = (lane == 0) && !left_own;
    const bool halo_right = (lane == 31) && !right_own;
    const bool need_halo = halo_left || halo_right;
    const int hx = halo_left ? (xl - 1) : (xl + kLanePx);
    const size_t row_bytes = (size_t)W * 3;

    auto row_ptr = [&](int k) -> const uint8_t* {     // k-th row of the walk: r0-1 .. r1
        int y = r0 - 1 + k;
        y = (y < 0) ? 1 : y;
        y = (y >= H) ? (H - 2) : y;
        return img + (size_t)y * row_bytes;
    };

    const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(fb_smem);
    const int kRound = a.gray_round;
    const int nrows = rb + 2;
    RowRegs cur, nx1, nx2;
    load_row<RGB>(cur, row_ptr(0), xl, need_halo, hx);
    load_row<RGB>(nx1, row_ptr(1), xl, need_halo, hx);

    __half2 g_pp[8], g_p[8], d_pp[8], d_p[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        g_pp[j] = g_p[j] = d_pp[j] = d_p[j] = __float2half2_rn(0.f);
    }
    float acc_lh = 0.f, acc_lh2 = 0.f;      // horizontal telescoped part of sum(L)
    float acc_lv = 0.f;                     // vertical telescoped part
    unsigned long long acc_l2 = 0ull;
    unsigned int acc_n = 0u;
    const __half2 kMinus2 = __float2half2_rn(-2.f);

    for (int k = 0; k < nrows; ++k) {
        if (k + 2 < nrows) load_row<RGB>(nx2, row_ptr(k + 2), xl, need_halo, hx);
        const bool owned = (k >= 1) && (k <= rb);

        uint32_t w[12] = {cur.q0.x, cur.q0.y, cur.q0.z, cur.q0.w, cur.q1.x, cur.q1.y,
                          cur.q1.z, cur.q1.w, cur.q2.x, cur.q2.y, cur.q2.z, cur.q2.w};
        int gr[16];
        uint32_t lum[4] = {0u, 0u, 0u, 0u};
        if (owned) {
#pragma unroll
            for (int p = 0; p < 16; ++p) {
                const int o = 3 * p;
                const int c0 = (int)__byte_perm(w[o >> 2], 0u, 0x4440 + (o & 3));
                const int c1 = (int)__byte_perm(w[(o + 1) >> 2], 0u, 0x4440 + ((o + 1) & 3));
                const int c2 = (int)__byte_perm(w[(o + 2) >> 2], 0u, 0x4440 + ((o + 2) & 3));
                const int b = RGB ? c2 : c0, g = c1, r = RGB ? c0 : c2;
                gr[p] = (3735 * b + kRound + 19235 * g + 9798 * r) >> 15;
                if (LUMA) {   // Pillow convert('L'): (19595 R + 38470 G + 7471 B + 2^15) >> 16
                    const int l = (7471 * b + a.luma_round + 38470 * g + 19595 * r) >> 16;
                    lum[p >> 2] |= (uint32_t)l << (8 * (p & 3));
                }
                // OpenCV RGB2HSV_b (hue range 180), branch-free
                const int v = max(max(b, g), r);
                const int d = v - min(min(b, g), r);
                const int sd = (int)fb_smem[kOffSdiv + v];
                const int hd = (int)fb_smem[kOffHdiv + d];
                const bool vr = (v == r), vg = (v == g);
                const int x = vr ? g : (vg ? b : r);
                const int y = vr ? b : (vg ? r : g);
                const int off = vr ? 0 : (vg ? 2 * d : 4 * d);
                const int sat = (d * sd + 2048) >> 12;
                int hue = ((x - y + off) * hd + 2048) >> 12;
                hue += (hue >> 31) & 180;
                if (FULL || active) {
                    smem_inc(smem_base, hue * 256 + sat);
                    smem_inc(smem_base, kOffH256 + gr[p] * kH256Copies + lane);
                }
            }
        } else {
#pragma unroll
            for (int p = 0; p < 16; ++p) {
                const int o = 3 * p;
                const int c0 = (int)__byte_perm(w[o >> 2], 0u, 0x4440 + (o & 3));
                const int c1 = (int)__byte_perm(w[(o + 1) >> 2], 0u, 0x4440 + ((o + 1) & 3));
                const int c2 = (int)__byte_perm(w[(o + 2) >> 2], 0u, 0x4440 + ((o + 2) & 3));
                gr[p] = (3735 * (RGB ? c2 : c0) + kRound + 19235 * c1 + 9798 * (RGB ? c0 : c2)) >> 15;
            }
        }
        if (LUMA && owned && (FULL || active)) {
            const int y = r0 - 1 + k;      // owned rows are never reflected
            *reinterpret_cast<uint4*>(a.luma + ((size_t)(img - a.img) / 3) + (size_t)y * W + xl) = make_uint4(lum[0], lum[1], lum[2], lum[3]);
        }
        int gh;
        {
            const int c0 = (int)cur.h0, c1 = (int)cur.h1, c2 = (int)cur.h2;
            gh = gray_of(RGB ? c2 : c0, c1, RGB ? c0 : c2);
        }

        __half2 g_c[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
            g_c[j] = u16x2_to_half2(__byte_perm((uint32_t)gr[2 * j], (uint32_t)gr[2 * j + 1], 0x5410));
        const uint32_t halo_h2 = h2_bits(u16x2_to_half2((uint32_t)gh | ((uint32_t)gh << 16)));

        // neighbours across lanes (both halves of the shuffled register are valid gray values)
        uint32_t from_left = __shfl_up_sync(0xffffffffu, h2_bits(g_c[7]), 1);     // .hi = g[15] of lane-1
        uint32_t from_right = __shfl_down_sync(0xffffffffu, h2_bits(g_c[0]), 1);  // .lo = g[0] of lane+1
        if (left_own) from_left = h2_bits(g_c[0]);                 // .hi = own g[1]
        if (halo_left) from_left = halo_h2;
        if (right_own) from_right = h2_bits(g_c[7]);               // .lo = own g[14]
        if (halo_right) from_right = halo_h2;

        // shifted pairs S_j = (g[2j-1], g[2j]), j = 0..8
        uint32_t S[9];
        S[0] = __byte_perm(from_left, h2_bits(g_c[0]), 0x5432);
#pragma unroll
        for (int j = 1; j < 8; ++j) S[j] = __byte_perm(h2_bits(g_c[j - 1]), h2_bits(g_c[j]), 0x5432);
        S[8] = __byte_perm(h2_bits(g_c[7]), from_right, 0x5432);

        __half2 d_c[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
            d_c[j] = __hfma2(g_c[j], kMinus2, __hadd2(bits_h2(S[j]), bits_h2(S[j + 1])));

        if (owned) {
            // sum_x dxx over the lane's span telescopes to (gl - g0) + (gr - g15)
            __half2 t = __hsub2(bits_h2(S[0]), bits_h2(S[8]));      // (gl - g15, g0 - gr)
            acc_lh = add_f32_f16((uint16_t)(h2_bits(t) & 0xffffu), acc_lh);
            acc_lh2 = add_f32_f16((uint16_t)(h2_bits(t) >> 16), acc_lh2);
        }
        if (k == 1 || k == rb + 1) {
            // sum_y dyy telescopes to (g[r0-1]-g[r0]) + (g[r1]-g[r1-1]) per column
            __half2 t = __float2half2_rn(0.f);
#pragma unroll
            for (int j = 0; j < 8; ++j) t = __hadd2(t, __hsub2(g_p[j], g_c[j]));
            float s = __low2float(t) + __high2float(t);
            acc_lv += (k == 1) ? s : -s;
        }
        if (k >= 2) {
            float row_l2 = 0.f, row_n = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                __half2 dyy = __hfma2(g_p[j], kMinus2, __hadd2(g_pp[j], g_c[j]));
                __half2 L = __hadd2(d_p[j], dyy);
                __half2 N = __habs2(__hfma2(d_p[j], kMinus2, __hadd2(d_pp[j], d_c[j])));
                uint32_t lb = h2_bits(L), nb = h2_bits(N);
                row_l2 = fma_f32_f16((uint16_t)(lb & 0xffffu), row_l2);
                row_l2 = fma_f32_f16((uint16_t)(lb >> 16), row_l2);
                row_n = add_f32_f16((uint16_t)(nb & 0xffffu), row_n);
                row_n = add_f32_f16((uint16_t)(nb >> 16), row_n);
            }
            // 16 * 1020^2 < 2^24 and 16 * 2040 < 2^24: both row sums are exact in fp32
            acc_l2 += (unsigned long long)__float2uint_rn(row_l2);
            acc_n += __float2uint_rn(row_n);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            g_pp[j] = g_p[j];
            g_p[j] = g_c[j];
            d_pp[j] = d_p[j];
            d_p[j] = d_c[j];
        }
        cur = nx1;
        nx1 = nx2;
    }
    if (active) {
        out_l2 += acc_l2;
        out_n += acc_n;
        out_l += (long long)__float2int_rn(acc_lh - acc_lh2) + (long long)__float2int_rn(acc_lv);
    }
}

template <bool RGB, bool LUMA>
__global__ void __launch_bounds__(kThreads, 1) tech_stats_kernel(TechArgs a) {
    unsigned int* const smem = fb_smem;
    unsigned int* const s_hs = fb_smem;
    unsigned int* const s_h256 = fb_smem + kOffH256;
    unsigned int* const s_sdiv = fb_smem + kOffSdiv;
    unsigned int* const s_hdiv = fb_smem + kOffHdiv;
    unsigned int* const s_next = fb_smem + kOffHdiv + 256;       // next unclaimed unit of the current segment

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    for (int i = tid; i < kHsBins + 256 * kH256Copies; i += kThreads) smem[i] = 0u;
    if (tid < 256) {
        s_sdiv[tid] = (unsigned int)sdiv_entry(tid);
        s_hdiv[tid] = (unsigned int)hdiv_entry(tid);
    }
    if (tid == 0) *s_next = 0u;
    __syncthreads();

    const long long upi = (long long)a.tiles_x * a.units_y;
    const long long total = upi * a.n;
    const long long u_begin = total * blockIdx.x / gridDim.x;
    const long long u_end = total * (blockIdx.x + 1) / gridDim.x;

    for (long long u0 = u_begin; u0 < u_end;) {
        const int img_idx = (int)(u0 / upi);
        const long long img_first = (long long)img_idx * upi;
        const long long seg_end = min(u_end, img_first + upi);
        const uint8_t* img = a.img + (size_t)img_idx * a.img_stride;

        unsigned long long acc_l2 = 0ull, acc_n = 0ull;
        long long acc_l = 0;
        // warps claim units of the segment dynamically (s_next is reset between segments)
        for (;;) {
            long long u = 0;
            if ((tid & 31) == 0) u = u0 + (long long)atomicAdd(s_next, 1u);
            u = __shfl_sync(0xffffffffu, u, 0);
            if (u >= seg_end) break;
            const int ul = (int)(u - img_first);
            const int tx = ul % a.tiles_x;
            if ((tx + 1) * kTileW <= a.W) process_unit<RGB, true, LUMA>(a, img, tx, ul / a.tiles_x, acc_l2, acc_n, acc_l);
            else process_unit<RGB, false, LUMA>(a, img, tx, ul / a.tiles_x, acc_l2, acc_n, acc_l);
        }
        acc_l2 = warp_sum_u64(acc_l2);
        acc_n = warp_sum_u64(acc_n);
        unsigned long long l_bits = warp_sum_u64((unsigned long long)acc_l);
        if ((tid & 31) == 0) {
            unsigned long long* s = a.sums + (size_t)img_idx * 4;
            atomicAdd(s + 0, l_bits);
            atomicAdd(s + 1, acc_l2);
            atomicAdd(s + 2, acc_n);
        }
        __syncthreads();
        // merge the CTA-private histograms into the image's global ones
        if (tid == 0) *s_next = 0u;
        unsigned int* g_hs = a.hs + (size_t)img_idx * kHsBins;
        for (int i = tid; i < kHsBins; i += kThreads) {
            unsigned int c = s_hs[i];
            if (c) {
                atomicAdd(g_hs + i, c);
                s_hs[i] = 0u;
            }
        }
        if (tid < 256) {
            unsigned int c = 0;
#pragma unroll 8
            for (int j = 0; j < kH256Copies; ++j) {
                int jj = (j + tid) & (kH256Copies - 1);
                c += s_h256[tid * kH256Copies + jj];
                s_h256[tid * kH256Copies + jj] = 0u;
            }
            if (c) atomicAdd(a.hist256 + (size_t)img_idx * 256 + tid, c);
        }
        __syncthreads();
        u0 = seg_end;
    }
}

// ---------------------------------------------------------------------------------------------
// Generic kernel: any shape >= 2x2, any alignment.  One thread per pixel, 27 byte loads.
// ---------------------------------------------------------------------------------------------
template <bool RGB>
__global__ void __launch_bounds__(256) tech_stats_generic_kernel(TechArgs a) {
    __shared__ unsigned int s_h256[256];
    __shared__ unsigned int s_sdiv[256], s_hdiv[256];
    const int tid = threadIdx.x;
    s_h256[tid] = 0u;
    s_sdiv[tid] = (unsigned int)sdiv_entry(tid);
    s_hdiv[tid] = (unsigned int)hdiv_entry(tid);
    __syncthreads();
    const int img_idx = blockIdx.y;
    const uint8_t* img = a.img + (size_t)img_idx * a.img_stride;
    const int W = a.W, H = a.H;
    const long long npx = (long long)W * H;
    long long acc_l = 0;
    unsigned long long acc_l2 = 0, acc_n = 0;
    unsigned int* g_hs = a.hs + (size_t)img_idx * kHsBins;
    for (long long i = (long long)blockIdx.x * blockDim.x + tid; i < npx; i += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(i / W), x = (int)(i % W);
        int g[3][3];
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
            int yy = y + dy;
            yy = yy < 0 ? 1 : (yy >= H ? H - 2 : yy);
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
                int xx = x + dx;
                xx = xx < 0 ? 1 : (xx >= W ? W - 2 : xx);
                const uint8_t* p = img + ((size_t)yy * W + xx) * 3;
                int c0 = p[0], c1 = p[1], c2 = p[2];
                g[dy + 1][dx + 1] = gray_of(RGB ? c2 : c0, c1, RGB ? c0 : c2);
                if (dy == 0 && dx == 0) {
                    atomicAdd(g_hs + hs_bin_of(RGB ? c2 : c0, c1, RGB ? c0 : c2, s_sdiv, s_hdiv), 1u);
                    if (a.luma)
                        a.luma[(size_t)img_idx * W * H + i] =
                            (uint8_t)((7471 * (RGB ? c2 : c0) + 38470 * c1 + 19595 * (RGB ? c0 : c2) + 0x8000) >> 16);
                }
            }
        }
        atomicAdd(&s_h256[g[1][1]], 1u);
        int L = g[0][1] + g[2][1] + g[1][0] + g[1][2] - 4 * g[1][1];
        int N = g[0][0] + g[0][2] + g[2][0] + g[2][2] - 2 * (g[0][1] + g[2][1] + g[1][0] + g[1][2]) + 4 * g[1][1];
        acc_l += L;
        acc_l2 += (unsigned long long)(L * L);
        acc_n += (unsigned long long)abs(N);
    }
    acc_l2 = warp_sum_u64(acc_l2);
    acc_n = warp_sum_u64(acc_n);
    unsigned long long l_bits = warp_sum_u64((unsigned long long)acc_l);
    if ((tid & 31) == 0) {
        unsigned long long* s = a.sums + (size_t)img_idx * 4;
        atomicAdd(s + 0, l_bits);
        atomicAdd(s + 1, acc_l2);
        atomicAdd(s + 2, acc_n);
    }
    __syncthreads();
    if (s_h256[tid]) atomicAdd(a.hist256 + (size_t)img_idx * 256 + tid, s_h256[tid]);
}

// ---------------------------------------------------------------------------------------------
// Per-image reductions of the HS histogram: entropy (technical.py:97-104) and sum of S
// (technical.py:237).  One CTA per image, fixed reduction order (deterministic).
// out[img] = { entropy_bits, sum_s, nonzero_bins, total_count }
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) hs_derive_kernel(const unsigned int* hs, double* out) {
    __shared__ double s_e[32], s_s[32], s_z[32], s_t[32];
    const unsigned int* h = hs + (size_t)blockIdx.x * kHsBins;
    const int tid = threadIdx.x;
    double tot = 0.0;
    for (int i = tid; i < kHsBins; i += 1024) tot += (double)h[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    if ((tid & 31) == 0) s_t[tid >> 5] = tot;
    __syncthreads();
    tot = 0.0;
    for (int i = 0; i < 32; ++i) tot += s_t[i];
    double e = 0.0, ss = 0.0, nz = 0.0;
    for (int i = tid; i < kHsBins; i += 1024) {
        unsigned int c = h[i];
        if (c) {
            double p = (double)c / tot;
            e -= p * log2(p);
            ss += (double)c * (double)(i & 255);
            nz += 1.0;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        e += __shfl_xor_sync(0xffffffffu, e, o);
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
        nz += __shfl_xor_sync(0xffffffffu, nz, o);
    }
    __syncthreads();
    if ((tid & 31) == 0) {
        s_e[tid >> 5] = e;
        s_s[tid >> 5] = ss;
        s_z[tid >> 5] = nz;
    }
    __syncthreads();
    if (tid == 0) {
        double E = 0, S = 0, Z = 0;
        for (int i = 0; i < 32; ++i) {
            E += s_e[i];
            S += s_s[i];
            Z += s_z[i];
        }
        double* o = out + (size_t)blockIdx.x * 4;
        o[0] = E;
        o[1] = S;
        o[2] = Z;
        o[3] = tot;
    }
}

// gray / hsv planes for callers that still want the arrays of image_cache.py:30-31.
template <bool RGB>
__global__ void __launch_bounds__(256) gray_hsv_kernel(const uint8_t* __restrict__ img, long long npx,
                                                       uint8_t* __restrict__ gray, uint8_t* __restrict__ hsv) {
    __shared__ unsigned int s_sdiv[256], s_hdiv[256];
    s_sdiv[threadIdx.x] = (unsigned int)sdiv_entry(threadIdx.x);
    s_hdiv[threadIdx.x] = (unsigned int)hdiv_entry(threadIdx.x);
    __syncthreads();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += (long long)gridDim.x * blockDim.x) {
        const int c0 = img[3 * i], c1 = img[3 * i + 1], c2 = img[3 * i + 2];
        const int b = RGB ? c2 : c0, g = c1, r = RGB ? c0 : c2;
        gray[i] = (uint8_t)gray_of(b, g, r);
        const int bin = hs_bin_of(b, g, r, s_sdiv, s_hdiv);
        hsv[3 * i] = (uint8_t)(bin >> 8);
        hsv[3 * i + 1] = (uint8_t)(bin & 255);
        hsv[3 * i + 2] = (uint8_t)max(max(b, g), r);
    }
}

}  // namespace

int launch_gray_hsv(const uint8_t* d_image, int H, int W, int rgb_order, uint8_t* d_gray, uint8_t* d_hsv,
                    cudaStream_t stream) {
    FB_REQUIRE(d_image && d_gray && d_hsv && H >= 1 && W >= 1, "fb_gray_hsv: bad arguments");
    const long long npx = (long long)H * W;
    int blocks = (int)((npx + 255) / 256);
    if (blocks > sm_count() * 16) blocks = sm_count() * 16;
    if (rgb_order) gray_hsv_kernel<true><<<blocks, 256, 0, stream>>>(d_image, npx, d_gray, d_hsv);
    else gray_hsv_kernel<false><<<blocks, 256, 0, stream>>>(d_image, npx, d_gray, d_hsv);
    FB_CUDA_OK(cudaGetLastError());
    return 0;
}

int tech_rows_per_unit(int n, int H, int W, int sms) {
    // aim for >= 8 units per warp so the static round-robin balances, cap the halo overhead at ~4 %
    const int tiles_x = (W + kTileW - 1) / kTileW;
    const long long want = (long long)sms * kWarps * 8;
    long long units_y = (want + (long long)n * tiles_x - 1) / ((long long)n * tiles_x);
    if (units_y < 1) units_y = 1;
    int rows = (int)((H + units_y - 1) / units_y);
    if (rows < 48) rows = 48;
    if (rows > 512) rows = 512;
    if (rows > H) rows = H;
    return rows;
}

int launch_tech_stats(const uint8_t* d_images, int n, int H, int W, long long image_stride, int rgb_order,
                      unsigned int* d_hist256, unsigned int* d_hs_hist, long long* d_sums, int force_generic,
                      uint8_t* d_luma, cudaStream_t stream) {
    FB_REQUIRE(d_images && d_hist256 && d_hs_hist && d_sums, "fb_tech_stats: null pointer");
    FB_REQUIRE(n >= 1 && H >= 2 && W >= 2, "fb_tech_stats: need n>=1 and images of at least 2x2 (got n=%d %dx%d)", n, H, W);
    FB_REQUIRE(image_stride >= (long long)H * W * 3, "fb_tech_stats: image_stride smaller than one image");
    FB_CUDA_OK(cudaMemsetAsync(d_hist256, 0, (size_t)n * 256 * sizeof(unsigned int), stream));
    FB_CUDA_OK(cudaMemsetAsync(d_hs_hist, 0, (size_t)n * kHsBins * sizeof(unsigned int), stream));
    FB_CUDA_OK(cudaMemsetAsync(d_sums, 0, (size_t)n * 4 * sizeof(long long), stream));

    TechArgs a;
    a.img = d_images;
    a.img_stride = image_stride;
    a.n = n;
    a.H = H;
    a.W = W;
    a.hist256 = d_hist256;
    a.hs = d_hs_hist;
    a.sums = reinterpret_cast<unsigned long long*>(d_sums);
    a.gray_round = 1 << 14;
    a.luma = d_luma;
    a.luma_round = 1 << 15;
    FB_REQUIRE(!d_luma || image_stride == (long long)H * W * 3, "fb_tech_stats: the luma plane needs a contiguous batch");
    FB_REQUIRE(!d_luma || (reinterpret_cast<uintptr_t>(d_luma) & 15) == 0, "fb_tech_stats: luma plane must be 16-byte aligned");
    const bool aligned = (W % 16 == 0) && ((reinterpret_cast<uintptr_t>(d_images) & 15) == 0) &&
                         (image_stride % 16 == 0);
    const int sms = sm_count();
    if (aligned && !force_generic) {
        a.tiles_x = (W + kTileW - 1) / kTileW;
        a.rows_per_unit = tech_rows_per_unit(n, H, W, sms);
        a.units_y = (H + a.rows_per_unit - 1) / a.rows_per_unit;
        const size_t smem = kSmemWords * sizeof(unsigned int);
        auto kern = d_luma ? (rgb_order ? tech_stats_kernel<true, true> : tech_stats_kernel<false, true>)
                           : (rgb_order ? tech_stats_kernel<true, false> : tech_stats_kernel<false, false>);
        FB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        long long total_units = (long long)a.tiles_x * a.units_y * n;
        int grid = (int)(total_units < sms ? total_units : sms);
        kern<<<grid, kThreads, smem, stream>>>(a);
    } else {
        a.tiles_x = a.units_y = a.rows_per_unit = 0;
        long long npx = (long long)H * W;
        int gx = (int)((npx + 255) / 256);
        if (gx > sms * 8) gx = sms * 8;
        dim3 grid(gx, n);
        if (rgb_order) tech_stats_generic_kernel<true><<<grid, 256, 0, stream>>>(a);
        else tech_stats_generic_kernel<false><<<grid, 256, 0, stream>>>(a);
    }
    FB_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_hs_derive(const unsigned int* d_hs_hist, int n, double* d_out, cudaStream_t stream) {
    FB_REQUIRE(d_hs_hist && d_out && n >= 1, "fb_tech_derive: bad arguments");
    hs_derive_kernel<<<n, 1024, 0, stream>>>(d_hs_hist, d_out);
    FB_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace fb
