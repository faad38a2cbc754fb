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
// Fast kernel (W % 8 == 0, 8-byte aligned rows):
//   * persistent grid = #SMs, 640 threads, 1 CTA/SM; the CTA owns a contiguous range of
//     "units" (256-px x R-row tiles, R <= 128) so that it mostly stays inside one image; warps claim
//     units dynamically
//   * a warp walks down its tile: lane = 8 consecutive pixels (3 x LDG.64 per row, the next row
//     prefetched behind the first use of the current one); gray straight from the interleaved bytes
//     with IDP.2A; gray, dxx and the two second-difference carries of the previous rows stay in
//     registers as exact fp16 pairs, so the 4-neighbour Laplacian and the Immerkaer response (outer
//     product of [1,-2,1]) cost one pass of packed half2 adds
//   * OpenCV RGB2HSV_b two pixels per instruction in packed fp16; its two fixed-point divisions as
//     one round-down packed fp32 FMA on reciprocal tables (bit-exact, see hsv_pair / fixed_products)
//   * HS histogram (180 rows x 257 words) and a lane-private luminance histogram (256 x 32 lanes,
//     conflict-free) live in shared memory at compile-time offsets; they are merged into the
//     per-image global histograms only when the CTA crosses an image boundary
// Generic kernel: any W/H >= 2, any alignment; one thread per pixel, global atomics.
#include <cuda_fp16.h>

#include <mutex>

#include "common.cuh"
#include "kernels.h"

namespace fb {

namespace {

constexpr int kHsBins = 180 * 256;
#ifndef FB_TECH_LANE_PX
#define FB_TECH_LANE_PX 8
#endif
#ifndef FB_TECH_THREADS
#define FB_TECH_THREADS 640
#endif
constexpr int kThreads = FB_TECH_THREADS;
constexpr int kWarps = kThreads / 32;
constexpr int kLanePx = FB_TECH_LANE_PX;       // pixels per lane per row: 8 (3 x LDG.64) or 16 (3 x LDG.128)
constexpr int kLaneWords = 3 * kLanePx / 4;
constexpr int kPairs = kLanePx / 2;
constexpr int kGroups = kLanePx / 4;
constexpr int kTileW = 32 * kLanePx;   // pixels per warp row (256 at 8 px per lane)
constexpr int kH256Copies = 32;        // one luminance-histogram column per lane
constexpr int kHsStride = 257;          // shared-memory row stride of the H-S histogram: bank = (h + s) mod 32, so
                                       // pixels of similar saturation and different hue do not collide
constexpr int kHsSmemWords = 180 * kHsStride + 12;   // padded to a multiple of 16 words
constexpr size_t kSmemWords = kHsSmemWords + 256 * kH256Copies + 512 + 16;   // + work counter
constexpr int kOffH256 = kHsSmemWords;                  // word offsets inside the dynamic smem block
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
    uint8_t* box;     // optional [n][ceil(H/4)][W/4][3] 4x4 box reduction (Pillow ImagingReduce, first step of the thumbnail)
    unsigned int box_mult_full, box_mult_bottom;   // Pillow's multipliers for 16-pixel boxes / the shorter boxes of the last rows
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

// Shared-window address of the dynamic shared memory block.  On sm_100 the first KiB of the window is reserved
// for the system, so a kernel without static shared memory sees its dynamic block at 0x400; the launcher
// verifies this once with a probe kernel (and falls back to the generic kernel if it ever differs), which
// lets every histogram / table access use a compile-time offset instead of an address add.
constexpr uint32_t kSmemBase = 0x400;

struct RowRegs {
    uint32_t w[kLaneWords];   // 3 * kLanePx bytes
    uint32_t halo;            // the aligned word that holds the one extra pixel lane 0 / lane 31 may need
};

// histogram increment in shared memory at [addr + OFF] (ptxas turns "+1" into the warp-aggregating
// ATOMS.POPC.INC, which must not be predicated: FULL tiles call it unconditionally).  No "memory" clobber:
// the increments only touch histogram words, which nothing else reads or writes before the __syncthreads()
// that precedes the merge, so table reads may be scheduled across them.
template <uint32_t OFF>
__device__ __forceinline__ void smem_inc(uint32_t addr) {
    asm volatile("red.shared.add.u32 [%0+%1], %2;" :: "r"(addr), "n"(OFF), "r"(1u));
}
template <uint32_t OFF>
__device__ __forceinline__ float smem_ld_f32(uint32_t addr) {
    float r;
    asm("ld.shared.f32 %0, [%1+%2];" : "=f"(r) : "r"(addr), "n"(OFF));
    return r;
}

// 16-bit coefficient x 8-bit pixel dot products (IDP.2A): .lo uses bytes 0,1 of px, .hi bytes 2,3
__device__ __forceinline__ uint32_t dp2a_lo(uint32_t coef, uint32_t px, uint32_t acc) {
    uint32_t r;
    asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(coef), "r"(px), "r"(acc));
    return r;
}
__device__ __forceinline__ uint32_t dp2a_hi(uint32_t coef, uint32_t px, uint32_t acc) {
    uint32_t r;
    asm("dp2a.hi.u32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(coef), "r"(px), "r"(acc));
    return r;
}

// Weighted channel sum of the 4 pixels held in three consecutive words (12 interleaved bytes):
// out[p] = c0 + k0*ch0 + k1*ch1 + k2*ch2, two IDP.2A per pixel, no byte extraction.
//   k01 = k0 | k1 << 16, k2_ = k2, k_0 = k0 << 16, k12 = k1 | k2 << 16
__device__ __forceinline__ void dot4(uint32_t a, uint32_t b, uint32_t c, uint32_t k01, uint32_t k2_, uint32_t k_0,
                                     uint32_t k12, uint32_t c0, uint32_t* out) {
    out[0] = dp2a_hi(k2_, a, dp2a_lo(k01, a, c0));     // a0 a1 a2
    out[1] = dp2a_lo(k12, b, dp2a_hi(k_0, a, c0));     // a3 b0 b1
    out[2] = dp2a_lo(k2_, c, dp2a_hi(k01, b, c0));     // b2 b3 c0
    out[3] = dp2a_hi(k12, c, dp2a_lo(k_0, c, c0));     // c1 c2 c3
}

// Per-lane constants of one unit.
struct LaneCtx {
    bool active, left_own, right_own, halo_left, halo_right;
    uint32_t h256_off;    // byte offset of this lane's luminance column, minus 0x6400 * 128
    uint32_t k64;         // 0x64646464 kept in a register (PRMT source of the fp16 exponent byte)
    uint32_t hs_unbias;   // minus the two float magic numbers (0x4B000000 * 4 + 0x4B400000 * 4 * kHsStride)
    int halo_delta;       // byte distance from the lane's first pixel to the aligned word with the halo pixel
    uint8_t* luma_lane;   // LUMA: address of (row 0, xl) in the luma plane
    uint8_t* box_lane;    // BOX: address of the lane's two boxes in box row 0 of the image
    unsigned int box_mult_full, box_mult_bottom;
    int W, H;
};

template <bool NEED_HALO_CHECK = true>
__device__ __forceinline__ void load_row(RowRegs& r, const uint8_t* img, uint32_t off, bool need_halo, int halo_delta) {
    const uint8_t* p = img + off;
    if (kLanePx == 16) {
#pragma unroll
        for (int i = 0; i < kLaneWords / 4; ++i) {
            const uint4 q = ldg_nc_v4(p + 16 * i);
            r.w[4 * i] = q.x, r.w[4 * i + 1] = q.y, r.w[4 * i + 2] = q.z, r.w[4 * i + 3] = q.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < kLaneWords / 2; ++i) {
            const uint2 q = ldg_nc_v2(p + 8 * i);
            r.w[2 * i] = q.x, r.w[2 * i + 1] = q.y;
        }
    }
    r.halo = 0u;
    if (need_halo) r.halo = __ldg(reinterpret_cast<const unsigned int*>(p + halo_delta));
}

// Running sums of a warp: sum L^2, sum |N| (exact integers) and the telescoped sum of L.
struct WarpAcc {
    unsigned long long l2, n;
    long long l;
    // per unit: integer-valued floats / a 32-bit integer, flushed into the 64-bit sums at the end of the unit
    // (128 rows x 16 px: sum L < 2^24, sum |N| < 2^24 exactly representable; sum L^2 < 2^32)
    float lf, nf;
    unsigned int l2u;
};

// OpenCV RGB2HSV_b (hue range 180) of two pixels held as biased fp16 pairs (bits 0x6400 | value).  All
// packed steps are exact in fp16 (|values| <= 1275).  Outputs: the table offsets of v and d for both pixels
// (fp16 bit patterns 0x6400 + 4 v), d and the hue numerator as fp16 pairs.
struct HsvPair {
    uint32_t v4, d4;
    __half2 d, hr;
};

template <bool RGB>
__device__ __forceinline__ HsvPair hsv_pair(uint32_t x0, uint32_t x1, uint32_t x2) {
    const __half2 B = bits_h2(RGB ? x2 : x0), G = bits_h2(x1), R = bits_h2(RGB ? x0 : x2);
    const __half2 v = __hmax2(__hmax2(B, G), R);
    const __half2 mn = __hmin2(__hmin2(B, G), R);
    HsvPair o;
    o.d = __hsub2(v, mn);
    const __half2 gb = __hsub2(G, B), br = __hsub2(B, R), rg = __hsub2(R, G);
    const __half2 two = __float2half2_rn(2.f), four = __float2half2_rn(4.f);
    const uint32_t cg = h2_bits(__hfma2(o.d, two, br));
    const uint32_t cb = h2_bits(__hfma2(o.d, four, rg));
    const uint32_t m_r = __heq2_mask(v, R), m_g = __heq2_mask(v, G);
    const uint32_t t = (m_g & cg) | (~m_g & cb);
    o.hr = bits_h2((m_r & h2_bits(gb)) | (~m_r & t));
    // table offsets straight from the fp16 bit patterns: 1024 + 4 v has bits 0x6400 + 4 v
    o.v4 = h2_bits(__hfma2(v, four, __float2half2_rn(-3072.f)));
    o.d4 = h2_bits(__hfma2(o.d, four, __float2half2_rn(1024.f)));
    return o;
}

// The two fixed-point products of RGB2HSV_b run as one round-down packed fp32 FMA (FFMA2) on tables
// pre-scaled by 2^-12, followed by a round-down add of 2^23 / 1.5 * 2^23 that leaves the integers in the
// low mantissa bits:   floor((d * sdiv[v] + 2048) / 4096) = floor(RD(d * sdiv[v] / 4096 + 0.5))
__device__ __forceinline__ void fixed_products(float d, float hr, float sdf, float hdf, uint32_t& sb, uint32_t& hb) {
    asm("{.reg .b64 a, b, c, m;\n\t"
        "mov.b64 a, {%2, %3};\n\t"
        "mov.b64 b, {%4, %5};\n\t"
        "mov.b64 c, {%6, %6};\n\t"
        "mov.b64 m, {%7, %8};\n\t"
        "fma.rm.f32x2 a, a, b, c;\n\t"
        "add.rm.f32x2 a, a, m;\n\t"
        "mov.b64 {%0, %1}, a;}"
        : "=r"(sb), "=r"(hb)
        : "f"(d), "f"(hr), "f"(sdf), "f"(hdf), "f"(0.5f), "f"(8388608.f), "f"(12582912.f));
}

// One H-S histogram increment per pixel; NP pairs per call so that their 4 NP table reads are in flight
// together.
template <bool FULL, int NP>
__device__ __forceinline__ void hsv_bins(const HsvPair (&hp)[NP], const LaneCtx& ln) {
    float sdf[2 * NP], hdf[2 * NP];
#pragma unroll
    for (int i = 0; i < 2 * NP; ++i) {
        const uint32_t v4 = hp[i >> 1].v4, d4 = hp[i >> 1].d4;
        const uint32_t iv = (i & 1) ? (v4 >> 16) : (v4 & 0xffffu);
        const uint32_t id = (i & 1) ? (d4 >> 16) : (d4 & 0xffffu);
        sdf[i] = smem_ld_f32<kSmemBase + 4 * kOffSdiv - 0x6400>(iv);
        hdf[i] = smem_ld_f32<kSmemBase + 4 * kOffHdiv - 0x6400>(id);
    }
#pragma unroll
    for (int i = 0; i < 2 * NP; ++i) {
        const __half2 d = hp[i >> 1].d, hr = hp[i >> 1].hr;
        const float df = (i & 1) ? __high2float(d) : __low2float(d);
        const float hf = (i & 1) ? __high2float(hr) : __low2float(hr);
        uint32_t sb, hb;      // 0x4B000000 + s, 0x4B400000 + h
        fixed_products(df, hf, sdf[i], hdf[i], sb, hb);
        // byte offset of the bin, negative (wrapped) when h < 0: the smaller of (a, a + 180 rows) is h mod 180
        // (IMAD + LEA + VIADDMNMX; written as PTX so that the three steps are not re-associated into four)
        uint32_t off;
        asm("{.reg .b32 t, u;\n\t"
            "mad.lo.u32 t, %2, %4, %3;\n\t"
            "shl.b32 u, %1, 2;\n\t"
            "add.u32 t, t, u;\n\t"
            "add.u32 u, t, %5;\n\t"
            "min.u32 %0, t, u;}"
            : "=r"(off) : "r"(sb), "r"(hb), "r"(ln.hs_unbias), "n"(4 * kHsStride), "n"(180 * 4 * kHsStride));
        if (FULL || ln.active) smem_inc<kSmemBase>(off);
    }
}

enum RowMode { ROW_FIRST = 0, ROW_SECOND = 1, ROW_STEADY = 2, ROW_LAST = 3 };

// Channel sums of the lane's 4-pixel-wide boxes over the rows of the current box row (BOX variant).
struct BoxAcc {
    unsigned int s[kGroups][3];
};

// One image row of the lane's pixel span.  Stencil state carried between rows (exact fp16 pairs; gray keeps
// its +1024 bias, which cancels in every difference):
//   g_p = gray of the previous row, p1 = g[y-2] - 2 g[y-1] - 1024, d_p = dxx of the previous row,
//   q1 = dxx[y-2] - 2 dxx[y-1]
// MODE: FIRST  = halo row above the unit (gray only)
//       SECOND = first owned row: histograms, vertical telescoped sum picks up g_p - g_c
//       STEADY = owned row: histograms + Laplacian / Immerkaer response of the previous row
//       LAST   = halo row below the unit: stencil of the last owned row, vertical sum picks up -(g_p - g_c)
template <bool RGB, bool FULL, bool LUMA, bool BOX, int MODE>
__device__ __forceinline__ void row_step(const RowRegs& cur, RowRegs& nxt, const uint8_t* img, uint32_t next_off,
                                         bool need_halo, const LaneCtx& ln, int y,
                                         const __half2 (&g_p)[kPairs], __half2 (&g_c)[kPairs],
                                         const __half2 (&d_p)[kPairs], __half2 (&d_c)[kPairs], __half2 (&p1)[kPairs],
                                         __half2 (&q1)[kPairs], WarpAcc& acc, BoxAcc& bx) {
    constexpr bool hist = (MODE == ROW_SECOND || MODE == ROW_STEADY);
    constexpr bool sten = (MODE == ROW_STEADY || MODE == ROW_LAST);
#ifdef FB_TECH_PREFETCH_EARLY
    if (MODE != ROW_LAST) load_row(nxt, img, next_off, need_halo, ln.halo_delta);
#endif
    const uint32_t (&w)[kLaneWords] = cur.w;
    // gray = (3735 B + 19235 G + 9798 R + 2^14) >> 15, computed with doubled coefficients so that the high
    // half of the accumulator is the fp16 bit pattern of 1024 + gray
    constexpr uint32_t kB = 2 * 3735, kG = 2 * 19235, kR = 2 * 9798;
    constexpr uint32_t k0 = RGB ? kR : kB, k2 = RGB ? kB : kR;
    constexpr uint32_t k01 = k0 | (kG << 16), k2_ = k2, k_0 = k0 << 16, k12 = kG | (k2 << 16);
    uint32_t ga[kLanePx];
#pragma unroll
    for (int q = 0; q < kGroups; ++q) dot4(w[3 * q], w[3 * q + 1], w[3 * q + 2], k01, k2_, k_0, k12, 0x64008000u, ga + 4 * q);
    // the halo pixel sits in bytes 1..3 (left: word before the span) or 0..2 (right: word after it)
    const uint32_t gha = ln.halo_left ? dp2a_hi(k12, cur.halo, dp2a_lo(k_0, cur.halo, 0x64008000u))
                                      : dp2a_hi(k2_, cur.halo, dp2a_lo(k01, cur.halo, 0x64008000u));
    const uint32_t halo_h2 = __byte_perm(gha, 0u, 0x3232);
    // prefetch the next row only now: issued before the first read of this row's registers, the new loads
    // would share a scoreboard with the loads that read waits for, and the read would wait for them too
#ifndef FB_TECH_PREFETCH_EARLY
    if (MODE != ROW_LAST) load_row(nxt, img, next_off, need_halo, ln.halo_delta);
#endif

    if (hist) {
#pragma unroll
        for (int p = 0; p < kLanePx; ++p) {
            uint32_t addr;     // (0x6400 + gray) * 128 + lane column; kept as SHF + IMAD (one per pipe)
            asm volatile("{.reg .b32 t; shr.u32 t, %1, 16; mad.lo.u32 %0, t, 128, %2;}" : "=r"(addr) : "r"(ga[p]), "r"(ln.h256_off));
            if (FULL || ln.active) smem_inc<kSmemBase + 4 * kOffH256>(addr);
        }
    }
    // gray pairs as biased fp16 (the accumulators die here)
#pragma unroll
    for (int j = 0; j < kPairs; ++j) g_c[j] = bits_h2(__byte_perm(ga[2 * j], ga[2 * j + 1], 0x7632));

    if (hist) {
        const uint32_t k64 = ln.k64;
#pragma unroll
        for (int q = 0; q < kGroups; ++q) {
            HsvPair hp[2];
            const uint32_t wa = w[3 * q], wb = w[3 * q + 1], wc = w[3 * q + 2];
            // channel pairs of pixels (0,1) and (2,3) of the group as biased fp16 pairs (0x64 in the odd bytes)
            const uint32_t p0 = __byte_perm(wa, k64, 0x4340);                              // a0 a3
            const uint32_t p1_ = __byte_perm(__byte_perm(wa, k64, 0x4441), wb, 0x1410);    // a1 b0
            const uint32_t p2 = __byte_perm(__byte_perm(wa, k64, 0x4442), wb, 0x1510);     // a2 b1
            hp[0] = hsv_pair<RGB>(p0, p1_, p2);
            const uint32_t r0 = __byte_perm(__byte_perm(wb, k64, 0x4442), wc, 0x1510);     // b2 c1
            const uint32_t r1 = __byte_perm(__byte_perm(wb, k64, 0x4443), wc, 0x1610);     // b3 c2
            const uint32_t r2 = __byte_perm(wc, k64, 0x4340);                              // c0 c3
            hp[1] = hsv_pair<RGB>(r0, r1, r2);
            hsv_bins<FULL, 2>(hp, ln);
        }
        if (LUMA) {   // Pillow convert('L'): (19595 R + 38470 G + 7471 B + 2^15) >> 16
            constexpr uint32_t l0 = RGB ? 19595u : 7471u, l2 = RGB ? 7471u : 19595u;
            uint32_t lw[kGroups];
#pragma unroll
            for (int q = 0; q < kGroups; ++q) {
                uint32_t la[4];
                dot4(w[3 * q], w[3 * q + 1], w[3 * q + 2], l0 | (38470u << 16), l2, l0 << 16, 38470u | (l2 << 16), 0x8000u, la);
                lw[q] = __byte_perm(__byte_perm(la[0], la[1], 0x0062), __byte_perm(la[2], la[3], 0x0062), 0x5410);
            }
            if (FULL || ln.active) {
                if (kLanePx == 16) *reinterpret_cast<uint4*>(ln.luma_lane + (size_t)y * ln.W) = make_uint4(lw[0], lw[1], lw[kGroups - 2], lw[kGroups - 1]);
                else *reinterpret_cast<uint2*>(ln.luma_lane + (size_t)y * ln.W) = make_uint2(lw[0], lw[1]);
            }
        }
    }

    if (BOX && hist) {
        // Pillow ImagingReduce(4, 4): channel sums of the 4-pixel groups with byte-selecting dot products, carried
        // over the (up to) four rows of a box row; ((sum + n/2) * multiplier) >> 24 leaves when the box row is complete
#pragma unroll
        for (int q = 0; q < kGroups; ++q) {
            const uint32_t w0 = w[3 * q], w1 = w[3 * q + 1], w2 = w[3 * q + 2];     // c0 c1 c2 c0 | c1 c2 c0 c1 | c2 c0 c1 c2
            bx.s[q][0] = __dp4a(w0, 0x01000001u, __dp4a(w1, 0x00010000u, __dp4a(w2, 0x00000100u, bx.s[q][0])));
            bx.s[q][1] = __dp4a(w0, 0x00000100u, __dp4a(w1, 0x01000001u, __dp4a(w2, 0x00010000u, bx.s[q][1])));
            bx.s[q][2] = __dp4a(w0, 0x00010000u, __dp4a(w1, 0x00000100u, __dp4a(w2, 0x01000001u, bx.s[q][2])));
        }
        const int yr = y & 3;
        if (yr == 3 || y == ln.H - 1) {          // warp-uniform
            const unsigned int half_n = 2u * (unsigned int)(yr + 1);
            const unsigned int m = yr == 3 ? ln.box_mult_full : ln.box_mult_bottom;
            uint8_t* o = ln.box_lane + (size_t)(y >> 2) * (size_t)(ln.W >> 2) * 3;
            uint16_t h16[3 * kGroups / 2];
#pragma unroll
            for (int i = 0; i < 3 * kGroups; i += 2) {
                const unsigned int b0 = ((bx.s[i / 3][i % 3] + half_n) * m) >> 24;
                const unsigned int b1 = ((bx.s[(i + 1) / 3][(i + 1) % 3] + half_n) * m) >> 24;
                h16[i / 2] = (uint16_t)(b0 | (b1 << 8));
            }
            if (FULL || ln.active) {
#pragma unroll
                for (int i = 0; i < 3 * kGroups / 2; ++i) reinterpret_cast<uint16_t*>(o)[i] = h16[i];
            }
#pragma unroll
            for (int q = 0; q < kGroups; ++q) bx.s[q][0] = bx.s[q][1] = bx.s[q][2] = 0u;
        }
    }

    // neighbours across lanes (both halves of the shuffled register are valid gray values)
    uint32_t from_left = __shfl_up_sync(0xffffffffu, h2_bits(g_c[kPairs - 1]), 1);     // .hi = last gray of lane-1
    uint32_t from_right = __shfl_down_sync(0xffffffffu, h2_bits(g_c[0]), 1);  // .lo = g[0] of lane+1
    if (ln.left_own) from_left = h2_bits(g_c[0]);                 // .hi = own g[1]
    if (ln.halo_left) from_left = halo_h2;
    if (ln.right_own) from_right = h2_bits(g_c[kPairs - 1]);      // .lo = own last-but-one gray
    if (ln.halo_right) from_right = halo_h2;

    // shifted pairs S_j = (g[2j-1], g[2j]), j = 0..kPairs
    uint32_t S[kPairs + 1];
    S[0] = __byte_perm(from_left, h2_bits(g_c[0]), 0x5432);
#pragma unroll
    for (int j = 1; j < kPairs; ++j) S[j] = __byte_perm(h2_bits(g_c[j - 1]), h2_bits(g_c[j]), 0x5432);
    S[kPairs] = __byte_perm(h2_bits(g_c[kPairs - 1]), from_right, 0x5432);

    // dxx = g[x-1] - 2 g[x] + g[x+1]; with biased gray (1024 + g) the fma leaves a - 2 g - 1024, exact below 2048
    const __half2 kMinus2 = __float2half2_rn(-2.f);
#pragma unroll
    for (int j = 0; j < kPairs; ++j) d_c[j] = __hadd2(__hfma2(g_c[j], kMinus2, bits_h2(S[j])), bits_h2(S[j + 1]));

    if (hist) {
        // sum_x dxx over the lane's span telescopes to (gl - g0) + (gr - g_last)
        __half2 t = __hsub2(bits_h2(S[0]), bits_h2(S[kPairs]));      // (gl - g_last, g0 - gr)
        float s = add_f32_f16((uint16_t)(h2_bits(t) & 0xffffu), 0.f);
        s = add_f32_f16((uint16_t)((h2_bits(t) >> 16) ^ 0x8000u), s);
        if (FULL || ln.active) acc.lf += s;
    }
    if (MODE == ROW_SECOND || MODE == ROW_LAST) {
        // sum_y dyy telescopes to (g[r0-1]-g[r0]) + (g[r1]-g[r1-1]) per column
        __half2 t = __float2half2_rn(0.f);
#pragma unroll
        for (int j = 0; j < kPairs; ++j) t = __hadd2(t, __hsub2(g_p[j], g_c[j]));
        const float s = __low2float(t) + __high2float(t);
        if (FULL || ln.active) acc.lf += (MODE == ROW_SECOND) ? s : -s;
    }
    if (sten) {
        float l2a = 0.f, l2b = 0.f, na = 0.f, nb = 0.f;
#pragma unroll
        for (int j = 0; j < kPairs; ++j) {
            const uint32_t lb = h2_bits(__hadd2(__hadd2(d_p[j], p1[j]), g_c[j]));      // dxx + dyy of row y-1
            const uint32_t nn = h2_bits(__habs2(__hadd2(q1[j], d_c[j])));              // |dyy(dxx)| of row y-1
            l2a = fma_f32_f16((uint16_t)(lb & 0xffffu), l2a);
            l2b = fma_f32_f16((uint16_t)(lb >> 16), l2b);
            na = add_f32_f16((uint16_t)(nn & 0xffffu), na);
            nb = add_f32_f16((uint16_t)(nn >> 16), nb);
        }
        // 16 * 1020^2 < 2^24 and 16 * 2040 < 2^24: both row sums are exact in fp32
        if (FULL || ln.active) {
            acc.l2u += __float2uint_rn(l2a + l2b);
            acc.nf += na + nb;
        }
    }
    if (MODE != ROW_LAST) {
#pragma unroll
        for (int j = 0; j < kPairs; ++j) {
            p1[j] = __hfma2(g_c[j], kMinus2, g_p[j]);      // g_p - 2 g_c - 1024
            q1[j] = __hfma2(d_c[j], kMinus2, d_p[j]);
        }
    }
}

template <bool RGB, bool FULL, bool LUMA, bool BOX>
__device__ __forceinline__ void process_unit(const TechArgs& a, const uint8_t* img, int tx, int uy, WarpAcc& acc) {
    const int lane = (int)lane_id();
    const int W = a.W, H = a.H;
    const int r0 = uy * a.rows_per_unit;
    const int rb = min(H, r0 + a.rows_per_unit) - r0;
    const int x0 = tx * kTileW + lane * kLanePx;
    LaneCtx ln;
    ln.active = x0 < W;
    const int xl = ln.active ? x0 : (W - kLanePx);
    // left / right neighbour of the lane's span
    ln.left_own = (xl == 0);                 // reflect-101: x=-1 -> x=1
    ln.right_own = (xl + kLanePx == W);      // x=W -> x=W-2
    ln.halo_left = (lane == 0) && !ln.left_own;
    ln.halo_right = (lane == 31) && !ln.right_own;
    const bool need_halo = ln.halo_left || ln.halo_right;
    ln.halo_delta = ln.halo_left ? -4 : 3 * kLanePx;
    ln.h256_off = 4u * (uint32_t)lane - 0x6400u * 128u;
    asm volatile("mov.b32 %0, 0x64646464;" : "=r"(ln.k64));
    asm volatile("mov.b32 %0, %1;" : "=r"(ln.hs_unbias) : "n"(0u - 4u * 0x4B000000u - (4u * kHsStride) * 0x4B400000u));
    ln.W = W;
    ln.H = H;
    ln.luma_lane = LUMA ? (a.luma + ((size_t)(img - a.img) / 3) + xl) : nullptr;
    // contiguous batch: image i's reduced plane starts at i * ceil(H/4) * (W/4) * 3
    ln.box_lane = BOX ? (a.box + ((size_t)((img - a.img) / ((size_t)H * W * 3)) * (size_t)((H + 3) >> 2) * (size_t)(W >> 2) + (size_t)(xl >> 2)) * 3) : nullptr;
    ln.box_mult_full = a.box_mult_full;
    ln.box_mult_bottom = a.box_mult_bottom;
    BoxAcc bx;
#pragma unroll
    for (int q = 0; q < kGroups; ++q) bx.s[q][0] = bx.s[q][1] = bx.s[q][2] = 0u;

    const uint32_t row_bytes = (uint32_t)W * 3u;
    const uint32_t lane_off = (uint32_t)xl * 3u;
    auto row_off = [&](int k) -> uint32_t {     // byte offset of the lane's span in the k-th row of the walk: r0-1 .. r1
        int y = r0 - 1 + k;
        y = (y < 0) ? 1 : y;
        y = (y >= H) ? (H - 2) : y;
        return (uint32_t)y * row_bytes + lane_off;
    };

    RowRegs LA, LB;
    __half2 GA[kPairs], GB[kPairs], DA[kPairs], DB[kPairs], P1[kPairs], Q1[kPairs];
#pragma unroll
    for (int j = 0; j < kPairs; ++j) GA[j] = GB[j] = DA[j] = DB[j] = P1[j] = Q1[j] = __float2half2_rn(0.f);
    acc.lf = acc.nf = 0.f;
    acc.l2u = 0u;

    // rows k = 0 .. rb+1 of the walk; row k prefetches row k+1.  Rows alternate between two register sets so
    // that (previous, current) swap by renaming instead of by moves.
    load_row(LA, img, row_off(0), need_halo, ln.halo_delta);
    row_step<RGB, FULL, LUMA, BOX, ROW_FIRST>(LA, LB, img, row_off(1), need_halo, ln, r0 - 1, GB, GA, DB, DA, P1, Q1, acc, bx);
    row_step<RGB, FULL, LUMA, BOX, ROW_SECOND>(LB, LA, img, row_off(2), need_halo, ln, r0, GA, GB, DA, DB, P1, Q1, acc, bx);
    bool odd_tail = false;
    int k = 2;
    for (; k <= rb; k += 2) {
        row_step<RGB, FULL, LUMA, BOX, ROW_STEADY>(LA, LB, img, row_off(k + 1), need_halo, ln, r0 - 1 + k, GB, GA, DB, DA, P1, Q1, acc, bx);
        if (k + 1 > rb) {
            odd_tail = true;
            break;
        }
        row_step<RGB, FULL, LUMA, BOX, ROW_STEADY>(LB, LA, img, row_off(k + 2), need_halo, ln, r0 + k, GA, GB, DA, DB, P1, Q1, acc, bx);
    }
    if (odd_tail) row_step<RGB, FULL, LUMA, BOX, ROW_LAST>(LB, LA, img, 0u, need_halo, ln, r0 + rb, GA, GB, DA, DB, P1, Q1, acc, bx);
    else row_step<RGB, FULL, LUMA, BOX, ROW_LAST>(LA, LB, img, 0u, need_halo, ln, r0 + rb, GB, GA, DB, DA, P1, Q1, acc, bx);

    acc.l += (long long)__float2int_rn(acc.lf);
    acc.n += (unsigned long long)__float2uint_rn(acc.nf);
    acc.l2 += (unsigned long long)acc.l2u;
}

__device__ unsigned int g_smem_base_probe;
__global__ void smem_base_probe_kernel() {
    if (threadIdx.x == 0) g_smem_base_probe = (unsigned int)__cvta_generic_to_shared(fb_smem);
}

template <bool RGB, bool LUMA, bool BOX>
__global__ void __launch_bounds__(kThreads, 1) tech_stats_kernel(TechArgs a) {
    unsigned int* const smem = fb_smem;
    unsigned int* const s_hs = fb_smem;
    unsigned int* const s_h256 = fb_smem + kOffH256;
    unsigned int* const s_sdiv = fb_smem + kOffSdiv;
    unsigned int* const s_hdiv = fb_smem + kOffHdiv;
    unsigned int* const s_next = fb_smem + kOffHdiv + 256;       // next unclaimed unit of the current segment

    const int tid = threadIdx.x;
    // every histogram / table access below uses compile-time offsets from kSmemBase: fail loudly (never silently
    // wrong) should the dynamic block of THIS launch start anywhere else (the launcher also probes once per device)
    if ((uint32_t)__cvta_generic_to_shared(fb_smem) != kSmemBase) __trap();
    for (int i = tid; i < kHsSmemWords + 256 * kH256Copies; i += kThreads) smem[i] = 0u;
    if (tid < 256) {
        // fixed-point reciprocals pre-scaled by 2^-12 (exact in fp32: integers below 2^21)
        s_sdiv[tid] = __float_as_uint((float)sdiv_entry(tid) * (1.f / 4096.f));
        s_hdiv[tid] = __float_as_uint((float)hdiv_entry(tid) * (1.f / 4096.f));
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

        WarpAcc acc;
        acc.l2 = acc.n = 0ull;
        acc.l = 0;
        // warps claim units of the segment dynamically (s_next is reset between segments)
        for (;;) {
            long long u = 0;
            if ((tid & 31) == 0) u = u0 + (long long)atomicAdd(s_next, 1u);
            u = __shfl_sync(0xffffffffu, u, 0);
            if (u >= seg_end) break;
            const int ul = (int)(u - img_first);
            const int tx = ul % a.tiles_x;
            if ((tx + 1) * kTileW <= a.W) process_unit<RGB, true, LUMA, BOX>(a, img, tx, ul / a.tiles_x, acc);
            else process_unit<RGB, false, LUMA, BOX>(a, img, tx, ul / a.tiles_x, acc);
        }
        const unsigned long long acc_l2 = warp_sum_u64(acc.l2);
        const unsigned long long acc_n = warp_sum_u64(acc.n);
        const unsigned long long l_bits = warp_sum_u64((unsigned long long)acc.l);
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
            const int si = (i >> 8) * kHsStride + (i & 255);
            unsigned int c = s_hs[si];
            if (c) {
                atomicAdd(g_hs + i, c);
                s_hs[si] = 0u;
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
    // aim for >= 32 units per warp: warps claim units dynamically and wait for each other at the end of an
    // image segment, so the tail is about one unit per segment; the two halo rows of a unit only cost the gray
    // conversion (about 1.5 % at 32 rows)
    const int tiles_x = (W + kTileW - 1) / kTileW;
    const long long want = (long long)sms * kWarps * 32;
    long long units_y = (want + (long long)n * tiles_x - 1) / ((long long)n * tiles_x);
    if (units_y < 1) units_y = 1;
    int rows = (int)((H + units_y - 1) / units_y);
    if (rows < 32) rows = 32;
    if (rows > 128) rows = 128;
    if (rows > H) rows = H;
    return rows;
}

// The fast kernel addresses shared memory with compile-time offsets from kSmemBase.  Checked once per device (not on
// the caller's stream: a private stream, no allocation, result cached under a mutex) with the real kernel's dynamic
// shared-memory size; if the base ever differs the generic kernel is used.  The kernel itself traps on a mismatch.
static bool smem_base_matches() {
    constexpr int kMaxDev = 64;
    static std::mutex mu;
    static int state[kMaxDev];          // 0 unknown, 1 ok, 2 mismatch
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev) return false;
    std::lock_guard<std::mutex> lk(mu);
    if (state[dev] == 0) {
        const int smem = (int)(kSmemWords * sizeof(unsigned int));
        unsigned int h_probe = 0;
        cudaStream_t s = nullptr;
        bool ok = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) == cudaSuccess &&
                  cudaFuncSetAttribute(smem_base_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) == cudaSuccess;
        if (ok) {
            smem_base_probe_kernel<<<1, 32, smem, s>>>();
            ok = cudaMemcpyFromSymbolAsync(&h_probe, g_smem_base_probe, sizeof(h_probe), 0, cudaMemcpyDeviceToHost, s) == cudaSuccess &&
                 cudaStreamSynchronize(s) == cudaSuccess;
        }
        if (s) cudaStreamDestroy(s);
        state[dev] = (ok && h_probe == kSmemBase) ? 1 : 2;
        if (state[dev] == 2)
            fprintf(stderr, "facet_b200: dynamic shared memory starts at 0x%x, not 0x%x: the technical pass uses its "
                            "generic (slow) kernel\n", h_probe, kSmemBase);
    }
    return state[dev] == 1;
}

int launch_tech_stats(const uint8_t* d_images, int n, int H, int W, long long image_stride, int rgb_order,
                      unsigned int* d_hist256, unsigned int* d_hs_hist, long long* d_sums, int force_generic,
                      uint8_t* d_luma, uint8_t* d_box4, const unsigned int* box_mult4, cudaStream_t stream) {
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
    a.box = d_box4;
    a.box_mult_full = box_mult4 ? box_mult4[0] : 0u;
    a.box_mult_bottom = box_mult4 ? box_mult4[2] : 0u;
    FB_REQUIRE(!d_box4 || box_mult4, "fb_tech_stats: the box reduction needs its multipliers");
    FB_REQUIRE(!d_box4 || image_stride == (long long)H * W * 3, "fb_tech_stats: the box reduction needs a contiguous batch");
    FB_REQUIRE(!d_box4 || (reinterpret_cast<uintptr_t>(d_box4) & 3) == 0, "fb_tech_stats: reduced plane must be 4-byte aligned");
    FB_REQUIRE(!d_luma || image_stride == (long long)H * W * 3, "fb_tech_stats: the luma plane needs a contiguous batch");
    FB_REQUIRE(!d_luma || (reinterpret_cast<uintptr_t>(d_luma) & 15) == 0, "fb_tech_stats: luma plane must be 16-byte aligned");
    constexpr int kAlign = kLanePx == 16 ? 16 : 8;     // vector width of the row loads
    const bool aligned = (W % kLanePx == 0) && ((reinterpret_cast<uintptr_t>(d_images) & (kAlign - 1)) == 0) &&
                         (image_stride % kAlign == 0) && (W >= kLanePx);
    const int sms = sm_count();
    const bool smem_base_ok = aligned && !force_generic && smem_base_matches();
    if (smem_base_ok) {
        a.tiles_x = (W + kTileW - 1) / kTileW;
        a.rows_per_unit = tech_rows_per_unit(n, H, W, sms);
        if (d_box4) a.rows_per_unit = (a.rows_per_unit + 3) & ~3;       // units start on box rows
        a.units_y = (H + a.rows_per_unit - 1) / a.rows_per_unit;
        const size_t smem = kSmemWords * sizeof(unsigned int);
        using Kern = void (*)(TechArgs);
        static const Kern table[8] = {tech_stats_kernel<false, false, false>, tech_stats_kernel<true, false, false>,
                                      tech_stats_kernel<false, true, false>,  tech_stats_kernel<true, true, false>,
                                      tech_stats_kernel<false, false, true>,  tech_stats_kernel<true, false, true>,
                                      tech_stats_kernel<false, true, true>,   tech_stats_kernel<true, true, true>};
        const Kern kern = table[(rgb_order ? 1 : 0) + (d_luma ? 2 : 0) + (d_box4 ? 4 : 0)];
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
        if (d_box4) {          // shapes the fast kernel does not take: the separate reduction pass of the thumbnail
            FB_CUDA_OK(cudaGetLastError());
            int rc = launch_box_reduce(d_images, n, H, W, image_stride, 4, 4, (H + 3) / 4, (W + 3) / 4, box_mult4, d_box4, stream);
            if (rc) return rc;
        }
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
