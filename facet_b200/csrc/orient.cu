// EXIF orientation + channel order of a decoded frame, on the device.
//
// Replaces the pixel work of utils/image_loading.py:101-106 of the reference for frames that are already in
// device memory: ImageOps.exif_transpose(pil_img) (one of PIL's seven transpose methods, chosen by EXIF tag
// 0x0112) followed by cv2.cvtColor(RGB2BGR).  Pure byte movement: 3 H W bytes read, 3 H W written.
//
// Every method is out(x', y') = in(sx, sy) with (u, v) = swap ? (y', x') : (x', y'), sx = flip_x ? W-1-u : u,
// sy = flip_y ? H-1-v : v.  A CTA moves one 64 x 64 pixel tile of the SOURCE through shared memory, so that the
// transposing methods read and write full sectors too.
//   fast path (full tile, 16-byte aligned rows; one instantiation per method): 16-byte loads into the input tile;
//     every thread gathers 4 output pixels (non-transposing methods: three word loads and compile-time byte
//     selections, mirror and channel swap included; transposing methods: 12 byte loads, one address per pixel)
//     into three words of an output tile whose 49-word pitch keeps the transposed writes conflict-free; the output
//     tile leaves as coalesced 16-byte (or 4-byte) stores.  Edge tiles whose sides are multiples of 16 x 4 pixels take
//     the same path with bounds.  For the transposing methods a warp takes 32 consecutive source columns of
//     the same four source rows, so its byte loads fall into 24 consecutive words.
//   generic path (edge tiles, odd shapes): same data flow with run-time indices, words where aligned, else bytes.
#include "common.cuh"
#include "kernels.h"

namespace fb {

namespace {

constexpr int kTile = 64;
constexpr int kPitch = kTile * 3 + 4;      // bytes per tile row in shared memory (a multiple of 4)
constexpr int kThreads = 256;

struct OrientArgs {
    const uint8_t* src;
    uint8_t* dst;
    long long src_stride, dst_stride;      // bytes between images
    int H, W;                              // source shape
    int tiles_x, tiles_y;
    int swap, flip_x, flip_y, swap_rb;
    int vec_ok;                            // rows of source and destination are 16-byte / 4-byte aligned
    int vec_out;                           // rows of the destination are 16-byte aligned as well
};

constexpr int kPitchIn = kTile * 3 + 16;     // fast path: 16-byte aligned rows
constexpr int kPitchOut = kTile * 3 + 4;     // 49 words: odd, so a column of the output tile spans all banks

// Byte j (0..11) of four output pixels = byte src_byte_of(j) of the four source pixels they come from (12 contiguous
// bytes, ascending x): mirrored pixel order when FX, reversed channel order when RB.
template <bool FX, bool RB>
__host__ __device__ constexpr int src_byte_of(int j) {
    const int i = j / 3, c = j % 3;
    return 3 * (FX ? 3 - i : i) + (RB ? 2 - c : c);
}

// One output word from bytes A, B, C, D (indices 0..11) of the three words (w0, w1, w2): one PRMT when they come from
// at most two of the words, two otherwise.
template <int A, int B, int C, int D>
__device__ __forceinline__ uint32_t pick4(uint32_t w0, uint32_t w1, uint32_t w2) {
    constexpr int wa[4] = {A >> 2, B >> 2, C >> 2, D >> 2};
    constexpr int ba[4] = {A & 3, B & 3, C & 3, D & 3};
    constexpr bool u0 = wa[0] == 0 || wa[1] == 0 || wa[2] == 0 || wa[3] == 0;
    constexpr bool u1 = wa[0] == 1 || wa[1] == 1 || wa[2] == 1 || wa[3] == 1;
    constexpr bool u2 = wa[0] == 2 || wa[1] == 2 || wa[2] == 2 || wa[3] == 2;
    if constexpr (u0 && u1 && u2) {
        // bytes of w0 / w1 into their final positions first (the positions fed by w2 are don't-cares), then w2
        constexpr uint32_t sel1 = (wa[0] == 1 ? 4 + ba[0] : ba[0]) | ((wa[1] == 1 ? 4 + ba[1] : ba[1]) << 4) |
                                  ((wa[2] == 1 ? 4 + ba[2] : ba[2]) << 8) | ((wa[3] == 1 ? 4 + ba[3] : ba[3]) << 12);
        constexpr uint32_t sel2 = (wa[0] == 2 ? 4 + ba[0] : 0) | ((wa[1] == 2 ? 4 + ba[1] : 1) << 4) |
                                  ((wa[2] == 2 ? 4 + ba[2] : 2) << 8) | ((wa[3] == 2 ? 4 + ba[3] : 3) << 12);
        return __byte_perm(__byte_perm(w0, w1, sel1), w2, sel2);
    } else {
        constexpr int X = u0 ? 0 : 1;                    // first word in use
        constexpr int Y = u2 ? 2 : 1;                    // last word in use (may equal X)
        const uint32_t x = X == 0 ? w0 : w1, y = Y == 2 ? w2 : w1;
        constexpr uint32_t sel = (wa[0] == X ? ba[0] : 4 + ba[0]) | ((wa[1] == X ? ba[1] : 4 + ba[1]) << 4) |
                                 ((wa[2] == X ? ba[2] : 4 + ba[2]) << 8) | ((wa[3] == X ? ba[3] : 4 + ba[3]) << 12);
        return __byte_perm(x, y, sel);
    }
}

// FULL: a 64 x 64 tile, no bounds checks.  Otherwise tw (a multiple of 16) x th (a multiple of 4) pixels.
// VEC: the destination rows are 16-byte aligned, so the output tile leaves as 16-byte stores.
template <bool SWAP, bool FX, bool FY, bool RB, bool FULL, bool VEC>
__device__ __forceinline__ void tile_fast(const uint8_t* __restrict__ src_tile, size_t src_pitch, uint8_t* __restrict__ dst_tile,
                                          size_t dst_pitch, uint8_t* s_in, uint8_t* s_out, int tid, int tw, int th) {
    if (FULL) tw = th = kTile;
    const int ow = SWAP ? th : tw, oh = SWAP ? tw : th;
    // th rows x (3 tw / 16) vectors in
    const int vin = (3 * tw) >> 4;
#pragma unroll
    for (int it = 0; it < 3; ++it) {
        const int i = tid + it * kThreads;
        const int r = i / 12, k = i - 12 * r;
        if (FULL || (r < th && k < vin)) {
            const uint4 v = ldg_nc_v4(src_tile + (size_t)r * src_pitch + 16 * k);
            *reinterpret_cast<uint4*>(s_in + r * kPitchIn + 16 * k) = v;
        }
    }
    __syncthreads();
    // groups of 4 output pixels: (row r, pixels 4q .. 4q+3) of the output tile
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const int g = tid + it * kThreads;
        const int r = SWAP ? (g & 63) : (g >> 4);
        const int q = SWAP ? (g >> 6) : (g & 15);
        if (!FULL && (r >= oh || 4 * q >= ow)) continue;
        uint32_t* o = reinterpret_cast<uint32_t*>(s_out + r * kPitchOut + 12 * q);
        if (!SWAP) {
            // the 4 source pixels are 12 contiguous, word-aligned bytes of one row (ascending x): three word loads,
            // then every output word is a compile-time byte selection of them (mirror in x and channel swap included)
            const int ly = FY ? th - 1 - r : r;
            const int s0 = FX ? tw - 4 - 4 * q : 4 * q;
            const uint32_t* w = reinterpret_cast<const uint32_t*>(s_in + ly * kPitchIn + s0 * 3);
            const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
            o[0] = pick4<src_byte_of<FX, RB>(0), src_byte_of<FX, RB>(1), src_byte_of<FX, RB>(2), src_byte_of<FX, RB>(3)>(w0, w1, w2);
            o[1] = pick4<src_byte_of<FX, RB>(4), src_byte_of<FX, RB>(5), src_byte_of<FX, RB>(6), src_byte_of<FX, RB>(7)>(w0, w1, w2);
            o[2] = pick4<src_byte_of<FX, RB>(8), src_byte_of<FX, RB>(9), src_byte_of<FX, RB>(10), src_byte_of<FX, RB>(11)>(w0, w1, w2);
        } else {
            // the 4 source pixels sit in 4 consecutive rows at the same x: byte loads at compile-time channel offsets
            // (a warp's 32 columns fall into 24 consecutive words).  Word loads + funnel shifts were measured too:
            // faster for TRANSPOSE, slower for the three rotating methods, so the byte form stays.
            const int lx = FX ? tw - 1 - r : r;
            uint32_t b[4][3];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int p = 4 * q + i;
                const int ly = FY ? th - 1 - p : p;
                const uint8_t* px = s_in + ly * kPitchIn + lx * 3;
#pragma unroll
                for (int c = 0; c < 3; ++c) b[i][c] = px[RB ? 2 - c : c];
            }
            o[0] = b[0][0] | (b[0][1] << 8) | (b[0][2] << 16) | (b[1][0] << 24);
            o[1] = b[1][1] | (b[1][2] << 8) | (b[2][0] << 16) | (b[2][1] << 24);
            o[2] = b[2][2] | (b[3][0] << 8) | (b[3][1] << 16) | (b[3][2] << 24);
        }
    }
    __syncthreads();
    if (VEC) {
        // oh rows x (3 ow / 16) vectors out (ow is a multiple of 16 here); the odd pitch means four word reads
        const int vout = (3 * ow) >> 4;
#pragma unroll
        for (int it = 0; it < 3; ++it) {
            const int i = tid + it * kThreads;
            const int r = i / 12, k = i - 12 * r;
            if (FULL || (r < oh && k < vout)) {
                const uint32_t* w = reinterpret_cast<const uint32_t*>(s_out + r * kPitchOut + 16 * k);
                *reinterpret_cast<uint4*>(dst_tile + (size_t)r * dst_pitch + 16 * k) = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
    } else {
        const int wout = (3 * ow) >> 2;
#pragma unroll
        for (int it = 0; it < 12; ++it) {
            const int i = tid + it * kThreads;
            const int r = i / 48, k = i - 48 * r;
            if (FULL || (r < oh && k < wout))
                *(reinterpret_cast<uint32_t*>(dst_tile + (size_t)r * dst_pitch) + k) = *reinterpret_cast<const uint32_t*>(s_out + r * kPitchOut + 4 * k);
        }
    }
}

template <bool SWAP, bool FX, bool FY, bool RB>
__global__ void __launch_bounds__(kThreads) orient_kernel(OrientArgs a) {
    __shared__ __align__(16) uint8_t s[kTile * kPitchIn];
    __shared__ __align__(16) uint8_t s_out[kTile * kPitchOut];
    const int tid = threadIdx.x;
    long long t = blockIdx.x;
    const int tx = (int)(t % a.tiles_x);
    t /= a.tiles_x;
    const int ty = (int)(t % a.tiles_y);
    const int img = (int)(t / a.tiles_y);
    const int x0 = tx * kTile, y0 = ty * kTile;
    const int tw = min(kTile, a.W - x0), th = min(kTile, a.H - y0);
    const uint8_t* src = a.src + (size_t)img * a.src_stride;
    uint8_t* dst = a.dst + (size_t)img * a.dst_stride;

    if (a.vec_ok && (tw & 15) == 0 && (th & 3) == 0) {      // block-uniform
        const int bx = FX ? a.W - x0 - tw : x0, by = FY ? a.H - y0 - th : y0;
        const int ox = SWAP ? by : bx, oy = SWAP ? bx : by;
        const int ow = SWAP ? th : tw;
        const size_t sp = (size_t)a.W * 3, dp = (size_t)(SWAP ? a.H : a.W) * 3;
        const uint8_t* st = src + (size_t)y0 * sp + (size_t)x0 * 3;
        uint8_t* dt = dst + (size_t)oy * dp + (size_t)ox * 3;
        const bool vec = a.vec_out && ((ox * 3) & 15) == 0 && (ow & 15) == 0;
        if (tw == kTile && th == kTile && vec) {
            tile_fast<SWAP, FX, FY, RB, true, true>(st, sp, dt, dp, s, s_out, tid, tw, th);
            return;
        }
        if (vec) {
            tile_fast<SWAP, FX, FY, RB, false, true>(st, sp, dt, dp, s, s_out, tid, tw, th);
            return;
        }
        if (((ox * 3) & 3) == 0) {
            tile_fast<SWAP, FX, FY, RB, false, false>(st, sp, dt, dp, s, s_out, tid, tw, th);
            return;
        }
    }

    // ---- generic path: load the source tile
    const size_t src_pitch = (size_t)a.W * 3;
    const int row_bytes = tw * 3;
    const bool src_words = ((src_pitch | (size_t)(x0 * 3) | (size_t)row_bytes | (size_t)(uintptr_t)src) & 3) == 0;
    if (src_words) {
        const int wpr = row_bytes >> 2;
        for (int i = tid; i < th * wpr; i += kThreads) {
            const int r = i / wpr, k = i - r * wpr;
            const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(src + (size_t)(y0 + r) * src_pitch + (size_t)x0 * 3) + k);
            *reinterpret_cast<uint32_t*>(s + r * kPitch + 4 * k) = v;
        }
    } else {
        for (int i = tid; i < th * row_bytes; i += kThreads) {
            const int r = i / row_bytes, k = i - r * row_bytes;
            s[r * kPitch + k] = __ldg(src + (size_t)(y0 + r) * src_pitch + (size_t)x0 * 3 + k);
        }
    }
    __syncthreads();

    // ---- the tile in the output: oh rows of ow pixels starting at (ox, oy)
    const int Wout = a.swap ? a.H : a.W;
    const int ow = a.swap ? th : tw, oh = a.swap ? tw : th;
    const int bx = a.flip_x ? a.W - x0 - tw : x0;        // where the tile's source-x range lands
    const int by = a.flip_y ? a.H - y0 - th : y0;        // where the tile's source-y range lands
    const int ox = a.swap ? by : bx, oy = a.swap ? bx : by;
    const size_t dst_pitch = (size_t)Wout * 3;
    const int out_row_bytes = ow * 3;
    // source-tile coordinates of output pixel (p, r) of the tile
    auto src_byte = [&](int r, int p, int c) -> uint32_t {
        const int u = a.swap ? r : p, v = a.swap ? p : r;
        const int lx = a.flip_x ? tw - 1 - u : u;
        const int ly = a.flip_y ? th - 1 - v : v;
        return s[ly * kPitch + lx * 3 + (a.swap_rb ? 2 - c : c)];
    };
    const bool dst_words = ((dst_pitch | (size_t)(ox * 3) | (size_t)out_row_bytes | (size_t)(uintptr_t)dst) & 3) == 0;
    if (dst_words) {
        const int wpr = out_row_bytes >> 2;
        for (int i = tid; i < oh * wpr; i += kThreads) {
            const int r = i / wpr, k = i - r * wpr;
            uint32_t v = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int b = 4 * k + j;
                const int p = (b * 171) >> 9;            // b / 3 for b < 768
                v |= src_byte(r, p, b - 3 * p) << (8 * j);
            }
            *(reinterpret_cast<uint32_t*>(dst + (size_t)(oy + r) * dst_pitch + (size_t)ox * 3) + k) = v;
        }
    } else {
        for (int i = tid; i < oh * out_row_bytes; i += kThreads) {
            const int r = i / out_row_bytes, b = i - r * out_row_bytes;
            const int p = (b * 171) >> 9;
            dst[(size_t)(oy + r) * dst_pitch + (size_t)ox * 3 + b] = (uint8_t)src_byte(r, p, b - 3 * p);
        }
    }
}

}  // namespace

int launch_orient(const uint8_t* d_src, int n, int H, int W, long long src_stride, int swap, int flip_x, int flip_y,
                  int swap_rb, uint8_t* d_dst, long long dst_stride, cudaStream_t stream) {
    FB_REQUIRE(d_src && d_dst, "fb_orient: null pointer");
    FB_REQUIRE(n >= 0 && H >= 1 && W >= 1, "fb_orient: bad shape n=%d H=%d W=%d", n, H, W);
    FB_REQUIRE(d_src != d_dst, "fb_orient: in-place operation is not supported");
    if (n == 0) return 0;
    OrientArgs a;
    a.src = d_src, a.dst = d_dst, a.src_stride = src_stride, a.dst_stride = dst_stride;
    a.H = H, a.W = W;
    a.tiles_x = (W + kTile - 1) / kTile, a.tiles_y = (H + kTile - 1) / kTile;
    a.swap = swap != 0, a.flip_x = flip_x != 0, a.flip_y = flip_y != 0, a.swap_rb = swap_rb != 0;
    a.vec_ok = (((size_t)W * 3) % 16 == 0) && ((uintptr_t)d_src % 16 == 0) && (src_stride % 16 == 0) &&
               (((size_t)(a.swap ? H : W) * 3) % 4 == 0) && ((uintptr_t)d_dst % 4 == 0) && (dst_stride % 4 == 0);
    a.vec_out = (((size_t)(a.swap ? H : W) * 3) % 16 == 0) && ((uintptr_t)d_dst % 16 == 0) && (dst_stride % 16 == 0);
    const long long tiles = (long long)a.tiles_x * a.tiles_y * n;
    FB_REQUIRE(tiles < (1ll << 31), "fb_orient: batch too large for one launch");
    const int method = (a.swap << 3) | (a.flip_x << 2) | (a.flip_y << 1) | a.swap_rb;
#define FB_ORIENT_CASE(M)                                                                                          \
    case M:                                                                                                        \
        orient_kernel<((M) >> 3) & 1, ((M) >> 2) & 1, ((M) >> 1) & 1, (M) & 1><<<(unsigned)tiles, kThreads, 0, stream>>>(a); \
        break;
    switch (method) {
        FB_ORIENT_CASE(0) FB_ORIENT_CASE(1) FB_ORIENT_CASE(2) FB_ORIENT_CASE(3) FB_ORIENT_CASE(4) FB_ORIENT_CASE(5)
        FB_ORIENT_CASE(6) FB_ORIENT_CASE(7) FB_ORIENT_CASE(8) FB_ORIENT_CASE(9) FB_ORIENT_CASE(10) FB_ORIENT_CASE(11)
        FB_ORIENT_CASE(12) FB_ORIENT_CASE(13) FB_ORIENT_CASE(14) FB_ORIENT_CASE(15)
    }
#undef FB_ORIENT_CASE
    FB_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace fb
