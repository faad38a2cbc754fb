// Baseline JPEG encoding of the photo thumbnails on the device, byte-exact with what the reference gets from Pillow.
//
// Replaces the encoder half of `generate_photo_thumbnail` (utils/image_transforms.py:32-50: `thumb.save(buf, format='JPEG',
// quality=80)`, called for every saved photo at processing/scorer.py:1681-1686).  Pillow drives libjpeg(-turbo) with its
// defaults: YCbCr 4:2:0, the standard Huffman tables (optimize off), the quantisation tables of the quality setting, no restart
// markers.  Restated here (published algorithm; pinned against Pillow's output byte for byte, oracle/jpeg_encode_np.py and
// tests/test_gpu_jpeg_encode.py):
//   colour   rgb_ycc_convert: 16-bit fixed point, Y = (19595 R + 38470 G + 7471 B + 2^15) >> 16, Cb / Cr with the 128 << 16 offset
//            and the 2^15 - 1 rounding term
//   chroma   h2v2_downsample: (a + b + c + d + bias) >> 2 with the bias alternating 1, 2 along a row; right / bottom edges
//            replicated (expand_right_edge / expand_bottom_edge) before and after the downsampling
//   blocks   jpeg_fdct_islow on level-shifted samples, quantisation = round-half-up division by 8 q (what libjpeg-turbo's
//            reciprocal tables compute), dummy blocks beyond a component's block grid (zero AC, DC = previous block's DC)
//   entropy  encode_one_block: DC difference / AC run-length categories, 0xF0 for runs of 16 zeros, EOB; the bit stream is padded
//            with one bits, every 0xFF byte is followed by a stuffed 0x00
// The header (SOI .. SOS) does not depend on the pixels: the caller passes the bytes Pillow writes for this size and quality.
// Launch sequence per batch of same-sized images:
//   jpeg_enc_blocks_kernel   one thread per 8x8 block: samples straight from the RGB pixels, FDCT, quantisation, zig-zag order
//   jpeg_enc_scan_kernel     one CTA per image: bit length of every block (DC predictors resolved through dummy blocks),
//                            exclusive prefix = the bit offset of every block
//   jpeg_enc_emit_kernel     one thread per block: its code words ORed into the zeroed bit stream at its offset
//   jpeg_enc_finish_kernel   one CTA per image: one-bit padding, 0xFF stuffing (count / prefix / expand), header + data + EOI
#include "common.cuh"
#include "kernels.h"

namespace fb {

namespace {

struct EncGeom {
    int H, W;
    int mcux, mcuy;
    int yw_blocks, yh_blocks;      // luma block grid (ceil(W / 8), ceil(H / 8))
    int cw_blocks, ch_blocks;      // chroma block grid of the downsampled planes
    int chh;                       // rows of the downsampled chroma planes, ceil(H / 2)
    int total_blocks;              // 6 per MCU
    long long raw_stride;          // bytes per image of the unstuffed bit stream (multiple of 4)
    long long out_stride;
};

// Encoder tables as packed by facet_b200/utils/jpeg.py `encoder_tables`:
//   uint16 q8[2][64]      8 * quantisation value, natural order, luma / chroma
//   uint16 code[4][256]   Huffman code of every symbol: DC luma, AC luma, DC chroma, AC chroma
//   uint8  size[4][256]   its length in bits (0 = symbol not in the table)
struct EncTables {
    uint16_t q8[2][64];
    uint16_t code[4][256];
    uint8_t size[4][256];
};
static_assert(sizeof(EncTables) == 256 + 2048 + 1024, "layout shared with facet_b200/utils/jpeg.py");

__constant__ uint8_t c_zigzag[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20,
                                     13, 6, 7, 14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45,
                                     38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

constexpr int kFix_0_298631336 = 2446, kFix_0_390180644 = 3196, kFix_0_541196100 = 4433, kFix_0_765366865 = 6270,
              kFix_0_899976223 = 7373, kFix_1_175875602 = 9633, kFix_1_501321110 = 12299, kFix_1_847759065 = 15137,
              kFix_1_961570560 = 16069, kFix_2_053119869 = 16819, kFix_2_562915447 = 20995, kFix_3_072711026 = 25172;
constexpr int kConstBits = 13, kPass1Bits = 2;

__device__ __forceinline__ int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

// one 1-D pass of jpeg_fdct_islow on eight values; PASS = 1 (rows, results scaled up by 2^PASS1_BITS) or 2 (columns)
template <int PASS>
__device__ __forceinline__ void fdct8(int (&d)[8]) {
    const int t0 = d[0] + d[7], t7 = d[0] - d[7], t1 = d[1] + d[6], t6 = d[1] - d[6];
    const int t2 = d[2] + d[5], t5 = d[2] - d[5], t3 = d[3] + d[4], t4 = d[3] - d[4];
    const int t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
    constexpr int sh = PASS == 1 ? kConstBits - kPass1Bits : kConstBits + kPass1Bits;
    d[0] = PASS == 1 ? (t10 + t11) << kPass1Bits : descale(t10 + t11, kPass1Bits);
    d[4] = PASS == 1 ? (t10 - t11) << kPass1Bits : descale(t10 - t11, kPass1Bits);
    int z1 = (t12 + t13) * kFix_0_541196100;
    d[2] = descale(z1 + t13 * kFix_0_765366865, sh);
    d[6] = descale(z1 + t12 * (-kFix_1_847759065), sh);
    z1 = t4 + t7;
    int z2 = t5 + t6, z3 = t4 + t6, z4 = t5 + t7;
    const int z5 = (z3 + z4) * kFix_1_175875602;
    const int a4 = t4 * kFix_0_298631336, a5 = t5 * kFix_2_053119869, a6 = t6 * kFix_3_072711026, a7 = t7 * kFix_1_501321110;
    z1 *= -kFix_0_899976223;
    z2 *= -kFix_2_562915447;
    z3 = z3 * (-kFix_1_961570560) + z5;
    z4 = z4 * (-kFix_0_390180644) + z5;
    d[7] = descale(a4 + z1 + z3, sh);
    d[5] = descale(a5 + z2 + z4, sh);
    d[3] = descale(a6 + z2 + z3, sh);
    d[1] = descale(a7 + z1 + z4, sh);
}

// block b of the scan (MCU by MCU: Y00 Y01 Y10 Y11 Cb Cr) -> component, block row / column, "beyond the component's grid"
__device__ __forceinline__ void block_place(const EncGeom& g, int b, int& comp, int& brow, int& bcol, bool& dummy) {
    const int mcu = b / 6, bi = b - 6 * mcu;
    const int my = mcu / g.mcux, mx = mcu - my * g.mcux;
    if (bi < 4) {
        comp = 0;
        brow = 2 * my + (bi >> 1);
        bcol = 2 * mx + (bi & 1);
        dummy = brow >= g.yh_blocks || bcol >= g.yw_blocks;
    } else {
        comp = bi - 3;
        brow = my;
        bcol = mx;
        dummy = brow >= g.ch_blocks || bcol >= g.cw_blocks;
    }
}

__global__ void __launch_bounds__(128) jpeg_enc_blocks_kernel(const uint8_t* __restrict__ rgb, long long image_stride, EncGeom g,
                                                              const EncTables* __restrict__ tab, int16_t* __restrict__ coef,
                                                              uint8_t* __restrict__ dummy_flag) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x, img = blockIdx.y;
    if (b >= g.total_blocks) return;
    int comp, brow, bcol;
    bool dummy;
    block_place(g, b, comp, brow, bcol, dummy);
    dummy_flag[(size_t)img * g.total_blocks + b] = dummy ? 1 : 0;
    int16_t* out = coef + ((size_t)img * g.total_blocks + b) * 64;
    if (dummy) return;                       // never read: the scan treats it as "difference 0, end of block"
    const uint8_t* px = rgb + (size_t)img * image_stride;
    const int W = g.W, H = g.H;
    int ws[8][8];
    constexpr int kHalf = 1 << 15, kOff = 128 << 16;
    for (int r = 0; r < 8; ++r) {
        int d[8];
        if (comp == 0) {
            const int yy = min(8 * brow + r, H - 1);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const uint8_t* p = px + ((size_t)yy * W + min(8 * bcol + c, W - 1)) * 3;
                d[c] = ((19595 * p[0] + 38470 * p[1] + 7471 * p[2] + kHalf) >> 16) - 128;
            }
        } else {
            const int cy = min(8 * brow + r, g.chh - 1);
            const int r0 = min(2 * cy, H - 1), r1 = min(2 * cy + 1, H - 1);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int cx = 8 * bcol + c;
                const int x0 = min(2 * cx, W - 1), x1 = min(2 * cx + 1, W - 1);
                int sum = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint8_t* p = px + ((size_t)((k & 2) ? r1 : r0) * W + ((k & 1) ? x1 : x0)) * 3;
                    const int R = p[0], G = p[1], B = p[2];
                    sum += comp == 1 ? (-11059 * R - 21709 * G + 32768 * B + kOff + kHalf - 1) >> 16
                                     : (32768 * R - 27439 * G - 5329 * B + kOff + kHalf - 1) >> 16;
                }
                d[c] = ((sum + ((cx & 1) ? 2 : 1)) >> 2) - 128;
            }
        }
        fdct8<1>(d);
#pragma unroll
        for (int c = 0; c < 8; ++c) ws[r][c] = d[c];
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        int d[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) d[r] = ws[r][c];
        fdct8<2>(d);
#pragma unroll
        for (int r = 0; r < 8; ++r) ws[r][c] = d[r];
    }
    const uint16_t* q8 = tab->q8[comp ? 1 : 0];
    for (int k = 0; k < 64; ++k) {
        const int nat = c_zigzag[k];
        const int v = ws[nat >> 3][nat & 7];
        const int qv = q8[nat];
        const int a = (abs(v) + (qv >> 1)) / qv;
        out[k] = (int16_t)(v < 0 ? -a : a);
    }
}

// DC of the last real block of the same component before block b of the scan (0 at the start): dummy blocks repeat their
// predecessor's DC, so they are skipped.
__device__ __forceinline__ int dc_predictor(const EncGeom& g, const int16_t* coef, const uint8_t* dummy, int b) {
    const int bi = b % 6;
    for (;;) {
        int p;
        if (bi < 4) p = (b % 6) > 0 ? b - 1 : b - 3;          // Y: previous block of the MCU, or Y11 of the previous MCU
        else p = b - 6;
        if (p < 0) return 0;
        if (!dummy[p]) return coef[(size_t)p * 64];
        b = p;
    }
}

__device__ __forceinline__ int nbits_of(int v) { return 32 - __clz(v); }      // v >= 0

// Walks one block: calls put(code, size) for every code word / value field in stream order.
template <class Put>
__device__ __forceinline__ void encode_block(const EncGeom& g, const EncTables& T, const int16_t* coef, const uint8_t* dummy, int b, Put put) {
    const int bi = b % 6, t_dc = bi < 4 ? 0 : 2, t_ac = t_dc + 1;
    if (dummy[b]) {
        put(T.code[t_dc][0], T.size[t_dc][0]);
        put(T.code[t_ac][0], T.size[t_ac][0]);
        return;
    }
    const int16_t* blk = coef + (size_t)b * 64;
    int diff = blk[0] - dc_predictor(g, coef, dummy, b);
    int t = diff, t2 = diff;
    if (t < 0) {
        t = -t;
        --t2;
    }
    int n = nbits_of(t);
    put(T.code[t_dc][n], T.size[t_dc][n]);
    if (n) put((uint32_t)t2 & ((1u << n) - 1), n);
    int run = 0;
    for (int k = 1; k < 64; ++k) {
        const int v = blk[k];
        if (v == 0) {
            ++run;
            continue;
        }
        while (run > 15) {
            put(T.code[t_ac][0xF0], T.size[t_ac][0xF0]);
            run -= 16;
        }
        t = v, t2 = v;
        if (t < 0) {
            t = -t;
            --t2;
        }
        n = nbits_of(t);
        put(T.code[t_ac][(run << 4) + n], T.size[t_ac][(run << 4) + n]);
        put((uint32_t)t2 & ((1u << n) - 1), n);
        run = 0;
    }
    if (run > 0) put(T.code[t_ac][0], T.size[t_ac][0]);
}

// bit length of every block and its exclusive prefix (the block's bit offset); total[img] = bits of the whole scan
__global__ void __launch_bounds__(1024) jpeg_enc_scan_kernel(EncGeom g, const EncTables* __restrict__ tab, const int16_t* __restrict__ coef,
                                                             const uint8_t* __restrict__ dummy, unsigned int* __restrict__ offset,
                                                             unsigned int* __restrict__ total) {
    __shared__ EncTables T;
    __shared__ unsigned int s_part[1024];
    __shared__ unsigned int s_carry;
    const int img = blockIdx.x, tid = threadIdx.x;
    for (int i = tid; i < (int)(sizeof(EncTables) / 4); i += 1024) reinterpret_cast<uint32_t*>(&T)[i] = reinterpret_cast<const uint32_t*>(tab)[i];
    if (tid == 0) s_carry = 0u;
    __syncthreads();
    const int16_t* cimg = coef + (size_t)img * g.total_blocks * 64;
    const uint8_t* dimg = dummy + (size_t)img * g.total_blocks;
    unsigned int* oimg = offset + (size_t)img * g.total_blocks;
    for (int base = 0; base < g.total_blocks; base += 1024) {
        const int b = base + tid;
        unsigned int bits = 0;
        if (b < g.total_blocks) encode_block(g, T, cimg, dimg, b, [&](uint32_t, int size) { bits += (unsigned int)size; });
        s_part[tid] = bits;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            const unsigned int v = tid >= o ? s_part[tid - o] : 0u;
            __syncthreads();
            s_part[tid] += v;
            __syncthreads();
        }
        if (b < g.total_blocks) oimg[b] = s_carry + s_part[tid] - bits;
        __syncthreads();
        if (tid == 1023) s_carry += s_part[1023];
        __syncthreads();
    }
    if (tid == 0) total[img] = s_carry;
}

// every block ORs its bits into the zeroed stream (big-endian bit order inside 32-bit words, stored byte-swapped)
__global__ void __launch_bounds__(128) jpeg_enc_emit_kernel(EncGeom g, const EncTables* __restrict__ tab, const int16_t* __restrict__ coef,
                                                            const uint8_t* __restrict__ dummy, const unsigned int* __restrict__ offset,
                                                            uint8_t* __restrict__ raw) {
    __shared__ EncTables T;
    for (int i = threadIdx.x; i < (int)(sizeof(EncTables) / 4); i += blockDim.x)
        reinterpret_cast<uint32_t*>(&T)[i] = reinterpret_cast<const uint32_t*>(tab)[i];
    __syncthreads();
    const int b = blockIdx.x * blockDim.x + threadIdx.x, img = blockIdx.y;
    if (b >= g.total_blocks) return;
    const unsigned int start = offset[(size_t)img * g.total_blocks + b];
    unsigned int* words = reinterpret_cast<unsigned int*>(raw + (size_t)img * g.raw_stride);
    unsigned int wi = start >> 5;
    int used = (int)(start & 31);              // bits already taken in word wi (from the most significant end)
    unsigned long long acc = 0;                // pending bits, right-aligned; `used` + `pend` bits of word wi are decided
    int pend = 0;
    auto flush = [&](bool all) {
        // move whole words out; with `all` also the partial last one
        while (used + pend >= 32) {
            const int take = 32 - used;                                  // bits that complete word wi
            const unsigned int v = (unsigned int)(acc >> (pend - take)) & (take == 32 ? 0xffffffffu : ((1u << take) - 1u));
            atomicOr(words + wi, __byte_perm(v, 0u, 0x0123));
            pend -= take;
            acc &= pend ? ((1ull << pend) - 1ull) : 0ull;
            ++wi;
            used = 0;
        }
        if (all && pend) {
            const unsigned int v = (unsigned int)acc << (32 - used - pend);
            atomicOr(words + wi, __byte_perm(v, 0u, 0x0123));
        }
    };
    encode_block(g, T, coef + (size_t)img * g.total_blocks * 64, dummy + (size_t)img * g.total_blocks, b, [&](uint32_t code, int size) {
        acc = (acc << size) | (unsigned long long)code;
        pend += size;
        if (pend > 32) flush(false);
    });
    flush(true);
}

// one-bit padding of the last byte, 0xFF stuffing, header + entropy-coded data + EOI into the output slot; length[img] = total bytes
__global__ void __launch_bounds__(1024) jpeg_enc_finish_kernel(EncGeom g, const unsigned int* __restrict__ total, uint8_t* __restrict__ raw,
                                                               const uint8_t* __restrict__ header, int header_len, uint8_t* __restrict__ out,
                                                               unsigned int* __restrict__ length) {
    __shared__ unsigned int s_part[1024];
    const int img = blockIdx.x, tid = threadIdx.x;
    uint8_t* r = raw + (size_t)img * g.raw_stride;
    uint8_t* o = out + (size_t)img * g.out_stride;
    const unsigned int bits = total[img];
    const unsigned int nbytes = (bits + 7) >> 3;
    if (tid == 0 && (bits & 7)) r[nbytes - 1] |= (uint8_t)(0xFFu >> (bits & 7));       // flush_bits: fill with ones
    for (int i = tid; i < header_len; i += 1024) o[i] = header[i];
    __syncthreads();
    // every thread owns a contiguous piece of the stream: count its 0xFF bytes, scan, copy with the stuffed zeros
    const unsigned int per = (nbytes + 1023) / 1024;
    const unsigned int lo = min(tid * per, nbytes), hi = min(lo + per, nbytes);
    unsigned int ff = 0;
    for (unsigned int i = lo; i < hi; ++i) ff += r[i] == 0xFF;
    s_part[tid] = ff;
    __syncthreads();
    for (int s = 1; s < 1024; s <<= 1) {
        const unsigned int v = tid >= s ? s_part[tid - s] : 0u;
        __syncthreads();
        s_part[tid] += v;
        __syncthreads();
    }
    unsigned int w = header_len + lo + (s_part[tid] - ff);
    for (unsigned int i = lo; i < hi; ++i) {
        const uint8_t v = r[i];
        o[w++] = v;
        if (v == 0xFF) o[w++] = 0;
    }
    if (tid == 1023) {
        const unsigned int end = header_len + nbytes + s_part[1023];
        o[end] = 0xFF;
        o[end + 1] = 0xD9;
        length[img] = end + 2;
    }
}

inline size_t al256(size_t v) { return (v + 255) & ~(size_t)255; }

EncGeom make_geom(int H, int W) {
    EncGeom g;
    g.H = H;
    g.W = W;
    g.mcux = (W + 15) / 16;
    g.mcuy = (H + 15) / 16;
    g.yw_blocks = (W + 7) / 8;
    g.yh_blocks = (H + 7) / 8;
    g.chh = (H + 1) / 2;
    g.cw_blocks = ((W + 1) / 2 + 7) / 8;
    g.ch_blocks = (g.chh + 7) / 8;
    g.total_blocks = g.mcux * g.mcuy * 6;
    // worst case per block: 64 coefficients x (16-bit code + 11 value bits) < 1800 bits; real streams stay far below
    g.raw_stride = (long long)al256((size_t)g.total_blocks * 232);
    g.out_stride = 0;
    return g;
}

}  // namespace

size_t jpeg_encode_workspace_bytes(int n, int H, int W) {
    const EncGeom g = make_geom(H, W);
    return al256((size_t)n * g.total_blocks * 128) + al256((size_t)n * g.total_blocks) + al256((size_t)n * g.total_blocks * 4) +
           al256((size_t)n * 4) + al256((size_t)n * g.raw_stride);
}

size_t jpeg_encode_out_stride(int H, int W, int header_len) {
    const EncGeom g = make_geom(H, W);
    return al256((size_t)header_len + 2 * (size_t)g.raw_stride + 2);      // every byte could be 0xFF
}

int launch_jpeg_encode(const uint8_t* d_rgb, int n, int H, int W, long long image_stride, const void* d_tables, const uint8_t* d_header,
                       int header_len, void* d_ws, size_t ws_bytes, uint8_t* d_out, long long out_stride, unsigned int* d_length,
                       cudaStream_t stream) {
    FB_REQUIRE(d_rgb && d_tables && d_header && d_ws && d_out && d_length, "fb_jpeg_encode: null pointer");
    FB_REQUIRE(n >= 1 && n <= 65535 && H >= 1 && W >= 1 && H <= 16384 && W <= 16384 && header_len >= 4, "fb_jpeg_encode: bad arguments");
    FB_REQUIRE(image_stride >= (long long)H * W * 3, "fb_jpeg_encode: image_stride smaller than one image");
    FB_REQUIRE(ws_bytes >= jpeg_encode_workspace_bytes(n, H, W) && (reinterpret_cast<uintptr_t>(d_ws) & 255) == 0,
               "fb_jpeg_encode: workspace too small or not 256-byte aligned");
    FB_REQUIRE(out_stride >= (long long)jpeg_encode_out_stride(H, W, header_len), "fb_jpeg_encode: output slots too small");
    EncGeom g = make_geom(H, W);
    g.out_stride = out_stride;
    uint8_t* w = static_cast<uint8_t*>(d_ws);
    int16_t* coef = reinterpret_cast<int16_t*>(w);
    w += al256((size_t)n * g.total_blocks * 128);
    uint8_t* dummy = w;
    w += al256((size_t)n * g.total_blocks);
    unsigned int* offset = reinterpret_cast<unsigned int*>(w);
    w += al256((size_t)n * g.total_blocks * 4);
    unsigned int* total = reinterpret_cast<unsigned int*>(w);
    w += al256((size_t)n * 4);
    uint8_t* raw = w;
    const EncTables* tab = reinterpret_cast<const EncTables*>(d_tables);
    FB_CUDA_OK(cudaMemsetAsync(raw, 0, (size_t)n * g.raw_stride, stream));
    const dim3 bgrid((g.total_blocks + 127) / 128, n);
    jpeg_enc_blocks_kernel<<<bgrid, 128, 0, stream>>>(d_rgb, image_stride, g, tab, coef, dummy);
    jpeg_enc_scan_kernel<<<n, 1024, 0, stream>>>(g, tab, coef, dummy, offset, total);
    jpeg_enc_emit_kernel<<<bgrid, 128, 0, stream>>>(g, tab, coef, dummy, offset, raw);
    jpeg_enc_finish_kernel<<<n, 1024, 0, stream>>>(g, total, raw, d_header, header_len, d_out, d_length);
    FB_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace fb
