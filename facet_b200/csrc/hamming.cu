// Duplicate / burst grouping kernels on 64-bit perceptual hashes.
//
//   hamming_pairs_kernel  replaces the all-pairs loop of utils/duplicate.py:94-119
//                         (XOR, 8 x byte-popcount LUT, np.where(d <= max_distance))
//   burst_links_kernel    evaluates the pairwise rule of processing/scorer.py:1943-1968 for
//                         every photo against the photos before it inside the time window
// Integer work, bit-exact by construction.  Union-Find and the sequential burst chain stay on
// the host (facet_b200/utils/duplicate.py, facet_b200/processing/bursts.py): they are O(pairs).
#include "common.cuh"
#include "kernels.h"

namespace fb {

namespace {

constexpr int kThreads = 256;

// Upper-triangle tile `lin` (row-major over the tiles (rt, ct), ct >= rt, of a T x T tile grid) -> (rt, ct).
__device__ __forceinline__ void tri_tile(long long lin, long long T, long long& rt, long long& ct) {
    // tiles before row r: r * T - r (r - 1) / 2
    const double b = 2.0 * (double)T + 1.0;
    long long r = (long long)((b - sqrt(b * b - 8.0 * (double)lin)) * 0.5);
    r = r < 0 ? 0 : (r >= T ? T - 1 : r);
    while (r > 0 && r * T - r * (r - 1) / 2 > lin) --r;
    while ((r + 1) * T - (r + 1) * r / 2 <= lin) ++r;
    rt = r;
    ct = r + (lin - (r * T - r * (r - 1) / 2));
}

// Square tiles of TILE = 256 * RPT hashes; thread = RPT rows, the column tile staged in shared memory.  The
// upper-triangle tiles are dealt to the parts round-robin by their linear index (tile `lin` belongs to part
// lin % nparts), so a triangular problem balances across GPUs to within one tile, and CTAs stride over the
// part's tiles (no grid-dimension limit).  RPT = 1 for small sets (enough tiles to fill the SMs), 8 otherwise.
template <int RPT>
__global__ void __launch_bounds__(kThreads) hamming_pairs_kernel(
    const unsigned long long* __restrict__ h, long long n, int thr, int part, int nparts, long long T,
    long long my_tiles, int* __restrict__ pairs, long long cap, unsigned long long* __restrict__ count) {
    constexpr int TILE = kThreads * RPT;
    __shared__ unsigned long long s_cols[TILE];
    const int tid = threadIdx.x;
    for (long long k = blockIdx.x; k < my_tiles; k += gridDim.x) {
        long long rt, ct;
        tri_tile(k * nparts + part, T, rt, ct);
        const long long row0 = rt * TILE, col0 = ct * TILE;
        __syncthreads();                                   // the previous tile's columns are no longer read
        for (int c = tid; c < TILE; c += kThreads) {
            long long j = col0 + c;
            s_cols[c] = (j < n) ? h[j] : 0ull;
        }
        unsigned long long hr[RPT];
        bool rv[RPT];
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            long long i = row0 + tid + (long long)r * kThreads;
            rv[r] = i < n;
            hr[r] = rv[r] ? h[i] : 0ull;
        }
        __syncthreads();
        const int ncols = (int)min((long long)TILE, n - col0);
        const bool diagonal = (ct == rt);
#pragma unroll 4
        for (int c = 0; c < ncols; ++c) {
            const unsigned long long hc = s_cols[c];
            int d[RPT];
            int dmin = 64;
#pragma unroll
            for (int r = 0; r < RPT; ++r) {
                d[r] = __popcll(hr[r] ^ hc);
                dmin = min(dmin, d[r]);
            }
            if (dmin <= thr) {
                const long long j = col0 + c;
#pragma unroll
                for (int r = 0; r < RPT; ++r) {
                    const long long i = row0 + tid + (long long)r * kThreads;
                    if (d[r] <= thr && rv[r] && (!diagonal || j > i)) {
                        unsigned long long pos = atomicAdd(count, 1ull);
                        if ((long long)pos < cap) {
                            pairs[2 * pos] = (int)i;
                            pairs[2 * pos + 1] = (int)j;
                        }
                    }
                }
            }
        }
    }
}

// flags: bit0 = date parsed, bit1 = hash present (scorer.py:1927-1929 returns 999 otherwise)
__global__ void __launch_bounds__(256) burst_links_kernel(
    const unsigned long long* __restrict__ h, const long long* __restrict__ t,
    const unsigned char* __restrict__ flags, const int* __restrict__ lo, long long n, int thr,
    long long window_s, double rapid_s, int* __restrict__ last_slow, int* __restrict__ rapid_pairs,
    long long rapid_cap, unsigned long long* __restrict__ rapid_count) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int best = -1;
    const unsigned char fi = flags[i];
    if (fi & 1) {
        const unsigned long long hi = h[i];
        const long long ti = t[i];
        for (long long b = i - 1; b >= (long long)lo[i]; --b) {
            const unsigned char fb_ = flags[b];
            if (!(fb_ & 1)) continue;
            long long dt = ti - t[b];
            dt = dt < 0 ? -dt : dt;
            const int d = ((fi & 2) && (fb_ & 2)) ? __popcll(hi ^ h[b]) : 999;
            if ((double)dt <= rapid_s && d <= 2 * thr) {
                unsigned long long pos = atomicAdd(rapid_count, 1ull);
                if ((long long)pos < rapid_cap) {
                    rapid_pairs[2 * pos] = (int)i;
                    rapid_pairs[2 * pos + 1] = (int)b;
                }
            }
            if (dt <= window_s && d <= thr && best < 0) best = (int)b;   // first hit walking down = largest b
        }
    }
    last_slow[i] = best;
}

}  // namespace

int launch_hamming_pairs(const unsigned long long* d_hashes, long long n, int max_distance, int part,
                         int nparts, int* d_pairs, long long cap, unsigned long long* d_count,
                         cudaStream_t stream) {
    FB_REQUIRE(d_hashes && d_count && (d_pairs || cap == 0), "fb_hamming_pairs: null pointer");
    FB_REQUIRE(n >= 0 && n < (1ll << 31), "fb_hamming_pairs: n out of range (int32 pair indices)");
    FB_REQUIRE(nparts >= 1 && part >= 0 && part < nparts, "fb_hamming_pairs: bad part %d of %d", part, nparts);
    FB_REQUIRE(max_distance >= 0 && max_distance <= 64, "fb_hamming_pairs: max_distance out of range");
    FB_CUDA_OK(cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), stream));
    if (n < 2) return 0;
    // small sets: 256-hash tiles so that there are enough tiles for every SM of every part
    const bool small = n <= (1ll << 16);
    const long long tile = small ? kThreads : kThreads * 8;
    const long long T = (n + tile - 1) / tile;
    const long long U = T * (T + 1) / 2;                         // upper-triangle tiles
    const long long my_tiles = (U - part + nparts - 1) / nparts;
    if (my_tiles <= 0) return 0;
    const long long max_grid = (long long)sm_count() * 64;
    const unsigned grid = (unsigned)(my_tiles < max_grid ? my_tiles : max_grid);
    if (small) hamming_pairs_kernel<1><<<grid, kThreads, 0, stream>>>(d_hashes, n, max_distance, part, nparts, T, my_tiles, d_pairs, cap, d_count);
    else hamming_pairs_kernel<8><<<grid, kThreads, 0, stream>>>(d_hashes, n, max_distance, part, nparts, T, my_tiles, d_pairs, cap, d_count);
    FB_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_burst_links(const unsigned long long* d_hashes, const long long* d_time_s, const unsigned char* d_flags,
                       const int* d_lo, long long n, int thr, long long window_s, double rapid_s,
                       int* d_last_slow, int* d_rapid_pairs, long long rapid_cap,
                       unsigned long long* d_rapid_count, cudaStream_t stream) {
    FB_REQUIRE(d_hashes && d_time_s && d_flags && d_lo && d_last_slow && d_rapid_count, "fb_burst_links: null pointer");
    FB_REQUIRE(n >= 0 && n < (1ll << 31), "fb_burst_links: n out of range");
    FB_CUDA_OK(cudaMemsetAsync(d_rapid_count, 0, sizeof(unsigned long long), stream));
    if (n == 0) return 0;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    burst_links_kernel<<<blocks, 256, 0, stream>>>(d_hashes, d_time_s, d_flags, d_lo, n, thr, window_s, rapid_s,
                                                  d_last_slow, d_rapid_pairs, rapid_cap, d_rapid_count);
    FB_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace fb
