// Shared helpers for the facet_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

namespace fb {

// A flag per device ordinal for "kernel attributes have been set" (cudaFuncSetAttribute is per device; one process may
// drive several GPUs).  get() / set() are lock-free; setting attributes twice from two racing threads is harmless.
struct PerDeviceFlag {
    std::atomic<unsigned long long> mask{0ull};
    static int device() {
        int d = 0;
        if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d > 63) d = 0;
        return d;
    }
    bool get() const { return (mask.load(std::memory_order_acquire) >> device()) & 1ull; }
    void set() { mask.fetch_or(1ull << device(), std::memory_order_release); }
};

// Thread-local last-error text behind fb_last_error().
void set_error(const char* fmt, ...);
const char* get_error();
int sm_count();

#define FB_CUDA_OK(expr)                                                              \
    do {                                                                              \
        cudaError_t _e = (expr);                                                      \
        if (_e != cudaSuccess) {                                                      \
            fb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),     \
                          __FILE__, __LINE__);                                        \
            return (int)_e;                                                           \
        }                                                                             \
    } while (0)

#define FB_REQUIRE(cond, ...)                                                         \
    do {                                                                              \
        if (!(cond)) {                                                                \
            fb::set_error(__VA_ARGS__);                                               \
            return -1;                                                                \
        }                                                                             \
    } while (0)

__device__ __forceinline__ uint32_t lane_id() {
    uint32_t l;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(l));
    return l;
}

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

__device__ __forceinline__ uint2 ldg_nc_v2(const void* p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}

}  // namespace fb
