// Throughput probe for the integer / packed-16 instructions the technical kernel could use (sm_100a).
// Prints warp-instructions per clock per SM for each op alone and for pairs (pipe sharing).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 2048
#define UNROLL 8

enum Op { IDP2A, VIMNMX3, VHMNMX, VIADD16, HSET2, HADD2, HFMA2, PRMT, LOP3, IMAD, FFMARM, CVTF16, SHF, IADD3, FFMA, LDS, ATOMS_PRIV, ATOMS_RAND, HMNMX2, SEL, MUFU, I2F, FFMA2, FADD2, FHFMA, FADD, VIADDMNMX, LEA, FRND, NOPS };
static const char* names[] = {"IDP.2A", "VIMNMX3.U16x2", "VHMNMX", "VIADD.16x2", "HSET2", "HADD2", "HFMA2", "PRMT", "LOP3", "IMAD", "FFMA.RM", "HADD2.F32(cvt)", "SHF", "IADD3", "FFMA", "LDS", "ATOMS lane-private", "ATOMS random", "HMNMX2", "SEL(ISETP+SEL)", "MUFU.RCP", "I2F.U8", "FFMA2.RM", "FADD2.RM", "FHFMA", "FADD.RM", "VIADDMNMX.U32", "LEA", "FRND.FLOOR"};

template <int OP>
__device__ __forceinline__ void one(uint32_t& a, uint32_t x, uint32_t y, uint32_t smem_base) {
    if (OP == IDP2A) asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %0;" : "+r"(a) : "r"(x), "r"(y));
    if (OP == VIMNMX3) asm volatile("{.reg .b32 t; max.u16x2 t, %0, %1; max.u16x2 %0, t, %2;}" : "+r"(a) : "r"(x), "r"(y));
    if (OP == VHMNMX) asm volatile("{.reg .b32 t; max.f16x2 t, %0, %1; max.f16x2 %0, t, %2;}" : "+r"(a) : "r"(x), "r"(y));
    if (OP == VIADD16) asm volatile("add.u16x2 %0, %0, %1;" : "+r"(a) : "r"(x));
    if (OP == HSET2) asm volatile("set.eq.u32.f16x2 %0, %0, %1;" : "+r"(a) : "r"(x));
    if (OP == HADD2) asm volatile("add.f16x2 %0, %0, %1;" : "+r"(a) : "r"(x));
    if (OP == HFMA2) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(a) : "r"(x), "r"(y));
    if (OP == PRMT) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(a) : "r"(x), "r"(y));
    if (OP == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(x), "r"(y));
    if (OP == IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(x), "r"(y));
    if (OP == FFMARM) asm volatile("fma.rm.f32 %0, %0, %1, %2;" : "+f"(*(float*)&a) : "f"(__uint_as_float(x)), "f"(__uint_as_float(y)));
    if (OP == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(*(float*)&a) : "f"(__uint_as_float(x)), "f"(__uint_as_float(y)));
    if (OP == CVTF16) asm volatile("{.reg .b16 lo, hi; mov.b32 {lo, hi}, %0; cvt.f32.f16 %0, lo;}" : "+r"(a));
    if (OP == SHF) asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(a) : "r"(x), "r"(y));
    if (OP == IADD3) asm volatile("{.reg .b32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(a) : "r"(x), "r"(y));
    if (OP == LDS) asm volatile("ld.shared.u32 %0, [%0];" : "+r"(a));
    if (OP == ATOMS_PRIV) asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(smem_base + ((a & 0xffu) << 7)) : "memory");
    if (OP == ATOMS_RAND) asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a) : "memory");
    if (OP == HMNMX2) asm volatile("max.f16x2 %0, %0, %1;" : "+r"(a) : "r"(x));
    if (OP == SEL) asm volatile("{.reg .pred p; setp.eq.u32 p, %0, %1; selp.u32 %0, %1, %2, p;}" : "+r"(a) : "r"(x), "r"(y));
    if (OP == MUFU) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(*(float*)&a));
    if (OP == FFMA2) { unsigned long long t = ((unsigned long long)a << 32) | x, u = ((unsigned long long)y << 32) | x; asm volatile("fma.rm.f32x2 %0, %0, %1, %1;" : "+l"(t) : "l"(u)); a = (uint32_t)(t >> 32); }
    if (OP == FADD2) { unsigned long long t = ((unsigned long long)a << 32) | x, u = ((unsigned long long)y << 32) | x; asm volatile("add.rm.f32x2 %0, %0, %1;" : "+l"(t) : "l"(u)); a = (uint32_t)(t >> 32); }
    if (OP == FHFMA) { unsigned short h = (unsigned short)x; asm volatile("fma.rn.f32.f16 %0, %1, %1, %0;" : "+f"(*(float*)&a) : "h"(h)); }
    if (OP == FADD) asm volatile("add.rm.f32 %0, %0, %1;" : "+f"(*(float*)&a) : "f"(__uint_as_float(x)));
    if (OP == VIADDMNMX) asm volatile("{.reg .b32 t; add.u32 t, %0, %1; min.u32 %0, t, %0;}" : "+r"(a) : "r"(x));
    if (OP == LEA) asm volatile("{.reg .b32 t; shl.b32 t, %0, 2; add.u32 %0, t, %1;}" : "+r"(a) : "r"(x));
    if (OP == FRND) asm volatile("cvt.rmi.f32.f32 %0, %0;" : "+f"(*(float*)&a));
    if (OP == I2F) asm volatile("{.reg .b32 t; and.b32 t, %0, 255; cvt.rn.f32.u32 %0, t;}" : "+r"(a));
}

template <int OPA, int OPB>
__global__ void __launch_bounds__(1024, 1) probe(uint32_t* out, long long* cycles, uint32_t seed) {
    __shared__ uint32_t sm[8192 + 32];
    for (int i = threadIdx.x; i < 8192 + 32; i += blockDim.x) sm[i] = (uint32_t)__cvta_generic_to_shared(sm) + 4u * ((i * 37 + 5) & 8191);
    __syncthreads();
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm);
    uint32_t a[UNROLL], b[UNROLL];
    uint32_t lane_base = base + 4u * (threadIdx.x & 31);
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) {
        a[j] = seed * (threadIdx.x + 1) * (j + 3);
        b[j] = seed * (threadIdx.x + 7) * (j + 11);
        if (OPA == LDS) a[j] = base + 4u * ((threadIdx.x * 33 + j * 131) & 8191);
        if (OPB == LDS) b[j] = base + 4u * ((threadIdx.x * 33 + j * 131) & 8191);
        if (OPA == ATOMS_RAND) a[j] = base + 4u * ((a[j] >> 7) & 8191);
        if (OPB == ATOMS_RAND) b[j] = base + 4u * ((b[j] >> 7) & 8191);
    }
    uint32_t x = seed | 0x00010001u, y = seed ^ 0x12345u;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int j = 0; j < UNROLL; ++j) {
            one<OPA>(a[j], x, y, lane_base);
            if (OPB != NOPS) one<OPB>(b[j], x, y, lane_base);
        }
    }
    long long t1 = clock64();
    __syncthreads();
    uint32_t acc = 0;
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) acc ^= a[j] ^ b[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + sm[threadIdx.x];
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OPA, int OPB>
void run(uint32_t* out, long long* cyc, int sms) {
    probe<OPA, OPB><<<sms, 1024>>>(out, cyc, 12345u);
    cudaDeviceSynchronize();
    probe<OPA, OPB><<<sms, 1024>>>(out, cyc, 12345u);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return; }
    long long h[256];
    cudaMemcpy(h, cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < sms; ++i) avg += (double)h[i];
    avg /= sms;
    const double winstr = 32.0 * ITERS * UNROLL * (OPB == NOPS ? 1 : 2);   // 32 warps, per SM (PTX-level ops)
    if (OPB == NOPS) printf("%-22s                        %6.3f PTX-ops/clk/SM  (%.0f cycles)\n", names[OPA], winstr / avg, avg);
    else printf("%-22s + %-22s %6.3f PTX-ops/clk/SM  (%.0f cycles)\n", names[OPA], names[OPB], winstr / avg, avg);
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint32_t* out;
    long long* cyc;
    cudaMalloc(&out, sms * 1024 * 4);
    cudaMalloc(&cyc, sms * 8);
    printf("warp-level PTX ops per clock per SM (4 SMSPs; 4.0 = full issue rate). Ops that expand to 2 SASS are noted by name.\n");
    run<FFMA2, NOPS>(out, cyc, sms);
    run<FADD2, NOPS>(out, cyc, sms);
    run<FHFMA, NOPS>(out, cyc, sms);
    run<FADD, NOPS>(out, cyc, sms);
    run<VIADDMNMX, NOPS>(out, cyc, sms);
    run<LEA, NOPS>(out, cyc, sms);
    run<FRND, NOPS>(out, cyc, sms);
    run<FFMA2, IMAD>(out, cyc, sms);
    run<FFMA2, PRMT>(out, cyc, sms);
    run<FFMA2, FFMA>(out, cyc, sms);
    run<FHFMA, IMAD>(out, cyc, sms);
    run<FHFMA, PRMT>(out, cyc, sms);
    run<FHFMA, FFMA>(out, cyc, sms);
    run<FADD, IMAD>(out, cyc, sms);
    run<FADD, FFMA>(out, cyc, sms);
    run<VIADDMNMX, IMAD>(out, cyc, sms);
    run<VIADDMNMX, PRMT>(out, cyc, sms);
    run<LEA, IMAD>(out, cyc, sms);
    run<LEA, PRMT>(out, cyc, sms);
    run<FRND, PRMT>(out, cyc, sms);
    run<FFMA, IMAD>(out, cyc, sms);
    run<FFMA, HFMA2>(out, cyc, sms);
    return 0;
}
