// common.cuh -- shared host/device helpers for libguidegen_sm100 (sm_100a only)
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include "../../include/guidegen_sm100.h"

namespace gg {

extern std::atomic<uint64_t> g_launches;

inline int launch_result() {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaPeekAtLastError();
    return e == cudaSuccess ? GG_OK : (int)e;
}

inline cudaStream_t as_stream(gg_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

inline int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
    }
    return n;
}

#define GG_REQUIRE(cond, code) do { if (!(cond)) return (code); } while (0)

// Programmatic dependent launch (GG_PDL=1; OFF by default).  With the stream-serialization attribute a kernel is scheduled while
// its predecessor still runs: its CTAs become resident as SM resources free up, run their prologue (barrier init, TMEM allocation,
// tensor-map prefetch) and block in pdl_wait() until the predecessor has COMPLETED and its memory is visible.  Rule for every
// kernel launched through launch_k: nothing before pdl_wait() reads or writes global memory another kernel touches; ordering
// stays transitive (kernel N cannot complete before its own wait has seen N-1 complete).  Captured into CUDA graphs as
// programmatic dependency edges.  MEASURED on B200 (profiles/r2_lanes_and_pdl.md): inside a CUDA graph the kernel-to-kernel
// gap is already ~1 us, and nothing was gained -- config 3: 4.81 ms (off) / 4.80 (completion trigger) / 4.99 (early trigger:
// dependents placed greedily on the SMs that free up first); config 4: 11.4 / 11.3 / 11.4; two lanes + early trigger 12.2.
// All GPU tests pass with it on.  Kept as a tested knob.
inline bool pdl_enabled() {
    static const int on = [] { const char* e = getenv("GG_PDL"); return e ? atoi(e) : 0; }();
    return on != 0;
}
template <typename... P, typename... A>
inline cudaError_t launch_k(void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, A&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<P>(args)...);
}

// ---------------------------------------------------------------------------- device helpers
#ifndef GG_PDL_TRIGGER
#define GG_PDL_TRIGGER 1        // 0: dependents are released only by this grid's completion (tuning builds)
#endif
__device__ __forceinline__ void pdl_launch_dependents() {
#if GG_PDL_TRIGGER
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }

__device__ __forceinline__ uint4 ldg_nc_u4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float4 ldg_nc_f4(const void* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_na_u4(void* p, uint4 v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// 256-bit global accesses (sm_100: STG / LDG .256): one full 32-byte sector per lane instead of two half-filled ones.  The
// epilogues of the conv kernels write (and read residuals) row by row -- a lane owns an output position, lanes of a warp are
// 256+ bytes apart -- so every 128-bit access touches 32 separate sectors; ncu's source page of the folded-upsample conv had the
// epilogue warps waiting on their own stores.  p must be 32-byte aligned.
__device__ __forceinline__ void stg_u8(void* p, uint4 a, uint4 b) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 :: "l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}
__device__ __forceinline__ void ldg_nc_u8(const void* p, uint4& a, uint4& b) {
    asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
}
__device__ __forceinline__ bool aligned32(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31u) == 0; }
__device__ __forceinline__ void stg_na_f4(void* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Philox4x32-10 (counter-based; key = seed, counter = global element index -> results do not
// depend on how voxels are distributed over blocks, ranks or slabs)
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0; key.y += W1;
    }
    return ctr;
}
// Exp(1) variate from 32 random bits: -log(u), u in (0, 1]
__device__ __forceinline__ float exp1_from_bits(uint32_t b) {
    float u = (float)(b >> 8) * (1.0f / 16777216.0f) + (0.5f / 16777216.0f);
    return -__logf(u);
}

}  // namespace gg
