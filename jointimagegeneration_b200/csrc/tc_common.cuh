// tc_common.cuh -- PTX wrappers shared by the tcgen05 kernels (mbarrier, TMA, tcgen05.mma / ld / commit)
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace gg {

constexpr int BM = 128;   // output positions per tile (TMEM lanes)
constexpr int BK = 64;    // channels per K step (128 bytes, SWIZZLE_128B)

// ------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t ok, spins = 0;
    uint64_t t0 = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
        // watchdog (failed polls only): a pipeline-protocol bug must surface as a launch error, not hang the GPU
        if (!ok && (++spins & 0xFFFu) == 0) {
            uint64_t now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000ull) __trap();
        }
    } while (!ok);
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3,
                                            int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major, SWIZZLE_128B operand tile: rows of 128 bytes, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);        // start address  [0, 14)
    d |= (uint64_t)1 << 16;                         // leading byte offset (ignored for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;               // stride byte offset [32, 46)
    d |= (uint64_t)1 << 46;                         // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                         // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
        "elect.sync rx|px, %1;\n\t"
        "@px mov.s32 %0, 1;\n\t}"
        : "+r"(pred)
        : "r"(0xffffffffu));
    return pred != 0;
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }


__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}

// Epilogue of one accumulator row (this thread's TMEM lane): 32 columns per step -- two tcgen05.ld in flight,
// the per-column additive vector (bias + timestep embedding) read from shared memory (broadcast), the residual
// row software-prefetched one step ahead.  Every lane of the warp must call this (tcgen05.ld is warp-collective);
// lanes whose output position is out of range pass valid = false.
__device__ __forceinline__ void epilogue_row(uint32_t t_addr, int BN, int ncols, const float* __restrict__ bvec,
                                             const __nv_bfloat16* __restrict__ res_row, void* y_row, int y_is_f32, bool valid,
                                             const float* __restrict__ emb_row = nullptr) {
    uint4 rn[4];
    const bool res32 = res_row != nullptr && aligned32(res_row);      // c0 is a multiple of 32 elements: every 16-channel piece is 32-byte aligned
    auto load_res = [&](int c0) {
#pragma unroll
        for (int g = 0; g < 4; g += 2) {
            if (res32 && valid && c0 + 8 * g + 8 < ncols) {
                ldg_nc_u8(res_row + c0 + 8 * g, rn[g], rn[g + 1]);
            } else {
                rn[g] = (res_row != nullptr && valid && c0 + 8 * g < ncols) ? ldg_nc_u4(res_row + c0 + 8 * g) : make_uint4(0, 0, 0, 0);
                rn[g + 1] = (res_row != nullptr && valid && c0 + 8 * g + 8 < ncols) ? ldg_nc_u4(res_row + c0 + 8 * g + 8) : make_uint4(0, 0, 0, 0);
            }
        }
    };
    const bool y32 = !y_is_f32 && aligned32(y_row), yf32 = y_is_f32 && aligned32(y_row);
    load_res(0);
    for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld16_nowait(t_addr + c0, r);
        if (c0 + 16 < BN) tmem_ld16_nowait(t_addr + c0 + 16, r + 16);
        uint4 rc[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) rc[g] = rn[g];
        if (c0 + 32 < BN) load_res(c0 + 32);
        tmem_ld_wait();
        if (!valid) continue;
        uint4 pk_even = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const int c = c0 + 8 * g;
            if (c >= ncols) break;
            float v[8];
            const float4 b0 = *reinterpret_cast<const float4*>(bvec + c), b1 = *reinterpret_cast<const float4*>(bvec + c + 4);
            v[0] = __uint_as_float(r[8 * g + 0]) + b0.x; v[1] = __uint_as_float(r[8 * g + 1]) + b0.y;
            v[2] = __uint_as_float(r[8 * g + 2]) + b0.z; v[3] = __uint_as_float(r[8 * g + 3]) + b0.w;
            v[4] = __uint_as_float(r[8 * g + 4]) + b1.x; v[5] = __uint_as_float(r[8 * g + 5]) + b1.y;
            v[6] = __uint_as_float(r[8 * g + 6]) + b1.z; v[7] = __uint_as_float(r[8 * g + 7]) + b1.w;
            if (emb_row != nullptr) {       // per-row additive vector (tiles that span samples): global, rare
                const float4 e0 = __ldg(reinterpret_cast<const float4*>(emb_row + c)), e1 = __ldg(reinterpret_cast<const float4*>(emb_row + c + 4));
                v[0] += e0.x; v[1] += e0.y; v[2] += e0.z; v[3] += e0.w; v[4] += e1.x; v[5] += e1.y; v[6] += e1.z; v[7] += e1.w;
            }
            if (res_row != nullptr) {
                const uint4 rr = rc[g];
                v[0] += bf16_lo(rr.x); v[1] += bf16_hi(rr.x); v[2] += bf16_lo(rr.y); v[3] += bf16_hi(rr.y);
                v[4] += bf16_lo(rr.z); v[5] += bf16_hi(rr.z); v[6] += bf16_lo(rr.w); v[7] += bf16_hi(rr.w);
            }
            if (y_is_f32) {
                float* yp = reinterpret_cast<float*>(y_row) + c;
                if (yf32) {         // eight floats = one 32-byte sector (split-K partial tiles, fp32 head logits)
                    stg_u8(yp, make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3])),
                           make_uint4(__float_as_uint(v[4]), __float_as_uint(v[5]), __float_as_uint(v[6]), __float_as_uint(v[7])));
                } else {
                    *reinterpret_cast<float4*>(yp) = make_float4(v[0], v[1], v[2], v[3]);
                    *reinterpret_cast<float4*>(yp + 4) = make_float4(v[4], v[5], v[6], v[7]);
                }
            } else {
                __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(y_row) + c;
                const uint4 pk = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
                // pairs of 8-channel pieces leave as one 256-bit store when the row allows it (g even: keep, g odd: store both)
                if (y32 && (g & 1) == 0 && c + 8 < ncols) { pk_even = pk; }
                else if (y32 && (g & 1) == 1) { stg_u8(yp - 8, pk_even, pk); }
                else { *reinterpret_cast<uint4*>(yp) = pk; }
            }
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn();
// 5-D activation map over a (possibly strided) sub-grid of a CL tensor; box = (64 ch, bw, bh, bd, bn)
bool encode_act_map(CUtensorMap* m, const void* base, int C, const int64_t dim[4], const int64_t stride_el[4], const int box[4]);
bool encode_w_map(CUtensorMap* m, const void* base, int64_t Ktot, int rows, int BN);
int conv_halo_fwd(const gg_conv_args* a, cudaStream_t stream);   // conv_halo.cu
int conv_halo_grid(const gg_conv_args* a, bool* pair_out);       // CTAs conv_halo_fwd launches for `a`
int conv_roll_fwd(const gg_conv_args* a, cudaStream_t stream);   // conv_roll.cu (algo 4)
int conv_roll_grid(const gg_conv_args* a);

}  // namespace gg
