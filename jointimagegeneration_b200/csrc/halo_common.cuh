// halo_common.cuh -- pieces shared by the halo-brick kernels (conv_halo.cu, conv_roll.cu): brick geometry,
// CTA-pair (cta_group::2) PTX wrappers, shifted-descriptor helper and the statistics epilogue
#pragma once
#include "tc_common.cuh"

namespace gg {

constexpr int H_BH = 16, H_BW = 8;            // output brick (one depth plane): 16 x 8 = 128 positions
constexpr int H_MAX_SEGS = 4;
constexpr int H_THREADS = 224;
constexpr int H_SMEM_BUDGET = 227 * 1024;

// Epilogue of one row for the BN = 64 statistics variant: as epilogue_row, plus per-thread running sums of the
// stored (bf16-rounded) values per column, kept in registers across all tiles of a sample.
__device__ __forceinline__ void epilogue_row_stats64(uint32_t t_addr, const float* __restrict__ bvec,
                                                     const __nv_bfloat16* __restrict__ res_row, __nv_bfloat16* y_row, bool valid,
                                                     float (&s1)[64], float (&s2)[64]) {
    uint32_t r[64];
    tmem_ld16_nowait(t_addr, r);
    tmem_ld16_nowait(t_addr + 16, r + 16);
    tmem_ld16_nowait(t_addr + 32, r + 32);
    tmem_ld16_nowait(t_addr + 48, r + 48);
    uint4 rr[8];
#pragma unroll
    for (int g = 0; g < 8; ++g) rr[g] = (res_row != nullptr && valid) ? ldg_nc_u4(res_row + 8 * g) : make_uint4(0, 0, 0, 0);
    tmem_ld_wait();
    if (!valid) return;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
        const float4 b0 = *reinterpret_cast<const float4*>(bvec + 8 * g), b1 = *reinterpret_cast<const float4*>(bvec + 8 * g + 4);
        float v[8];
        v[0] = __uint_as_float(r[8 * g + 0]) + b0.x + bf16_lo(rr[g].x); v[1] = __uint_as_float(r[8 * g + 1]) + b0.y + bf16_hi(rr[g].x);
        v[2] = __uint_as_float(r[8 * g + 2]) + b0.z + bf16_lo(rr[g].y); v[3] = __uint_as_float(r[8 * g + 3]) + b0.w + bf16_hi(rr[g].y);
        v[4] = __uint_as_float(r[8 * g + 4]) + b1.x + bf16_lo(rr[g].z); v[5] = __uint_as_float(r[8 * g + 5]) + b1.y + bf16_hi(rr[g].z);
        v[6] = __uint_as_float(r[8 * g + 6]) + b1.z + bf16_lo(rr[g].w); v[7] = __uint_as_float(r[8 * g + 7]) + b1.w + bf16_hi(rr[g].w);
        const uint4 pk = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
        *reinterpret_cast<uint4*>(y_row + 8 * g) = pk;
        const uint32_t w[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float lo = bf16_lo(w[e]), hi = bf16_hi(w[e]);
            s1[8 * g + 2 * e] += lo; s2[8 * g + 2 * e] = fmaf(lo, lo, s2[8 * g + 2 * e]);
            s1[8 * g + 2 * e + 1] += hi; s2[8 * g + 2 * e + 1] = fmaf(hi, hi, s2[8 * g + 2 * e + 1]);
        }
    }
}

// sum the 64 per-thread column sums over the 32 lanes of a warp (butterfly reduce-scatter: 62 shuffles per
// quantity, once per (CTA, sample)); lane l ends with columns 2l, 2l+1 and writes them to the partial row
__device__ __forceinline__ void flush_stats64(float (&s1)[64], float (&s2)[64], float* __restrict__ dst, int lane) {
#pragma unroll
    for (int half = 32; half >= 2; half >>= 1) {
        const int m = half >> 1;          // lane mask 16, 8, 4, 2, 1
        const bool up = (lane & m) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float k1 = up ? s1[i + half] : s1[i], x1 = up ? s1[i] : s1[i + half];
            const float k2 = up ? s2[i + half] : s2[i], x2 = up ? s2[i] : s2[i + half];
            s1[i] = k1 + __shfl_xor_sync(0xffffffffu, x1, m);
            s2[i] = k2 + __shfl_xor_sync(0xffffffffu, x2, m);
        }
    }
    *reinterpret_cast<float4*>(dst + 4 * lane) = make_float4(s1[0], s2[0], s1[1], s2[1]);
#pragma unroll
    for (int i = 0; i < 64; ++i) { s1[i] = 0.f; s2[i] = 0.f; }
}

__device__ __forceinline__ uint64_t make_sw128_desc_sbo(uint32_t saddr, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(sbo_bytes >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// ------------------------------------------------------------------------------ CTA-pair (cta_group::2) helpers
// PAIR mode (protocol probed on B200: tools/umma2_probe.cu): a cluster of two CTAs computes two w-adjacent bricks
// with ONE M = 256 MMA stream issued by the leader (rank 0).  Each CTA keeps its own halo planes and HALF of the
// weight rows in its own shared memory, so weight traffic (L2 -> smem and smem -> tensor core) per brick halves.
// TMA loads of both CTAs count bytes on the leader's "full" barrier; tcgen05.commit multicasts to the same-offset
// "empty"/"accumulator full" barriers of both CTAs; both epilogues arrive on the leader's "accumulator empty".
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t leader_addr(const void* smem_ptr) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(smem_ptr)), "r"(0u));
    return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_5d_pair(void* dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1, int c2,
                                                 int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1)
        : "memory");
}
template <bool PAIR>
__device__ __forceinline__ void umma_bf16_t(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if constexpr (PAIR) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    } else {
        umma_bf16(tmem_d, adesc, bdesc, idesc, accumulate);
    }
}
template <bool PAIR>
__device__ __forceinline__ void umma_commit_t(uint64_t* bar) {
    if constexpr (PAIR) {
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
    } else {
        umma_commit(bar);
    }
}

}  // namespace gg
