// halo_common.cuh -- pieces shared by the halo-brick kernels (conv_halo.cu, conv_roll.cu): brick geometry,
// CTA-pair (cta_group::2) PTX wrappers, shifted-descriptor helper and the statistics epilogue
#pragma once
#include "tc_common.cuh"

namespace gg {

constexpr int H_BH = 16, H_BW = 8;            // output brick (one depth plane): 16 x 8 = 128 positions
constexpr int H_MAX_SEGS = 4;
constexpr int H_THREADS = 224;
constexpr int H_SMEM_BUDGET = 227 * 1024;

// Epilogue of one row with fused GroupNorm statistics (bf16 outputs, any N tile): as epilogue_row, 32 columns per
// step; after each step the warp's 32 rows x 32 columns of STORED (bf16-rounded) values are summed per column by a
// butterfly reduce-scatter (31 shuffles per quantity) and lane l adds (sum, sum of squares) of column c0 + l to the
// partial row `stat_row` ([Cout8][2] floats owned by this (CTA, warp), zeroed by the launch).  Every lane must call.
__device__ __forceinline__ void epilogue_row_stats(uint32_t t_addr, int BN, int ncols, const float* __restrict__ bvec,
                                                   const __nv_bfloat16* __restrict__ res_row, __nv_bfloat16* y_row, bool valid,
                                                   float* __restrict__ stat_row, int lane) {
    for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld16_nowait(t_addr + c0, r);
        if (c0 + 16 < BN) tmem_ld16_nowait(t_addr + c0 + 16, r + 16);
        uint4 rr[4];
        const bool res32 = res_row != nullptr && aligned32(res_row), y32 = aligned32(y_row);
#pragma unroll
        for (int g = 0; g < 4; g += 2) {
            if (res32 && valid && c0 + 8 * g + 8 < ncols) {
                ldg_nc_u8(res_row + c0 + 8 * g, rr[g], rr[g + 1]);
            } else {
                rr[g] = (res_row != nullptr && valid && c0 + 8 * g < ncols) ? ldg_nc_u4(res_row + c0 + 8 * g) : make_uint4(0, 0, 0, 0);
                rr[g + 1] = (res_row != nullptr && valid && c0 + 8 * g + 8 < ncols) ? ldg_nc_u4(res_row + c0 + 8 * g + 8) : make_uint4(0, 0, 0, 0);
            }
        }
        tmem_ld_wait();
        float a[32];
        uint4 pk_even = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const int c = c0 + 8 * g;
            const bool on = valid && c < ncols;
            const float4 b0 = on ? *reinterpret_cast<const float4*>(bvec + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 b1 = on ? *reinterpret_cast<const float4*>(bvec + c + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            float v[8];
            v[0] = __uint_as_float(r[8 * g + 0]) + b0.x + bf16_lo(rr[g].x); v[1] = __uint_as_float(r[8 * g + 1]) + b0.y + bf16_hi(rr[g].x);
            v[2] = __uint_as_float(r[8 * g + 2]) + b0.z + bf16_lo(rr[g].y); v[3] = __uint_as_float(r[8 * g + 3]) + b0.w + bf16_hi(rr[g].y);
            v[4] = __uint_as_float(r[8 * g + 4]) + b1.x + bf16_lo(rr[g].z); v[5] = __uint_as_float(r[8 * g + 5]) + b1.y + bf16_hi(rr[g].z);
            v[6] = __uint_as_float(r[8 * g + 6]) + b1.z + bf16_lo(rr[g].w); v[7] = __uint_as_float(r[8 * g + 7]) + b1.w + bf16_hi(rr[g].w);
            const uint4 pk = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
            // pairs of 8-channel pieces leave as one 256-bit store when the row allows it
            if (on && y32 && (g & 1) == 0 && c + 8 < ncols) pk_even = pk;
            else if (on && y32 && (g & 1) == 1) stg_u8(y_row + c - 8, pk_even, pk);
            else if (on) *reinterpret_cast<uint4*>(y_row + c) = pk;
            const uint32_t w[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                a[8 * g + 2 * e] = on ? bf16_lo(w[e]) : 0.f;
                a[8 * g + 2 * e + 1] = on ? bf16_hi(w[e]) : 0.f;
            }
        }
        float b[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) b[i] = a[i] * a[i];
#pragma unroll
        for (int half = 16; half >= 1; half >>= 1) {      // lane mask 16, 8, 4, 2, 1: lane l ends with column l
            const bool up = (lane & half) != 0;
#pragma unroll
            for (int i = 0; i < half; ++i) {
                const float ka = up ? a[i + half] : a[i], xa = up ? a[i] : a[i + half];
                const float kb = up ? b[i + half] : b[i], xb = up ? b[i] : b[i + half];
                a[i] = ka + __shfl_xor_sync(0xffffffffu, xa, half);
                b[i] = kb + __shfl_xor_sync(0xffffffffu, xb, half);
            }
        }
        if (c0 + lane < ncols) {        // this lane is the only writer of its two floats: plain RMW keeps the sum order fixed
            float2* dst = reinterpret_cast<float2*>(stat_row) + c0 + lane;
            const float2 o = *dst;
            *dst = make_float2(o.x + a[0], o.y + b[0]);
        }
    }
}

// Transform of one bf16 pair (packed f32x2 arithmetic, two FMAs per instruction).  q = (s0, s1, b0, b1) of the two
// channels.  SiLU: q is pre-halved so that h = x s/2 + b/2 = v/2 exactly, and silu(v) = v sigmoid(v) = h + h tanh(h)
// (one MUFU per element; the same formulation and rounding points as gg_gn_apply).
__device__ __forceinline__ uint32_t xf_pair(uint32_t w, float4 q, bool silu) {
    uint64_t x, sc, sh, h;
    asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "r"(w << 16), "r"(w & 0xffff0000u));
    asm("mov.b64 %0, {%1, %2};" : "=l"(sc) : "f"(q.x), "f"(q.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(sh) : "f"(q.z), "f"(q.w));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(h) : "l"(x), "l"(sc), "l"(sh));
    float h0, h1;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(h0), "=f"(h1) : "l"(h));
    if (silu) {
        float t0, t1;
        asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(h0));
        asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(h1));
        uint64_t t, y;
        asm("mov.b64 %0, {%1, %2};" : "=l"(t) : "f"(t0), "f"(t1));
        asm("fma.rn.f32x2 %0, %1, %2, %1;" : "=l"(y) : "l"(h), "l"(t));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(h0), "=f"(h1) : "l"(y));
    }
    return pack_bf16(h0, h1);
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
// Arrive on a barrier of the cluster (the leader's).  NOT `.release.cluster`: that form compiles to MEMBAR.ALL.GPU + ERRBAR +
// CGAERRBAR in front of every arrive -- the epilogue warps of a pair kernel then wait for all their global stores to be
// acknowledged before the accumulator goes back to the tensor core (ncu source page of the folded-upsample conv: 25 % of all
// warp samples sat on that fence), and every transformed plane pays a GPU-scope fence on its way to the MMA issuer.  What these
// arrives order is shared memory of the arriving CTA written through the generic proxy (made visible to the tensor core by the
// fence.proxy.async in front) and TMEM reads (tcgen05.wait::ld + tcgen05.fence::before_thread_sync in front): the default
// release at CTA scope is what that needs (the form CUTLASS's ClusterBarrier::arrive uses for the same hand-overs).
// GG_ARRIVE_CLUSTER_RELEASE=1 (build define) restores the old form.
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
#ifdef GG_ARRIVE_CLUSTER_RELEASE
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
#else
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
#endif
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
