// norm_small.cu -- GroupNorm32 (+SiLU), LayerNorm, GEGLU, nearest upsample, timestep embedding,
// small-M linear.  All HBM- or latency-bound; channels-last bf16 activations, fp32 statistics.
#include <type_traits>

#include "common.cuh"

namespace gg {

// ------------------------------------------------------------------------------------------
// GroupNorm step 1: per-(n, chunk, channel) partial (sum, sum of squares); deterministic
// ------------------------------------------------------------------------------------------
static inline int64_t gn_chunk_positions(int64_t S, int32_t C) {
    int64_t cs = (32768 + C - 1) / C;           // ~64 KB of bf16 per block
    cs = (cs + 7) / 8 * 8;
    int64_t n = (S + cs - 1) / cs;
    if (n > 2048) { cs = (S + 2047) / 2048; }
    return cs;
}

__global__ void __launch_bounds__(256) gn_partial_kernel(const __nv_bfloat16* __restrict__ x, int64_t S, int C, int64_t cs,
                                                         int nchunks, float* __restrict__ partial) {
    extern __shared__ float sm[];  // [R][2*C]
    pdl_launch_dependents();
    pdl_wait();
    const int P8 = C >> 3;
    const int R = 256 / P8;
    const int oct = threadIdx.x % P8, row = threadIdx.x / P8;
    const int chunk = blockIdx.x, n = blockIdx.y;
    const int64_t p0 = (int64_t)chunk * cs, p1 = min(S, p0 + cs);
    float s[8], ss[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { s[e] = 0.f; ss[e] = 0.f; }
    if (row < R) {
        const __nv_bfloat16* base = x + ((int64_t)n * S) * C + oct * 8;
        int64_t p = p0 + row;
        // four loads in flight per thread
        auto acc = [&](const uint4& a) {
            const uint32_t wa[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float f0 = bf16_lo(wa[e]), f1 = bf16_hi(wa[e]);
                s[2 * e] += f0; s[2 * e + 1] += f1;
                ss[2 * e] = fmaf(f0, f0, ss[2 * e]); ss[2 * e + 1] = fmaf(f1, f1, ss[2 * e + 1]);
            }
        };
        for (; p + 3 * R < p1; p += 4 * R) {
            const uint4 a = ldg_nc_u4(base + p * C), b = ldg_nc_u4(base + (p + R) * C);
            const uint4 c = ldg_nc_u4(base + (p + 2 * R) * C), d = ldg_nc_u4(base + (p + 3 * R) * C);
            acc(a); acc(b); acc(c); acc(d);
        }
        for (; p < p1; p += R) acc(ldg_nc_u4(base + p * C));
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            sm[row * 2 * C + 2 * (oct * 8 + e)] = s[e];
            sm[row * 2 * C + 2 * (oct * 8 + e) + 1] = ss[e];
        }
    }
    __syncthreads();
    float* out = partial + ((int64_t)n * nchunks + chunk) * 2 * C;
    for (int j = threadIdx.x; j < 2 * C; j += 256) {
        float acc = 0.f;
        for (int r = 0; r < R; ++r) acc += sm[r * 2 * C + j];
        out[j] = acc;
    }
}

// step 2: group statistics in fp64 from the fp32 partials -> per-(n, channel) scale / shift
// (sum, sum of squares) of group g of sample n over the partial rows, in fp64, identical in every thread of the block
__device__ __forceinline__ void gn_group_sums(const gg_gn_finalize_args& a, int g, int n, double& s_out, double& ss_out) {
    const int C = a.C1 + a.C2;
    const int cpg = C / a.groups;
    double s = 0.0, ss = 0.0;
    // One flat index space over (partial row, channel of the group): 128 threads stride over it with four independent
    // loads in flight, instead of walking the group's channels one after the other (a launch used to cost cpg
    // dependent load latencies, ~13 us for the LDM widths).  Channels of the group may live in either source; the
    // summation order per thread is fixed, so results stay reproducible.
    const int c0 = g * cpg;
    auto load = [&](int e) -> float2 {
        int c, k;
        if (c0 + cpg <= a.C1 || c0 >= a.C1) {       // the whole group sits in one source (always true when C1 % cpg == 0)
            const bool first = c0 < a.C1;
            const int nch = first ? a.nchunks1 : a.nchunks2, Cs = first ? a.C1 : a.C2;
            k = e / cpg; c = e - k * cpg;
            if (k >= nch) return make_float2(0.f, 0.f);
            const float* part = first ? a.partial1 : a.partial2;
            return __ldg(reinterpret_cast<const float2*>(part + (((int64_t)n * nch + k) * Cs + (c0 - (first ? 0 : a.C1) + c)) * 2));
        }
        // group straddles the two sources: channel-major walk over max(nchunks) rows
        const int nmax = max(a.nchunks1, a.nchunks2);
        c = e / nmax; k = e - c * nmax;
        const int ch = c0 + c;
        const bool first = ch < a.C1;
        const int nch = first ? a.nchunks1 : a.nchunks2, Cs = first ? a.C1 : a.C2;
        if (k >= nch) return make_float2(0.f, 0.f);
        const float* part = first ? a.partial1 : a.partial2;
        return __ldg(reinterpret_cast<const float2*>(part + (((int64_t)n * nch + k) * Cs + (ch - (first ? 0 : a.C1))) * 2));
    };
    const bool one_src = (c0 + cpg <= a.C1 || c0 >= a.C1);
    const int total = cpg * (one_src ? (c0 < a.C1 ? a.nchunks1 : a.nchunks2) : max(a.nchunks1, a.nchunks2));
    int e = threadIdx.x;
    for (; e + 3 * 128 < total; e += 4 * 128) {
        const float2 v0 = load(e), v1 = load(e + 128), v2 = load(e + 256), v3 = load(e + 384);
        s += (double)v0.x; ss += (double)v0.y; s += (double)v1.x; ss += (double)v1.y;
        s += (double)v2.x; ss += (double)v2.y; s += (double)v3.x; ss += (double)v3.y;
    }
    for (; e < total; e += 128) {
        const float2 v = load(e);
        s += (double)v.x; ss += (double)v.y;
    }
    __shared__ double sh[2][4];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
    }
    if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s; sh[1][threadIdx.x >> 5] = ss; }
    __syncthreads();
    s_out = sh[0][0] + sh[0][1] + sh[0][2] + sh[0][3];
    ss_out = sh[1][0] + sh[1][1] + sh[1][2] + sh[1][3];
    __syncthreads();
}

__device__ __forceinline__ void gn_write_scale_shift(const gg_gn_finalize_args& a, int g, int n, double s, double ss) {
    const int C = a.C1 + a.C2;
    const int cpg = C / a.groups;
    const double cnt = (double)a.S * cpg;
    const double mean = s / cnt;
    double var = ss / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)a.eps));
    for (int cc = threadIdx.x; cc < cpg; cc += 128) {
        const int c = g * cpg + cc;
        const float ga = a.gamma ? a.gamma[c] : 1.0f, be = a.beta ? a.beta[c] : 0.0f;
        const float sc = ga * rstd;
        a.scale_shift[((int64_t)n * C + c) * 2] = sc;
        a.scale_shift[((int64_t)n * C + c) * 2 + 1] = be - (float)mean * sc;
    }
}

__global__ void __launch_bounds__(128) gn_finalize_kernel(const gg_gn_finalize_args a) {
    pdl_launch_dependents();
    pdl_wait();
    double s, ss;
    gn_group_sums(a, blockIdx.x, blockIdx.y, s, ss);
    gn_write_scale_shift(a, blockIdx.x, blockIdx.y, s, ss);
}

// Depth-slab mode over NVLink peer memory: the statistics of the WHOLE volume from the ranks' group sums.  push: this
// rank's (sum, sum sq) per (sample, group) -- 16 bytes each, not the partial-row tables -- stored straight into every
// rank's table, then the epoch into the peers' flags (last block); combine: wait for the peers' flags, add the R entries
// in rank order (identical statistics on every rank) and write scale / shift.
__global__ void __launch_bounds__(128) gn_slab_push_kernel(const gg_gn_finalize_args a) {
    const int g = blockIdx.x, n = blockIdx.y;
    double s, ss;
    gn_group_sums(a, g, n, s, ss);
    if (threadIdx.x < a.slab_world) {
        double* t = reinterpret_cast<double*>(a.slab_tables[threadIdx.x]) + (((int64_t)a.slab_rank * a.N + n) * a.groups + g) * 2;
        t[0] = s; t[1] = ss;
    }
    __threadfence_system();
    __syncthreads();
    __shared__ int is_last;
    if (threadIdx.x == 0) {
        const unsigned int total = gridDim.x * gridDim.y;
        is_last = atomicInc(a.slab_done_counter, total - 1u) == total - 1u;
    }
    __syncthreads();
    if (is_last && threadIdx.x < a.slab_world && threadIdx.x != a.slab_rank) {
        __threadfence_system();
        const uint32_t epoch = *reinterpret_cast<const volatile uint32_t*>(a.slab_epoch);
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(a.slab_flag_out[threadIdx.x]), "r"(epoch) : "memory");
    }
}

__global__ void __launch_bounds__(128) gn_slab_combine_kernel(const gg_gn_finalize_args a) {
    const int g = blockIdx.x, n = blockIdx.y;
    const uint32_t epoch = *reinterpret_cast<const volatile uint32_t*>(a.slab_epoch);
    if (threadIdx.x < a.slab_world && threadIdx.x != a.slab_rank) {
        const uint32_t* f = a.slab_flag_in[threadIdx.x];
        uint64_t t0 = 0;
        uint32_t spins = 0, v;
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
            if ((int32_t)(v - epoch) < 0 && (++spins & 0x3FFu) == 0) {
                uint64_t now;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                if (t0 == 0) t0 = now;
                else if (now - t0 > 10000000000ull) __trap();
            }
        } while ((int32_t)(v - epoch) < 0);
    }
    __syncthreads();
    __threadfence_system();
    double s = 0.0, ss = 0.0;
    const double* mine = reinterpret_cast<const double*>(a.slab_tables[a.slab_rank]);
    for (int r = 0; r < a.slab_world; ++r) {
        const double* t = mine + (((int64_t)r * a.N + n) * a.groups + g) * 2;
        s += __ldcg(t); ss += __ldcg(t + 1);
    }
    gn_write_scale_shift(a, g, n, s, ss);
}

// step 3: y = act(x * scale + shift), concat of two sources written as one CL tensor.
// A thread owns one 8-channel octet of a fixed sample for a strip of positions: its 16 affine
// coefficients stay in registers, loads are issued 4 deep.  SiLU(v) = h + h*tanh(h), h = v/2 (the
// 1/2 is folded into the coefficients): one MUFU per element instead of two (ex2 + rcp) -- at
// HBM speed the two-MUFU form alone needs ~70 % of the SFU rate.
__device__ __forceinline__ float tanh_fast(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <bool SILU>
__global__ void __launch_bounds__(256) gn_apply_kernel(const __nv_bfloat16* __restrict__ x1, int C1,
                                                       const __nv_bfloat16* __restrict__ x2, int C2,
                                                       const float* __restrict__ ss, __nv_bfloat16* __restrict__ y,
                                                       int64_t S, int64_t cs) {
    pdl_launch_dependents();
    pdl_wait();
    const int C = C1 + C2, P8 = C >> 3;
    const int R = 256 / P8;
    const int oct = threadIdx.x % P8, row = threadIdx.x / P8;
    if (row >= R) return;
    const int n = blockIdx.y;
    const int64_t p0 = (int64_t)blockIdx.x * cs, p1 = min(S, p0 + cs);
    const int c0 = oct * 8;
    float sc[8], sh[8];
    {
        const float4* sp = reinterpret_cast<const float4*>(ss + ((int64_t)n * C + c0) * 2);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float4 k = __ldg(sp + e);
            const float f = SILU ? 0.5f : 1.0f;
            sc[2 * e] = k.x * f; sh[2 * e] = k.y * f; sc[2 * e + 1] = k.z * f; sh[2 * e + 1] = k.w * f;
        }
    }
    const __nv_bfloat16* src;
    int Cs;
    if (c0 < C1) { src = x1 + ((int64_t)n * S) * C1 + c0; Cs = C1; }
    else { src = x2 + ((int64_t)n * S) * C2 + (c0 - C1); Cs = C2; }
    __nv_bfloat16* dst = y + ((int64_t)n * S) * C + c0;
    auto apply = [&](uint4 v, int64_t p) {
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        uint32_t o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float f0 = fmaf(bf16_lo(w[e]), sc[2 * e], sh[2 * e]), f1 = fmaf(bf16_hi(w[e]), sc[2 * e + 1], sh[2 * e + 1]);
            if (SILU) { f0 = fmaf(f0, tanh_fast(f0), f0); f1 = fmaf(f1, tanh_fast(f1), f1); }
            o[e] = pack_bf16(f0, f1);
        }
        stg_na_u4(dst + p * C, make_uint4(o[0], o[1], o[2], o[3]));
    };
    int64_t p = p0 + row;
    for (; p + 3 * R < p1; p += 4 * R) {
        const uint4 a = ldg_nc_u4(src + p * Cs), b = ldg_nc_u4(src + (p + R) * Cs);
        const uint4 c = ldg_nc_u4(src + (p + 2 * R) * Cs), d = ldg_nc_u4(src + (p + 3 * R) * Cs);
        apply(a, p); apply(b, p + R); apply(c, p + 2 * R); apply(d, p + 3 * R);
    }
    for (; p < p1; p += R) apply(ldg_nc_u4(src + p * Cs), p);
}

// ------------------------------------------------------------------------------------------
// GroupNorm (+SiLU) of SMALL tensors in one launch: statistics + apply.  The three-kernel form above costs three
// dependent launches of ~10 us each on tensors that fit in L2 (the 64 x 64 ... 8 x 8 latents of the LDM configs, the
// deep levels of the CCDM network).  Here a cluster of GNF_CL CTAs owns one sample: every CTA sums its slice of the
// positions per channel (fp32 per thread, fixed order), the per-CTA channel sums are exchanged through distributed
// shared memory and combined in fp64 in rank order (identical in every CTA of the cluster), then each CTA normalises
// its own slice, which it re-reads from L2.  Same apply formula and rounding points as gn_apply_kernel.
// ------------------------------------------------------------------------------------------
constexpr int GNF_CL = 8;
template <bool SILU>
__global__ void __launch_bounds__(256) gn_fused_kernel(const __nv_bfloat16* __restrict__ x1, int C1, const __nv_bfloat16* __restrict__ x2,
                                                       int C2, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       __nv_bfloat16* __restrict__ y, int64_t S, int groups, float eps) {
    extern __shared__ float2 gsm[];
    const int C = C1 + C2, P8 = C >> 3, R = 256 / P8;
    float2* part = gsm;                 // [R][C]  per-thread-row channel sums of this CTA
    float2* csum = gsm + R * C;         // [C]     this CTA's channel sums (read by the whole cluster)
    float2* ssm = csum + C;             // [C]     (scale, shift)
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int n = blockIdx.x / GNF_CL;
    const int oct = threadIdx.x % P8, row = threadIdx.x / P8;
    const bool active = row < R;
    const int c0 = oct * 8;
    const int64_t chunk = (S + GNF_CL - 1) / GNF_CL, p0 = (int64_t)rank * chunk, p1 = min(S, p0 + chunk);
    const __nv_bfloat16* src = nullptr;
    int Cs = 0;
    if (active) {
        if (c0 < C1) { src = x1 + ((int64_t)n * S) * C1 + c0; Cs = C1; }
        else { src = x2 + ((int64_t)n * S) * C2 + (c0 - C1); Cs = C2; }
    }
    // ---- pass 1: per-channel (sum, sum of squares) over this CTA's positions
    if (active) {
        float s1[8], s2[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) { s1[e] = 0.f; s2[e] = 0.f; }
        auto acc = [&](uint4 v) {
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float lo = bf16_lo(w[e]), hi = bf16_hi(w[e]);
                s1[2 * e] += lo; s2[2 * e] = fmaf(lo, lo, s2[2 * e]);
                s1[2 * e + 1] += hi; s2[2 * e + 1] = fmaf(hi, hi, s2[2 * e + 1]);
            }
        };
        int64_t p = p0 + row;
        for (; p + 3 * R < p1; p += 4 * R) {
            const uint4 a = ldg_nc_u4(src + p * Cs), b = ldg_nc_u4(src + (p + R) * Cs);
            const uint4 c = ldg_nc_u4(src + (p + 2 * R) * Cs), d = ldg_nc_u4(src + (p + 3 * R) * Cs);
            acc(a); acc(b); acc(c); acc(d);
        }
        for (; p < p1; p += R) acc(ldg_nc_u4(src + p * Cs));
#pragma unroll
        for (int e = 0; e < 8; ++e) part[row * C + c0 + e] = make_float2(s1[e], s2[e]);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
        float a = 0.f, b = 0.f;
        for (int r = 0; r < R; ++r) { const float2 v = part[r * C + c]; a += v.x; b += v.y; }
        csum[c] = make_float2(a, b);
    }
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    // ---- group statistics: fp64 over the cluster's CTAs and the group's channels.  A WARP per group: its lanes stride over the
    // (rank, channel) pairs -- independent DSMEM loads in flight instead of one thread walking GNF_CL * cpg dependent ones
    // (that walk alone cost ~10 us per launch) -- then a fixed-order butterfly in fp64: identical in every CTA of the cluster.
    const int cpg = C / groups;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int g = warp; g < groups; g += 8) {
        double s = 0.0, q = 0.0;
        const uint32_t local = (uint32_t)__cvta_generic_to_shared(csum + g * cpg);
        for (int e = lane; e < GNF_CL * cpg; e += 32) {
            const int r = e / cpg, c = e - r * cpg;
            uint32_t remote;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(r));
            float2 v;
            asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(remote + 8u * (uint32_t)c));
            s += (double)v.x; q += (double)v.y;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, o);
            q += __shfl_xor_sync(0xffffffffu, q, o);
        }
        const double cnt = (double)S * cpg, mean = s / cnt;
        double var = q / cnt - mean * mean;
        if (var < 0.0) var = 0.0;
        const float rstd = (float)(1.0 / sqrt(var + (double)eps));
        for (int c = g * cpg + lane; c < (g + 1) * cpg; c += 32) {
            const float sc = (gamma ? __ldg(gamma + c) : 1.0f) * rstd;
            ssm[c] = make_float2(sc, (beta ? __ldg(beta + c) : 0.0f) - (float)mean * sc);
        }
    }
    // nobody may leave (or reuse csum) while a peer still reads it; also publishes ssm within the CTA
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    // ---- pass 2: apply to this CTA's slice (L2-resident re-read)
    if (!active) return;
    float sc[8], sh[8];
    const float f = SILU ? 0.5f : 1.0f;
#pragma unroll
    for (int e = 0; e < 8; ++e) { const float2 k = ssm[c0 + e]; sc[e] = k.x * f; sh[e] = k.y * f; }
    __nv_bfloat16* dst = y + ((int64_t)n * S) * C + c0;
    auto apply = [&](uint4 v, int64_t p) {
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        uint32_t o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float f0 = fmaf(bf16_lo(w[e]), sc[2 * e], sh[2 * e]), f1 = fmaf(bf16_hi(w[e]), sc[2 * e + 1], sh[2 * e + 1]);
            if (SILU) { f0 = fmaf(f0, tanh_fast(f0), f0); f1 = fmaf(f1, tanh_fast(f1), f1); }
            o[e] = pack_bf16(f0, f1);
        }
        *reinterpret_cast<uint4*>(dst + p * C) = make_uint4(o[0], o[1], o[2], o[3]);
    };
    int64_t p = p0 + row;
    for (; p + 3 * R < p1; p += 4 * R) {
        const uint4 a = ldg_nc_u4(src + p * Cs), b = ldg_nc_u4(src + (p + R) * Cs);
        const uint4 c = ldg_nc_u4(src + (p + 2 * R) * Cs), d = ldg_nc_u4(src + (p + 3 * R) * Cs);
        apply(a, p); apply(b, p + R); apply(c, p + 2 * R); apply(d, p + 3 * R);
    }
    for (; p < p1; p += R) apply(ldg_nc_u4(src + p * Cs), p);
}

// ------------------------------------------------------------------------------------------
// GroupNorm (+SiLU) of SMALL tensors, shared-memory resident (round 2, second design).  The cluster form above re-reads its
// slice from L2 and keeps only 256 threads x 4 loads in flight per CTA -- measured slower than the three launches it
// replaces.  Here a sample's slice per CTA (<= ~190 KB) is brought into shared memory ONCE by bulk async copies (cp.async.bulk:
// the bytes in flight do not depend on the thread count; eight chunks, one mbarrier each, so the statistics start on the first
// chunk while the rest streams in), the statistics and the normalisation both read shared memory, and the tensor crosses
// L2 exactly once in each direction.  Cluster of 1..8 CTAs per sample (the host picks the smallest that fits, at least one CTA
// per ~48 KB); statistics combined over the cluster in fp64 through distributed shared memory as above.
// Same apply formula and rounding points as gn_apply_kernel.
// ------------------------------------------------------------------------------------------
constexpr int GNR_THREADS = 512;
constexpr int GNR_CHUNKS = 8;
__device__ __forceinline__ void gnr_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void gnr_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void gnr_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void gnr_mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(bar);
    uint32_t ok, spins = 0;
    uint64_t t0 = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (!ok && (++spins & 0xFFFu) == 0) {       // watchdog: a protocol bug must surface as a launch error, not a hang
            uint64_t now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000ull) __trap();
        }
    } while (!ok);
}
__device__ __forceinline__ void gnr_bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}

template <bool SILU>
__global__ void __launch_bounds__(GNR_THREADS, 1) gn_resident_kernel(const __nv_bfloat16* __restrict__ x1, int C1, const __nv_bfloat16* __restrict__ x2,
                                                                    int C2, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                    __nv_bfloat16* __restrict__ y, int64_t S, int groups, float eps, int rows_per_cta) {
    extern __shared__ __align__(128) uint8_t rsm[];
    const int C = C1 + C2, P8 = C >> 3, R = GNR_THREADS / P8;
    uint64_t* bars = reinterpret_cast<uint64_t*>(rsm);                   // [GNR_CHUNKS]
    float2* part = reinterpret_cast<float2*>(rsm + 128);                 // [R][C]  per-thread-row channel sums of this CTA
    float2* csum = part + R * C;                                         // [C]     this CTA's channel sums (read by the whole cluster)
    float2* ssm = csum + C;                                              // [C]     (scale, shift)
    uint8_t* buf1 = reinterpret_cast<uint8_t*>(ssm + C);                 // [rows_per_cta][C1] bf16
    uint8_t* buf2 = buf1 + (size_t)rows_per_cta * C1 * 2;                // [rows_per_cta][C2] bf16
    uint32_t rank, cl;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(cl));
    pdl_launch_dependents();
    const int n = blockIdx.x / cl;
    const int64_t p0 = (int64_t)rank * rows_per_cta;
    const int rows = (int)max((int64_t)0, min(S, p0 + rows_per_cta) - p0);
    const int rpc = max(1, (rows + GNR_CHUNKS - 1) / GNR_CHUNKS);
    if (threadIdx.x == 0) {
        for (int k = 0; k < GNR_CHUNKS; ++k) gnr_mbar_init(&bars[k], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    pdl_wait();
    if (threadIdx.x == 0) {
        for (int k = 0; k < GNR_CHUNKS; ++k) {
            const int r0 = k * rpc, r1 = min(rows, r0 + rpc);
            if (r1 <= r0) { gnr_mbar_arrive(&bars[k]); continue; }
            const uint32_t b1 = (uint32_t)(r1 - r0) * (uint32_t)C1 * 2u, b2 = (uint32_t)(r1 - r0) * (uint32_t)C2 * 2u;
            gnr_mbar_expect_tx(&bars[k], b1 + b2);
            gnr_bulk_load(buf1 + (size_t)r0 * C1 * 2, x1 + ((int64_t)n * S + p0 + r0) * C1, b1, &bars[k]);
            if (C2 > 0) gnr_bulk_load(buf2 + (size_t)r0 * C2 * 2, x2 + ((int64_t)n * S + p0 + r0) * C2, b2, &bars[k]);
        }
    }
    const int oct = threadIdx.x % P8, row = threadIdx.x / P8;
    const bool active = row < R;
    const int c0 = oct * 8;
    const uint8_t* base = c0 < C1 ? buf1 + c0 * 2 : buf2 + (c0 - C1) * 2;
    const int pitch = (c0 < C1 ? C1 : C2) * 2;
    // ---- pass 1: per-channel (sum, sum of squares) over this CTA's rows, chunk by chunk as they land
    {
        float s1[8], s2[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) { s1[e] = 0.f; s2[e] = 0.f; }
        auto acc = [&](uint4 v) {
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float lo = bf16_lo(w[e]), hi = bf16_hi(w[e]);
                s1[2 * e] += lo; s2[2 * e] = fmaf(lo, lo, s2[2 * e]);
                s1[2 * e + 1] += hi; s2[2 * e + 1] = fmaf(hi, hi, s2[2 * e + 1]);
            }
        };
        int p = row;
        for (int k = 0; k < GNR_CHUNKS; ++k) {
            gnr_mbar_wait(&bars[k], 0);
            if (!active) continue;
            const int r1 = min(rows, (k + 1) * rpc);
            for (; p + 3 * R < r1; p += 4 * R) {
                const uint4 a = *reinterpret_cast<const uint4*>(base + (size_t)p * pitch), b = *reinterpret_cast<const uint4*>(base + (size_t)(p + R) * pitch);
                const uint4 c = *reinterpret_cast<const uint4*>(base + (size_t)(p + 2 * R) * pitch), d = *reinterpret_cast<const uint4*>(base + (size_t)(p + 3 * R) * pitch);
                acc(a); acc(b); acc(c); acc(d);
            }
            for (; p < r1; p += R) acc(*reinterpret_cast<const uint4*>(base + (size_t)p * pitch));
        }
        if (active) {
#pragma unroll
            for (int e = 0; e < 8; ++e) part[row * C + c0 + e] = make_float2(s1[e], s2[e]);
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += GNR_THREADS) {
        float a = 0.f, b = 0.f;
        for (int r = 0; r < R; ++r) { const float2 v = part[r * C + c]; a += v.x; b += v.y; }
        csum[c] = make_float2(a, b);
    }
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    // ---- group statistics: a warp per group, lanes stride over the (rank, channel) pairs of the cluster, fp64, fixed order
    const int cpg = C / groups;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int g = warp; g < groups; g += GNR_THREADS / 32) {
        double s = 0.0, q = 0.0;
        const uint32_t local = (uint32_t)__cvta_generic_to_shared(csum + g * cpg);
        const int total = (int)cl * cpg;
        for (int e = lane; e < total; e += 64) {            // two independent remote loads in flight per lane
            const int e2 = e + 32;
            const int r = e / cpg, c = e - r * cpg;
            const int r2 = e2 < total ? e2 / cpg : r, c2 = e2 < total ? e2 - r2 * cpg : c;
            uint32_t ra, rb;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local), "r"(r));
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rb) : "r"(local), "r"(r2));
            float2 va, vb;
            asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(va.x), "=f"(va.y) : "r"(ra + 8u * (uint32_t)c));
            asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(vb.x), "=f"(vb.y) : "r"(rb + 8u * (uint32_t)c2));
            s += (double)va.x; q += (double)va.y;
            if (e2 < total) { s += (double)vb.x; q += (double)vb.y; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, o);
            q += __shfl_xor_sync(0xffffffffu, q, o);
        }
        const double cnt = (double)S * cpg, mean = s / cnt;
        double var = q / cnt - mean * mean;
        if (var < 0.0) var = 0.0;
        const float rstd = (float)(1.0 / sqrt(var + (double)eps));
        for (int c = g * cpg + lane; c < (g + 1) * cpg; c += 32) {
            const float sc = (gamma ? __ldg(gamma + c) : 1.0f) * rstd;
            ssm[c] = make_float2(sc, (beta ? __ldg(beta + c) : 0.0f) - (float)mean * sc);
        }
    }
    // nobody may leave while a peer still reads its csum; also publishes ssm within the CTA
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    // ---- pass 2: normalise this CTA's rows out of shared memory
    if (!active) return;
    float sc[8], sh[8];
    const float f = SILU ? 0.5f : 1.0f;
#pragma unroll
    for (int e = 0; e < 8; ++e) { const float2 k = ssm[c0 + e]; sc[e] = k.x * f; sh[e] = k.y * f; }
    __nv_bfloat16* dst = y + ((int64_t)n * S + p0) * C + c0;
    auto apply = [&](uint4 v, int p) {
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        uint32_t o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float f0 = fmaf(bf16_lo(w[e]), sc[2 * e], sh[2 * e]), f1 = fmaf(bf16_hi(w[e]), sc[2 * e + 1], sh[2 * e + 1]);
            if (SILU) { f0 = fmaf(f0, tanh_fast(f0), f0); f1 = fmaf(f1, tanh_fast(f1), f1); }
            o[e] = pack_bf16(f0, f1);
        }
        *reinterpret_cast<uint4*>(dst + (int64_t)p * C) = make_uint4(o[0], o[1], o[2], o[3]);
    };
    int p = row;
    for (; p + 3 * R < rows; p += 4 * R) {
        const uint4 a = *reinterpret_cast<const uint4*>(base + (size_t)p * pitch), b = *reinterpret_cast<const uint4*>(base + (size_t)(p + R) * pitch);
        const uint4 c = *reinterpret_cast<const uint4*>(base + (size_t)(p + 2 * R) * pitch), d = *reinterpret_cast<const uint4*>(base + (size_t)(p + 3 * R) * pitch);
        apply(a, p); apply(b, p + R); apply(c, p + 2 * R); apply(d, p + 3 * R);
    }
    for (; p < rows; p += R) apply(*reinterpret_cast<const uint4*>(base + (size_t)p * pitch), p);
}

// cluster size and rows per CTA of the shared-memory-resident form; 0 = the sample does not fit (cluster of 8 x ~190 KB)
static int gn_resident_plan(int64_t S, int C, int* rows_per_cta, size_t* smem) {
    if (C <= 0 || C % 8 != 0 || C > 2048 || S <= 0) return 0;
    const int R = GNR_THREADS / (C / 8);
    if (R < 1) return 0;
    const size_t overhead = 128 + (size_t)(R * C + 2 * C) * sizeof(float2);
    const size_t cap = 226 * 1024 - overhead;
    const int64_t sample = S * C * 2;
    int cl = 1;
    while (cl < 8 && sample > (int64_t)cl * 48 * 1024) cl *= 2;
    while (cl < 8 && (uint64_t)((S + cl - 1) / cl) * C * 2 > cap) cl *= 2;
    const int64_t rows = (S + cl - 1) / cl;
    if ((uint64_t)rows * C * 2 > cap) return 0;
    *rows_per_cta = (int)rows;
    *smem = overhead + (size_t)rows * C * 2;
    return cl;
}

// ------------------------------------------------------------------------------------------
// LayerNorm (one warp per row), GEGLU, nearest x2 upsample
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) layernorm_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, __nv_bfloat16* __restrict__ y,
                                                        int64_t rows, int C, float eps) {
    pdl_launch_dependents();
    pdl_wait();
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const __nv_bfloat16* xr = x + row * C;
    float s = 0.f;
    for (int c = lane * 8; c < C; c += 256) {
        uint4 v = *reinterpret_cast<const uint4*>(xr + c);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) s += bf16_lo(w[e]) + bf16_hi(w[e]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)C;
    float q = 0.f;
    for (int c = lane * 8; c < C; c += 256) {
        uint4 v = *reinterpret_cast<const uint4*>(xr + c);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float a = bf16_lo(w[e]) - mean, b = bf16_hi(w[e]) - mean;
            q += a * a + b * b;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q / (float)C + eps);
    __nv_bfloat16* yr = y + row * C;
    for (int c = lane * 8; c < C; c += 256) {
        uint4 v = *reinterpret_cast<const uint4*>(xr + c);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        uint32_t o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int cc = c + 2 * e;
            const float a = (bf16_lo(w[e]) - mean) * rstd * gamma[cc] + beta[cc];
            const float b = (bf16_hi(w[e]) - mean) * rstd * gamma[cc + 1] + beta[cc + 1];
            o[e] = pack_bf16(a, b);
        }
        *reinterpret_cast<uint4*>(yr + c) = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

__global__ void __launch_bounds__(256) geglu_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                    int64_t rows, int inner) {
    pdl_launch_dependents();
    pdl_wait();
    const int P8 = inner >> 3;
    const int64_t total = rows * P8;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / P8;
        const int c0 = (int)(i - r * P8) * 8;
        uint4 a = ldg_nc_u4(x + r * 2 * inner + c0), g = ldg_nc_u4(x + r * 2 * inner + inner + c0);
        const uint32_t wa[4] = {a.x, a.y, a.z, a.w}, wg[4] = {g.x, g.y, g.z, g.w};
        uint32_t o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e)
            o[e] = pack_bf16(bf16_lo(wa[e]) * gelu_erf(bf16_lo(wg[e])), bf16_hi(wa[e]) * gelu_erf(bf16_hi(wg[e])));
        stg_na_u4(y + r * inner + c0, make_uint4(o[0], o[1], o[2], o[3]));
    }
}

__global__ void __launch_bounds__(256) upsample2x_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                         int N, int D, int H, int W, int C, int fd, int fh, int fw) {
    pdl_launch_dependents();
    pdl_wait();
    const int P8 = C >> 3;
    const int Do = D * fd, Ho = H * fh, Wo = W * fw;
    const int64_t total = (int64_t)N * Do * Ho * Wo * P8;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t t = i;
        const int oct = (int)(t % P8); t /= P8;
        const int w = (int)(t % Wo); t /= Wo;
        const int h = (int)(t % Ho); t /= Ho;
        const int d = (int)(t % Do); t /= Do;
        const int64_t src = ((((int64_t)t * D + d / fd) * H + h / fh) * W + w / fw) * C + oct * 8;
        stg_na_u4(y + (i / P8) * C + oct * 8, __ldg(reinterpret_cast<const uint4*>(x + src)));
    }
}

// ------------------------------------------------------------------------------------------
// timestep embedding + small-M linear (fp32; accurate sinf/cosf/expf -- no fast-math here)
// ------------------------------------------------------------------------------------------
__global__ void timestep_embedding_kernel(const float* __restrict__ t, float* __restrict__ emb, int B, int dim, float max_period) {
    pdl_launch_dependents();
    pdl_wait();
    const int half = dim / 2;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * half) return;
    const int b = i / half, j = i - b * half;
    const float freq = expf(-logf(max_period) * (float)j / (float)half);
    const float arg = t[b] * freq;
    emb[(int64_t)b * dim + j] = cosf(arg);
    emb[(int64_t)b * dim + half + j] = sinf(arg);
    if ((dim & 1) && j == 0) emb[(int64_t)b * dim + dim - 1] = 0.f;
}

// one warp per NF consecutive output features: every weight row is read once (128-bit when K % 4 == 0); a chunk of up to MT rows
// of x is loaded ONCE per warp and used for all NF features (NF = 4 for the wide projection of the timestep embedding onto every
// ResBlock, [16, 640] x [~14 k, 640]^T: with one feature per warp each of the 14 k warps re-read all 40 KB of x through L1).
// The arithmetic per output is the same expression in the same order for every NF: results do not depend on NF.
// ACT_IN (SiLU on x) is a template parameter: as a run-time flag the compiler evaluated expf + a precise division for every
// element of x and selected afterwards -- ~25 instructions next to each FMA.
template <int MT, bool VEC, int NF, bool ACT_IN>
__global__ void __launch_bounds__(256) small_linear_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                           const float* __restrict__ bias, float* __restrict__ y, int M, int N,
                                                           int K, int act_out) {
    pdl_launch_dependents();
    pdl_wait();
    const int n0 = (blockIdx.x * 8 + (threadIdx.x >> 5)) * NF;
    const int lane = threadIdx.x & 31;
    if (n0 >= N) return;
    auto act = [&](float v) { return ACT_IN ? v / (1.0f + expf(-v)) : v; };
    for (int m0 = 0; m0 < M; m0 += MT) {
        float acc[MT][NF];
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int f = 0; f < NF; ++f) acc[m][f] = 0.f;
        if (VEC) {
            for (int k = lane * 4; k < K; k += 128) {
                float4 wv[NF];
#pragma unroll
                for (int f = 0; f < NF; ++f)
                    wv[f] = n0 + f < N ? __ldg(reinterpret_cast<const float4*>(w + (int64_t)(n0 + f) * K + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int m = 0; m < MT; ++m) {
                    if (m0 + m < M) {
                        const float4 xv = __ldg(reinterpret_cast<const float4*>(x + (int64_t)(m0 + m) * K + k));
                        const float a0 = act(xv.x), a1 = act(xv.y), a2 = act(xv.z), a3 = act(xv.w);
#pragma unroll
                        for (int f = 0; f < NF; ++f) acc[m][f] += a0 * wv[f].x + a1 * wv[f].y + a2 * wv[f].z + a3 * wv[f].w;
                    }
                }
            }
        } else {
            for (int k = lane; k < K; k += 32) {
                float wv[NF];
#pragma unroll
                for (int f = 0; f < NF; ++f) wv[f] = n0 + f < N ? __ldg(w + (int64_t)(n0 + f) * K + k) : 0.f;
#pragma unroll
                for (int m = 0; m < MT; ++m)
                    if (m0 + m < M) {
                        const float a = act(__ldg(x + (int64_t)(m0 + m) * K + k));
#pragma unroll
                        for (int f = 0; f < NF; ++f) acc[m][f] += a * wv[f];
                    }
            }
        }
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int f = 0; f < NF; ++f) {
                float v = acc[m][f];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0 && m0 + m < M && n0 + f < N) {
                    v += bias ? bias[n0 + f] : 0.f;
                    if (act_out) v = v / (1.0f + expf(-v));
                    y[(int64_t)(m0 + m) * N + n0 + f] = v;
                }
            }
    }
}

}  // namespace gg

using namespace gg;

extern "C" int32_t gg_gn_num_chunks(int64_t S, int32_t C) {
    if (S <= 0 || C <= 0) return 0;
    const int64_t cs = gn_chunk_positions(S, C);
    return (int32_t)((S + cs - 1) / cs);
}

extern "C" int gg_gn_partial(const void* x_cl, int32_t N, int64_t S, int32_t C, float* partial, gg_stream_t stream) {
    GG_REQUIRE(x_cl && partial && N > 0 && S > 0 && C > 0, GG_ERR_BAD_ARG);
    GG_REQUIRE(C % 8 == 0 && C <= 2048, GG_ERR_UNSUPPORTED);
    GG_REQUIRE(aligned(x_cl, 16), GG_ERR_ALIGNMENT);
    const int64_t cs = gn_chunk_positions(S, C);
    const int nchunks = (int)((S + cs - 1) / cs);
    const int R = 256 / (C / 8);
    const size_t smem = (size_t)R * 2 * C * sizeof(float);
    dim3 grid(nchunks, N);
    const cudaError_t e = launch_k(gn_partial_kernel, grid, dim3(256), smem, as_stream(stream), reinterpret_cast<const __nv_bfloat16*>(x_cl), S,
                                   (int)C, cs, nchunks, partial);
    if (e != cudaSuccess) return (int)e;
    return launch_result();
}

extern "C" int gg_gn_finalize(const gg_gn_finalize_args* a, gg_stream_t stream) {
    GG_REQUIRE(a && a->partial1 && a->scale_shift && a->N > 0 && a->groups > 0 && a->S > 0, GG_ERR_BAD_ARG);
    GG_REQUIRE(a->C2 == 0 || a->partial2, GG_ERR_BAD_ARG);
    GG_REQUIRE((a->C1 + a->C2) % a->groups == 0, GG_ERR_UNSUPPORTED);
    if (a->slab_world > 1) {
        GG_REQUIRE(a->slab_world <= 8 && a->slab_rank >= 0 && a->slab_rank < a->slab_world && a->slab_epoch && a->slab_done_counter &&
                   (a->slab_phase & 3) != 0, GG_ERR_BAD_ARG);
        for (int r = 0; r < a->slab_world; ++r)
            GG_REQUIRE(a->slab_tables[r] != nullptr && (r == a->slab_rank || (a->slab_flag_out[r] && a->slab_flag_in[r])), GG_ERR_BAD_ARG);
        dim3 sgrid(a->groups, a->N);
        if (a->slab_phase & 1) {
            gn_slab_push_kernel<<<sgrid, 128, 0, as_stream(stream)>>>(*a);
            const int st = launch_result();
            if (st != GG_OK) return st;
        }
        if (a->slab_phase & 2) {
            gn_slab_combine_kernel<<<sgrid, 128, 0, as_stream(stream)>>>(*a);
            return launch_result();
        }
        return GG_OK;
    }
    dim3 grid(a->groups, a->N);
    const cudaError_t e = launch_k(gn_finalize_kernel, grid, dim3(128), 0, as_stream(stream), *a);
    if (e != cudaSuccess) return (int)e;
    return launch_result();
}

extern "C" int gg_gn_apply(const void* x1_cl, int32_t C1, const void* x2_cl, int32_t C2, const float* scale_shift, void* y_cl,
                           int32_t N, int64_t S, int32_t silu, gg_stream_t stream) {
    GG_REQUIRE(x1_cl && scale_shift && y_cl && N > 0 && S > 0 && C1 > 0 && C2 >= 0 && (C2 == 0 || x2_cl), GG_ERR_BAD_ARG);
    GG_REQUIRE(C1 % 8 == 0 && C2 % 8 == 0, GG_ERR_UNSUPPORTED);
    GG_REQUIRE(aligned(x1_cl, 16) && aligned(y_cl, 16) && aligned(scale_shift, 16) && (!x2_cl || aligned(x2_cl, 16)),
               GG_ERR_ALIGNMENT);
    GG_REQUIRE(C1 + C2 <= 2048, GG_ERR_UNSUPPORTED);
    // ~64 KB of output per block, at least 2 waves of blocks when the tensor allows it
    const int C = C1 + C2;
    const int R = 256 / (C / 8);
    int64_t cs = std::max<int64_t>(4 * R, (32768 + C - 1) / C);
    cs = (cs + 4 * R - 1) / (4 * R) * (4 * R);
    int64_t nblk = (S + cs - 1) / cs;
    if (nblk > 65535 * 16) return GG_ERR_UNSUPPORTED;
    dim3 grid((unsigned)nblk, (unsigned)N);
    const cudaError_t e = launch_k(silu ? gn_apply_kernel<true> : gn_apply_kernel<false>, grid, dim3(256), 0, as_stream(stream),
                                   reinterpret_cast<const __nv_bfloat16*>(x1_cl), (int)C1, reinterpret_cast<const __nv_bfloat16*>(x2_cl), (int)C2,
                                   scale_shift, reinterpret_cast<__nv_bfloat16*>(y_cl), S, cs);
    if (e != cudaSuccess) return (int)e;
    return launch_result();
}

extern "C" int32_t gg_gn_fused_resident(int64_t S, int32_t C) {
    int rows = 0;
    size_t smem = 0;
    return gn_resident_plan(S, C, &rows, &smem);
}

extern "C" int gg_gn_fused(const void* x1_cl, int32_t C1, const void* x2_cl, int32_t C2, const float* gamma, const float* beta,
                           void* y_cl, int32_t N, int64_t S, int32_t groups, float eps, int32_t silu, gg_stream_t stream) {
    GG_REQUIRE(x1_cl && y_cl && N > 0 && S > 0 && C1 > 0 && C2 >= 0 && (C2 == 0 || x2_cl) && groups > 0, GG_ERR_BAD_ARG);
    const int C = C1 + C2;
    GG_REQUIRE(C1 % 8 == 0 && C2 % 8 == 0 && C % groups == 0 && C <= 2048 && groups <= 256, GG_ERR_UNSUPPORTED);
    GG_REQUIRE(aligned(x1_cl, 16) && aligned(y_cl, 16) && (!x2_cl || aligned(x2_cl, 16)), GG_ERR_ALIGNMENT);
    const __nv_bfloat16* a1 = reinterpret_cast<const __nv_bfloat16*>(x1_cl);
    const __nv_bfloat16* a2 = reinterpret_cast<const __nv_bfloat16*>(x2_cl);
    __nv_bfloat16* yy = reinterpret_cast<__nv_bfloat16*>(y_cl);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gn_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(gn_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(gn_resident_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(gn_resident_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.stream = as_stream(stream);
    int rows = 0;
    size_t rsmem = 0;
    const int cl = gn_resident_plan(S, C, &rows, &rsmem);
    if (cl > 0) {       // the sample fits the shared memory of a cluster: one pass over L2 in each direction
        cfg.gridDim = dim3((unsigned)(N * cl)); cfg.blockDim = dim3(GNR_THREADS); cfg.dynamicSmemBytes = rsmem;
        attr[0].val.clusterDim.x = (unsigned)cl;
        cfg.numAttrs = pdl_enabled() ? 2 : 1;
        cudaError_t e = silu ? cudaLaunchKernelEx(&cfg, gn_resident_kernel<true>, a1, (int)C1, a2, (int)C2, gamma, beta, yy, S, (int)groups, eps, rows)
                             : cudaLaunchKernelEx(&cfg, gn_resident_kernel<false>, a1, (int)C1, a2, (int)C2, gamma, beta, yy, S, (int)groups, eps, rows);
        if (e != cudaSuccess) return (int)e;
        return launch_result();
    }
    const int R = 256 / (C / 8);
    const size_t smem = (size_t)(R * C + 2 * C) * sizeof(float2);
    GG_REQUIRE(smem <= 64 * 1024, GG_ERR_UNSUPPORTED);
    cfg.gridDim = dim3((unsigned)(N * GNF_CL)); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem;
    attr[0].val.clusterDim.x = GNF_CL;
    cfg.numAttrs = 1;
    cudaError_t e = silu ? cudaLaunchKernelEx(&cfg, gn_fused_kernel<true>, a1, (int)C1, a2, (int)C2, gamma, beta, yy, S, (int)groups, eps)
                         : cudaLaunchKernelEx(&cfg, gn_fused_kernel<false>, a1, (int)C1, a2, (int)C2, gamma, beta, yy, S, (int)groups, eps);
    if (e != cudaSuccess) return (int)e;
    return launch_result();
}

extern "C" int gg_layernorm(const void* x, const float* gamma, const float* beta, void* y, int64_t rows, int32_t C, float eps,
                            gg_stream_t stream) {
    GG_REQUIRE(x && gamma && beta && y && rows > 0 && C > 0, GG_ERR_BAD_ARG);
    GG_REQUIRE(C % 8 == 0, GG_ERR_UNSUPPORTED);
    GG_REQUIRE(aligned(x, 16) && aligned(y, 16), GG_ERR_ALIGNMENT);
    const unsigned blocks = (unsigned)((rows + 7) / 8);
    const cudaError_t e = launch_k(layernorm_kernel, dim3(blocks), dim3(256), 0, as_stream(stream), reinterpret_cast<const __nv_bfloat16*>(x), gamma,
                                   beta, reinterpret_cast<__nv_bfloat16*>(y), rows, (int)C, eps);
    if (e != cudaSuccess) return (int)e;
    return launch_result();
}

// row softmax of scale * x (fp32 in, bf16 out): one 256-thread block per row, the row held in registers
// (n <= 256 * 4 * SM_MAXV), max and sum by warp shuffles + one smem exchange
constexpr int SM_MAXV = 8;      // float4 per thread -> n <= 8192
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int n, float scale) {
    const float* xr = x + (int64_t)blockIdx.x * n;
    __nv_bfloat16* yr = y + (int64_t)blockIdx.x * n;
    __shared__ float red[2][8];
    float4 v[SM_MAXV];
    const int nv = n >> 2;
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < SM_MAXV; ++i) {
        const int j = threadIdx.x + i * 256;
        if (j < nv) {
            v[i] = ldg_nc_f4(xr + 4 * j);
            v[i].x *= scale; v[i].y *= scale; v[i].z *= scale; v[i].w *= scale;
            m = fmaxf(fmaxf(fmaxf(m, v[i].x), fmaxf(v[i].y, v[i].z)), v[i].w);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[0][threadIdx.x >> 5] = m;
    __syncthreads();
    m = red[0][0];
#pragma unroll
    for (int i = 1; i < 8; ++i) m = fmaxf(m, red[0][i]);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < SM_MAXV; ++i) {
        const int j = threadIdx.x + i * 256;
        if (j < nv) {
            v[i].x = __expf(v[i].x - m); v[i].y = __expf(v[i].y - m); v[i].z = __expf(v[i].z - m); v[i].w = __expf(v[i].w - m);
            s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[1][threadIdx.x >> 5] = s;
    __syncthreads();
    s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[1][i];
    const float inv = 1.0f / s;
#pragma unroll
    for (int i = 0; i < SM_MAXV; ++i) {
        const int j = threadIdx.x + i * 256;
        if (j < nv) *reinterpret_cast<uint2*>(yr + 4 * j) = make_uint2(pack_bf16(v[i].x * inv, v[i].y * inv), pack_bf16(v[i].z * inv, v[i].w * inv));
    }
}

// bf16 [R, C] -> [C, R] through a padded 64 x 64 shared-memory tile
__global__ void __launch_bounds__(256) transpose_bf16_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int R, int C) {
    __shared__ __nv_bfloat16 tile[64][66];
    const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
    for (int i = threadIdx.x; i < 64 * 64; i += 256) {
        const int r = i >> 6, c = i & 63;
        tile[r][c] = (r0 + r < R && c0 + c < C) ? x[(int64_t)(r0 + r) * C + c0 + c] : __float2bfloat16(0.f);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * 64; i += 256) {
        const int c = i >> 6, r = i & 63;
        if (c0 + c < C && r0 + r < R) y[(int64_t)(c0 + c) * R + r0 + r] = tile[r][c];
    }
}

extern "C" int gg_softmax_rows(const float* x, void* y, int64_t rows, int32_t n, float scale, gg_stream_t stream) {
    GG_REQUIRE(x && y && rows > 0 && n > 0, GG_ERR_BAD_ARG);
    GG_REQUIRE(n % 4 == 0 && n <= 256 * 4 * SM_MAXV && rows <= 0x7fffffffll, GG_ERR_UNSUPPORTED);
    GG_REQUIRE(aligned(x, 16) && aligned(y, 8), GG_ERR_ALIGNMENT);
    softmax_rows_kernel<<<(unsigned)rows, 256, 0, as_stream(stream)>>>(x, reinterpret_cast<__nv_bfloat16*>(y), n, scale);
    return launch_result();
}

extern "C" int gg_transpose_bf16(const void* x, void* y, int32_t R, int32_t C, gg_stream_t stream) {
    GG_REQUIRE(x && y && R > 0 && C > 0, GG_ERR_BAD_ARG);
    GG_REQUIRE((R + 63) / 64 <= 65535, GG_ERR_UNSUPPORTED);
    const dim3 grid((unsigned)((C + 63) / 64), (unsigned)((R + 63) / 64));
    transpose_bf16_kernel<<<grid, 256, 0, as_stream(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(x),
                                                             reinterpret_cast<__nv_bfloat16*>(y), R, C);
    return launch_result();
}

extern "C" int gg_geglu(const void* x, void* y, int64_t rows, int32_t inner, gg_stream_t stream) {
    GG_REQUIRE(x && y && rows > 0 && inner > 0, GG_ERR_BAD_ARG);
    GG_REQUIRE(inner % 8 == 0, GG_ERR_UNSUPPORTED);
    GG_REQUIRE(aligned(x, 16) && aligned(y, 16), GG_ERR_ALIGNMENT);
    const int64_t total = rows * (inner / 8);
    const unsigned blocks = (unsigned)std::min<int64_t>((total + 255) / 256, (int64_t)num_sms() * 32);
    const cudaError_t e = launch_k(geglu_kernel, dim3(blocks), dim3(256), 0, as_stream(stream), reinterpret_cast<const __nv_bfloat16*>(x),
                                   reinterpret_cast<__nv_bfloat16*>(y), rows, (int)inner);
    if (e != cudaSuccess) return (int)e;
    return launch_result();
}

extern "C" int gg_upsample2x(const void* x_cl, void* y_cl, int32_t N, int32_t D, int32_t H, int32_t W, int32_t C, int32_t dims,
                             gg_stream_t stream) {
    GG_REQUIRE(x_cl && y_cl && N > 0 && D > 0 && H > 0 && W > 0 && C > 0 && dims >= 1 && dims <= 3, GG_ERR_BAD_ARG);
    GG_REQUIRE(C % 8 == 0, GG_ERR_UNSUPPORTED);
    GG_REQUIRE(aligned(x_cl, 16) && aligned(y_cl, 16), GG_ERR_ALIGNMENT);
    const int fd = dims >= 3 ? 2 : 1, fh = dims >= 2 ? 2 : 1, fw = 2;
    const int64_t total = (int64_t)N * D * fd * H * fh * W * fw * (C / 8);
    const unsigned blocks = (unsigned)std::min<int64_t>((total + 255) / 256, (int64_t)num_sms() * 32);
    const cudaError_t e = launch_k(upsample2x_kernel, dim3(blocks), dim3(256), 0, as_stream(stream), reinterpret_cast<const __nv_bfloat16*>(x_cl),
                                   reinterpret_cast<__nv_bfloat16*>(y_cl), (int)N, (int)D, (int)H, (int)W, (int)C, fd, fh, fw);
    if (e != cudaSuccess) return (int)e;
    return launch_result();
}

extern "C" int gg_timestep_embedding(const float* t, float* emb, int32_t B, int32_t dim, float max_period, gg_stream_t stream) {
    GG_REQUIRE(t && emb && B > 0 && dim >= 2, GG_ERR_BAD_ARG);
    const int total = B * (dim / 2);
    const cudaError_t e = launch_k(timestep_embedding_kernel, dim3((total + 127) / 128), dim3(128), 0, as_stream(stream), t, emb, (int)B, (int)dim,
                                   max_period);
    if (e != cudaSuccess) return (int)e;
    return launch_result();
}

extern "C" int gg_small_linear(const float* x, const float* w, const float* b, float* y, int32_t M, int32_t N, int32_t K,
                               int32_t act_in, int32_t act_out, gg_stream_t stream) {
    GG_REQUIRE(x && w && y && M > 0 && N > 0 && K > 0, GG_ERR_BAD_ARG);
    const bool vec = (K % 4 == 0) && aligned(x, 16) && aligned(w, 16);
    cudaStream_t s = as_stream(stream);
    // four features per warp once there are enough features to fill the GPU that way (>= 2 warps per SM scheduler)
    const bool wide = N >= 4 * 8 * 2 * num_sms();
    auto pick = [&](auto act_tag) {
        constexpr bool A = decltype(act_tag)::value;
        return wide ? (M <= 4 ? (vec ? small_linear_kernel<4, true, 4, A> : small_linear_kernel<4, false, 4, A>)
                              : (vec ? small_linear_kernel<16, true, 4, A> : small_linear_kernel<16, false, 4, A>))
                    : (M <= 4 ? (vec ? small_linear_kernel<4, true, 1, A> : small_linear_kernel<4, false, 1, A>)
                              : (vec ? small_linear_kernel<16, true, 1, A> : small_linear_kernel<16, false, 1, A>));
    };
    auto fn = act_in ? pick(std::true_type{}) : pick(std::false_type{});
    const unsigned blocks = (unsigned)((N + 8 * (wide ? 4 : 1) - 1) / (8 * (wide ? 4 : 1)));
    const cudaError_t e = launch_k(fn, dim3(blocks), dim3(256), 0, s, x, w, b, y, (int)M, (int)N, (int)K, (int)act_out);
    if (e != cudaSuccess) return (int)e;
    return launch_result();
}
