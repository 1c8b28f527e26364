// conv_halo.cu -- stride-1 convolution with the input brick (plus halo) resident in shared memory
//
// conv_tcgen05.cu reloads the 128-position A tile from L2 for every filter tap.  On B200 that kernel is
// bound by shared-memory bandwidth (TMA writes + UMMA operand reads ~ 125 B/clk/SM measured), and for the
// full-resolution Cout = 64 layers two thirds of that traffic is the A operand.  Here ONE TMA box brings
// the 16 x 8 output brick's input window (18 x 10 positions per depth tap plane for a 3^3 filter, 64
// channels = 128-byte rows, SWIZZLE_128B) and all kd*kh*kw taps read it in place: tcgen05.mma applies
// the 128-byte swizzle to the ABSOLUTE shared-memory address (probed on B200: tools/umma_probe.cu),
// so a descriptor whose start address is shifted by (tap offset) rows and whose 8-row group stride is
// the halo pitch (10 rows = 1280 B) addresses exactly the rows TMA wrote.  A-operand smem writes drop
// from 16 KB to ~2.6 KB per tap.
//
// The MMA-issuing thread needs ~430 clk per loop iteration (mbarrier wait, election, descriptors, commit),
// more than the tensor time of four N <= 128 MMAs, so one B stage carries the weights of a whole kw row of
// taps (3 taps = 12 MMAs per iteration) whenever shared memory allows.
//
// The A ring holds depth planes of the halo window (18 x 10 rows = 23 KB each, used by the kh*kw taps of one
// depth tap), three planes deep, so the next chunk's planes stream in while the current ones are consumed.
//
// Roles: warp 0 = A (halo) producer, warp 1 = MMA issuer, warps 2..5 = epilogue, warp 6 = B (weight) producer, warps 7..10 = a
// second epilogue group (352 threads; with the fused input transform these are transform warps instead, 480 threads).
// K order of the packed weights: source -> 64-channel chunk -> tap.
//
// Two epilogue groups (round 2): a tile with a short reduction (the 2 x 2 x 2 taps of the folded upsample: 4096 tensor-core
// clocks) used to wait for its four epilogue warps (~9000 clocks per 128 x 128 tile with statistics, ncu: tensor pipe 47 %);
// warps w and w + 5 share a TMEM lane quadrant (w % 4) and take half of the tile's columns each.
//
// Stationary weights (round 2): when the weight ring holds exactly the B stages of ONE tile (SB == stages per tile, one
// N tile), stage i carries the same weights for every tile of the CTA: they are loaded for the first tile only.  The
// upsample phases streamed 128 KB of weights per CTA and tile from L2 next to 80 KB of input windows.
#include <cstdlib>

#include "halo_common.cuh"

namespace gg {

constexpr int H_ACC_COLS = 256;
// Template parameter XW = transform warps: 0 = no fused input GroupNorm (warps 7..10 are a second epilogue group, 352 threads);
// 8 = warps 7..14 transform, one epilogue group (3-D layers: three depth taps re-normalise every window, long reductions hide
// the epilogue); 4 = warps 7..10 transform, warps 11..14 a second epilogue group (2-D layers: one window per chunk, short
// reductions).  Measured: 8 vs 4 transform warps with one epilogue group (tools/run_ad.sh): config 2 41.7 / 42.1 ms, config 4
// 10.02 / 9.84 ms; 8 vs 4 + second epilogue group (tools/run_ae.sh): config 3 4.44 / 4.48, config 4 10.41 / 10.38.  Default 8.
constexpr int H_THREADS_XF = H_THREADS + 256;        // XW = 4 or 8
constexpr int H_THREADS_E2 = H_THREADS + 128;        // XW = 0
constexpr int H_MAX_SB = 8;
constexpr int H_MAX_SA = 4;

struct HaloSeg {
    int nchunks;
    int kd, kh, kw;        // taps
    int od, oh, ow;        // input offset of tap 0
    int dshift;
    int pitch;             // bw + kw - 1 rows between consecutive h lines of the halo brick
    int plane;             // (bh + kh - 1) * pitch rows between depth planes
    uint32_t a_bytes;      // bytes of ONE depth plane of the halo window (= one A stage load)
    int g;                 // taps (along kw) whose weights share one B stage: kw or 1
    int C;                 // channels
    int inv_pitch;         // ceil(2^16 / pitch)
    const float* ss;       // XFORM: (scale, shift) of channel 0, sample 0 of this source; nullptr = the source is used as it is
};

struct alignas(64) HaloParams {
    CUtensorMap amap[H_MAX_SEGS];
    CUtensorMap wmap;
    HaloSeg seg[H_MAX_SEGS];
    int nseg, BN, SA, SB;
    uint32_t a_stage_bytes, b_stage_bytes, b_tap_bytes;
    int No, Do, Ho, Wo;
    int th, tw;                    // tiles along h, w (d and n are 1 per tile)
    int n_tiles_n, total_tiles;
    int w_stat;                    // stationary weights: SB == B stages per tile, loaded once per CTA
    int Cout8;
    const float* bias;
    const float* emb;
    int emb_stride;
    const __nv_bfloat16* residual;
    int res_stride;
    void* y;
    long long y_sn, y_sd, y_sh, y_sw;
    int y_is_f32;
    float* gn_partial;             // fused GroupNorm statistics (BN == Cout8 == 64): [N, gn_nchunks_total, 64, 2]
    int gn_chunk_base, gn_nchunks_total;
    int ss_stride, xf_silu, z_lo, z_hi;      // XFORM (fused GroupNorm + SiLU on the input planes, as conv_roll.cu)
    int Hi, Wi;                    // input extents (= output extents: stride 1)
};

template <int G, bool STATS, bool PAIR, int XW>
__global__ void __launch_bounds__(XW > 0 ? H_THREADS_XF : H_THREADS_E2, 1) conv_halo_kernel(const __grid_constant__ HaloParams p) {
    constexpr bool XFORM = XW > 0;
    constexpr int H_XW = XW > 0 ? XW : 1;          // transform warps
    constexpr int EGROUPS = XW == 8 ? 1 : 2;       // epilogue warp groups
    constexpr int EG1_WARP = XW == 4 ? 11 : 7;     // first warp of the second group
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    const int SA = p.SA, SB = p.SB, BN = p.BN;
    uint8_t* smem_b = smem + (size_t)SA * p.a_stage_bytes;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(smem_b + (size_t)SB * p.b_stage_bytes);
    uint64_t* a_empty = a_full + H_MAX_SA;
    uint64_t* b_full = a_empty + H_MAX_SA;
    uint64_t* b_empty = b_full + H_MAX_SB;
    uint64_t* tfull = b_empty + H_MAX_SB;
    uint64_t* tempty = tfull + 2;
    uint64_t* a_ready = tempty + 2;                // XFORM (PAIR: the leader's): plane landed AND transformed in every CTA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_ready + H_MAX_SA);
    float* bvec = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(a_full) + 512);      // [BN] bias + emb[n]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = PAIR ? (int)cluster_ctarank() : 0;
    // persistent schedule: in PAIR mode a "tile" is a pair of w-adjacent bricks and the two CTAs of a cluster walk
    // the same tile sequence (p.tw = pairs along w); CTA `rank` owns brick 2 iw + rank
    const int tile0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, tstep = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    pdl_launch_dependents();
    if (warp == 0 && lane == 0) {
        for (int i = 0; i < p.nseg; ++i) prefetch_tmap(&p.amap[i]);
        prefetch_tmap(&p.wmap);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int i = 0; i < SA; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); mbar_init(&a_ready[i], PAIR ? 2 * H_XW : H_XW); }
            for (int i = 0; i < SB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
            for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], (PAIR ? 8 : 4) * EGROUPS); }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        if constexpr (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (PAIR) cluster_sync_all();       // the peer's barriers are initialised before anything targets them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();                     // the prologue above touched only shared memory / TMEM / kernel parameters (common.cuh)

    if (warp == 0) {
        // ================================================================ A (halo brick) producer
        int sa = 0;
        uint32_t pha = 0;
        for (int tile = tile0; tile < p.total_tiles; tile += tstep) {
            int mt = tile / p.n_tiles_n;
            const int iw = (PAIR ? 2 : 1) * (mt % p.tw) + rank; mt /= p.tw;
            const int ih = mt % p.th; mt /= p.th;
            const int d0 = mt % p.Do, n0 = mt / p.Do;
            const int h0 = ih * H_BH, w0 = iw * H_BW;
            for (int s = 0; s < p.nseg; ++s) {
                const HaloSeg sg = p.seg[s];
                for (int j = 0; j < sg.nchunks; ++j)
                    for (int a = 0; a < sg.kd; ++a) {
                        mbar_wait(&a_empty[sa], pha ^ 1u);
                        if (elect_one()) {
                            if constexpr (XFORM) {       // each CTA's transform warps watch their own plane land
                                mbar_expect_tx(&a_full[sa], sg.a_bytes);
                                tma_load_5d(smem + (size_t)sa * p.a_stage_bytes, &p.amap[s], &a_full[sa], j * BK, w0 + sg.ow, h0 + sg.oh,
                                            d0 + sg.od + sg.dshift + a, n0);
                            } else if constexpr (PAIR) {
                                if (rank == 0) mbar_expect_tx(&a_full[sa], 2u * sg.a_bytes);
                                tma_load_5d_pair(smem + (size_t)sa * p.a_stage_bytes, &p.amap[s], leader_addr(&a_full[sa]), j * BK,
                                                 w0 + sg.ow, h0 + sg.oh, d0 + sg.od + sg.dshift + a, n0);
                            } else {
                                mbar_expect_tx(&a_full[sa], sg.a_bytes);
                                tma_load_5d(smem + (size_t)sa * p.a_stage_bytes, &p.amap[s], &a_full[sa], j * BK, w0 + sg.ow, h0 + sg.oh,
                                            d0 + sg.od + sg.dshift + a, n0);
                            }
                        }
                        __syncwarp();
                        if (++sa == SA) { sa = 0; pha ^= 1u; }
                    }
            }
        }
    } else if (warp == 6) {
        // ================================================================ B (weights) producer
        int sb = 0;
        uint32_t phb = 0;
        for (int tile = tile0; tile < p.total_tiles; tile += tstep) {
            if (p.w_stat && tile != tile0) break;       // stationary weights: every stage already holds what this tile needs
            const int nt = tile % p.n_tiles_n;
            int kb = 0;
            for (int s = 0; s < p.nseg; ++s) {
                const HaloSeg sg = p.seg[s];
                const int nk = sg.nchunks * sg.kd * sg.kh * (sg.kw / sg.g);
                for (int i = 0; i < nk; ++i) {
                    mbar_wait(&b_empty[sb], phb ^ 1u);
                    if (elect_one()) {
                        if constexpr (PAIR) {       // this CTA's half of the weight rows
                            if (rank == 0) mbar_expect_tx(&b_full[sb], 2u * p.b_tap_bytes * (uint32_t)sg.g);
                            const uint32_t bar = leader_addr(&b_full[sb]);
                            for (int t = 0; t < sg.g; ++t)
                                tma_load_2d_pair(smem_b + (size_t)sb * p.b_stage_bytes + (size_t)t * p.b_tap_bytes, &p.wmap, bar,
                                                 (kb + t) * BK, nt * BN + rank * (BN >> 1));
                        } else {
                            mbar_expect_tx(&b_full[sb], p.b_tap_bytes * (uint32_t)sg.g);
                            for (int t = 0; t < sg.g; ++t)
                                tma_load_2d(smem_b + (size_t)sb * p.b_stage_bytes + (size_t)t * p.b_tap_bytes, &p.wmap, &b_full[sb],
                                            (kb + t) * BK, nt * BN);
                        }
                    }
                    __syncwarp();
                    kb += sg.g;
                    if (++sb == SB) { sb = 0; phb ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================================================ MMA issuer (PAIR: the leader CTA only)
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) |
                               ((uint32_t)((PAIR ? 2 * BM : BM) >> 4) << 24);
        const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem_b);
        int sa = 0, sb = 0;
        uint32_t pha = 0, phb = 0, acc = 0, acc_phase = 0;
        const bool w_stat = p.w_stat != 0;
        for (int tile = tile0; tile < (rank == 0 ? p.total_tiles : 0); tile += tstep) {
            const bool b_wait = !w_stat || tile == tile0;       // stationary weights: landed during the first tile, never released
            mbar_wait(&tempty[acc], acc_phase ^ 1u);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * H_ACC_COLS;
            uint32_t first = 1;
            for (int s = 0; s < p.nseg; ++s) {
                const HaloSeg sg = p.seg[s];
                // descriptor templates: only the 14-bit start-address field changes per tap / K step
                const uint64_t a_tmpl = make_sw128_desc_sbo(0, (uint32_t)sg.pitch * 128u);
                const uint64_t b_tmpl = make_sw128_desc_sbo(0, 1024u);
                const uint32_t b_tap16 = p.b_tap_bytes >> 4;
                for (int j = 0; j < sg.nchunks; ++j) {
                    const bool last_chunk = (s == p.nseg - 1) && (j == sg.nchunks - 1);
                    for (int a = 0; a < sg.kd; ++a) {
                        mbar_wait(XFORM ? &a_ready[sa] : &a_full[sa], pha);
                        if constexpr (XFORM) tc_fence_after();
                        const uint32_t a_stage16 = (a_base + (uint32_t)sa * p.a_stage_bytes) >> 4;
                        for (int b = 0; b < sg.kh; ++b) {
                            const uint32_t row16 = a_stage16 + (uint32_t)(b * sg.pitch) * 8u;      // 128 B rows -> 8 x 16 B
                            if (G > 1 && sg.g == G) {
                                if (b_wait) mbar_wait(&b_full[sb], phb);
                                tc_fence_after();
                                if (elect_one()) {
                                    const uint32_t b16 = (b_base + (uint32_t)sb * p.b_stage_bytes) >> 4;
#pragma unroll
                                    for (int t = 0; t < G; ++t) {
                                        const uint64_t ad = a_tmpl | (uint64_t)(row16 + 8u * t), bd = b_tmpl | (uint64_t)(b16 + b_tap16 * t);
                                        umma_bf16_t<PAIR>(d_tmem, ad, bd, idesc, (t == 0 && first) ? 0u : 1u);
                                        umma_bf16_t<PAIR>(d_tmem, ad + 2, bd + 2, idesc, 1u);
                                        umma_bf16_t<PAIR>(d_tmem, ad + 4, bd + 4, idesc, 1u);
                                        umma_bf16_t<PAIR>(d_tmem, ad + 6, bd + 6, idesc, 1u);
                                    }
                                    if (!w_stat) umma_commit_t<PAIR>(&b_empty[sb]);
                                    if (b == sg.kh - 1) {
                                        umma_commit_t<PAIR>(&a_empty[sa]);
                                        if (last_chunk && a == sg.kd - 1) umma_commit_t<PAIR>(&tfull[acc]);
                                    }
                                }
                                __syncwarp();
                                first = 0;
                                if (++sb == SB) { sb = 0; phb ^= 1u; }
                            } else {
                                for (int c = 0; c < sg.kw; ++c) {
                                    if (b_wait) mbar_wait(&b_full[sb], phb);
                                    tc_fence_after();
                                    if (elect_one()) {
                                        const uint64_t ad = a_tmpl | (uint64_t)(row16 + 8u * c);
                                        const uint64_t bd = b_tmpl | (uint64_t)((b_base + (uint32_t)sb * p.b_stage_bytes) >> 4);
                                        umma_bf16_t<PAIR>(d_tmem, ad, bd, idesc, first ? 0u : 1u);
                                        umma_bf16_t<PAIR>(d_tmem, ad + 2, bd + 2, idesc, 1u);
                                        umma_bf16_t<PAIR>(d_tmem, ad + 4, bd + 4, idesc, 1u);
                                        umma_bf16_t<PAIR>(d_tmem, ad + 6, bd + 6, idesc, 1u);
                                        if (!w_stat) umma_commit_t<PAIR>(&b_empty[sb]);
                                        if (b == sg.kh - 1 && c == sg.kw - 1) {
                                            umma_commit_t<PAIR>(&a_empty[sa]);
                                            if (last_chunk && a == sg.kd - 1) umma_commit_t<PAIR>(&tfull[acc]);
                                        }
                                    }
                                    __syncwarp();
                                    first = 0;
                                    if (++sb == SB) { sb = 0; phb ^= 1u; }
                                }
                            }
                        }
                        if (++sa == SA) { sa = 0; pha ^= 1u; }
                    }
                }
            }
            acc ^= 1u;
            if (acc == 0) acc_phase ^= 1u;
        }
    } else if (XFORM && warp >= 7 && warp < 7 + H_XW) {
        // ================================================================ transform (warps 7..10): GroupNorm (+ SiLU) in place on
        // every landed halo plane of a normalised source (same arithmetic and rounding points as gg_gn_apply: xf_pair).  Rows
        // are 128 B = 64 channels, SWIZZLE_128B: the 16-byte chunk at physical slot jp of stage row r holds channels
        // 8 (jp ^ (r & 7)) .. + 7.  Rows outside the tensor stay zero (the reference pads the NORMALISED tensor); the
        // (scale, shift) pairs of a thread's eight channels come straight from gg_gn_finalize's table (L1 / L2 hits).
        const int xt = (int)threadIdx.x - H_THREADS;        // 0..32 H_XW - 1
        constexpr int XR = 4 * H_XW, XB = 3;                // rows per pass of all transform threads, rows per thread and batch
        const int jl = xt & 7;
        int sa = 0;
        uint32_t pha = 0;
        const bool silu = p.xf_silu != 0;
        const float kk = silu ? 0.5f : 1.f;
        const uint32_t ready_r = PAIR ? leader_addr(&a_ready[0]) : 0u;
        for (int tile = tile0; tile < p.total_tiles; tile += tstep) {
            int mt = tile / p.n_tiles_n;
            const int iw = (PAIR ? 2 : 1) * (mt % p.tw) + rank; mt /= p.tw;
            const int ih = mt % p.th; mt /= p.th;
            const int d0 = mt % p.Do, n0 = mt / p.Do;
            const int h0 = ih * H_BH, w0 = iw * H_BW;
            for (int s = 0; s < p.nseg; ++s) {
                const HaloSeg sg = p.seg[s];
                for (int j = 0; j < sg.nchunks; ++j) {
                    float4 q0 = make_float4(0.f, 0.f, 0.f, 0.f), q1 = q0, q2 = q0, q3 = q0;
                    if (sg.ss != nullptr) {      // (s0, b0, s1, b1) per channel pair -> (s0, s1, b0, b1), halved when SiLU follows
                        const int c0 = j * BK + 8 * jl;
                        const float4* src = reinterpret_cast<const float4*>(sg.ss + (long long)n0 * p.ss_stride + 2 * c0);
                        float4 v[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) v[e] = (c0 + 2 * e < sg.C) ? __ldg(src + e) : make_float4(0.f, 0.f, 0.f, 0.f);
                        q0 = make_float4(kk * v[0].x, kk * v[0].z, kk * v[0].y, kk * v[0].w);
                        q1 = make_float4(kk * v[1].x, kk * v[1].z, kk * v[1].y, kk * v[1].w);
                        q2 = make_float4(kk * v[2].x, kk * v[2].z, kk * v[2].y, kk * v[2].w);
                        q3 = make_float4(kk * v[3].x, kk * v[3].z, kk * v[3].y, kk * v[3].w);
                    }
                    for (int a = 0; a < sg.kd; ++a) {
                        mbar_wait(&a_full[sa], pha);
                        const int z = d0 + sg.od + a;              // local depth plane (without the slab shift)
                        if (sg.ss != nullptr && z >= p.z_lo && z < p.z_hi) {
                            uint8_t* stg = smem + (size_t)sa * p.a_stage_bytes;
                            const int nrow = sg.plane;
#pragma unroll 1
                            for (int r0 = xt >> 3; r0 < nrow; r0 += XR * XB) {
                                uint4 v[XB];
                                uint4* ptr[XB];
                                bool ok[XB];
#pragma unroll
                                for (int u = 0; u < XB; ++u) {
                                    const int r = r0 + XR * u, rc = min(r, nrow - 1);
                                    const int hh = (int)(((uint32_t)rc * (uint32_t)sg.inv_pitch) >> 16), ww = rc - hh * sg.pitch;
                                    const int gh = h0 + sg.oh + hh, gw = w0 + sg.ow + ww;
                                    ok[u] = r < nrow && gh >= 0 && gh < p.Hi && gw >= 0 && gw < p.Wi;
                                    ptr[u] = reinterpret_cast<uint4*>(stg + rc * 128 + ((jl ^ (rc & 7)) << 4));
                                    v[u] = *ptr[u];
                                }
#pragma unroll
                                for (int u = 0; u < XB; ++u) {
                                    v[u].x = xf_pair(v[u].x, q0, silu); v[u].y = xf_pair(v[u].y, q1, silu);
                                    v[u].z = xf_pair(v[u].z, q2, silu); v[u].w = xf_pair(v[u].w, q3, silu);
                                }
#pragma unroll
                                for (int u = 0; u < XB; ++u)
                                    if (ok[u]) *ptr[u] = v[u];
                            }
                            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic writes -> tensor-core reads
                        }
                        __syncwarp();
                        if (lane == 0) {
                            if constexpr (PAIR) mbar_arrive_remote(ready_r + (uint32_t)sa * 8u);
                            else mbar_arrive(&a_ready[sa]);
                        }
                        if (++sa == SA) { sa = 0; pha ^= 1u; }
                    }
                }
            }
        }
    } else {
        // ================================================================ epilogue (warps 2..5; with two groups also warps 7..10
        // (XW = 0) or 11..14 (XW = 4): warps of the same TMEM lane quadrant (warp % 4) split the tile's columns at a multiple of 32)
        constexpr int ET = 128 * EGROUPS;              // epilogue threads
        const int eg = (EGROUPS == 2 && warp >= EG1_WARP) ? 1 : 0;       // column group
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int rh = row >> 3, rw = row & 7;
        const int etid = eg ? (int)threadIdx.x - 32 * EG1_WARP + 128 : (int)threadIdx.x - 64;       // 0..ET-1 among the epilogue warps
        const int c_split = EGROUPS == 1 ? BN : min(BN, (BN / 2 + 31) / 32 * 32);
        const int c_lo = eg ? c_split : 0, c_n = eg ? BN - c_split : c_split;      // this warp's columns [c_lo, c_lo + c_n) of the tile
        uint32_t acc = 0, acc_phase = 0;
        int cur_n = -1, cur_nt = -1;
        const uint32_t tempty_r0 = PAIR ? leader_addr(&tempty[0]) : 0u, tempty_r1 = PAIR ? leader_addr(&tempty[1]) : 0u;
        for (int tile = tile0; tile < p.total_tiles; tile += tstep) {
            const int nt = tile % p.n_tiles_n;
            int mt = tile / p.n_tiles_n;
            const int iw = (PAIR ? 2 : 1) * (mt % p.tw) + rank; mt /= p.tw;
            const int ih = mt % p.th; mt /= p.th;
            const int d = mt % p.Do, n = mt / p.Do;
            const int h = ih * H_BH + rh, w = iw * H_BW + rw;
            const bool valid = h < p.Ho && w < p.Wo;
            if (n != cur_n || nt != cur_nt) {          // uniform over the four epilogue warps (same tile sequence)
                asm volatile("bar.sync 1, %0;" ::"n"(ET) : "memory");      // everyone is done with the previous vector
                for (int c = etid; c < BN; c += ET) {
                    const int ch = nt * BN + c;
                    float v = 0.f;
                    if (ch < p.Cout8) {
                        if (p.bias) v += __ldg(p.bias + ch);
                        if (p.emb) v += __ldg(p.emb + (long long)n * p.emb_stride + ch);
                    }
                    bvec[c] = v;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(ET) : "memory");
                cur_n = n; cur_nt = nt;
            }
            const long long yoff = (long long)n * p.y_sn + (long long)d * p.y_sd + (long long)h * p.y_sh + (long long)w * p.y_sw;
            const long long lin = (((long long)n * p.Do + d) * p.Ho + h) * p.Wo + w;
            const int ncols = max(0, min(c_n, p.Cout8 - nt * BN - c_lo));       // real output columns among this warp's c_n
            const int ch0 = nt * BN + c_lo;
            const __nv_bfloat16* res_row = p.residual ? p.residual + lin * p.res_stride + ch0 : nullptr;
            void* y_row = p.y_is_f32 ? static_cast<void*>(reinterpret_cast<float*>(p.y) + yoff + ch0)
                                     : static_cast<void*>(reinterpret_cast<__nv_bfloat16*>(p.y) + yoff + ch0);
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + acc * H_ACC_COLS + (uint32_t)c_lo + ((uint32_t)(q * 32) << 16);
            if (c_n > 0) {
                if constexpr (STATS) {
                    float* stat_row = p.gn_partial + (((long long)n * p.gn_nchunks_total + p.gn_chunk_base + (long long)blockIdx.x * 4 + q) *
                                                          p.Cout8 + ch0) * 2;
                    epilogue_row_stats(t_addr, c_n, ncols, bvec + c_lo, res_row, reinterpret_cast<__nv_bfloat16*>(y_row), valid, stat_row, lane);
                } else {
                    epilogue_row(t_addr, c_n, ncols, bvec + c_lo, res_row, y_row, p.y_is_f32, valid);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (PAIR) mbar_arrive_remote(acc ? tempty_r1 : tempty_r0);
                else mbar_arrive(&tempty[acc]);
            }
            acc ^= 1u;
            if (acc == 0) acc_phase ^= 1u;
        }
    }

    tc_fence_before();
    __syncthreads();
    if constexpr (PAIR) cluster_sync_all();       // the leader's MMAs / multicast commits no longer touch the peer
    if (warp == 1) {
        tc_fence_after();
        if constexpr (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

template <int G, bool STATS, bool PAIR, int XW>
static int launch_halo(const HaloParams& p, int grid, size_t smem, cudaStream_t stream) {
    constexpr bool XFORM = XW > 0;
    auto* fn = conv_halo_kernel<G, STATS, PAIR, XW>;
    static bool attr_set = false;      // one flag per instantiation
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, H_SMEM_BUDGET);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(XFORM ? H_THREADS_XF : H_THREADS_E2); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = PAIR ? 2 : 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, fn, p);
    if (e != cudaSuccess) return (int)e;
    return launch_result();
}
template <bool STATS, bool PAIR>
static int launch_halo_g(int G, int xw, const HaloParams& p, int grid, size_t smem, cudaStream_t stream) {
    if (xw == 8) {
        if (G == 3) return launch_halo<3, STATS, PAIR, 8>(p, grid, smem, stream);
        if (G == 2) return launch_halo<2, STATS, PAIR, 8>(p, grid, smem, stream);
        return launch_halo<1, STATS, PAIR, 8>(p, grid, smem, stream);
    }
    if (xw == 4) {
        if (G == 3) return launch_halo<3, STATS, PAIR, 4>(p, grid, smem, stream);
        if (G == 2) return launch_halo<2, STATS, PAIR, 4>(p, grid, smem, stream);
        return launch_halo<1, STATS, PAIR, 4>(p, grid, smem, stream);
    }
    if (G == 3) return launch_halo<3, STATS, PAIR, 0>(p, grid, smem, stream);
    if (G == 2) return launch_halo<2, STATS, PAIR, 0>(p, grid, smem, stream);
    return launch_halo<1, STATS, PAIR, 0>(p, grid, smem, stream);
}

// ---------------------------------------------------------------------------------------- host
// launch geometry shared with gg_conv_stats_chunks: CTA pairs (algo 3 forces, algo 2 forbids) take two w-adjacent
// bricks per cluster; an odd brick count along w would idle one CTA of the last pair, so auto mode wants it even
int conv_halo_grid(const gg_conv_args* a, bool* pair_out) {
    const int BN = a->block_n > 0 ? a->block_n : gg_conv_pick_block_n(a->Cout);
    const int th = (a->Ho + H_BH - 1) / H_BH, tw = (a->Wo + H_BW - 1) / H_BW;
    bool pair = a->algo == 3 || (a->algo == 1 && tw % 2 == 0);
    static const int knob_pair = [] { const char* e = getenv("GG_HALO_PAIR"); return e ? atoi(e) : 1; }();      // tuning knob, read once
    if (a->algo == 1) pair = pair && knob_pair != 0;
    const int64_t tiles = (int64_t)a->N * a->Do * th * (pair ? (tw + 1) / 2 : tw) * ((a->Cout + BN - 1) / BN);
    if (pair_out) *pair_out = pair;
    return pair ? 2 * (int)std::min<int64_t>(tiles, num_sms() / 2) : (int)std::min<int64_t>(tiles, num_sms());
}

int conv_halo_fwd(const gg_conv_args* a, cudaStream_t stream) {
    GG_REQUIRE(a->stride == 1, GG_ERR_UNSUPPORTED);
    GG_REQUIRE(a->nsrc >= 1 && a->nsrc <= H_MAX_SEGS, GG_ERR_BAD_ARG);
    if (!encode_fn()) return GG_ERR_DRIVER;
    HaloParams p;
    memset(&p, 0, sizeof(p));
    const int BN = a->block_n > 0 ? a->block_n : gg_conv_pick_block_n(a->Cout);
    GG_REQUIRE(BN % 16 == 0 && BN >= 16 && BN <= 256, GG_ERR_UNSUPPORTED);
    p.BN = BN;
    p.No = a->N; p.Do = a->Do; p.Ho = a->Ho; p.Wo = a->Wo;
    p.th = (a->Ho + H_BH - 1) / H_BH; p.tw = (a->Wo + H_BW - 1) / H_BW;
    bool pair = false;
    const int grid = conv_halo_grid(a, &pair);
    GG_REQUIRE(!pair || BN % 16 == 0, GG_ERR_UNSUPPORTED);
    if (pair) p.tw = (p.tw + 1) / 2;
    p.n_tiles_n = (a->Cout + BN - 1) / BN;
    const int64_t total = (int64_t)a->N * a->Do * p.th * p.tw * p.n_tiles_n;
    GG_REQUIRE(total < (1ll << 31), GG_ERR_UNSUPPORTED);
    p.total_tiles = (int)total;
    p.Cout8 = (a->Cout + 7) / 8 * 8;
    const int64_t W = a->W, H = a->H, D = a->D, N = a->N;
    uint32_t max_a = 0;
    int num_kb = 0;
    bool xform = false;
    for (int s = 0; s < a->nsrc; ++s) {
        const gg_conv_src& src = a->src[s];
        GG_REQUIRE(src.x != nullptr && src.C > 0 && src.C % 8 == 0, GG_ERR_BAD_ARG);
        GG_REQUIRE(aligned(src.x, 16), GG_ERR_ALIGNMENT);
        HaloSeg& sg = p.seg[s];
        sg.nchunks = (src.C + BK - 1) / BK;
        sg.dshift = src.d_shift;
        if (src.centre_only) { sg.kd = sg.kh = sg.kw = 1; sg.od = sg.oh = sg.ow = 0; }
        else { sg.kd = a->kd; sg.kh = a->kh; sg.kw = a->kw; sg.od = a->od; sg.oh = a->oh; sg.ow = a->ow; }
        sg.pitch = H_BW + sg.kw - 1;
        sg.inv_pitch = (65536 + sg.pitch - 1) / sg.pitch;
        sg.C = src.C;
        sg.ss = a->src_ss[s];
        if (sg.ss != nullptr) {
            GG_REQUIRE(aligned(sg.ss, 16) && a->ss_stride % 4 == 0 && !src.centre_only, GG_ERR_ALIGNMENT);
            xform = true;
        }
        sg.plane = (H_BH + sg.kh - 1) * sg.pitch;
        sg.a_bytes = (uint32_t)sg.plane * 128u;
        max_a = std::max(max_a, sg.a_bytes);
        const int64_t C = src.C;
        const int64_t dim[4] = {W, H, D, N};
        const int64_t str[4] = {C, W * C, H * W * C, D * H * W * C};
        const int box[4] = {sg.pitch, H_BH + sg.kh - 1, 1, 1};
        if (!encode_act_map(&p.amap[s], src.x, src.C, dim, str, box)) return GG_ERR_DRIVER;
        num_kb += sg.nchunks * sg.kd * sg.kh * sg.kw;
    }
    p.nseg = a->nsrc;
    p.ss_stride = a->ss_stride; p.xf_silu = a->xf_silu; p.z_lo = a->xf_z_lo; p.z_hi = a->xf_z_hi; p.Hi = a->H; p.Wi = a->W;
    const int b_rows = pair ? BN / 2 : BN;          // weight rows per CTA and tap
    if (!encode_w_map(&p.wmap, a->w_packed, (int64_t)num_kb * BK, a->Cout, b_rows)) return GG_ERR_DRIVER;
    p.a_stage_bytes = (max_a + 1023u) & ~1023u;
    p.b_tap_bytes = (uint32_t)b_rows * 128u;
    const int bar_bytes = 512 + 1024;        // barriers + the epilogue's [BN] additive vector
    const int avail = H_SMEM_BUDGET - 1024 - bar_bytes;
    int kwmax = 1;
    for (int s = 0; s < a->nsrc; ++s) kwmax = std::max(kwmax, p.seg[s].kw);
    // a B stage holds a whole kw row of taps when three A planes and >= 3 such B stages still fit
    int G = kwmax;
    static const int knob_maxg = [] { const char* e = getenv("GG_HALO_MAXG"); return e ? std::max(1, atoi(e)) : 3; }();     // tuning knob, read once
    G = std::min(G, knob_maxg);
    int SA = 3;
    int SB = (avail - SA * (int)p.a_stage_bytes) / (G * (int)p.b_tap_bytes);
    if (SB < 3) {
        G = 1;
        SB = (avail - SA * (int)p.a_stage_bytes) / (int)p.b_tap_bytes;
    }
    GG_REQUIRE(SB >= 2, GG_ERR_UNSUPPORTED);
    if (SB > H_MAX_SB) {            // spare shared memory: deepen the A ring instead
        SA = std::min(H_MAX_SA, (avail - H_MAX_SB * G * (int)p.b_tap_bytes) / (int)p.a_stage_bytes);
        SB = H_MAX_SB;
    }
    p.b_stage_bytes = (uint32_t)G * p.b_tap_bytes;
    for (int s = 0; s < a->nsrc; ++s) p.seg[s].g = (G > 1 && p.seg[s].kw == G) ? G : 1;
    {   // stationary weights: the ring holds exactly one tile's B stages (and the CTA keeps them for all its tiles)
        int nst = 0;
        for (int s = 0; s < a->nsrc; ++s) nst += p.seg[s].nchunks * p.seg[s].kd * p.seg[s].kh * (p.seg[s].kw / p.seg[s].g);
        static const bool allow = [] { const char* e = getenv("GG_HALO_WSTAT"); return e == nullptr || atoi(e) != 0; }();
        if (allow && p.n_tiles_n == 1 && nst <= H_MAX_SB && p.total_tiles > grid / (pair ? 2 : 1) &&
            nst * (int)p.b_stage_bytes + 3 * (int)p.a_stage_bytes <= avail) {
            SB = nst;
            SA = std::min(H_MAX_SA, (avail - nst * (int)p.b_stage_bytes) / (int)p.a_stage_bytes);
            p.w_stat = 1;
        }
    }
    p.SA = SA; p.SB = SB;
    const size_t smem = (size_t)p.SA * p.a_stage_bytes + (size_t)p.SB * p.b_stage_bytes + bar_bytes + 1024;
    // transform warps of the fused input GroupNorm: eight (one epilogue group); GG_HALO_XW=4 selects four + a second epilogue
    // group (see the XW template parameter) -- measured equal within the box-to-box noise on the 2-D networks
    // (config 3: 4.44 / 4.48 ms, config 4: 10.41 / 10.38 ms, tools/run_ae.sh), so the simpler shape stays the default
    static const int knob_xw = [] { const char* e = getenv("GG_HALO_XW"); return e ? atoi(e) : 0; }();
    const int xw = !xform ? 0 : (knob_xw == 4 ? 4 : 8);

    p.bias = a->bias; p.emb = a->emb; p.emb_stride = a->emb_stride;
    p.residual = reinterpret_cast<const __nv_bfloat16*>(a->residual); p.res_stride = a->res_stride;
    p.y = a->y; p.y_sn = a->y_sn; p.y_sd = a->y_sd; p.y_sh = a->y_sh; p.y_sw = a->y_sw; p.y_is_f32 = a->y_is_f32;

    if (a->gn_partial != nullptr) {
        // fused GroupNorm statistics: one partial row per (CTA, epilogue warp) and sample, accumulated tile by tile by its
        // owner lanes (plain read-modify-write: deterministic)
        // after a per-tile shuffle reduction, so the rows are cleared first (a memset node under graph capture)
        GG_REQUIRE(!a->y_is_f32, GG_ERR_UNSUPPORTED);
        GG_REQUIRE(aligned(a->gn_partial, 16) && a->gn_chunk_base >= 0 && a->gn_chunk_base + grid * 4 <= a->gn_nchunks_total, GG_ERR_BAD_ARG);
        p.gn_partial = a->gn_partial; p.gn_chunk_base = a->gn_chunk_base; p.gn_nchunks_total = a->gn_nchunks_total;
        {   // one 2-D memset: this launch's rows of every sample
            const size_t row_bytes = (size_t)(p.Cout8 * 2) * sizeof(float);
            cudaError_t e = cudaMemset2DAsync(a->gn_partial + (size_t)a->gn_chunk_base * (p.Cout8 * 2), (size_t)a->gn_nchunks_total * row_bytes, 0,
                                              (size_t)(grid * 4) * row_bytes, (size_t)a->N, stream);
            if (e != cudaSuccess) return (int)e;
        }
        return pair ? launch_halo_g<true, true>(G, xw, p, grid, smem, stream) : launch_halo_g<true, false>(G, xw, p, grid, smem, stream);
    }
    return pair ? launch_halo_g<false, true>(G, xw, p, grid, smem, stream) : launch_halo_g<false, false>(G, xw, p, grid, smem, stream);
}

}  // namespace gg
