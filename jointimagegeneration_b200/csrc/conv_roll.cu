// conv_roll.cu -- "depth-rolling" 3x3x3 convolution for narrow outputs (3 * Cout <= 256), CTA pairs only
//
// With Cout = 64 the halo-brick kernel issues 128 x 64 x 16 MMAs: every MMA re-reads its 4 KB A slice from
// shared memory to fill only 64 accumulator columns, and the A reads alone (128 B/clk) are the SM's whole
// shared-memory bandwidth.  This kernel walks the INPUT depth planes of a brick column instead of the output
// planes.  Input plane z feeds output planes z+1, z, z-1 through the depth taps kd = 0, 1, 2, so the three
// tap-weight tiles are stacked along N and ONE MMA stream (N = 3 Cout) updates three accumulator slots:
//
//     step t (input plane z = d0 - 1 + t):   slot (t+1-kd) mod 3  +=  X[z] (shifted by kh, kw) * W[kd, kh, kw]
//
// Each halo plane is loaded once (not three times) and each A slice read from shared memory feeds 3x the
// columns.  After step t the slot of output plane z-1 is complete: the epilogue stores it, ZEROES the slot
// (tcgen05.st; every MMA accumulates) and hands it back; it becomes the kd = 0 target of step t+1.  So that
// the tensor core never waits for that hand-over, a CTA interleaves two bricks (h-adjacent, `wi` = 0/1) with
// their own slot triples: X(t), Y(t), X(t+1), ...  A depth segment of L output planes costs L + 2 steps (the
// first/last step feed one real output plane each); the host picks the segment length that balances that
// overhead against the number of work items per CTA pair.
//
// CTA pair (cta_group::2, M = 256): two w-adjacent bricks per cluster, each CTA holds its own planes and half
// of the stacked weight rows (N/2 = 1.5 slots).  Sources flagged centre_only (the fused 1x1x1 skip) are
// N = Cout MMAs into the slot of output plane z.  K order of the packed weights: as conv_halo.cu (algo 1).
//
// Roles (224 threads): warp 0 = A (plane) producer, warp 6 = B (weights) producer, warp 1 = MMA issuer
// (leader CTA), warps 2..5 = epilogue.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "halo_common.cuh"
#include "cat_row.cuh"

namespace gg {

// GG_ROLL_TIMING (build define, with GG_ROLL_DBG=1 at run time): GG_CLK() around the waits of the MMA issuer and the transform
// warps.  Off in release builds: the issuer is a single thread whose instruction stream IS the critical path of the kernel
// (ncu, skip-source layer: 65 % of its time goes to issuing, 30 % to waiting for weights), six clock reads per step do not belong there.
#ifdef GG_ROLL_TIMING
#define GG_CLK() clock64()
#else
#define GG_CLK() 0ll
#endif

constexpr int R_MAX_SA = 6, R_MAX_SB = 8;
constexpr int R_THREADS = 352;      // warps 0, 1, 6 = producers / MMA; warps 2..5 drain brick 0, warps 7..10 drain brick 1
// XFORM adds warps 11..: GroupNorm/SiLU transform of the landed halo planes (RollThreads below)
// Transform warps.  Wide outputs (3 x 64 stacked columns) spend ~3500 clocks of MMAs per plane and four warps keep up;
// the 16-column head conv (64 -> 12 classes) spends ~900, so its planes wait for the transform: eight warps there.
template <int XWARPS> struct RollThreads { static constexpr int value = R_THREADS + 32 * XWARPS; };       // XWARPS = 0: no transform
constexpr int R_STAGE_BYTES = 16384; // 128 rows x 128 B output / residual staging tile
constexpr int R_SS_BYTES = 4096;    // (scale, shift) table: up to 512 channels over all sources

// Tuning knobs (environment), read ONCE at first use: the launch path itself never calls getenv.
struct RollKnobs {
    int no_tma_epi, g, sa, dbg, skip_first, xw, cw;
    RollKnobs() {
        auto geti = [](const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; };
        no_tma_epi = getenv("GG_ROLL_NO_TMA_EPI") != nullptr;
        g = geti("GG_ROLL_G", 0);
        sa = geti("GG_ROLL_SA", 0);
        dbg = getenv("GG_ROLL_DBG") != nullptr;
        skip_first = geti("GG_ROLL_SKIP_FIRST", 0);
        xw = geti("GG_ROLL_XW", 0);
        cw = geti("GG_ROLL_CW", 1);
    }
};
static const RollKnobs& roll_knobs() { static const RollKnobs k; return k; }

struct RollSeg {
    int nchunks, centre;
    int kh, kw, oh, ow, dshift;
    int pitch, inv_pitch;      // rows per h line of the halo plane; ceil(2^16 / pitch)
    uint32_t a_bytes;
    int g;           // kw taps per B stage (kw or 1)
    int kb_base;     // first 64-wide K block of this source in the packed weights
    int taps;        // K blocks per chunk: 3 kh kw, or 1 (centre_only)
    int C;           // channels
    int cw_idx0;     // centre_only sources with stationary weights: index of this source's first tile in the resident region
    int ss_off;      // XFORM: first entry of this source in the shared (scale, shift) table, -1 = used as is
    const float* ss; // XFORM: (scale, shift) of channel 0, sample 0
};

struct alignas(64) RollParams {
    CUtensorMap amap[H_MAX_SEGS];
    CUtensorMap wmap;              // box = 64 k x (BNs / 2) weight rows: one "unit" = half a slot
    CUtensorMap ymap, rmap;        // tma_epi: output / residual bricks (64 ch x 8 w x 16 h), SWIZZLE_128B
    int tma_epi;                   // bf16 64-channel outputs leave (and residuals arrive) through a staging tile + TMA
    RollSeg seg[H_MAX_SEGS];
    int nseg, BNs, SA, SB;
    int cw_tiles;                  // stationary 1x1x1 weights: tiles (one per 64-channel chunk of the centre_only sources) resident in smem, 0 = off
    int order[H_MAX_SEGS];         // sequence in which a step walks the sources (1x1x1 sources first, see conv_roll_fwd)
    uint32_t a_stage_bytes, b_stage_bytes, b_unit_bytes;
    int No, Do, Ho, Wo;
    int thp, twp, nsd, Lseg, total_items;      // brick pairs along h / w, depth segments and their length
    int Cout8;
    const float* bias;
    const float* emb;
    int emb_stride;
    const __nv_bfloat16* residual;
    int res_stride;
    void* y;
    long long y_sn, y_sd, y_sh, y_sw;
    int y_is_f32;
    float* gn_partial;
    int gn_chunk_base, gn_nchunks_total;
    int ss_stride, xf_silu, z_lo, z_hi, ss_entries;      // XFORM (fused GroupNorm + SiLU on the input planes)
    unsigned long long* dbg;       // tuning aid (GG_ROLL_DBG=1): per leader CTA, clocks the MMA issuer spent in each wait
    // sampler epilogue (gg_conv_args.cat, 16-column head conv only): the accumulator row of a voxel holds all its class
    // logits, so softmax + posterior + clamp + draw + next-input row happen here and no logits tensor is written
    int cat_on, cat_C, cat_ncond, cat_cin_pad;
    const uint8_t* cat_lab_in;
    uint8_t* cat_lab_out;
    uint16_t* cat_next_x;
    const uint16_t* cat_cond;
    const float* cat_coef;
    float cat_clamp;
    unsigned long long cat_seed, cat_offset;
    long long cat_vox_base;
};

__device__ __forceinline__ void tmem_zero16(uint32_t taddr) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
        ::"r"(taddr), "r"(0u) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tma_store_5d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
                 ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

// accumulator slot -> registers (the slot is handed back to the tensor core BEFORE the row is finished and stored)
template <int BNS>
__device__ __forceinline__ void load_slot(uint32_t t_addr, uint32_t (&r)[BNS]) {
#pragma unroll
    for (int c = 0; c < BNS; c += 16) tmem_ld16_nowait(t_addr + c, r + c);
    tmem_ld_wait();
}
// bias / embedding / residual / rounding / store of one output row from registers.
template <int BNS, bool STATS>
__device__ __forceinline__ void finish_row(uint32_t (&r)[BNS], const uint4 (&rr)[BNS / 8], int ncols, const float* __restrict__ bvec,
                                           void* y_row, int y_is_f32, bool valid, bool to_stage = false, int swz = 0) {
#pragma unroll
    for (int g = 0; g < BNS / 8; ++g) {
        const float4 b0 = *reinterpret_cast<const float4*>(bvec + 8 * g), b1 = *reinterpret_cast<const float4*>(bvec + 8 * g + 4);
        float v[8];
        v[0] = __uint_as_float(r[8 * g + 0]) + b0.x + bf16_lo(rr[g].x); v[1] = __uint_as_float(r[8 * g + 1]) + b0.y + bf16_hi(rr[g].x);
        v[2] = __uint_as_float(r[8 * g + 2]) + b0.z + bf16_lo(rr[g].y); v[3] = __uint_as_float(r[8 * g + 3]) + b0.w + bf16_hi(rr[g].y);
        v[4] = __uint_as_float(r[8 * g + 4]) + b1.x + bf16_lo(rr[g].z); v[5] = __uint_as_float(r[8 * g + 5]) + b1.y + bf16_hi(rr[g].z);
        v[6] = __uint_as_float(r[8 * g + 6]) + b1.z + bf16_lo(rr[g].w); v[7] = __uint_as_float(r[8 * g + 7]) + b1.w + bf16_hi(rr[g].w);
        if (y_is_f32) {
            if (valid && 8 * g < ncols) {
                float* yp = reinterpret_cast<float*>(y_row) + 8 * g;
                *reinterpret_cast<float4*>(yp) = make_float4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<float4*>(yp + 4) = make_float4(v[4], v[5], v[6], v[7]);
            }
        } else {
            const uint4 pk = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
            // to_stage: the row goes to the swizzled staging tile (every row, TMA clips), else straight to global memory.
            // STATS: the column sums are taken from the staged tile, so rows outside the tensor are staged as zeros
            if (to_stage) *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(y_row) + 8 * (g ^ swz)) =
                              (!STATS || valid) ? pk : make_uint4(0u, 0u, 0u, 0u);
            else if (valid && 8 * g < ncols) *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(y_row) + 8 * g) = pk;
        }
    }
}
// The same for a row that lives in the swizzled staging tile (bf16, 64 columns): the residual chunk is read from the tile
// where TMA put it and the result overwrites it in place -- no residual registers (the 64 accumulator values are the
// only large live range, so the kernel fits 96 registers when more transform warps share the SM).
// STATS: rows outside the tensor are staged as zeros (the column sums are taken from the tile).
template <bool STATS>
__device__ __forceinline__ void finish_row_staged(const uint32_t (&r)[64], bool has_res, const float* __restrict__ bvec, uint8_t* srow, int swz,
                                                  bool valid) {
#pragma unroll
    for (int g = 0; g < 8; ++g) {
        uint4* slot = reinterpret_cast<uint4*>(srow + ((g ^ swz) << 4));
        const uint4 rr = has_res ? *slot : make_uint4(0u, 0u, 0u, 0u);
        const float4 b0 = *reinterpret_cast<const float4*>(bvec + 8 * g), b1 = *reinterpret_cast<const float4*>(bvec + 8 * g + 4);
        float v[8];
        v[0] = __uint_as_float(r[8 * g + 0]) + b0.x + bf16_lo(rr.x); v[1] = __uint_as_float(r[8 * g + 1]) + b0.y + bf16_hi(rr.x);
        v[2] = __uint_as_float(r[8 * g + 2]) + b0.z + bf16_lo(rr.y); v[3] = __uint_as_float(r[8 * g + 3]) + b0.w + bf16_hi(rr.y);
        v[4] = __uint_as_float(r[8 * g + 4]) + b1.x + bf16_lo(rr.z); v[5] = __uint_as_float(r[8 * g + 5]) + b1.y + bf16_hi(rr.z);
        v[6] = __uint_as_float(r[8 * g + 6]) + b1.z + bf16_lo(rr.w); v[7] = __uint_as_float(r[8 * g + 7]) + b1.w + bf16_hi(rr.w);
        const uint4 pk = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
        *slot = (!STATS || valid) ? pk : make_uint4(0u, 0u, 0u, 0u);
    }
}

struct RollItem { int n, d0, L, ihp, iwp; };
__device__ __forceinline__ RollItem roll_item(const RollParams& p, int item) {
    RollItem it;
    it.iwp = item % p.twp; item /= p.twp;
    it.ihp = item % p.thp; item /= p.thp;
    const int sd = item % p.nsd;
    it.n = item / p.nsd;
    it.d0 = sd * p.Lseg;
    it.L = min(p.Do, it.d0 + p.Lseg) - it.d0;
    return it;
}

template <int G, bool STATS, int BNS, int XW>
__global__ void __launch_bounds__(RollThreads<XW>::value, 1) conv_roll_kernel(const __grid_constant__ RollParams p) {
    constexpr bool XFORM = XW > 0;
    constexpr int XT = 32 * (XW > 0 ? XW : 1);        // transform threads
    // rows a transform thread handles per batch: a plane has 180 rows, a pass of all transform threads covers XT / 8
    // of them -> 4 warps: 3 batches of 4 x 16 rows; 5 warps: 3 batches of 3 x 20 (exact); 8 warps: 2 batches of 3 x 32
    constexpr int XB = XW == 4 ? 4 : 3;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    const int SA = p.SA, SB = p.SB;
    constexpr int BNs = BNS;
    uint8_t* smem_b = smem + (size_t)SA * p.a_stage_bytes;
    uint8_t* smem_cw = smem_b + (size_t)SB * p.b_stage_bytes;          // stationary 1x1x1 weights: cw_tiles x b_unit_bytes
    uint8_t* smem_c = smem_cw + (size_t)p.cw_tiles * p.b_unit_bytes;    // tma_epi: one 16 KB staging tile per epilogue group
    uint64_t* a_full = reinterpret_cast<uint64_t*>(smem_c + (p.tma_epi ? 2 * R_STAGE_BYTES : 0));
    uint64_t* a_empty = a_full + R_MAX_SA;
    uint64_t* b_full = a_empty + R_MAX_SA;
    uint64_t* b_empty = b_full + R_MAX_SB;
    uint64_t* step_done = b_empty + R_MAX_SB;       // [wi]: the MMAs of one step of brick wi have completed
    uint64_t* slot_free = step_done + 2;            // [wi] (leader's): finished slot of brick wi drained and zeroed
    uint64_t* a_ready = slot_free + 2;              // XFORM (leader's): plane landed AND transformed in both CTAs
    uint64_t* res_full = a_ready + R_MAX_SA;        // [wi] tma_epi: residual brick landed in the staging tile
    uint64_t* cw_full = res_full + 2;               // (leader's) the stationary 1x1x1 weights of both CTAs have landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(cw_full + 1);
    // XFORM: per channel PAIR (2k, 2k+1) the quad (s_2k, s_2k+1, b_2k, b_2k+1), halved when SiLU follows: [ss_entries / 2]
    float4* ss_tab = reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(a_full) + 1024);
    float* bvec_all = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(a_full) + 512);  // [2][BNs] bias + emb[n], per epilogue group

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)cluster_ctarank();
    const int item0 = (int)(blockIdx.x >> 1), istep = (int)(gridDim.x >> 1);
    pdl_launch_dependents();
    if (warp == 0 && lane == 0) {
        for (int i = 0; i < p.nseg; ++i) prefetch_tmap(&p.amap[i]);
        prefetch_tmap(&p.wmap);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int i = 0; i < SA; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); mbar_init(&a_ready[i], 2 * (XW > 0 ? XW : 1)); }
            for (int i = 0; i < SB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
            for (int i = 0; i < 2; ++i) { mbar_init(&step_done[i], 1); mbar_init(&slot_free[i], 8); mbar_init(&res_full[i], 1); }
            mbar_init(cw_full, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();                     // the prologue above touched only shared memory / TMEM / kernel parameters (common.cuh)

    if (warp == 0) {
        // ================================================================ A (halo plane) producer
        int sa = 0;
        uint32_t pha = 0;
        for (int item = item0; item < p.total_items; item += istep) {
            const RollItem it = roll_item(p, item);
            const int w0 = (2 * it.iwp + rank) * H_BW;
            for (int t = 0; t < it.L + 2; ++t)
                for (int wi = 0; wi < 2; ++wi) {
                    const int h0 = (2 * it.ihp + wi) * H_BH, z = it.d0 - 1 + t;
                    for (int si = 0; si < p.nseg; ++si) {
                        const int s = p.order[si];
                        const RollSeg sg = p.seg[s];
                        if (sg.centre && (t == 0 || t == it.L + 1)) continue;     // would only feed planes outside the segment
                        for (int j = 0; j < sg.nchunks; ++j) {
                            mbar_wait(&a_empty[sa], pha ^ 1u);
                            if (elect_one()) {
                                if constexpr (XFORM) {      // each CTA's transform warps watch their own plane land
                                    mbar_expect_tx(&a_full[sa], sg.a_bytes);
                                    tma_load_5d(smem + (size_t)sa * p.a_stage_bytes, &p.amap[s], &a_full[sa], j * BK, w0 + sg.ow,
                                                h0 + sg.oh, z + sg.dshift, it.n);
                                    // the transform is one more hop between HBM and the tensor core and the plane ring is
                                    // only four deep: pull the plane of two steps ahead into L2 now
                                    if (t + 2 < it.L + 2)
                                        asm volatile("cp.async.bulk.prefetch.tensor.5d.L2.global.tile [%0, {%1, %2, %3, %4, %5}];"
                                                     ::"l"(&p.amap[s]), "r"(j * BK), "r"(w0 + sg.ow), "r"(h0 + sg.oh), "r"(z + 2 + sg.dshift), "r"(it.n)
                                                     : "memory");
                                } else {
                                    if (rank == 0) mbar_expect_tx(&a_full[sa], 2u * sg.a_bytes);
                                    tma_load_5d_pair(smem + (size_t)sa * p.a_stage_bytes, &p.amap[s], leader_addr(&a_full[sa]), j * BK,
                                                     w0 + sg.ow, h0 + sg.oh, z + sg.dshift, it.n);
                                }
                            }
                            __syncwarp();
                            if (++sa == SA) { sa = 0; pha ^= 1u; }
                        }
                    }
                }
        }
    } else if (warp == 6) {
        // ================================================================ B (stacked weights) producer
        int sb = 0;
        uint32_t phb = 0;
        const int half_rows = BNs >> 1;
        if (p.cw_tiles > 0 && item0 < p.total_items) {
            // the 1x1x1 (skip) weights are the same in every step: one resident tile per 64-channel chunk, loaded once, so
            // that they never take a slot of the weight ring (a two-stage ring next to five plane stages is all that fits)
            if (elect_one()) {
                if (rank == 0) mbar_expect_tx(cw_full, 2u * (uint32_t)p.cw_tiles * p.b_unit_bytes);
                const uint32_t bar = leader_addr(cw_full);
                for (int s = 0; s < p.nseg; ++s) {
                    const RollSeg sg = p.seg[s];
                    if (!sg.centre) continue;
                    for (int j = 0; j < sg.nchunks; ++j)
                        tma_load_2d_pair(smem_cw + (size_t)(sg.cw_idx0 + j) * p.b_unit_bytes, &p.wmap, bar, (sg.kb_base + j * sg.taps) * BK,
                                         rank * half_rows);
                }
            }
            __syncwarp();
        }
        for (int item = item0; item < p.total_items; item += istep) {
            const RollItem it = roll_item(p, item);
            for (int t = 0; t < it.L + 2; ++t)
                for (int wi = 0; wi < 2; ++wi)
                    for (int si = 0; si < p.nseg; ++si) {
                        const int s = p.order[si];
                        const RollSeg sg = p.seg[s];
                        if (sg.centre && (p.cw_tiles > 0 || t == 0 || t == it.L + 1)) continue;
                        for (int j = 0; j < sg.nchunks; ++j) {
                            const int kb0 = sg.kb_base + j * sg.taps;
                            if (sg.centre) {
                                mbar_wait(&b_empty[sb], phb ^ 1u);
                                if (elect_one()) {
                                    if (rank == 0) mbar_expect_tx(&b_full[sb], 2u * p.b_unit_bytes);
                                    tma_load_2d_pair(smem_b + (size_t)sb * p.b_stage_bytes, &p.wmap, leader_addr(&b_full[sb]), kb0 * BK,
                                                     rank * half_rows);
                                }
                                __syncwarp();
                                if (++sb == SB) { sb = 0; phb ^= 1u; }
                                continue;
                            }
                            const int khw = sg.kh * sg.kw;
                            for (int b = 0; b < sg.kh; ++b)
                                for (int cg = 0; cg < sg.kw / sg.g; ++cg) {
                                    mbar_wait(&b_empty[sb], phb ^ 1u);
                                    if (elect_one()) {
                                        if (rank == 0) mbar_expect_tx(&b_full[sb], 2u * 3u * p.b_unit_bytes * (uint32_t)sg.g);
                                        const uint32_t bar = leader_addr(&b_full[sb]);
                                        uint8_t* dst = smem_b + (size_t)sb * p.b_stage_bytes;
                                        for (int tt = 0; tt < sg.g; ++tt)
                                            for (int ul = 0; ul < 3; ++ul) {
                                                // this CTA's rows of the stacked tile: units 3 rank .. 3 rank + 2 of six half-slots
                                                const int u = 3 * rank + ul, slot = u >> 1, half = u & 1;
                                                const int kd = (t + 4 - slot) % 3;
                                                const int kb = kb0 + kd * khw + b * sg.kw + cg * sg.g + tt;
                                                tma_load_2d_pair(dst + (size_t)(tt * 3 + ul) * p.b_unit_bytes, &p.wmap, bar, kb * BK,
                                                                 half * half_rows);
                                            }
                                    }
                                    __syncwarp();
                                    if (++sb == SB) { sb = 0; phb ^= 1u; }
                                }
                        }
                    }
        }
    } else if (warp == 1) {
        // ================================================================ MMA issuer (leader CTA only)
        if (rank == 0) {
            const uint32_t idesc_full = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((3 * BNs) >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
            const uint32_t idesc_one = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BNs >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
            const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem_b), cw_base = smem_u32(smem_cw);
            const bool cw_on = p.cw_tiles > 0;
            bool cw_ready = false;
            const uint64_t b_tmpl = make_sw128_desc_sbo(0, 1024u);
            const uint32_t tap16 = (3u * p.b_unit_bytes) >> 4;
            int sa = 0, sb = 0;
            uint32_t pha = 0, phb = 0, phf[2] = {0, 0};
            long long w_slot = 0, w_a = 0, w_b = 0, t_begin = GG_CLK(), tq;
            for (int item = item0; item < p.total_items; item += istep) {
                const RollItem it = roll_item(p, item);
                for (int t = 0; t < it.L + 2; ++t)
                    for (int wi = 0; wi < 2; ++wi) {
                        tq = GG_CLK();
                        mbar_wait(&slot_free[wi], phf[wi]);
                        w_slot += GG_CLK() - tq;
                        phf[wi] ^= 1u;
                        tc_fence_after();
                        const uint32_t d_tmem = tmem_base + (uint32_t)(wi * 3 * BNs);
                        for (int si = 0; si < p.nseg; ++si) {
                            const int s = p.order[si];
                            const RollSeg sg = p.seg[s];
                            if (sg.centre && (t == 0 || t == it.L + 1)) continue;
                            const uint64_t a_tmpl = make_sw128_desc_sbo(0, (uint32_t)sg.pitch * 128u);
                            for (int j = 0; j < sg.nchunks; ++j) {
                                tq = GG_CLK();
                                mbar_wait(XFORM ? &a_ready[sa] : &a_full[sa], pha);
                                w_a += GG_CLK() - tq;
                                if constexpr (XFORM) tc_fence_after();
                                const uint32_t a_stage16 = (a_base + (uint32_t)sa * p.a_stage_bytes) >> 4;
                                if (sg.centre) {
                                    if (cw_on) {
                                        if (!cw_ready) { mbar_wait(cw_full, 0u); cw_ready = true; }
                                    } else {
                                        mbar_wait(&b_full[sb], phb);
                                    }
                                    tc_fence_after();
                                    if (elect_one()) {
                                        const uint64_t ad = a_tmpl | (uint64_t)a_stage16;
                                        const uint64_t bd = b_tmpl | (uint64_t)((cw_on ? cw_base + (uint32_t)(sg.cw_idx0 + j) * p.b_unit_bytes
                                                                                       : b_base + (uint32_t)sb * p.b_stage_bytes) >> 4);
                                        const uint32_t dc = d_tmem + (uint32_t)((t % 3) * BNs);      // slot of output plane z
#pragma unroll
                                        for (int k = 0; k < 4; ++k) umma_bf16_t<true>(dc, ad + 2 * k, bd + 2 * k, idesc_one, 1u);
                                        if (!cw_on) umma_commit_t<true>(&b_empty[sb]);
                                        umma_commit_t<true>(&a_empty[sa]);
                                    }
                                    __syncwarp();
                                    if (!cw_on && ++sb == SB) { sb = 0; phb ^= 1u; }
                                } else {
                                    for (int b = 0; b < sg.kh; ++b) {
                                        const uint32_t row16 = a_stage16 + (uint32_t)(b * sg.pitch) * 8u;
                                        for (int cg = 0; cg < sg.kw / sg.g; ++cg) {
                                            tq = GG_CLK();
                                            mbar_wait(&b_full[sb], phb);
                                            w_b += GG_CLK() - tq;
                                            tc_fence_after();
                                            if (elect_one()) {
                                                const uint32_t b16 = (b_base + (uint32_t)sb * p.b_stage_bytes) >> 4;
                                                if (G > 1 && sg.g == G) {
#pragma unroll
                                                    for (int tt = 0; tt < G; ++tt) {
                                                        const uint64_t ad = a_tmpl | (uint64_t)(row16 + 8u * tt), bd = b_tmpl | (uint64_t)(b16 + tap16 * tt);
#pragma unroll
                                                        for (int k = 0; k < 4; ++k) umma_bf16_t<true>(d_tmem, ad + 2 * k, bd + 2 * k, idesc_full, 1u);
                                                    }
                                                } else {
                                                    const uint64_t ad = a_tmpl | (uint64_t)(row16 + 8u * cg), bd = b_tmpl | (uint64_t)b16;
#pragma unroll
                                                    for (int k = 0; k < 4; ++k) umma_bf16_t<true>(d_tmem, ad + 2 * k, bd + 2 * k, idesc_full, 1u);
                                                }
                                                umma_commit_t<true>(&b_empty[sb]);
                                                if (b == sg.kh - 1 && cg == sg.kw / sg.g - 1) umma_commit_t<true>(&a_empty[sa]);
                                            }
                                            __syncwarp();
                                            if (++sb == SB) { sb = 0; phb ^= 1u; }
                                        }
                                    }
                                }
                                if (++sa == SA) { sa = 0; pha ^= 1u; }
                            }
                        }
                        if (elect_one()) umma_commit_t<true>(&step_done[wi]);
                        __syncwarp();
                    }
            }
            if (p.dbg != nullptr && elect_one()) {
                unsigned long long* o = p.dbg + (size_t)(blockIdx.x >> 1) * 4;
                o[0] = (unsigned long long)(GG_CLK() - t_begin); o[1] = (unsigned long long)w_slot;
                o[2] = (unsigned long long)w_a; o[3] = (unsigned long long)w_b;
            }
        }
    } else if (XFORM && warp >= 11) {
        // ================================================================ transform (warps 11..14): GroupNorm (+ SiLU) in place
        // on each landed halo plane.  Rows are 128 B = 64 channels, SWIZZLE_128B: the 16-byte chunk at physical
        // slot jp of stage row r holds channels 8 (jp ^ (r & 7)) .. +7.  Rows outside the tensor stay zero.
        const int xt = (int)threadIdx.x - R_THREADS;        // 0..XT-1
        int sa = 0, cur_n = -1;
        uint32_t pha = 0;
        const bool silu = p.xf_silu != 0;
        long long x_wait = 0, x_begin = GG_CLK(), xq;
        for (int item = item0; item < p.total_items; item += istep) {
            const RollItem it = roll_item(p, item);
            if (it.n != cur_n) {        // (scale, shift) of this sample; uniform over the four warps
                asm volatile("bar.sync 3, %0;" ::"n"(XT) : "memory");
                for (int si = 0; si < p.nseg; ++si) {
                    const int s = p.order[si];
                    const RollSeg sg = p.seg[s];
                    if (sg.ss_off < 0) continue;
                    const float4* src = reinterpret_cast<const float4*>(sg.ss + (long long)it.n * p.ss_stride);
                    const float k = silu ? 0.5f : 1.f;
                    for (int c = xt; c < sg.nchunks * (BK / 2); c += XT) {
                        const float4 v = 2 * c < sg.C ? __ldg(src + c) : make_float4(0.f, 0.f, 0.f, 0.f);      // (s0, b0, s1, b1)
                        ss_tab[(sg.ss_off >> 1) + c] = make_float4(k * v.x, k * v.z, k * v.y, k * v.w);
                    }
                }
                asm volatile("bar.sync 3, %0;" ::"n"(XT) : "memory");
                cur_n = it.n;
            }
            const int w0 = (2 * it.iwp + rank) * H_BW;
            for (int t = 0; t < it.L + 2; ++t)
                for (int wi = 0; wi < 2; ++wi) {
                    const int h0 = (2 * it.ihp + wi) * H_BH, z = it.d0 - 1 + t;
                    for (int si = 0; si < p.nseg; ++si) {
                        const int s = p.order[si];
                        const RollSeg sg = p.seg[s];
                        if (sg.centre && (t == 0 || t == it.L + 1)) continue;
                        for (int j = 0; j < sg.nchunks; ++j) {
                            xq = GG_CLK();
                            mbar_wait(&a_full[sa], pha);
                            x_wait += GG_CLK() - xq;
                            if (sg.ss_off >= 0 && z >= p.z_lo && z < p.z_hi) {
                                uint8_t* stg = smem + (size_t)sa * p.a_stage_bytes;
                                const int nrow = (H_BH + sg.kh - 1) * sg.pitch;
                                // thread -> (logical 8-channel group jl, row group): its four channel-pair quads stay in
                                // registers for the whole plane; 8 consecutive threads cover one 128-byte row
                                const int jl = xt & 7;
                                const float4* tb = ss_tab + ((sg.ss_off + j * BK) >> 1) + 4 * jl;
                                const float4 q0 = tb[0], q1 = tb[1], q2 = tb[2], q3 = tb[3];
                                // XB rows per batch, loads first: a lone warp per scheduler needs the ILP (a per-row
                                // branch would serialise LDS -> FMA -> MUFU -> FMA -> STS chains)
#pragma unroll 1
                                for (int r0 = xt >> 3; r0 < nrow; r0 += (XT / 8) * XB) {
                                    uint4 v[XB];
                                    uint4* ptr[XB];
                                    bool ok[XB];
#pragma unroll
                                    for (int u = 0; u < XB; ++u) {
                                        const int r = r0 + (XT / 8) * u, rc = min(r, nrow - 1);
                                        const int hh = (int)(((uint32_t)rc * (uint32_t)sg.inv_pitch) >> 16), ww = rc - hh * sg.pitch;    // rc < 2^10
                                        const int gh = h0 + sg.oh + hh, gw = w0 + sg.ow + ww;
                                        ok[u] = r < nrow && gh >= 0 && gh < p.Ho && gw >= 0 && gw < p.Wo;      // stride 1: input extent = output extent
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
                            if (lane == 0) mbar_arrive_remote(leader_addr(&a_ready[sa]));
                            if (++sa == SA) { sa = 0; pha ^= 1u; }
                        }
                    }
                }
        }
        if (p.dbg != nullptr && rank == 0 && xt == 0) {
            unsigned long long* o = p.dbg + 2048 + (size_t)(blockIdx.x >> 1) * 2;
            o[0] = (unsigned long long)(GG_CLK() - x_begin); o[1] = (unsigned long long)x_wait;
        }
    } else {
        // ================================================================ epilogue: warps 2..5 drain brick 0, warps 7..10 brick 1
        const int wi = warp >= 7 ? 1 : 0;
        const int q = warp & 3;                        // TMEM lane quadrant this warp may access
        const int row = q * 32 + lane;
        const int rh = row >> 3, rw = row & 7;
        const int etid = ((warp - (wi ? 7 : 2)) << 5) + lane;       // 0..127 within the group
        float* bvec = bvec_all + wi * BNs;
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
        const uint32_t free_r = leader_addr(&slot_free[wi]);
        // every MMA accumulates: the three slots of this group's brick start as zeros
        for (int c = 0; c < 3 * BNs; c += 16) tmem_zero16(lane_base + (uint32_t)(wi * 3 * BNs + c));
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(free_r);
        uint32_t phd = 0, ph_res = 0;
        int cur_n = -1;
        // STATS: thread (q, lane) owns column (q & 1) * 32 + lane over rows (q >> 1) * 64 .. + 63 of every staged brick:
        // (sum, sum of squares) of the STORED (bf16-rounded) values, accumulated in a fixed order
        float2 st = make_float2(0.f, 0.f);
        const int st_col = (q & 1) * 32 + lane, st_row0 = (q >> 1) * 64;
        int stat_n = -1;
        CatRowCoef cat_cf = {};
        auto flush = [&](int n) {        // one partial row per (CTA, epilogue warp), its 32 columns (the rest stays zero): once per sample
            *reinterpret_cast<float2*>(p.gn_partial + (((long long)n * p.gn_nchunks_total + p.gn_chunk_base + (long long)blockIdx.x * 8 + wi * 4 + q) * 64) * 2 +
                                       2 * st_col) = st;
            st = make_float2(0.f, 0.f);
        };
        for (int item = item0; item < p.total_items; item += istep) {
            const RollItem it = roll_item(p, item);
            const int n = it.n;
            if (n != cur_n) {          // uniform over the four warps of the group (same item sequence)
                asm volatile("bar.sync %0, 128;" ::"r"(1 + wi) : "memory");
                for (int c = etid; c < BNs; c += 128) {
                    float v = 0.f;
                    if (c < p.Cout8) {
                        if (p.bias) v += __ldg(p.bias + c);
                        if (p.emb) v += __ldg(p.emb + (long long)n * p.emb_stride + c);
                    }
                    bvec[c] = v;
                }
                asm volatile("bar.sync %0, 128;" ::"r"(1 + wi) : "memory");
                cur_n = n;
                if constexpr (BNS == 16) {
                    if (p.cat_on) cat_cf = cat_row_coef(__ldg(p.cat_coef + 2 * n), __ldg(p.cat_coef + 2 * n + 1), p.cat_C);
                }
            }
            if constexpr (STATS) {
                if (n != stat_n) {
                    if (stat_n >= 0) flush(stat_n);
                    stat_n = n;
                }
            }
            const int w = (2 * it.iwp + rank) * H_BW + rw;
            for (int t = 0; t < it.L + 2; ++t) {
                {
                    const int slot = (t + 2) % 3, d = it.d0 - 2 + t;
                    const uint32_t t_addr = lane_base + (uint32_t)((wi * 3 + slot) * BNs);
                    const bool store = d >= it.d0 && d < it.d0 + it.L;
                    const int h = (2 * it.ihp + wi) * H_BH + rh;
                    const bool valid = store && h < p.Ho && w < p.Wo;
                    const long long lin = (((long long)n * p.Do + d) * p.Ho + h) * p.Wo + w;
                    const bool tma_epi = BNS == 64 && p.tma_epi != 0;
                    uint8_t* stage = smem_c + wi * R_STAGE_BYTES;
                    const int hb = (2 * it.ihp + wi) * H_BH, wb = (2 * it.iwp + rank) * H_BW;
                    // the residual is requested before waiting for the accumulator: its latency hides behind the MMAs
                    uint4 rr[BNS == 64 ? 1 : BNS / 8];      // 64-column rows that are not staged (rare) prefetch per chunk instead
                    if (tma_epi) {
                        if (store && p.residual != nullptr) {
                            if (etid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");     // previous brick left the tile
                            asm volatile("bar.sync %0, 128;" ::"r"(1 + wi) : "memory");
                            if (etid == 0) {
                                mbar_expect_tx(&res_full[wi], R_STAGE_BYTES);
                                tma_load_5d(stage, &p.rmap, &res_full[wi], 0, wb, hb, d, n);
                            }
                        }
                    } else if constexpr (BNS != 64) {
#pragma unroll
                        for (int g = 0; g < BNS / 8; ++g)
                            rr[g] = (p.residual != nullptr && valid && 8 * g < p.Cout8) ? ldg_nc_u4(p.residual + lin * p.res_stride + 8 * g)
                                                                                        : make_uint4(0, 0, 0, 0);
                    }
                    int cat_lab = 0;
                    if constexpr (BNS == 16) {      // x_t's label of this voxel: requested before the accumulator is waited for
                        if (p.cat_on && valid) cat_lab = (int)__ldg(p.cat_lab_in + lin);
                    }
                    mbar_wait(&step_done[wi], phd);
                    phd ^= 1u;
                    tc_fence_after();
                    uint32_t r[BNS];
                    if (store) load_slot<BNS>(t_addr, r);
                    // hand the slot back first: zero it (and, at the end of the segment, the two slots holding partial
                    // sums of planes outside it); the global-memory half of the epilogue is off the MMA's critical path
                    if (t == it.L + 1) {
                        for (int c = 0; c < 3 * BNs; c += 16) tmem_zero16(lane_base + (uint32_t)(wi * 3 * BNs + c));
                    } else {
                        for (int c = 0; c < BNs; c += 16) tmem_zero16(t_addr + c);
                    }
                    tmem_st_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_remote(free_r);
                    if (store) {
                        if (tma_epi) {
                          if constexpr (BNS == 64) {
                            // row `row` of the tile is 128 bytes; SWIZZLE_128B puts its 16-byte chunk g at slot g ^ (row & 7)
                            uint8_t* srow = stage + row * 128;
                            if (p.residual != nullptr) {
                                mbar_wait(&res_full[wi], ph_res);
                                ph_res ^= 1u;
                            } else {
                                // no residual: the tile is first touched here, AFTER the slot went back to the tensor core
                                if (etid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");     // previous brick left the tile
                                asm volatile("bar.sync %0, 128;" ::"r"(1 + wi) : "memory");
                            }
                            finish_row_staged<STATS>(r, p.residual != nullptr, bvec, srow, row & 7, valid);
                            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic writes -> TMA store reads
                            asm volatile("bar.sync %0, 128;" ::"r"(1 + wi) : "memory");
                            if (etid == 0) tma_store_5d(&p.ymap, stage, 0, wb, hb, d, n);
                          }
                        } else if (BNS == 16 && p.cat_on) {
                            if constexpr (BNS == 16) {
                                // ---- sampler epilogue: this thread's 16 accumulator columns are the class logits of voxel `lin`
                                float lg[16];
#pragma unroll
                                for (int c = 0; c < 16; ++c) lg[c] = __uint_as_float(r[c]) + bvec[c] + 0.f;     // as finish_row (no residual)
                                const unsigned long long gv = (unsigned long long)lin + (unsigned long long)p.cat_vox_base;
                                const unsigned long long gq = gv >> 2;             // one Philox call per group of 4 voxels (pervoxel.cu)
                                const uint4 rr4 = philox4x32_10(make_uint4((uint32_t)gq, (uint32_t)(gq >> 32), (uint32_t)p.cat_offset,
                                                                           (uint32_t)(p.cat_offset >> 32)),
                                                                make_uint2((uint32_t)p.cat_seed, (uint32_t)(p.cat_seed >> 32)));
                                const uint32_t sel = (uint32_t)gv & 3u;
                                const uint32_t rbits = sel == 0 ? rr4.x : sel == 1 ? rr4.y : sel == 2 ? rr4.z : rr4.w;
                                const int bi = cat_row_draw(lg, p.cat_C, cat_lab, cat_cf, p.cat_clamp, rbits);
                                // labels: the 8 lanes of a brick row hold 8 w-consecutive voxels -> one 8-byte store per row
                                uint32_t pk = (uint32_t)bi;
                                pk |= __shfl_down_sync(0xffffffffu, pk, 1) << 8;
                                pk |= __shfl_down_sync(0xffffffffu, pk, 2) << 16;
                                const uint32_t hi = __shfl_down_sync(0xffffffffu, pk, 4);
                                const bool row_full = (2 * it.iwp + rank) * H_BW + H_BW <= p.Wo && (p.Wo & 7) == 0;
                                if (row_full) {
                                    if (rw == 0 && valid) *reinterpret_cast<uint2*>(p.cat_lab_out + lin) = make_uint2(pk, hi);
                                } else if (valid) {
                                    p.cat_lab_out[lin] = (uint8_t)bi;
                                }
                                if (valid && p.cat_next_x != nullptr) {
                                    // next UNet input row: one-hot(C) | condition channel(s) | zero padding, bf16
                                    uint16_t* rowp = p.cat_next_x + lin * p.cat_cin_pad;
                                    for (int c0 = 0; c0 < p.cat_cin_pad; c0 += 8) {
                                        uint32_t wv[4];
#pragma unroll
                                        for (int e = 0; e < 4; ++e) {
                                            const int c = c0 + 2 * e;
                                            uint32_t bits = (c == (bi & ~1)) ? ((bi & 1) ? 0x3F800000u : 0x3F80u) : 0u;
                                            if (c + 1 >= p.cat_C && p.cat_cond != nullptr) {
#pragma unroll
                                                for (int z = 0; z < 2; ++z) {
                                                    const int cc = c + z - p.cat_C;
                                                    if (cc >= 0 && cc < p.cat_ncond) bits |= (uint32_t)p.cat_cond[lin * p.cat_ncond + cc] << (16 * z);
                                                }
                                            }
                                            wv[e] = bits;
                                        }
                                        *reinterpret_cast<uint4*>(rowp + c0) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
                                    }
                                }
                            }
                        } else {
                            const long long yoff = (long long)n * p.y_sn + (long long)d * p.y_sd + (long long)h * p.y_sh + (long long)w * p.y_sw;
                            void* y_row = p.y_is_f32 ? static_cast<void*>(reinterpret_cast<float*>(p.y) + yoff)
                                                     : static_cast<void*>(reinterpret_cast<__nv_bfloat16*>(p.y) + yoff);
                            if constexpr (BNS == 64) {
                                uint4 rq[8];
#pragma unroll
                                for (int g = 0; g < 8; ++g)
                                    rq[g] = (p.residual != nullptr && valid && 8 * g < p.Cout8) ? ldg_nc_u4(p.residual + lin * p.res_stride + 8 * g)
                                                                                                : make_uint4(0, 0, 0, 0);
                                finish_row<64, false>(r, rq, p.Cout8, bvec, y_row, p.y_is_f32, valid);
                            } else {
                                finish_row<BNS, false>(r, rr, p.Cout8, bvec, y_row, p.y_is_f32, valid);
                            }
                        }
                        if constexpr (STATS) {
                            // column sums from the staged tile (complete after the bar.sync above; the TMA store only reads
                            // it, and the next brick is written after the group's next bar.sync): 64 two-byte loads per
                            // thread, 32 consecutive columns of one row per warp -> conflict-free, and no accumulator
                            // registers beyond the row itself
                            const uint8_t* colp = stage + ((st_col & 7) << 1);
                            const int chunk = st_col >> 3;
                            float s1 = 0.f, s2 = 0.f;
#pragma unroll 8
                            for (int i = 0; i < 64; ++i) {
                                const int rw2 = st_row0 + i;
                                const float f = __uint_as_float((uint32_t)(*reinterpret_cast<const uint16_t*>(colp + rw2 * 128 + ((chunk ^ (rw2 & 7)) << 4))) << 16);
                                s1 += f;
                                s2 = fmaf(f, f, s2);
                            }
                            st.x += s1; st.y += s2;
                        }
                    }
                }
            }
        }
        if constexpr (STATS) {
            if (stat_n >= 0) flush(stat_n);
        }
        if (BNS == 64 && p.tma_epi != 0 && etid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // bricks are in global memory
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ---------------------------------------------------------------------------------------- host
template <int G, bool STATS, int BNS, int XW>
static int launch_roll(const RollParams& p, int grid, size_t smem, cudaStream_t stream) {
    auto* fn = conv_roll_kernel<G, STATS, BNS, XW>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, H_SMEM_BUDGET);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(RollThreads<XW>::value); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, fn, p);
    if (e != cudaSuccess) return (int)e;
    return launch_result();
}

template <bool STATS, int BNS>
static int dispatch_roll(int G, int xw, const RollParams& p, int grid, size_t smem, cudaStream_t stream) {
#define GG_ROLL_CASE(g_, xw_) if (G == g_ && xw == xw_) return launch_roll<g_, STATS, BNS, xw_>(p, grid, smem, stream)
    GG_ROLL_CASE(1, 0); GG_ROLL_CASE(3, 0);
    GG_ROLL_CASE(1, 4); GG_ROLL_CASE(3, 4);
    GG_ROLL_CASE(1, 5); GG_ROLL_CASE(3, 5);
    GG_ROLL_CASE(1, 8); GG_ROLL_CASE(3, 8);
#undef GG_ROLL_CASE
    return GG_ERR_UNSUPPORTED;
}

// geometry shared with gg_conv_stats_chunks: brick pairs, depth segmentation, grid
struct RollGeom { int BNs, thp, twp, nsd, Lseg, items, grid; };
static RollGeom roll_geom(const gg_conv_args* a) {
    RollGeom g;
    g.BNs = (a->Cout + 15) / 16 * 16;
    const int th = (a->Ho + H_BH - 1) / H_BH, tw = (a->Wo + H_BW - 1) / H_BW;
    g.thp = (th + 1) / 2; g.twp = (tw + 1) / 2;
    const int npairs = std::max(1, num_sms() / 2);
    const int64_t cols = (int64_t)a->N * g.thp * g.twp;
    // a segment of L output planes costs L + 2 steps; pick the segmentation with the shortest critical path
    int64_t best = -1;
    g.nsd = 1; g.Lseg = a->Do;
    for (int nsd = 1; nsd <= a->Do; ++nsd) {
        const int L = (a->Do + nsd - 1) / nsd;
        const int real = (a->Do + L - 1) / L;
        const int64_t rounds = (cols * real + npairs - 1) / npairs;
        const int64_t cost = rounds * (L + 2);
        if (best < 0 || cost < best) { best = cost; g.nsd = real; g.Lseg = L; }
    }
    const int64_t items = cols * g.nsd;
    g.items = (int)std::min<int64_t>(items, INT32_MAX);
    g.grid = 2 * (int)std::min<int64_t>(items, npairs);
    return g;
}
int conv_roll_grid(const gg_conv_args* a) { return roll_geom(a).grid; }

int conv_roll_fwd(const gg_conv_args* a, cudaStream_t stream) {
    GG_REQUIRE(a->stride == 1 && a->dims == 3, GG_ERR_UNSUPPORTED);
    GG_REQUIRE(a->kd == 3 && a->od == -1, GG_ERR_UNSUPPORTED);
    GG_REQUIRE(a->nsrc >= 1 && a->nsrc <= H_MAX_SEGS && !a->src[0].centre_only, GG_ERR_BAD_ARG);
    if (!encode_fn()) return GG_ERR_DRIVER;
    RollParams p;
    memset(&p, 0, sizeof(p));
    const RollGeom g = roll_geom(a);
    GG_REQUIRE(g.BNs == 16 || g.BNs == 64, GG_ERR_UNSUPPORTED);      // instantiated slot widths (3 BNs <= 256)
    GG_REQUIRE((int64_t)a->N * g.thp * g.twp * g.nsd < (1ll << 31), GG_ERR_UNSUPPORTED);
    p.BNs = g.BNs;
    p.No = a->N; p.Do = a->Do; p.Ho = a->Ho; p.Wo = a->Wo;
    p.thp = g.thp; p.twp = g.twp; p.nsd = g.nsd; p.Lseg = g.Lseg; p.total_items = g.items;
    p.Cout8 = (a->Cout + 7) / 8 * 8;
    const int64_t W = a->W, H = a->H, D = a->D, N = a->N;
    uint32_t max_a = 0;
    int num_kb = 0, kwmax = 1, ss_entries = 0;
    bool xform = false;
    for (int s = 0; s < a->nsrc; ++s) {
        const gg_conv_src& src = a->src[s];
        GG_REQUIRE(src.x != nullptr && src.C > 0 && src.C % 8 == 0, GG_ERR_BAD_ARG);
        GG_REQUIRE(aligned(src.x, 16), GG_ERR_ALIGNMENT);
        RollSeg& sg = p.seg[s];
        sg.nchunks = (src.C + BK - 1) / BK;
        sg.dshift = src.d_shift;
        sg.centre = src.centre_only ? 1 : 0;
        sg.C = src.C;
        sg.ss = a->src_ss[s];
        sg.ss_off = -1;
        if (sg.ss != nullptr) {
            GG_REQUIRE(aligned(sg.ss, 16) && a->ss_stride % 4 == 0, GG_ERR_ALIGNMENT);
            sg.ss_off = ss_entries;
            ss_entries += sg.nchunks * BK;
            xform = true;
        }
        if (src.centre_only) { sg.kh = sg.kw = 1; sg.oh = sg.ow = 0; sg.taps = 1; }
        else { sg.kh = a->kh; sg.kw = a->kw; sg.oh = a->oh; sg.ow = a->ow; sg.taps = 3 * a->kh * a->kw; }
        sg.pitch = H_BW + sg.kw - 1;
        sg.inv_pitch = (65536 + sg.pitch - 1) / sg.pitch;
        sg.a_bytes = (uint32_t)((H_BH + sg.kh - 1) * sg.pitch) * 128u;
        max_a = std::max(max_a, sg.a_bytes);
        kwmax = std::max(kwmax, sg.kw);
        sg.kb_base = num_kb;
        const int64_t C = src.C;
        const int64_t dim[4] = {W, H, D, N};
        const int64_t str[4] = {C, W * C, H * W * C, D * H * W * C};
        const int box[4] = {sg.pitch, H_BH + sg.kh - 1, 1, 1};
        if (!encode_act_map(&p.amap[s], src.x, src.C, dim, str, box)) return GG_ERR_DRIVER;
        num_kb += sg.nchunks * sg.taps;
    }
    p.nseg = a->nsrc;
    if (!encode_w_map(&p.wmap, a->w_packed, (int64_t)num_kb * BK, a->Cout, g.BNs / 2)) return GG_ERR_DRIVER;
    p.a_stage_bytes = (max_a + 1023u) & ~1023u;
    p.b_unit_bytes = (uint32_t)(g.BNs / 2) * 128u;
    GG_REQUIRE(ss_entries * (int)sizeof(float2) <= R_SS_BYTES, GG_ERR_UNSUPPORTED);
    p.ss_stride = a->ss_stride; p.xf_silu = a->xf_silu; p.z_lo = a->xf_z_lo; p.z_hi = a->xf_z_hi; p.ss_entries = ss_entries;
    // bf16 64-channel outputs (and their residuals) move through a swizzled staging tile by TMA: coalesced, clipped
    // by the tensor map, and none of the epilogue's global-memory instructions left to contend with the transform
    p.tma_epi = (g.BNs == 64 && !a->y_is_f32 && p.Cout8 == 64 && aligned(a->y, 16) && a->y_sw % 8 == 0 && a->y_sh % 8 == 0 &&
                 a->y_sd % 8 == 0 && a->y_sn % 8 == 0 && (a->residual == nullptr || (aligned(a->residual, 16) && a->res_stride % 8 == 0)))
                    ? 1 : 0;
    const RollKnobs& knobs = roll_knobs();
    if (knobs.no_tma_epi) p.tma_epi = 0;
    if (p.tma_epi) {
        const int64_t dim[4] = {a->Wo, a->Ho, a->Do, a->N};
        const int box[4] = {H_BW, H_BH, 1, 1};
        const int64_t ystr[4] = {a->y_sw, a->y_sh, a->y_sd, a->y_sn};
        if (!encode_act_map(&p.ymap, a->y, p.Cout8, dim, ystr, box)) return GG_ERR_DRIVER;
        if (a->residual != nullptr) {
            const int64_t rs = a->res_stride;
            const int64_t rstr[4] = {rs, (int64_t)a->Wo * rs, (int64_t)a->Ho * a->Wo * rs, (int64_t)a->Do * a->Ho * a->Wo * rs};
            if (!encode_act_map(&p.rmap, a->residual, p.Cout8, dim, rstr, box)) return GG_ERR_DRIVER;
        }
    }
    const int bar_bytes = 1024 + (xform ? R_SS_BYTES : 0) + (p.tma_epi ? 2 * R_STAGE_BYTES : 0);   // + barriers, additive vectors, (scale, shift) table
    // stationary 1x1x1 weights (GG_ROLL_CW=0 sends them through the weight ring again): one resident tile per chunk of the centre_only sources
    p.cw_tiles = 0;
    if (knobs.cw != 0) {
        for (int s = 0; s < a->nsrc; ++s)
            if (a->src[s].centre_only) { p.seg[s].cw_idx0 = p.cw_tiles; p.cw_tiles += p.seg[s].nchunks; }
        if (p.cw_tiles > 16) p.cw_tiles = 0;
    }
    const int avail = H_SMEM_BUDGET - 1024 - bar_bytes - p.cw_tiles * (int)p.b_unit_bytes;
    const int tap_bytes = 3 * (int)p.b_unit_bytes;
    GG_REQUIRE(kwmax == 3, GG_ERR_UNSUPPORTED);
    // Ring shapes.  Default: a weight stage = one kw row of stacked tap tiles (G = 3, 12 MMAs per stage), three plane
    // stages (four when the transform adds a hop between "landed" and "usable").  With 1x1x1 extra sources every step
    // pushes 2-3 more (small) planes through the plane ring and each is held for a whole TMA latency but only four
    // MMAs: those convs take single-tap weight stages (G = 1) and spend the shared memory on a fourth plane stage.
    bool has_centre = false;
    for (int s = 0; s < a->nsrc; ++s) has_centre = has_centre || a->src[s].centre_only;
    // Ring shapes (measured on B200, N = 8 x 64 x 128 x 128, tools/bench_conv.py; profiles/r2_conv_roll_tuning.md).
    // A weight stage = one kw row of stacked tap tiles (G = 3: 12 MMAs per issue-loop iteration -- with single-tap
    // stages the loop itself, ~450 clocks per iteration, is slower than its four MMAs).  Plane stages: three, or four
    // when the transform adds a hop between "landed" and "usable".  With fused 1x1x1 skip sources every step pushes
    // 2-3 more (small) planes through the plane ring, each held for a whole TMA latency but only four MMAs: five plane
    // stages and two weight stages (2.60 -> 2.28 ms on the 64 + 192 -> 64 layer; four plane stages + eight single-tap
    // weight stages, the round-1 choice: 2.43-2.60 ms).
    int G = 3;
    if (knobs.g == 1 || knobs.g == 3) G = knobs.g;
    // transform warps (0 = no fused input normalisation): wide outputs keep up with four; the 16-column head conv spends
    // few clocks of MMAs per plane and the skip-source convs have the shallowest look-ahead -> eight; GG_ROLL_XW overrides
    int xw = xform ? ((g.BNs == 16 || has_centre) ? 8 : 4) : 0;
    if (xform && (knobs.xw == 4 || knobs.xw == 5 || knobs.xw == 8)) xw = knobs.xw;
    // order of the sources within a step: as given, or (GG_ROLL_SKIP_FIRST=1, measured slower) 1x1x1 sources first
    const bool skip_first = has_centre && knobs.skip_first != 0;
    {
        int n = 0;
        for (int pass = 0; pass < 2; ++pass)
            for (int s2 = 0; s2 < a->nsrc; ++s2) {
                const bool c = a->src[s2].centre_only != 0;
                if (skip_first ? (c == (pass == 0)) : (pass == 0)) p.order[n++] = s2;
            }
    }
    // (with stationary 1x1x1 weights the ring carries only the 3^3 taps; the resident tiles take 4 KB per chunk, i.e. half a plane stage)
    int SA = has_centre ? (p.cw_tiles > 0 ? 4 : 5) : xform ? 4 : 3;
    int SB = (avail - SA * (int)p.a_stage_bytes) / (G * tap_bytes);
    if (SB < 3 && SA > 3 && !has_centre) { SA = 3; SB = (avail - SA * (int)p.a_stage_bytes) / (G * tap_bytes); }
    if (SB < 2 && SA > 3) { SA = 4; SB = (avail - SA * (int)p.a_stage_bytes) / (G * tap_bytes); }       // no transform table / staging: smaller budget
    GG_REQUIRE(SB >= 2, GG_ERR_UNSUPPORTED);
    if (SB > R_MAX_SB) {
        SA = std::min(R_MAX_SA, (avail - R_MAX_SB * G * tap_bytes) / (int)p.a_stage_bytes);
        SB = R_MAX_SB;
    }
    if (knobs.sa > 0) {       // plane stages (weight stages take the rest)
        const int sa = knobs.sa, sb = (avail - sa * (int)p.a_stage_bytes) / (G * tap_bytes);
        if (sa >= 2 && sa <= R_MAX_SA && sb >= 2) { SA = sa; SB = std::min(sb, R_MAX_SB); }
    }
    p.b_stage_bytes = (uint32_t)(G * tap_bytes);
    for (int s = 0; s < a->nsrc; ++s) p.seg[s].g = (G > 1 && p.seg[s].kw == G) ? G : 1;
    p.SA = SA; p.SB = SB;
    const size_t smem = (size_t)p.SA * p.a_stage_bytes + (size_t)p.SB * p.b_stage_bytes + (size_t)p.cw_tiles * p.b_unit_bytes + bar_bytes + 1024;
    GG_REQUIRE(smem <= (size_t)H_SMEM_BUDGET, GG_ERR_UNSUPPORTED);

    p.bias = a->bias; p.emb = a->emb; p.emb_stride = a->emb_stride;
    p.residual = reinterpret_cast<const __nv_bfloat16*>(a->residual); p.res_stride = a->res_stride;
    p.y = a->y; p.y_sn = a->y_sn; p.y_sd = a->y_sd; p.y_sh = a->y_sh; p.y_sw = a->y_sw; p.y_is_f32 = a->y_is_f32;

    if (a->cat != nullptr) {
        const gg_cat_epilogue& c = *a->cat;
        GG_REQUIRE(g.BNs == 16 && a->gn_partial == nullptr && a->residual == nullptr, GG_ERR_UNSUPPORTED);
        GG_REQUIRE(c.labels_in && c.labels_out && c.coef && c.C >= 2 && c.C <= 16 && c.C <= a->Cout && c.mode == GG_CAT_SAMPLE, GG_ERR_BAD_ARG);
        GG_REQUIRE(aligned(c.labels_out, 8) && c.vox_base % 4 == 0 && ((int64_t)a->Do * a->Ho * a->Wo) % 4 == 0, GG_ERR_ALIGNMENT);
        if (c.next_x) GG_REQUIRE(c.Cin_pad % 8 == 0 && c.Cin_pad >= c.C + c.n_cond && aligned(c.next_x, 16), GG_ERR_BAD_ARG);
        p.cat_on = 1; p.cat_C = c.C; p.cat_ncond = c.n_cond; p.cat_cin_pad = c.Cin_pad;
        p.cat_lab_in = c.labels_in; p.cat_lab_out = c.labels_out;
        p.cat_next_x = reinterpret_cast<uint16_t*>(c.next_x); p.cat_cond = reinterpret_cast<const uint16_t*>(c.cond);
        p.cat_coef = c.coef; p.cat_clamp = c.clamp_min; p.cat_seed = c.seed; p.cat_offset = c.offset; p.cat_vox_base = c.vox_base;
    }

    static unsigned long long* dbg_buf = nullptr;
    const bool dbg = knobs.dbg != 0;
    if (dbg) {
        // the debug path allocates and synchronises: refuse it while the stream is being captured into a CUDA graph
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(stream, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return GG_ERR_UNSUPPORTED;
        if (!dbg_buf && cudaMalloc(&dbg_buf, 4096 * sizeof(unsigned long long)) != cudaSuccess) return GG_ERR_DRIVER;
        cudaMemsetAsync(dbg_buf, 0, 4096 * sizeof(unsigned long long), stream);
        p.dbg = dbg_buf;
    }
    struct DbgPrint {
        bool on; int pairs; cudaStream_t st; unsigned long long* buf;
        ~DbgPrint() {
            if (!on) return;
            cudaStreamSynchronize(st);
            std::vector<unsigned long long> h((size_t)pairs * 4);
            cudaMemcpy(h.data(), buf, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
            double s[4] = {0, 0, 0, 0};
            for (int i = 0; i < pairs; ++i) for (int k = 0; k < 4; ++k) s[k] += (double)h[(size_t)i * 4 + k] / pairs;
            fprintf(stderr, "[conv_roll] MMA issuer clocks: total %.0f, waiting slot_free %.0f, a_full %.0f, b_full %.0f (avg of %d pairs)\n",
                    s[0], s[1], s[2], s[3], pairs);
            std::vector<unsigned long long> hx((size_t)pairs * 2);
            cudaMemcpy(hx.data(), buf + 2048, hx.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
            double x[2] = {0, 0};
            for (int i = 0; i < pairs; ++i) for (int k = 0; k < 2; ++k) x[k] += (double)hx[(size_t)i * 2 + k] / pairs;
            fprintf(stderr, "[conv_roll] transform warp clocks: total %.0f, waiting for planes %.0f\n", x[0], x[1]);
        }
    } dbg_print{dbg, g.grid / 2, stream, dbg_buf};
    if (a->gn_partial != nullptr) {
        GG_REQUIRE(g.BNs == 64 && p.Cout8 == 64 && !a->y_is_f32 && p.tma_epi, GG_ERR_UNSUPPORTED);   // statistics are read from the staged tile
        GG_REQUIRE(aligned(a->gn_partial, 16) && a->gn_chunk_base >= 0 && a->gn_chunk_base + g.grid * 8 <= a->gn_nchunks_total, GG_ERR_BAD_ARG);
        p.gn_partial = a->gn_partial; p.gn_chunk_base = a->gn_chunk_base; p.gn_nchunks_total = a->gn_nchunks_total;
        {   // one 2-D memset: this launch's rows of every sample
            const size_t row_bytes = (size_t)(128) * sizeof(float);
            cudaError_t e = cudaMemset2DAsync(a->gn_partial + (size_t)a->gn_chunk_base * (128), (size_t)a->gn_nchunks_total * row_bytes, 0,
                                              (size_t)(g.grid * 8) * row_bytes, (size_t)a->N, stream);
            if (e != cudaSuccess) return (int)e;
        }
        return dispatch_roll<true, 64>(G, xw, p, g.grid, smem, stream);
    }
    if (g.BNs == 16) return dispatch_roll<false, 16>(G, xw, p, g.grid, smem, stream);
    return dispatch_roll<false, 64>(G, xw, p, g.grid, smem, stream);
}

}  // namespace gg
