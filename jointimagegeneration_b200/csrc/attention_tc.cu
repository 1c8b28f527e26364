// attention_tc.cu -- flash-style softmax(Q K^T * scale) V on the 5th-generation tensor cores
//
// Replaces QKVAttentionLegacy (ccdm/ddpm/models/unet_openai/unet.py:334-360, ldm openaimodel.py:350-379) and CrossAttention
// (ldm/modules/attention.py:152-193) for head dims 32 / 64.  The first-generation kernel (attention.cu) issues mma.sync
// from every warp and is bound by that; here the two contractions are tcgen05.mma streams issued by one thread with
// TMEM accumulators, K / V / Q arrive by TMA, and the CUDA cores are left with what only they can do: the exponentials.
//
// One CTA = one (batch, head) and TWO 128-row query tiles A and B ("ping-pong": while the eight softmax warps of one tile
// exponentiate, the tensor core computes the other tile's scores), walking the keys in tiles of 128:
//
//     S_g(j) = Q_g K_j^T                M = 128, N = 128, K = d        -> TMEM, 128 columns per tile g
//     P_g(j) = exp2(S_g(j) c - m_g)     softmax warps: TMEM -> registers -> bf16 -> shared memory (K-major, SWIZZLE_128B)
//     O_g   += P_g(j) V_j               M = 128, N = d,   K = 128      -> TMEM, d columns per tile g
//
// * Operands are the K-major SWIZZLE_128B tiles the convolution kernels use.  Q and K rows are loaded as 64-element
//   (128-byte) boxes whose channel extent is d: for d = 32 the upper half is TMA zero fill and the MMAs simply stop at
//   K = 32.  V is needed with the KEYS contiguous (B operand of P V), so a small pre-pass writes V^T [B, H, d, Tk] into a
//   caller-provided workspace (8 MB for the largest site of config 5) and its [d x 64-key] boxes are ordinary K-major tiles.
// * Online softmax with LAZY rescaling: a row keeps the maximum it last rescaled to and only when a new tile raises it by
//   more than 2^8 are l and the O accumulator (tcgen05.ld / st) rescaled; P stays <= 256, exact in fp32 / bf16 terms.
// * K_j / V_j^T are loaded once per CTA and feed both query tiles (256 queries per key tile).
//
// Roles (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..5 = softmax of tile A, warps 6..9 = tile B
// (warp w touches TMEM lanes 32 (w % 4) ..).  The exponentials bound the kernel: 16 MUFU / clock / SM.
#include <cstdlib>
#include <cstring>

#include "tc_common.cuh"

namespace gg {

// GG_ATTN_RS = 1 (tuning builds; OFF): every score row shared by TWO softmax threads (64 keys each; 16 softmax warps, four per
// scheduler, instead of 8; the halves exchange their maxima / row sums through shared memory and a 64-thread named barrier per
// key tile).  The idea was to hide the per-tile chain  TMEM load -> row maximum -> exponentials -> P store -> fence  behind more
// warps; measured SLOWER on every site (tools/run_al.sh: T = 16384 0.856 -> 1.126 ms, T = 4096 0.238 -> 0.307, T = 1024
// 0.103 -> 0.126): the kernel is bound by MUFU / FMA issue, not by latency, and the extra barrier + exchange per tile (and 96
// registers per thread at 576 threads) cost more than the added warps hide.
#ifndef GG_ATTN_RS
#define GG_ATTN_RS 0
#endif
constexpr int AT_HS = GG_ATTN_RS ? 2 : 1;          // threads per score row
constexpr int AT_NC = 128 / AT_HS;                 // score columns per thread
constexpr int AT_BM = 128, AT_BN = 128, AT_ST = 3, AT_THREADS = 64 + 256 * AT_HS;
constexpr int AT_XCH_BYTES = 2 * 2 * 2 * 128 * 4 + 2 * 2 * 128 * 4;      // maxima [parity][tile][half][row] + row sums [tile][half][row]
constexpr int AT_TILE = 16384;              // 128 rows x 128 B

struct alignas(64) AttnTcParams {
    CUtensorMap qmap, kmap, vtmap;
    __nv_bfloat16* o;
    long long o_bs;
    int o_rs, o_hs;
    int H, Tq, Tk, n_kt;
    int two;                // 1: two 128-row query tiles per CTA (ping-pong); 0: one (more CTAs when the grid would not fill the GPU)
    float scale_log2;
};

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ float ex2_fast(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// exp2 of a PAIR of scores on the FMA pipe instead of the MUFU (16 / clk / SM: the unit this kernel is bound by -- ncu: XU pipe
// 62 % of peak while active, tensor pipe 15 %).  Round to the nearest integer with the 1.5 * 2^23 trick, degree-3 polynomial of
// 2^f on f in [-1/2, 1/2] (relative error <= 6e-4, below the bf16 rounding of P: 2^-9), the integer part shifted straight into
// the exponent field.  The argument is clamped to >= -125 (masked keys are -inf): such a term is 2^-125, nothing next to the
// row maximum's 2^0 in the fp32 row sum and 0 after the bf16 rounding of P relative to it.
// Every GG_ATTN_POLY-th pair takes this path (default 3; measured, tools/run_af.sh: T = 16384 site 1.003 ms with none, 0.884
// with every 4th, 0.856 with every 3rd, 0.941 with every 2nd pair on the FMA pipe; T = 4096: 0.272 / 0.244 / 0.238 / 0.256).
__device__ __forceinline__ void ex2_poly_x2(uint64_t x, float& e0, float& e1) {
    float x0, x1;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(x));
    x0 = fmaxf(x0, -125.0f); x1 = fmaxf(x1, -125.0f);
    const uint64_t xc = [&] { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x0), "f"(x1)); return r; }();
    const uint64_t magic = [&] { uint64_t r; asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(12582912.0f)); return r; }();
    const uint64_t nmagic = [&] { uint64_t r; asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(-12582912.0f)); return r; }();
    uint64_t rr, nf, f, pp;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rr) : "l"(xc), "l"(magic));            // low mantissa bits = round(x)
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(nf) : "l"(rr), "l"(nmagic));           // round(x) as a float
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(f) : "l"(xc), "l"(nf));                // f = x - round(x)
    const uint64_t c3 = [&] { uint64_t r; asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(0.05550411f)); return r; }();
    const uint64_t c2 = [&] { uint64_t r; asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(0.24022651f)); return r; }();
    const uint64_t c1 = [&] { uint64_t r; asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(0.69314718f)); return r; }();
    const uint64_t c0 = [&] { uint64_t r; asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(1.0f)); return r; }();
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(pp) : "l"(c3), "l"(f), "l"(c2));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(pp) : "l"(pp), "l"(f), "l"(c1));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(pp) : "l"(pp), "l"(f), "l"(c0));
    uint32_t p0, p1, r0, r1;
    asm("mov.b64 {%0, %1}, %2;" : "=r"(p0), "=r"(p1) : "l"(pp));
    asm("mov.b64 {%0, %1}, %2;" : "=r"(r0), "=r"(r1) : "l"(rr));
    e0 = __uint_as_float(p0 + (r0 << 23));         // magic's low 23 mantissa bits are 0x400000: << 23 leaves round(x) mod 2^9
    e1 = __uint_as_float(p1 + (r1 << 23));
}
#ifndef GG_ATTN_POLY
#define GG_ATTN_POLY 3          // k > 0: every k-th pair of scores takes the polynomial; 0: all exponentials on the MUFU
#endif
__device__ __forceinline__ float max3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ uint64_t fma_f32x2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t add_f32x2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
          "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}

// V [B, Tk, H, d] (strided) -> V^T [B, H, d, Tkp] with the keys contiguous (Tkp = Tk rounded up to 8; pad columns zero)
template <int D>
__global__ void __launch_bounds__(256) attn_vt_kernel(const __nv_bfloat16* __restrict__ v, long long v_bs, int v_rs, int v_hs,
                                                      __nv_bfloat16* __restrict__ vt, int H, int Tk, int Tkp) {
    __shared__ __nv_bfloat16 tile[64][D + 2];
    pdl_launch_dependents();
    pdl_wait();
    const int b = blockIdx.z, h = blockIdx.y, k0 = blockIdx.x * 64;
    const __nv_bfloat16* src = v + (long long)b * v_bs + (long long)h * v_hs;
    for (int i = threadIdx.x; i < 64 * (D / 2); i += 256) {
        const int key = i / (D / 2), c2 = i % (D / 2);
        uint32_t w = 0;
        if (k0 + key < Tk) w = *reinterpret_cast<const uint32_t*>(src + (long long)(k0 + key) * v_rs + 2 * c2);
        *reinterpret_cast<uint32_t*>(&tile[key][2 * c2]) = w;
    }
    __syncthreads();
    __nv_bfloat16* dst = vt + ((long long)b * H + h) * D * Tkp;
    for (int i = threadIdx.x; i < D * 32; i += 256) {
        const int c = i / 32, kp = i % 32;              // two keys per thread
        const int key = k0 + 2 * kp;
        if (key < Tkp) {
            __nv_bfloat162 o;
            o.x = tile[2 * kp][c];
            o.y = tile[2 * kp + 1][c];
            *reinterpret_cast<__nv_bfloat162*>(dst + (long long)c * Tkp + key) = o;
        }
    }
}

template <int D>
__global__ void __launch_bounds__(AT_THREADS, 1) attention_tc_kernel(const __grid_constant__ AttnTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    constexpr int VT_BYTES = 2 * D * 128;                // two 64-key chunks of [D rows x 128 B]
    constexpr int KV_BYTES = AT_TILE + VT_BYTES;
    uint8_t* smem_q = smem;                              // [2][128 x 128 B]
    uint8_t* smem_kv = smem_q + 2 * AT_TILE;             // [AT_ST][K tile | V^T chunks]
    uint8_t* smem_p = smem_kv + AT_ST * KV_BYTES;        // [2][2 chunks x 128 x 128 B]
    uint64_t* q_full = reinterpret_cast<uint64_t*>(smem_p + 4 * AT_TILE);
    uint64_t* kv_full = q_full + 1;
    uint64_t* kv_empty = kv_full + AT_ST;
    uint64_t* s_full = kv_empty + AT_ST;                 // [2]
    uint64_t* p_full = s_full + 2;                       // [2]
    uint64_t* pv_done = p_full + 2;                      // [2]
    uint64_t* s_free = pv_done + 2;                      // [2]  the softmax warps hold S_g(j) in registers: its TMEM columns may be rewritten
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_free + 2);
    float* xch_max = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(q_full) + 256);      // [2][2][2][128]
    float* xch_sum = xch_max + 2 * 2 * 2 * 128;                                                // [2][2][128]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool two = p.two != 0;
    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * (two ? 2 * AT_BM : AT_BM);
    const int n_kt = p.n_kt;
    pdl_launch_dependents();
    if (warp == 0 && lane == 0) { prefetch_tmap(&p.qmap); prefetch_tmap(&p.kmap); prefetch_tmap(&p.vtmap); }
    if (warp == 1) {
        if (lane == 0) {
            mbar_init(q_full, 1);
            for (int i = 0; i < AT_ST; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
            for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 4 * AT_HS); mbar_init(&pv_done[i], 1); mbar_init(&s_free[i], 4 * AT_HS); }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();                     // prologue above: shared memory / TMEM / kernel parameters only (common.cuh)
    // TMEM columns: S_A [0, 128), S_B [128, 256), O_A [256, 256 + D), O_B [320, 320 + D)
    constexpr uint32_t COL_S = 0, COL_O = 256, O_STRIDE = 64;

    if (warp == 0) {
        // ================================================================ TMA producer
        if (elect_one()) {
            mbar_expect_tx(q_full, two ? 2 * AT_TILE : AT_TILE);
            tma_load_4d(smem_q, &p.qmap, q_full, 0, h, q0, b);
            if (two) tma_load_4d(smem_q + AT_TILE, &p.qmap, q_full, 0, h, q0 + AT_BM, b);
        }
        __syncwarp();
        for (int j = 0; j < n_kt; ++j) {
            const int st = j % AT_ST;
            mbar_wait(&kv_empty[st], (((uint32_t)(j / AT_ST)) & 1u) ^ 1u);
            if (elect_one()) {
                uint8_t* dst = smem_kv + (size_t)st * KV_BYTES;
                mbar_expect_tx(&kv_full[st], KV_BYTES);
                tma_load_4d(dst, &p.kmap, &kv_full[st], 0, h, j * AT_BN, b);
                tma_load_4d(dst + AT_TILE, &p.vtmap, &kv_full[st], j * AT_BN, 0, h, b);
                tma_load_4d(dst + AT_TILE + D * 128, &p.vtmap, &kv_full[st], j * AT_BN + 64, 0, h, b);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ================================================================ MMA issuer
        const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(AT_BN >> 3) << 17) | ((uint32_t)(AT_BM >> 4) << 24);
        const uint32_t idesc_pv = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(D >> 3) << 17) | ((uint32_t)(AT_BM >> 4) << 24);
        const uint64_t qd[2] = {make_sw128_desc(smem_u32(smem_q)), make_sw128_desc(smem_u32(smem_q + AT_TILE))};
        auto issue_s = [&](int g, int st) {
            const uint64_t kd = make_sw128_desc(smem_u32(smem_kv + (size_t)st * KV_BYTES));
#pragma unroll
            for (int k = 0; k < D / 16; ++k) umma_bf16(tmem_base + COL_S + (uint32_t)g * AT_BN, qd[g] + 2 * k, kd + 2 * k, idesc_s, k > 0 ? 1u : 0u);
            umma_commit(&s_full[g]);
        };
        auto issue_pv = [&](int g, int st, int j) {
            const uint8_t* vt = smem_kv + (size_t)st * KV_BYTES + AT_TILE;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const uint64_t pd = make_sw128_desc(smem_u32(smem_p + (size_t)(2 * g + c) * AT_TILE));
                const uint64_t vd = make_sw128_desc(smem_u32(vt + (size_t)c * D * 128));
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16(tmem_base + COL_O + (uint32_t)g * O_STRIDE, pd + 2 * k, vd + 2 * k, idesc_pv, (j > 0 || c > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit(&pv_done[g]);
        };
        mbar_wait(q_full, 0);
        mbar_wait(&kv_full[0], 0);
        tc_fence_after();
        if (elect_one()) { issue_s(0, 0); if (two) issue_s(1, 0); }
        __syncwarp();
        for (int j = 0; j < n_kt; ++j) {
            const int st = j % AT_ST, stn = (j + 1) % AT_ST;
            const uint32_t ph = (uint32_t)j & 1u;
            const bool more = j + 1 < n_kt;
            // ---- next scores first: a tile's S columns are free as soon as its softmax warps have pulled S(j) into registers,
            // long before they finish exponentiating it, so S(j + 1) is waiting for them when they do
            if (more) mbar_wait(&kv_full[stn], ((uint32_t)((j + 1) / AT_ST)) & 1u);
            mbar_wait(&s_free[0], ph);
            tc_fence_after();
            if (more) { if (elect_one()) issue_s(0, stn); __syncwarp(); }
            if (two) {
                mbar_wait(&s_free[1], ph);
                tc_fence_after();
                if (more) { if (elect_one()) issue_s(1, stn); __syncwarp(); }
            }
            // ---- O_g += P_g(j) V_j once P_g(j) is in shared memory
            mbar_wait(&p_full[0], ph);
            tc_fence_after();
            if (elect_one()) {
                issue_pv(0, st, j);
                if (!two) umma_commit(&kv_empty[st]);
            }
            __syncwarp();
            if (two) {
                mbar_wait(&p_full[1], ph);
                tc_fence_after();
                if (elect_one()) {
                    issue_pv(1, st, j);
                    umma_commit(&kv_empty[st]);      // every MMA that reads this K / V^T stage has been issued before this commit
                }
                __syncwarp();
            }
        }
    } else {
        // ================================================================ softmax + output: 4 AT_HS warps per query tile (tile A first);
        // within a tile, warps [0, 4) take the first AT_NC keys of every row, warps [4, 8) the rest (AT_HS = 2)
        const int sw = warp - 2;
        const int g = sw / (4 * AT_HS);
        const int hf = (sw >> 2) % AT_HS;
        const int qd4 = warp & 3;
        const int row = qd4 * 32 + lane;
        const uint32_t lane_base = tmem_base + ((uint32_t)(qd4 * 32) << 16);
        const uint32_t s_addr = lane_base + COL_S + (uint32_t)g * AT_BN + (uint32_t)(hf * AT_NC);
        const uint32_t o_addr = lane_base + COL_O + (uint32_t)g * O_STRIDE;
        uint8_t* prow = smem_p + (size_t)(2 * g) * AT_TILE + row * 128;
        const int swz = row & 7;
        const float sl2 = p.scale_log2;
        const int pair_bar = 1 + g * 4 + qd4;                      // named barrier of the two warps that share these 32 rows
        float m_used = -INFINITY, l = 0.f;
        const int n_mine = (g == 1 && !two) ? 0 : n_kt;          // one-tile CTAs: the warps of tile B have nothing to do
        for (int j = 0; j < n_mine; ++j) {
            const uint32_t ph = (uint32_t)j & 1u;
            mbar_wait(&s_full[g], ph);
            tc_fence_after();
            const int key0 = j * AT_BN + hf * AT_NC;
            const bool tail = key0 + AT_NC > p.Tk;
            // ---- this thread's part of the score row into registers (one TMEM round trip per tile), then hand the S columns back
            uint32_t r[AT_NC];
#pragma unroll
            for (int c = 0; c < AT_NC / 16; ++c) tmem_ld16_nowait(s_addr + 16 * c, r + 16 * c);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_free[g]);
            if (tail) {
#pragma unroll
                for (int i = 0; i < AT_NC; ++i)
                    if (key0 + i >= p.Tk) r[i] = 0xff800000u;       // -inf: never the maximum, exp2 = 0
            }
            // ---- row maximum (the scale is positive: maximum of the raw scores, scaled once; four independent three-input chains)
            float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int i = 0; i < AT_NC / 2; ++i) mx4[i & 3] = max3(mx4[i & 3], __uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
            float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * sl2;
            if constexpr (AT_HS == 2) {         // the other half of the row: slots alternate with the tile parity, one barrier per tile
                float* slot = xch_max + (((j & 1) * 2 + g) * 2) * 128;
                slot[hf * 128 + row] = mx;
                asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
                mx = fmaxf(mx, slot[(hf ^ 1) * 128 + row]);
            }
            const float m_new = fmaxf(m_used, mx);
            const bool need = m_new > m_used + 8.0f;           // first tile: m_used = -inf
            bool waited = false;
            if (__any_sync(0xffffffffu, need)) {
                const float alpha = need ? ex2_fast(m_used - m_new) : 1.0f;
                if (need) { m_used = m_new; l *= alpha; }
                if (j > 0 && hf == 0) {        // O_g += ... of tile j - 1 must have landed before it is rescaled (j = 0: PV overwrites)
                    mbar_wait(&pv_done[g], ph ^ 1u);
                    tc_fence_after();
                    waited = true;
#pragma unroll
                    for (int c = 0; c < D / 16; ++c) {
                        uint32_t o16[16];
                        tmem_ld16(o_addr + 16 * c, o16);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i) o16[i] = __float_as_uint(__uint_as_float(o16[i]) * alpha);
                        tmem_st16(o_addr + 16 * c, o16);
                    }
                    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                }
            }
            // ---- P = exp2(s c - m_used) as bf16.  Packed f32x2 arithmetic: per PAIR of scores one FMA, two MUFU.EX2 (or the FMA-pipe
            // polynomial), one pack and one add (the row sum, kept in independent partial sums); the packed P words replace the scores
            const uint64_t sl2x2 = pack_f32x2(sl2, sl2), negm = pack_f32x2(-m_used, -m_used);
            uint64_t ls2[4] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll
            for (int i = 0; i < AT_NC / 2; ++i) {
                const uint64_t x = fma_f32x2(pack_f32x2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), sl2x2, negm);
                float x0, x1;
                asm("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(x));
                float e0, e1;
                if (GG_ATTN_POLY > 0 && (i % (GG_ATTN_POLY > 0 ? GG_ATTN_POLY : 1)) == (GG_ATTN_POLY - 1)) ex2_poly_x2(x, e0, e1);
                else { e0 = ex2_fast(x0); e1 = ex2_fast(x1); }
                r[i] = pack_bf16(e0, e1);
                ls2[i & 3] = add_f32x2(ls2[i & 3], pack_f32x2(e0, e1));
            }
            if (j > 0 && !waited) {          // P_g(j - 1) is being read by its MMAs until they complete
                mbar_wait(&pv_done[g], ph ^ 1u);
                tc_fence_after();
            }
            // keys 8 q .. 8 q + 7 of this row = 16-byte piece q & 7 of chunk tile q >> 3 (64 keys per tile), swizzled by the row
#pragma unroll
            for (int q16 = 0; q16 < AT_NC / 8; ++q16) {
                const int qa = q16 + hf * (AT_NC / 8);             // piece index within the 128-key row
                uint8_t* base = prow + (size_t)(qa >> 3) * AT_TILE;
                *reinterpret_cast<uint4*>(base + (((qa & 7) ^ swz) << 4)) = make_uint4(r[4 * q16], r[4 * q16 + 1], r[4 * q16 + 2], r[4 * q16 + 3]);
            }
            float lsum;
            {
                float a0, a1, b0, b1, c0, c1, d0, d1;
                asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(ls2[0]));
                asm("mov.b64 {%0, %1}, %2;" : "=f"(b0), "=f"(b1) : "l"(ls2[1]));
                asm("mov.b64 {%0, %1}, %2;" : "=f"(c0), "=f"(c1) : "l"(ls2[2]));
                asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(ls2[3]));
                lsum = ((a0 + a1) + (b0 + b1)) + ((c0 + c1) + (d0 + d1));
            }
            l += lsum;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic writes -> tensor-core reads
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[g]);
        }
        // ---- output: O / l  (AT_HS = 2: the halves add their row sums, first half first, and take half of the channels each)
        if (n_mine > 0) {
        mbar_wait(&pv_done[g], ((uint32_t)(n_kt - 1)) & 1u);
        tc_fence_after();
        if constexpr (AT_HS == 2) {
            float* slot = xch_sum + (g * 2) * 128;
            slot[hf * 128 + row] = l;
            asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
            l = slot[row] + slot[128 + row];
        }
        const int t = q0 + g * AT_BM + row;
        const float inv = 1.0f / l;
        constexpr int DC = D / AT_HS;           // output channels per thread
        __nv_bfloat16* orow = p.o + (long long)b * p.o_bs + (long long)t * p.o_rs + (long long)h * p.o_hs + hf * DC;
#pragma unroll
        for (int c = 0; c < DC / 16; ++c) {
            uint32_t r[16];
            tmem_ld16(o_addr + (uint32_t)(hf * DC) + 16 * c, r);
            tmem_ld_wait();
            if (t < p.Tq) {
                uint32_t w[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) w[i] = pack_bf16(__uint_as_float(r[2 * i]) * inv, __uint_as_float(r[2 * i + 1]) * inv);
                if (aligned32(orow)) {          // 16 channels = one 32-byte sector per lane
                    stg_u8(orow + 16 * c, make_uint4(w[0], w[1], w[2], w[3]), make_uint4(w[4], w[5], w[6], w[7]));
                } else {
                    *reinterpret_cast<uint4*>(orow + 16 * c) = make_uint4(w[0], w[1], w[2], w[3]);
                    *reinterpret_cast<uint4*>(orow + 16 * c + 8) = make_uint4(w[4], w[5], w[6], w[7]);
                }
            }
        }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ---------------------------------------------------------------------------------------- host
static bool encode_map4(CUtensorMap* m, const void* base, const int64_t dim[4], const int64_t stride_el[3], const int box[4]) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t gdim[4] = {(cuuint64_t)dim[0], (cuuint64_t)dim[1], (cuuint64_t)dim[2], (cuuint64_t)dim[3]};
    cuuint64_t gstr[3] = {(cuuint64_t)stride_el[0] * 2, (cuuint64_t)stride_el[1] * 2, (cuuint64_t)stride_el[2] * 2};
    cuuint32_t bx[4] = {(cuuint32_t)box[0], (cuuint32_t)box[1], (cuuint32_t)box[2], (cuuint32_t)box[3]};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

static inline int64_t attn_tkp(const gg_attn_args* a) { return ((int64_t)a->Tk + 7) / 8 * 8; }

// shapes the tensor-core kernel takes; everything else stays on attention.cu's kernel
bool attention_tc_ok(const gg_attn_args* a) {
    if (!(a->d == 32 || a->d == 64)) return false;
    // measured (tools/bench_attn.py, B200): 1.25-1.45x the mma.sync kernel from 1024 keys up; below that the V^T pre-pass
    // and the 256-query CTA granularity cost more than the tensor core saves (T = 256: 0.025 vs 0.015 ms)
    static const int min_tk = [] { const char* e = getenv("GG_ATTN_TC_MIN_TK"); return e ? atoi(e) : 1024; }();
    if (a->Tk < min_tk || a->Tk < 64 || a->Tq < 64) return false;
    if (!aligned(a->o, 16) || a->o_rs % 8 || a->o_hs % 8 || a->o_bs % 8) return false;
    if (a->q_rs % 8 || a->k_rs % 8 || a->q_hs % 8 || a->k_hs % 8 || a->q_bs % 8 || a->k_bs % 8) return false;
    if (a->v_rs % 2 || a->v_hs % 2 || a->v_bs % 2 || !aligned(a->v, 4)) return false;
    // TMA: the smallest stride of each map must be the element stride of the next dimension up; heads before rows
    return true;
}
int64_t attention_tc_workspace(const gg_attn_args* a) { return attention_tc_ok(a) ? (int64_t)a->B * a->H * a->d * attn_tkp(a) * 2 : 0; }

template <int D>
static int launch_attention_tc(const gg_attn_args* a, cudaStream_t stream) {
    AttnTcParams p;
    memset(&p, 0, sizeof(p));
    const int64_t Tkp = attn_tkp(a);
    {   // Q / K: (channel [extent d, box 64: zero fill above d], head, token, batch)
        const int box[4] = {64, 1, AT_BM, 1};
        const int64_t qdim[4] = {a->d, a->H, a->Tq, a->B}, qstr[3] = {a->q_hs, a->q_rs, a->q_bs};
        const int64_t kdim[4] = {a->d, a->H, a->Tk, a->B}, kstr[3] = {a->k_hs, a->k_rs, a->k_bs};
        if (!encode_map4(&p.qmap, a->q, qdim, qstr, box) || !encode_map4(&p.kmap, a->k, kdim, kstr, box)) return GG_ERR_DRIVER;
        // V^T workspace: (key, channel, head, batch)
        const int vbox[4] = {64, D, 1, 1};
        const int64_t vdim[4] = {Tkp, a->d, a->H, a->B}, vstr[3] = {Tkp, (int64_t)a->d * Tkp, (int64_t)a->H * a->d * Tkp};
        if (!encode_map4(&p.vtmap, a->workspace, vdim, vstr, vbox)) return GG_ERR_DRIVER;
    }
    p.o = reinterpret_cast<__nv_bfloat16*>(a->o); p.o_bs = a->o_bs; p.o_rs = a->o_rs; p.o_hs = a->o_hs;
    p.H = a->H; p.Tq = a->Tq; p.Tk = a->Tk; p.n_kt = (a->Tk + AT_BN - 1) / AT_BN;
    p.scale_log2 = a->scale * 1.4426950408889634f;
    {
        const dim3 grid((unsigned)((a->Tk + 63) / 64), (unsigned)a->H, (unsigned)a->B);
        const cudaError_t e = launch_k(attn_vt_kernel<D>, grid, dim3(256), 0, stream, reinterpret_cast<const __nv_bfloat16*>(a->v),
                                       (long long)a->v_bs, (int)a->v_rs, (int)a->v_hs, reinterpret_cast<__nv_bfloat16*>(a->workspace), (int)a->H,
                                       (int)a->Tk, (int)Tkp);
        if (e != cudaSuccess) return (int)e;
        const int st = launch_result();
        if (st != GG_OK) return st;
    }
    constexpr size_t smem = 1024 + 2 * AT_TILE + AT_ST * (AT_TILE + 2 * D * 128) + 4 * AT_TILE + 256 + AT_XCH_BYTES;
    auto* fn = attention_tc_kernel<D>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    // two query tiles per CTA share every K / V^T tile and hide each other's tensor-core latency; when that grid would leave
    // SMs idle (a depth slab's share of the queries), one tile per CTA doubles the CTA count instead
    const int64_t ctas2 = (int64_t)((a->Tq + 2 * AT_BM - 1) / (2 * AT_BM)) * a->H * a->B;
    p.two = 2 * ctas2 > num_sms() ? 1 : 0;          // one tile per CTA only while the doubled grid still is a single wave
    const int rows = p.two ? 2 * AT_BM : AT_BM;
    const dim3 grid((unsigned)((a->Tq + rows - 1) / rows), (unsigned)a->H, (unsigned)a->B);
    const cudaError_t le = launch_k(fn, grid, dim3(AT_THREADS), smem, stream, p);
    if (le != cudaSuccess) return (int)le;
    return launch_result();
}

int attention_tc_fwd(const gg_attn_args* a, cudaStream_t stream) {
    if (!encode_fn()) return GG_ERR_DRIVER;
    GG_REQUIRE(a->workspace != nullptr && a->workspace_bytes >= attention_tc_workspace(a) && aligned(a->workspace, 128), GG_ERR_BAD_ARG);
    return a->d == 32 ? launch_attention_tc<32>(a, stream) : launch_attention_tc<64>(a, stream);
}

}  // namespace gg
