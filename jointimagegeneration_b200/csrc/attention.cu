// attention.cu -- flash-style attention (no T x T buffer), bf16 in/out, fp32 online softmax.
//   QKVAttentionLegacy (unet.py:343-360) and CrossAttention (ldm/modules/attention.py:170-193)
// Generic strided heads: element (b, t, h, i) of q sits at q[b*q_bs + t*q_rs + h*q_hs + i].
// One CTA = 128 query rows of one (batch, head); 8 warps x 16 rows; K/V streamed in 64-key tiles through
// double-buffered shared memory (cp.async); S = QK^T and O += PV on mma.sync.m16n8k16 (bf16, fp32 acc).
// Attention is ~0.3 % of the step's FLOPs at the reference shapes.
#include <cstdlib>

#include "common.cuh"

namespace gg {

constexpr int ATT_BM = 64, ATT_BN = 64;
constexpr int ATT_THREADS = 256;      // 8 warps x 16 query rows = 2 * ATT_BM rows per CTA

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ float ex2_fast(float x) {      // ex2.approx: 2 ulp, ex2(-inf) = 0
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool pred) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
    const int bytes = pred ? 16 : 0;      // src-size 0 -> the 16 destination bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// 128 query rows per CTA (8 warps x 16 rows); K/V tiles of 64 keys double-buffered with cp.async so the next
// tile streams in while the current one is multiplied; rows past the sequence end are zero-filled.
template <int D>
__global__ void __launch_bounds__(ATT_THREADS) attention_kernel(const gg_attn_args a) {
    constexpr int LD = D + 8;  // padded row (elements): conflict-free ldmatrix
    constexpr int NT = ATT_THREADS;
    extern __shared__ __align__(16) uint8_t att_smem[];
    __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(att_smem);
    __nv_bfloat16* sKb = sQ + 2 * ATT_BM * LD;
    __nv_bfloat16* sVb = sKb + 2 * ATT_BN * LD;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    pdl_launch_dependents();
    pdl_wait();
    const int q0 = blockIdx.x * (2 * ATT_BM), h = blockIdx.y, b = blockIdx.z;
    const __nv_bfloat16* qg = reinterpret_cast<const __nv_bfloat16*>(a.q) + (int64_t)b * a.q_bs + (int64_t)h * a.q_hs;
    const __nv_bfloat16* kg = reinterpret_cast<const __nv_bfloat16*>(a.k) + (int64_t)b * a.k_bs + (int64_t)h * a.k_hs;
    const __nv_bfloat16* vg = reinterpret_cast<const __nv_bfloat16*>(a.v) + (int64_t)b * a.v_bs + (int64_t)h * a.v_hs;
    constexpr int CPR = D / 8;  // 16-byte chunks per row

    auto load_tile = [&](__nv_bfloat16* dst, const __nv_bfloat16* src, int row0, int rows, int nrows, int rs) {
        for (int i = tid; i < rows * CPR; i += NT) {
            const int r = i / CPR, c = i - r * CPR;
            const bool ok = row0 + r < nrows;
            cp_async16(dst + r * LD + c * 8, src + (int64_t)(ok ? row0 + r : 0) * rs + c * 8, ok);
        }
    };
    load_tile(sQ, qg, q0, 2 * ATT_BM, a.Tq, a.q_rs);
    load_tile(sKb, kg, 0, ATT_BN, a.Tk, a.k_rs);
    load_tile(sVb, vg, 0, ATT_BN, a.Tk, a.v_rs);
    cp_async_commit();
    cp_async_wait_all();
    __syncthreads();

    // Q fragments for this warp's 16 rows, all of D
    uint32_t qf[D / 16][4];
    {
        const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
        for (int ks = 0; ks < D / 16; ++ks) {
            const int c = ks * 16 + (lane >> 4) * 8;
            ldsm_x4(qf[ks], (uint32_t)__cvta_generic_to_shared(sQ + r * LD + c));
        }
    }
    float o[D / 8][4];
#pragma unroll
    for (int i = 0; i < D / 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    const float sl2 = a.scale * 1.4426950408889634f;

    int buf = 0;
    for (int k0 = 0; k0 < a.Tk; k0 += ATT_BN) {
        if (k0 > 0) {
            cp_async_wait_all();   // this tile's K/V have landed ...
            __syncthreads();       // ... for every thread, and everyone is done with the buffer refilled next
        }
        const __nv_bfloat16* sK = sKb + buf * ATT_BN * LD;
        const __nv_bfloat16* sV = sVb + buf * ATT_BN * LD;
        if (k0 + ATT_BN < a.Tk) {
            load_tile(sKb + (buf ^ 1) * ATT_BN * LD, kg, k0 + ATT_BN, ATT_BN, a.Tk, a.k_rs);
            load_tile(sVb + (buf ^ 1) * ATT_BN * LD, vg, k0 + ATT_BN, ATT_BN, a.Tk, a.v_rs);
            cp_async_commit();
        }
        buf ^= 1;
        // ---- S = Q K^T  (16 x 64 per warp)
        float s[ATT_BN / 8][4];
#pragma unroll
        for (int j = 0; j < ATT_BN / 8; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; }
#pragma unroll
        for (int ks = 0; ks < D / 16; ++ks) {
#pragma unroll
            for (int jp = 0; jp < ATT_BN / 16; ++jp) {
                // x4: (keys jp*16 + 0..7, d lo), (same keys, d hi), (keys +8, d lo), (keys +8, d hi)
                uint32_t kf[4];
                const int r = jp * 16 + (lane & 7) + (lane >> 4) * 8;
                const int c = ks * 16 + ((lane >> 3) & 1) * 8;
                ldsm_x4(kf, (uint32_t)__cvta_generic_to_shared(sK + r * LD + c));
                mma_bf16(s[2 * jp], qf[ks], kf[0], kf[1]);
                mma_bf16(s[2 * jp + 1], qf[ks], kf[2], kf[3]);
            }
        }
        // ---- mask + online softmax (rows lane/4 and lane/4 + 8)
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int j = 0; j < ATT_BN / 8; ++j) {
            const int key = k0 + j * 8 + 2 * (lane & 3);
            if (key >= a.Tk) { s[j][0] = -INFINITY; s[j][2] = -INFINITY; }
            if (key + 1 >= a.Tk) { s[j][1] = -INFINITY; s[j][3] = -INFINITY; }
            mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
            mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
        const float c0 = ex2_fast((m0 - mn0) * sl2), c1 = ex2_fast((m1 - mn1) * sl2);
        m0 = mn0; m1 = mn1;
        float rs0 = 0.f, rs1 = 0.f;
        uint32_t pf[ATT_BN / 16][4];
#pragma unroll
        for (int j = 0; j < ATT_BN / 8; ++j) {
            const float p0 = ex2_fast((s[j][0] - mn0) * sl2), p1 = ex2_fast((s[j][1] - mn0) * sl2);
            const float p2 = ex2_fast((s[j][2] - mn1) * sl2), p3 = ex2_fast((s[j][3] - mn1) * sl2);
            rs0 += p0 + p1; rs1 += p2 + p3;
            pf[j >> 1][(j & 1) * 2] = pack_bf16(p0, p1);
            pf[j >> 1][(j & 1) * 2 + 1] = pack_bf16(p2, p3);
        }
        l0 = l0 * c0 + rs0; l1 = l1 * c1 + rs1;
#pragma unroll
        for (int i = 0; i < D / 8; ++i) { o[i][0] *= c0; o[i][1] *= c0; o[i][2] *= c1; o[i][3] *= c1; }
        // ---- O += P V
#pragma unroll
        for (int kk = 0; kk < ATT_BN / 16; ++kk) {
#pragma unroll
            for (int np = 0; np < D / 16; ++np) {
                // x4.trans: (keys kk*16 + 0..7, d np*16 + 0..7), (keys +8, same d), (keys 0..7, d +8), (keys +8, d +8)
                uint32_t vf[4];
                const int r = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                const int c = np * 16 + (lane >> 4) * 8;
                ldsm_x4_t(vf, (uint32_t)__cvta_generic_to_shared(sV + r * LD + c));
                mma_bf16(o[2 * np], pf[kk], vf[0], vf[1]);
                mma_bf16(o[2 * np + 1], pf[kk], vf[2], vf[3]);
            }
        }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
    __nv_bfloat16* og = reinterpret_cast<__nv_bfloat16*>(a.o) + (int64_t)b * a.o_bs + (int64_t)h * a.o_hs;
    const int r0 = q0 + warp * 16 + (lane >> 2), r1 = r0 + 8;
#pragma unroll
    for (int i = 0; i < D / 8; ++i) {
        const int c = i * 8 + 2 * (lane & 3);
        if (r0 < a.Tq) *reinterpret_cast<uint32_t*>(og + (int64_t)r0 * a.o_rs + c) = pack_bf16(o[i][0] * i0, o[i][1] * i0);
        if (r1 < a.Tq) *reinterpret_cast<uint32_t*>(og + (int64_t)r1 * a.o_rs + c) = pack_bf16(o[i][2] * i1, o[i][3] * i1);
    }
}

}  // namespace gg

using namespace gg;

namespace gg {
bool attention_tc_ok(const gg_attn_args* a);                    // attention_tc.cu
int64_t attention_tc_workspace(const gg_attn_args* a);
int attention_tc_fwd(const gg_attn_args* a, cudaStream_t stream);
static bool attn_tc_enabled() {
    static const bool on = [] { const char* e = getenv("GG_ATTN_TC"); return !(e && e[0] == '0'); }();     // tuning knob, read once
    return on;
}
}  // namespace gg

extern "C" int64_t gg_attention_workspace_bytes(const gg_attn_args* a) {
    if (!a || a->B <= 0 || a->H <= 0 || a->Tq <= 0 || a->Tk <= 0 || !attn_tc_enabled()) return 0;
    return attention_tc_workspace(a);
}

extern "C" int gg_attention_fwd(const gg_attn_args* a, gg_stream_t stream) {
    GG_REQUIRE(a && a->q && a->k && a->v && a->o, GG_ERR_BAD_ARG);
    GG_REQUIRE(a->B > 0 && a->H > 0 && a->Tq > 0 && a->Tk > 0, GG_ERR_BAD_ARG);
    GG_REQUIRE(a->d == 32 || a->d == 64, GG_ERR_UNSUPPORTED);
    GG_REQUIRE(a->H <= 65535 && a->B <= 65535, GG_ERR_UNSUPPORTED);
    if (a->workspace != nullptr && attn_tc_enabled() && aligned(a->q, 16) && aligned(a->k, 16) && attention_tc_ok(a))
        return attention_tc_fwd(a, as_stream(stream));
    GG_REQUIRE(aligned(a->q, 16) && aligned(a->k, 16) && aligned(a->v, 16) && aligned(a->o, 4), GG_ERR_ALIGNMENT);
    GG_REQUIRE(a->q_rs % 8 == 0 && a->k_rs % 8 == 0 && a->v_rs % 8 == 0 && a->o_rs % 2 == 0, GG_ERR_ALIGNMENT);
    GG_REQUIRE(a->q_hs % 8 == 0 && a->k_hs % 8 == 0 && a->v_hs % 8 == 0 && a->o_hs % 2 == 0, GG_ERR_ALIGNMENT);
    GG_REQUIRE(a->q_bs % 8 == 0 && a->k_bs % 8 == 0 && a->v_bs % 8 == 0 && a->o_bs % 2 == 0, GG_ERR_ALIGNMENT);
    GG_REQUIRE(a->H <= 65535 && a->B <= 65535, GG_ERR_UNSUPPORTED);
    dim3 grid((a->Tq + 2 * ATT_BM - 1) / (2 * ATT_BM), a->H, a->B);
    auto smem_for = [](int D) { return (size_t)(2 * ATT_BM + 4 * ATT_BN) * (D + 8) * sizeof(__nv_bfloat16); };
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(attention_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_for(64));
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    const cudaError_t le = a->d == 32 ? launch_k(attention_kernel<32>, grid, dim3(ATT_THREADS), smem_for(32), as_stream(stream), *a)
                                      : launch_k(attention_kernel<64>, grid, dim3(ATT_THREADS), smem_for(64), as_stream(stream), *a);
    if (le != cudaSuccess) return (int)le;
    return launch_result();
}
