// umma2_probe.cu -- CTA-pair (cta_group::2) tcgen05.mma experiment behind the paired halo-brick conv.
//   D[256 x N] = A[256 x K] * B[N x K]^T, bf16 in / fp32 accumulate, one cluster of two CTAs:
//   CTA r holds A rows [128r, 128r+128) and B rows [N/2 r, N/2 (r+1)) in ITS OWN shared memory,
//   the leader (rank 0) issues M=256 MMAs, each CTA drains its own 128 TMEM lanes.
// What it checks (each was an open question before writing conv_halo's PAIR path):
//   * both CTAs run tcgen05.alloc.cta_group::2 and get the same base;
//   * TMA loads with .cta_group::2 issued by either CTA complete_tx on the LEADER's mbarrier (mapa address);
//   * the A descriptor may carry a row shift (halo trick) and is applied at the same offset in both CTAs;
//   * tcgen05.commit .multicast::cluster arrives on the same-offset barrier of both CTAs;
//   * remote mbarrier.arrive (shared::cluster) from the peer's epilogue.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma2_probe tools/umma2_probe.cu -lcuda
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_rank(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {      // bounded: a protocol bug traps instead of hanging
    const uint32_t addr = smem_u32(bar);
    uint32_t ok;
    long long spins = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (!ok && ++spins > 20000000ll) { printf("TIMEOUT bar %u parity %u block %d thread %d\n", addr, parity, blockIdx.x, threadIdx.x); __trap(); }
    } while (!ok);
}
// 2-D TMA load in pair mode: data lands in THIS CTA's smem, bytes are counted on the mbarrier at cluster address `bar`
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar) {        // arrive on the same-offset barrier in both CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint64_t sw128_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}

constexpr int STAGES = 2;
constexpr int A_ROWS = 128 + 16;          // the A box carries `shift` extra leading rows (halo-style shifted descriptor)
struct Params {
    CUtensorMap amap, bmap;
    float* d;          // [256, N]
    int N, KB, shift, tiles;
};

// warps: 0 = TMA producer, 1 = MMA issuer (leader only), 2..5 = epilogue
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1) probe2(const __grid_constant__ Params p) {
    extern __shared__ uint8_t raw[];
    const uint32_t r0 = smem_u32(raw);
    uint8_t* smem = raw + (((r0 + 1023u) & ~1023u) - r0);
    const uint32_t a_stage = (A_ROWS * 128 + 1023) & ~1023;
    const uint32_t b_stage = (uint32_t)(p.N / 2) * 128;
    uint8_t* As = smem;
    uint8_t* Bs = smem + STAGES * a_stage;
    uint64_t* full = reinterpret_cast<uint64_t*>(Bs + STAGES * b_stage);     // leader's are the live ones
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = empty + STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_rank();
    const int N = p.N;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 2 * 4); }   // 4 epilogue warps x 2 CTAs
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) printf("rank %u tmem_base 0x%x\n", rank, tmem_base);

    if (warp == 0 && lane == 0) {
        uint32_t st = 0, ph = 0;
        for (int t = 0; t < p.tiles; ++t)
            for (int kb = 0; kb < p.KB; ++kb) {
                mbar_wait(&empty[st], ph ^ 1u);
                if (rank == 0) mbar_expect_tx(&full[st], 2u * (A_ROWS * 128u + b_stage));
                const uint32_t bar = map_to_rank(smem_u32(&full[st]), 0);
                // A rows [128 rank - shift, +A_ROWS): out-of-range rows are zero filled
                tma_load_2d_pair(As + st * a_stage, &p.amap, bar, kb * 64, (t * 256 + 128 * (int)rank) - p.shift);
                tma_load_2d_pair(Bs + st * b_stage, &p.bmap, bar, kb * 64, (N / 2) * (int)rank);
                if (++st == STAGES) { st = 0; ph ^= 1u; }
            }
    } else if (warp == 1 && rank == 0) {
        // instruction descriptor: D fp32, A/B bf16 K-major, N>>3 at [17,23), M>>4 at [24,29) with M = 256
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
        uint32_t st = 0, ph = 0, acc = 0, acc_ph = 0;
        for (int t = 0; t < p.tiles; ++t) {
            mbar_wait(&tempty[acc], acc_ph ^ 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int kb = 0; kb < p.KB; ++kb) {
                mbar_wait(&full[st], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (lane == 0) {
                    const uint32_t a0 = smem_u32(As + st * a_stage) + (uint32_t)p.shift * 128u, b0 = smem_u32(Bs + st * b_stage);
                    for (int k = 0; k < 4; ++k)
                        umma2_bf16(tmem_base + acc * 256, sw128_desc(a0 + k * 32), sw128_desc(b0 + k * 32), idesc, (kb | k) ? 1u : 0u);
                    umma2_commit_mc(&empty[st]);
                    if (kb == p.KB - 1) umma2_commit_mc(&tfull[acc]);
                }
                __syncwarp();
                if (++st == STAGES) { st = 0; ph ^= 1u; }
            }
            acc ^= 1u;
            if (acc == 0) acc_ph ^= 1u;
        }
    } else if (warp >= 2) {
        const int q = warp & 3;
        uint32_t acc = 0, acc_ph = 0;
        for (int t = 0; t < p.tiles; ++t) {
            mbar_wait(&tfull[acc], acc_ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int row = t * 256 + 128 * (int)rank + q * 32 + lane;
            for (int c0 = 0; c0 < N; c0 += 16) {
                uint32_t r[16];
                tmem_ld16(tmem_base + acc * 256 + c0 + ((uint32_t)(q * 32) << 16), r);
                for (int j = 0; j < 16; ++j) p.d[(size_t)row * N + c0 + j] = __uint_as_float(r[j]);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(map_to_rank(smem_u32(&tempty[acc]), 0));
            acc ^= 1u;
            if (acc == 0) acc_ph ^= 1u;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static bool make_map(EncodeFn fn, CUtensorMap* m, void* base, int K, int rows, int box_rows) {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t el[2] = {1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, el, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int main(int argc, char** argv) {
    const int N = argc > 1 ? atoi(argv[1]) : 64, KB = argc > 2 ? atoi(argv[2]) : 5, shift = argc > 3 ? atoi(argv[3]) : 3;
    const int tiles = argc > 4 ? atoi(argv[4]) : 5;
    const int K = KB * 64, M = 256 * tiles;
    void* fnp = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &qres));
    EncodeFn fn = (EncodeFn)fnp;
    std::vector<__nv_bfloat16> hA((size_t)M * K), hB((size_t)N * K);
    std::vector<float> fA(hA.size()), fB(hB.size());
    srand(1);
    for (size_t i = 0; i < hA.size(); ++i) { hA[i] = __float2bfloat16((rand() % 17 - 8) / 8.f); fA[i] = __bfloat162float(hA[i]); }
    for (size_t i = 0; i < hB.size(); ++i) { hB[i] = __float2bfloat16((rand() % 13 - 6) / 4.f); fB[i] = __bfloat162float(hB[i]); }
    __nv_bfloat16 *dA, *dB;
    float* dD;
    CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dD, (size_t)M * N * 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xff, (size_t)M * N * 4));
    Params p;
    if (!make_map(fn, &p.amap, dA, K, M, A_ROWS) || !make_map(fn, &p.bmap, dB, K, N, N / 2)) { printf("encode failed\n"); return 1; }
    p.d = dD; p.N = N; p.KB = KB; p.shift = shift; p.tiles = tiles;
    const size_t smem = 1024 + STAGES * (((A_ROWS * 128 + 1023) & ~1023) + (N / 2) * 128) + 256;
    CK(cudaFuncSetAttribute(probe2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    probe2<<<2, 192, smem>>>(p);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<float> hD((size_t)M * N);
    CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
    // with the shifted box the data row behind MMA row m is still global row m (box starts `shift` rows early,
    // descriptor skips them) -- so the expected result is the plain GEMM
    double maxerr = 0;
    int bad = 0;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double s = 0;
            for (int k = 0; k < K; ++k) s += (double)fA[(size_t)m * K + k] * fB[(size_t)n * K + k];
            const double e = fabs(s - hD[(size_t)m * N + n]);
            if (!(e <= 1e-3)) { if (bad < 5) printf("mismatch m %d n %d want %f got %f\n", m, n, s, hD[(size_t)m * N + n]); ++bad; }
            if (e > maxerr) maxerr = e;
        }
    printf("N %d KB %d shift %d tiles %d: max err %g, mismatches %d -> %s\n", N, KB, shift, tiles, maxerr, bad, bad ? "FAIL" : "OK");
    return bad != 0;
}
