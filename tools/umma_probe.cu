// umma_probe.cu -- which address does tcgen05.mma swizzle?  (experiment behind the halo-brick conv)
// A: 512 rows x 64 bf16 (128 B rows) written to smem exactly as a SWIZZLE_128B TMA box would
// (16-byte chunk index XOR (absolute row & 7), base 1024-aligned).  B = 64 x 64 identity in the same
// layout, so D[m][n] = A_row(m)[n].  The A descriptor is given a start address shifted by `s` rows
// and a group stride of `pitch` rows; D tells us which smem row each MMA row actually read.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdint>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(float* out, int s, int pitch, int base_off_mode) {
    extern __shared__ uint8_t raw[];
    const uint32_t r0 = smem_u32(raw);
    uint8_t* smem = raw + (((r0 + 1023u) & ~1023u) - r0);
    uint8_t* A = smem;                 // 512 rows * 128 B = 64 KB
    uint8_t* B = smem + 65536;         // 64 rows * 128 B
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const int tid = threadIdx.x;
    // A[r][c] = r + c/64 (exactly representable in bf16 for r < 256: use r % 251 and c separately)
    for (int i = tid; i < 512 * 64; i += blockDim.x) {
        const int r = i / 64, c = i % 64;
        const float v = (float)((r * 3 + c) % 255);          // integers < 256: exact in bf16
        const int chunk = (c / 8) ^ (r & 7);
        *reinterpret_cast<__nv_bfloat16*>(A + r * 128 + chunk * 16 + (c % 8) * 2) = __float2bfloat16(v);
    }
    for (int i = tid; i < 64 * 64; i += blockDim.x) {
        const int r = i / 64, c = i % 64;
        const int chunk = (c / 8) ^ (r & 7);
        *reinterpret_cast<__nv_bfloat16*>(B + r * 128 + chunk * 16 + (c % 8) * 2) = __float2bfloat16(r == c ? 1.0f : 0.0f);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // make generic-proxy smem writes visible to the async (tensor) proxy
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tslot;
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t a_start = smem_u32(A) + (uint32_t)s * 128u;
        uint64_t adesc = 0, bdesc = 0;
        adesc |= (uint64_t)((a_start & 0x3FFFF) >> 4);
        adesc |= (uint64_t)1 << 16;
        adesc |= (uint64_t)((pitch * 128) >> 4) << 32;
        adesc |= (uint64_t)1 << 46;
        adesc |= (uint64_t)2 << 61;
        if (base_off_mode) adesc |= (uint64_t)((a_start >> 7) & 7) << 49;
        const uint32_t b_start = smem_u32(B);
        bdesc |= (uint64_t)((b_start & 0x3FFFF) >> 4);
        bdesc |= (uint64_t)1 << 16;
        bdesc |= (uint64_t)(1024 >> 4) << 32;
        bdesc |= (uint64_t)1 << 46;
        bdesc |= (uint64_t)2 << 61;
        for (int k = 0; k < 4; ++k) {
            const uint32_t accum = k ? 1u : 0u;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem), "l"(adesc + 2 * k), "l"(bdesc + 2 * k), "r"(idesc), "r"(accum) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    // everyone waits for the MMAs
    {
        uint32_t ok;
        do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        } while (!ok);
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid < 128) {
        const int warp = tid >> 5;
        for (int c0 = 0; c0 < 64; c0 += 16) {
            uint32_t r[16];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                           "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                         : "r"(tmem + c0 + ((uint32_t)(warp * 32) << 16)));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 16; ++j) out[tid * 64 + c0 + j] = __uint_as_float(r[j]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64) : "memory");
}

int main() {
    float* d;
    cudaMalloc(&d, 128 * 64 * 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 8192 + 1024);
    std::vector<float> h(128 * 64);
    const int cases[][3] = {{0, 8, 0}, {1, 8, 0}, {3, 8, 0}, {0, 10, 0}, {1, 10, 0}, {11, 10, 0}, {21, 10, 0}, {1, 8, 1}, {3, 10, 1}, {11, 10, 1},
                            {0, 16, 0}, {2, 16, 0}, {2, 16, 1}, {18, 16, 1}};
    for (auto& cs : cases) {
        const int s = cs[0], pitch = cs[1], bo = cs[2];
        cudaMemset(d, 0, 128 * 64 * 4);
        probe<<<1, 128, 65536 + 8192 + 1024>>>(d, s, pitch, bo);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("s=%d pitch=%d bo=%d: CUDA error %s\n", s, pitch, bo, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h.data(), d, 128 * 64 * 4, cudaMemcpyDeviceToHost);
        // expected under the 'absolute address' semantic: MMA row m reads smem row s + (m/8)*pitch + (m%8)
        int ok_rows = 0, ok_cols_total = 0;
        for (int m = 0; m < 128; ++m) {
            const int r = s + (m / 8) * pitch + (m % 8);
            int okc = 0;
            for (int c = 0; c < 64; ++c) okc += (h[m * 64 + c] == (float)((r * 3 + c) % 255));
            ok_rows += (okc == 64);
            ok_cols_total += okc;
        }
        printf("s=%2d pitch=%2d base_offset=%d : rows matching absolute-address semantic %3d/128 (elements %4d/8192)   D[0][0..3]=%g %g %g %g  D[9][0..1]=%g %g\n",
               s, pitch, bo, ok_rows, ok_cols_total, h[0], h[1], h[2], h[3], h[9 * 64], h[9 * 64 + 1]);
    }
    return 0;
}
