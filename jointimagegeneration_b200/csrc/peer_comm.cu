// peer_comm.cu -- collectives of the depth-slab decomposition as plain kernels over NVLink peer memory
//
// BASELINE config 5 splits ONE volume into depth slabs over the GPUs of a box; per UNet forward every rank exchanges a
// halo plane with its neighbours before each 3-tap-in-depth convolution (65 exchanges), the GroupNorm partial sums of
// every normalisation (81) and the keys / values of every attention site (11) -- SURVEY.md section 8e.  As NCCL calls
// enqueued from the host between kernels these cost more than the kernels they separate (round 1: 7 of 20 ms at 4 GPUs)
// and keep the forward out of a CUDA graph.  Here every rank owns a PEER-VISIBLE communication arena (cudaMalloc +
// cudaIpc handles, identical layout on every rank), and one kernel per collective does
//
//     phase 1  push: copy my payloads into the destination slots of the peers' arenas (16-byte stores over NVLink),
//              __threadfence_system(), then store the current EPOCH into my flag on every peer (last CTA to finish);
//     phase 2  wait until every peer that writes to me has stored the epoch into my flags, then run the local follow-up
//              copies (staging slot -> halo plane of the activation) / zero fills (halo at the end of the volume).
//
// No NCCL, no host involvement: the kernels are ordinary launches, so the whole slab forward is CUDA-graph capturable.
// Every collective site has its own slots and flags (no reuse inside a forward); the epoch is a device counter bumped
// once per forward, flags are monotone, so a rank that runs ahead can never overwrite a slot its peer has not consumed
// (it would first need that peer's flags of the current forward, which the peer raises only after consuming the
// previous one -- stream order).  A wait that exceeds 10 s traps (a protocol bug must not hang the GPU).
#include <cstring>

#include "common.cuh"

namespace gg {

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ void copy_slice(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int64_t bytes, int cta, int nctas) {
    // 16-byte pieces, grid-strided, FOUR loads in flight per thread (a lone load per thread leaves the copy latency-bound:
    // 35 us for 1.2 MB measured); sources are read past L1 (they may have been written by a peer)
    const int64_t n16 = bytes >> 4, stride = (int64_t)nctas * blockDim.x;
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    int64_t i = (int64_t)cta * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n16; i += 4 * stride) {
        const uint4 a = __ldcg(s4 + i), b = __ldcg(s4 + i + stride), c = __ldcg(s4 + i + 2 * stride), d = __ldcg(s4 + i + 3 * stride);
        d4[i] = a; d4[i + stride] = b; d4[i + 2 * stride] = c; d4[i + 3 * stride] = d;
    }
    for (; i < n16; i += stride) d4[i] = __ldcg(s4 + i);
    if (cta == 0) for (int64_t k = (n16 << 4) + threadIdx.x; k < bytes; k += blockDim.x) dst[k] = src[k];
}

__global__ void __launch_bounds__(256) peer_exchange_kernel(const gg_peer_xchg_args a) {
    const int cta = blockIdx.x, nctas = gridDim.x;
    const uint32_t epoch = *reinterpret_cast<const volatile uint32_t*>(a.epoch);
    __shared__ int is_last;
    if (a.phase & 1) {
        for (int s = 0; s < a.nsend; ++s)
            copy_slice(reinterpret_cast<const uint8_t*>(a.src[s]), reinterpret_cast<uint8_t*>(a.dst[s]), a.bytes[s], cta, nctas);
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            // the last CTA to arrive raises the flags (atomicInc wraps the counter back to 0 for the next launch)
            const unsigned int old = atomicInc(a.done_counter, (unsigned int)nctas - 1u);
            is_last = old == (unsigned int)nctas - 1u;
        }
        __syncthreads();
        if (is_last && threadIdx.x < a.nflag_out) {
            __threadfence_system();
            st_release_sys(a.flag_out[threadIdx.x], epoch);
        }
    }
    if (a.phase & 2) {
        if (threadIdx.x < a.nflag_in) {
            const uint32_t* f = a.flag_in[threadIdx.x];
            uint64_t t0 = 0;
            uint32_t spins = 0;
            while ((int32_t)(ld_acquire_sys(f) - epoch) < 0) {
                if ((++spins & 0x3FFu) == 0) {
                    uint64_t now;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                    if (t0 == 0) t0 = now;
                    else if (now - t0 > 10000000000ull) __trap();
                }
            }
        }
        __syncthreads();
        __threadfence_system();
        for (int s = 0; s < a.ncopy; ++s)
            copy_slice(reinterpret_cast<const uint8_t*>(a.csrc[s]), reinterpret_cast<uint8_t*>(a.cdst[s]), a.cbytes[s], cta, nctas);
        for (int s = 0; s < a.nzero; ++s) {
            uint4* z = reinterpret_cast<uint4*>(a.zdst[s]);
            const int64_t n16 = a.zbytes[s] >> 4;
            for (int64_t i = (int64_t)cta * blockDim.x + threadIdx.x; i < n16; i += (int64_t)nctas * blockDim.x) z[i] = make_uint4(0u, 0u, 0u, 0u);
        }
    }
}

__global__ void peer_epoch_inc_kernel(uint32_t* epoch) { *epoch = *epoch + 1u; }

}  // namespace gg

using namespace gg;

extern "C" {

int gg_peer_alloc(int64_t bytes, void** out) {
    GG_REQUIRE(out != nullptr && bytes > 0, GG_ERR_BAD_ARG);
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, (size_t)bytes);          // plain cudaMalloc: the only kind of allocation cudaIpc can export
    if (e != cudaSuccess) return (int)e;
    e = cudaMemset(p, 0, (size_t)bytes);
    if (e != cudaSuccess) { cudaFree(p); return (int)e; }
    *out = p;
    return GG_OK;
}
int gg_peer_free(void* p) { return p ? (int)cudaFree(p) : GG_OK; }

int gg_peer_export(void* p, uint8_t* handle64) {
    GG_REQUIRE(p != nullptr && handle64 != nullptr, GG_ERR_BAD_ARG);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) return (int)e;
    memcpy(handle64, &h, 64);
    return GG_OK;
}
int gg_peer_open(const uint8_t* handle64, void** out) {
    GG_REQUIRE(handle64 != nullptr && out != nullptr, GG_ERR_BAD_ARG);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return (int)e;
    *out = p;
    return GG_OK;
}
int gg_peer_close(void* p) { return p ? (int)cudaIpcCloseMemHandle(p) : GG_OK; }

int gg_peer_epoch_inc(uint32_t* epoch, gg_stream_t stream) {
    GG_REQUIRE(epoch != nullptr, GG_ERR_BAD_ARG);
    peer_epoch_inc_kernel<<<1, 1, 0, as_stream(stream)>>>(epoch);
    return launch_result();
}

int gg_peer_exchange(const gg_peer_xchg_args* a, gg_stream_t stream) {
    GG_REQUIRE(a != nullptr && a->epoch != nullptr && a->done_counter != nullptr, GG_ERR_BAD_ARG);
    GG_REQUIRE(a->nsend >= 0 && a->nsend <= 8 && a->nflag_out >= 0 && a->nflag_out <= 8 && a->nflag_in >= 0 && a->nflag_in <= 8 &&
               a->ncopy >= 0 && a->ncopy <= 4 && a->nzero >= 0 && a->nzero <= 2 && (a->phase & 3) != 0, GG_ERR_BAD_ARG);
    int64_t most = 0;
    for (int s = 0; s < a->nsend; ++s) {
        GG_REQUIRE(a->src[s] && a->dst[s] && a->bytes[s] >= 0 && aligned(a->src[s], 16) && aligned(a->dst[s], 16), GG_ERR_ALIGNMENT);
        most = most > a->bytes[s] ? most : a->bytes[s];
    }
    for (int s = 0; s < a->ncopy; ++s) {
        GG_REQUIRE(a->csrc[s] && a->cdst[s] && aligned(a->csrc[s], 16) && aligned(a->cdst[s], 16), GG_ERR_ALIGNMENT);
        most = most > a->cbytes[s] ? most : a->cbytes[s];
    }
    for (int s = 0; s < a->nzero; ++s) {
        GG_REQUIRE(a->zdst[s] && aligned(a->zdst[s], 16) && a->zbytes[s] % 16 == 0, GG_ERR_ALIGNMENT);
        most = most > a->zbytes[s] ? most : a->zbytes[s];
    }
    // enough CTAs to fill the NVLink ports for MB-sized planes, one for the few-hundred-byte GroupNorm sums
    int ctas = (int)((most + 16383) / 16384);          // one CTA iteration = 256 threads x 4 x 16 B
    ctas = ctas < 1 ? 1 : ctas > 128 ? 128 : ctas;
    if (a->ctas > 0) ctas = a->ctas;
    peer_exchange_kernel<<<ctas, 256, 0, as_stream(stream)>>>(*a);
    return launch_result();
}

}  // extern "C"
