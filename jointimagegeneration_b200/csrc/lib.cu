// lib.cu -- library-level entry points of libguidegen_sm100 (version, status, launch counter)
#include "common.cuh"

namespace gg {
std::atomic<uint64_t> g_launches{0};
}

extern "C" {

int gg_version(void) { return 200; }

int gg_abi_sizes(int32_t* out, int n) {
    const int32_t sz[11] = {(int32_t)sizeof(gg_cat_args), (int32_t)sizeof(gg_cat_step_cl_args), (int32_t)sizeof(gg_ddim_args),
                           (int32_t)sizeof(gg_plms_args), (int32_t)sizeof(gg_ddpm_args), (int32_t)sizeof(gg_gn_finalize_args),
                           (int32_t)sizeof(gg_conv_src), (int32_t)sizeof(gg_conv_args), (int32_t)sizeof(gg_attn_args),
                            (int32_t)sizeof(gg_cat_epilogue), (int32_t)sizeof(gg_peer_xchg_args)};
    for (int i = 0; i < 11 && i < n; ++i)
        if (out) out[i] = sz[i];
    return 11;
}

const char* gg_status_string(int status) {
    switch (status) {
        case GG_OK: return "ok";
        case GG_ERR_BAD_ARG: return "bad argument";
        case GG_ERR_UNSUPPORTED: return "unsupported shape or configuration (no fallback path exists)";
        case GG_ERR_ALIGNMENT: return "pointer or stride alignment";
        case GG_ERR_NO_DEVICE: return "no sm_100 device";
        case GG_ERR_DRIVER: return "CUDA driver entry point unavailable";
        default: break;
    }
    if (status > 0) return cudaGetErrorString((cudaError_t)status);
    return "unknown status";
}

int gg_device_check(void) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return GG_ERR_NO_DEVICE;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return GG_ERR_NO_DEVICE;
    return major == 10 ? GG_OK : GG_ERR_NO_DEVICE;
}

uint64_t gg_launch_count(void) { return gg::g_launches.load(); }
void gg_launch_count_reset(void) { gg::g_launches.store(0); }

}  // extern "C"
