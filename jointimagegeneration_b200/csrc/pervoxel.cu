// pervoxel.cu -- HBM-bound per-voxel kernels of the two samplers (sm_100a)
//
//   gg_cat_posterior_sample : CCDM categorical posterior + clamp + categorical draw at the
//                             reference's tensor interface (fp32 [B, C, V])
//   gg_cat_step_cl          : the same step in the device-resident sampler loop (channels-last)
//   gg_ddim_update          : LDM DDIM update
//   gg_nchw_to_cl / gg_cl_to_nchw : layout bridges at the drop-in boundary
//
// Arithmetic contract (bit-exact with oracle/diffusion.py::theta_post_prob_closed and
// categorical_sample, oracle/ddim.py::ddim_update): every operation is an explicit
// round-to-nearest fp32 intrinsic (__fmul_rn/__fadd_rn/__fdiv_rn/__fsqrt_rn), which the
// compiler never contracts into FMAs; class-axis sums run left to right.
#include "common.cuh"

namespace gg {

// ------------------------------------------------------------------------------------------
// K11-K13 at the reference interface
// ------------------------------------------------------------------------------------------
template <int VPT> struct VecF;
template <> struct VecF<4> { using T = float4; };
template <> struct VecF<2> { using T = float2; };
template <> struct VecF<1> { using T = float; };

template <int VPT>
__device__ __forceinline__ void load_plane(const float* p, float (&dst)[VPT]) {
    if constexpr (VPT == 4) {
        float4 t = ldg_nc_f4(p);
        dst[0] = t.x; dst[1] = t.y; dst[2] = t.z; dst[3] = t.w;
    } else if constexpr (VPT == 2) {
        float2 t = __ldg(reinterpret_cast<const float2*>(p));
        dst[0] = t.x; dst[1] = t.y;
    } else {
        dst[0] = __ldg(p);
    }
}
template <int VPT>
__device__ __forceinline__ void store_plane(float* p, const float (&src)[VPT]) {
    if constexpr (VPT == 4) {
        stg_na_f4(p, make_float4(src[0], src[1], src[2], src[3]));
    } else if constexpr (VPT == 2) {
        *reinterpret_cast<float2*>(p) = make_float2(src[0], src[1]);
    } else {
        p[0] = src[0];
    }
}

template <int C, int VPT, bool VEC>
__global__ void __launch_bounds__(256) cat_posterior_kernel(const gg_cat_args a, const int64_t groups_per_sample) {
    // grid = (groups of VPT voxels, sample): no 64-bit division on the per-thread path
    const int64_t gi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gi >= groups_per_sample) return;
    const int b = blockIdx.y;
    const int64_t v0 = gi * VPT;
    const int64_t V = a.V;
    const int mode = a.mode;
    const bool given = (mode >= GG_CAT_SAMPLE_GIVEN);
    const bool draw = (mode == GG_CAT_SAMPLE || mode == GG_CAT_SAMPLE_GIVEN);
    int nv = VPT;
    if (!VEC) nv = (int)min((int64_t)VPT, V - v0);

    float p[C][VPT];
    const float* x0p = a.x0 + ((int64_t)b * C) * V + v0;
    // ---- load x0 (and xt) planes: 128-bit coalesced along the voxel axis
    float x0v[C][VPT];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        if (VEC) load_plane<VPT>(x0p + (int64_t)c * V, x0v[c]);
        else {
#pragma unroll
            for (int j = 0; j < VPT; ++j) x0v[c][j] = j < nv ? __ldg(x0p + (int64_t)c * V + j) : 1.0f;
        }
    }
    if (!given) {
        const float* xtp = a.xt + ((int64_t)b * C) * V + v0;
        const float al = __ldg(a.coef + 2 * b), g = __ldg(a.coef + 2 * b + 1);
        const float k = __fdiv_rn(__fsub_rn(1.0f, al), (float)C);
        const float h = __fdiv_rn(__fsub_rn(1.0f, g), (float)C);
        float u[C][VPT];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            float xt[VPT];
            if (VEC) load_plane<VPT>(xtp + (int64_t)c * V, xt);
            else {
#pragma unroll
                for (int j = 0; j < VPT; ++j) xt[j] = j < nv ? __ldg(xtp + (int64_t)c * V + j) : 0.0f;
            }
#pragma unroll
            for (int j = 0; j < VPT; ++j) u[c][j] = __fadd_rn(__fmul_rn(al, xt[j]), k);
        }
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
            float U = u[0][j];
#pragma unroll
            for (int c = 1; c < C; ++c) U = __fadd_rn(U, u[c][j]);
            const float hU = __fmul_rn(h, U);
            float r[C];
#pragma unroll
            for (int c = 0; c < C; ++c) r[c] = __fdiv_rn(x0v[c][j], __fadd_rn(__fmul_rn(g, u[c][j]), hU));
            float R = r[0];
#pragma unroll
            for (int c = 1; c < C; ++c) R = __fadd_rn(R, r[c]);
            const float hR = __fmul_rn(h, R);
#pragma unroll
            for (int c = 0; c < C; ++c) p[c][j] = __fmul_rn(u[c][j], __fadd_rn(__fmul_rn(g, r[c]), hR));
        }
    } else {
#pragma unroll
        for (int c = 0; c < C; ++c)
#pragma unroll
            for (int j = 0; j < VPT; ++j) p[c][j] = x0v[c][j];
    }

    float* outp = a.out ? a.out + ((int64_t)b * C) * V + v0 : nullptr;
    if (mode == GG_CAT_POSTERIOR) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
            if (VEC) store_plane<VPT>(outp + (int64_t)c * V, p[c]);
            else {
#pragma unroll
                for (int j = 0; j < VPT; ++j) if (j < nv) outp[(int64_t)c * V + j] = p[c][j];
            }
        }
        return;
    }
    // ---- clamp (diffusion_denoising.py:216) and normalise (Categorical.__init__)
    const float cm = a.clamp_min;
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
        if (cm > 0.0f) {
#pragma unroll
            for (int c = 0; c < C; ++c) p[c][j] = fmaxf(p[c][j], cm);
        }
        float P = p[0][j];
#pragma unroll
        for (int c = 1; c < C; ++c) P = __fadd_rn(P, p[c][j]);
#pragma unroll
        for (int c = 0; c < C; ++c) p[c][j] = __fdiv_rn(p[c][j], P);
    }
    if (mode == GG_CAT_PROBS || mode == GG_CAT_PROBS_GIVEN) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
            if (VEC) store_plane<VPT>(outp + (int64_t)c * V, p[c]);
            else {
#pragma unroll
                for (int j = 0; j < VPT; ++j) if (j < nv) outp[(int64_t)c * V + j] = p[c][j];
            }
        }
        return;
    }
    // ---- draw: first argmax of p/q (torch.multinomial(.., 1, True)) or plain first argmax
    int idx[VPT];
    if (draw) {
        const int64_t gv0 = (int64_t)b * V + v0;
        float qv[VPT * C];
        if (a.q != nullptr) {
            const float* qp = a.q + gv0 * C;
            if (VEC && (VPT * C) % 4 == 0) {
#pragma unroll
                for (int i = 0; i < VPT * C / 4; ++i) {
                    float4 t = ldg_nc_f4(qp + 4 * i);
                    qv[4 * i] = t.x; qv[4 * i + 1] = t.y; qv[4 * i + 2] = t.z; qv[4 * i + 3] = t.w;
                }
            } else {
#pragma unroll
                for (int i = 0; i < VPT * C; ++i) qv[i] = (i / C) < nv ? __ldg(qp + i) : 1.0f;
            }
        } else {
            constexpr int NB = (C + 3) / 4;
#pragma unroll
            for (int j = 0; j < VPT; ++j) {
                const uint64_t ctr = (uint64_t)(gv0 + j) * NB;
#pragma unroll
                for (int i = 0; i < NB; ++i) {
                    const uint64_t cc = ctr + i;
                    uint4 r = philox4x32_10(make_uint4((uint32_t)cc, (uint32_t)(cc >> 32), (uint32_t)a.offset,
                                                       (uint32_t)(a.offset >> 32)),
                                            make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32)));
                    const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (4 * i + e < C) qv[j * C + 4 * i + e] = exp1_from_bits(rr[e]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
            float best = __fdiv_rn(p[0][j], qv[j * C]);
            int bi = 0;
#pragma unroll
            for (int c = 1; c < C; ++c) {
                const float r = __fdiv_rn(p[c][j], qv[j * C + c]);
                if (r > best) { best = r; bi = c; }
            }
            idx[j] = bi;
        }
    } else {
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
            float best = p[0][j];
            int bi = 0;
#pragma unroll
            for (int c = 1; c < C; ++c)
                if (p[c][j] > best) { best = p[c][j]; bi = c; }
            idx[j] = bi;
        }
    }
    // ---- outputs
    if (outp != nullptr) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
            float oh[VPT];
#pragma unroll
            for (int j = 0; j < VPT; ++j) oh[j] = idx[j] == c ? 1.0f : 0.0f;
            if (VEC) store_plane<VPT>(outp + (int64_t)c * V, oh);
            else {
#pragma unroll
                for (int j = 0; j < VPT; ++j) if (j < nv) outp[(int64_t)c * V + j] = oh[j];
            }
        }
    }
    if (a.out_i64 != nullptr) {
        int64_t* o64 = a.out_i64 + ((int64_t)b * C) * V + v0;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            if (VEC && VPT % 2 == 0) {
#pragma unroll
                for (int j = 0; j < VPT; j += 2) {
                    longlong2 t = make_longlong2(idx[j] == c, idx[j + 1] == c);
                    *reinterpret_cast<longlong2*>(o64 + (int64_t)c * V + j) = t;
                }
            } else {
#pragma unroll
                for (int j = 0; j < VPT; ++j) if (j < nv) o64[(int64_t)c * V + j] = idx[j] == c;
            }
        }
    }
    if (a.labels != nullptr) {
        uint8_t* lp = a.labels + (int64_t)b * V + v0;
        bool done = false;
        if constexpr (VPT == 4) {
            if (VEC) {
                *reinterpret_cast<uchar4*>(lp) = make_uchar4(idx[0], idx[1], idx[2], idx[3]);
                done = true;
            }
        }
        if (!done) {
#pragma unroll
            for (int j = 0; j < VPT; ++j) if (j < nv) lp[j] = (uint8_t)idx[j];
        }
    }
}

template <int C>
static int launch_cat(const gg_cat_args& a, cudaStream_t s) {
    constexpr int VPT = C <= 12 ? 4 : (C <= 24 ? 2 : 1);
    const bool vec = (a.V % VPT == 0) && aligned(a.x0, 16) && (a.xt == nullptr || aligned(a.xt, 16)) &&
                     (a.out == nullptr || aligned(a.out, 16)) && (a.q == nullptr || aligned(a.q, 16)) &&
                     (a.out_i64 == nullptr || aligned(a.out_i64, 16)) && (a.labels == nullptr || aligned(a.labels, 4));
    const int64_t gps = (a.V + VPT - 1) / VPT;
    if (a.B > 65535) return GG_ERR_UNSUPPORTED;
    const dim3 blocks((unsigned)((gps + 255) / 256), (unsigned)a.B);
    if (vec) cat_posterior_kernel<C, VPT, true><<<blocks, 256, 0, s>>>(a, gps);
    else cat_posterior_kernel<C, VPT, false><<<blocks, 256, 0, s>>>(a, gps);
    return launch_result();
}

// ------------------------------------------------------------------------------------------
// channels-last sampler-loop form: 4 lanes per voxel (Cpad = 16 fp32 = one 64-byte row),
// class-axis reductions by warp shuffles (xor 1, 2)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
}
__device__ __forceinline__ float quad_max(float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
    return v;
}

__global__ void __launch_bounds__(256) cat_step_cl_kernel(const gg_cat_step_cl_args a) {
    // Production form: fast intrinsics (ex2/lg2/rcp approximations); the draw is arg-max of p_c / q_c
    // -- the common normaliser 1/sum(p) cannot change the arg-max, so it is only applied to probs_out.
    // grid = (64-voxel blocks of one sample, sample)
    const int b = blockIdx.y;
    const int64_t vl = (int64_t)blockIdx.x * 64 + (threadIdx.x >> 2);
    const int sub = threadIdx.x & 3;
    const bool live = vl < a.V;
    const int64_t vx = (int64_t)b * a.V + (live ? vl : a.V - 1);
    const int C = a.C;
    // softmax over the head conv's logits (unet.py:720), classes [4*sub, 4*sub+4)
    float4 lg = ldg_nc_f4(a.logits + vx * a.Cpad + 4 * sub);
    const int lab = a.labels_in[vx];
    float l[4] = {lg.x, lg.y, lg.z, lg.w};
    float m = -INFINITY;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        if (4 * sub + e >= C) l[e] = -INFINITY;
        m = fmaxf(m, l[e]);
    }
    m = quad_max(m);
    float ex[4], se = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) { ex[e] = __expf(l[e] - m); se += ex[e]; }       // exp(-inf) = 0 for padding classes
    se = quad_sum(se);
    const float inv = __fdividef(1.0f, se);
    // posterior (closed form) with x_t given as a label
    const float al = __ldg(a.coef + 2 * b), g = __ldg(a.coef + 2 * b + 1);
    const float k = (1.0f - al) / (float)C, h = (1.0f - g) / (float)C;
    // u_c = alpha [c == lab] + k; U = sum_c u_c = alpha + C k = 1
    float u[4], r[4], R = 0.f;
    const float hU = h * (al + (float)C * k);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int c = 4 * sub + e;
        u[e] = (c == lab ? al : 0.f) + k;
        r[e] = c < C ? __fdividef(ex[e] * inv, fmaf(g, u[e], hU)) : 0.f;
        R += r[e];
    }
    R = quad_sum(R);
    const float hR = h * R;
    float p[4], P = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int c = 4 * sub + e;
        p[e] = c < C ? fmaxf(u[e] * fmaf(g, r[e], hR), a.clamp_min) : 0.f;
        P += p[e];
    }
    float best = -1.f;
    int bi = 0;
    if (a.mode == GG_CAT_SAMPLE) {
        float q[4] = {1.f, 1.f, 1.f, 1.f};
        if (a.q != nullptr) {
#pragma unroll
            for (int e = 0; e < 4; ++e) if (4 * sub + e < C) q[e] = __ldg(a.q + vx * C + 4 * sub + e);
        } else {
            const uint64_t cc = ((uint64_t)vx + (uint64_t)a.vox_base) * 4 + sub;    // global voxel index: slab-invariant
            uint4 rr = philox4x32_10(make_uint4((uint32_t)cc, (uint32_t)(cc >> 32), (uint32_t)a.offset,
                                                (uint32_t)(a.offset >> 32)),
                                     make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32)));
            q[0] = exp1_from_bits(rr.x); q[1] = exp1_from_bits(rr.y);
            q[2] = exp1_from_bits(rr.z); q[3] = exp1_from_bits(rr.w);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float v = (4 * sub + e < C) ? __fdividef(p[e], q[e]) : -1.f;
            if (v > best) { best = v; bi = 4 * sub + e; }
        }
    } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float v = (4 * sub + e < C) ? p[e] : -1.f;
            if (v > best) { best = v; bi = 4 * sub + e; }
        }
    }
    // first-index argmax across the 4 lanes of the voxel
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (!live) return;
    if (a.probs_out != nullptr) {
        const int64_t v = vl;
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (4 * sub + e < C) a.probs_out[((int64_t)b * C + 4 * sub + e) * a.V + v] = p[e];      // clamped, un-normalised
    }
    if (sub == 0) a.labels_out[vx] = (uint8_t)bi;
    if (a.next_x != nullptr) {
        // next UNet input row: one-hot(C) | cond | zero pad, bf16, 8 channels (16 bytes) per lane
        __nv_bfloat16* row = reinterpret_cast<__nv_bfloat16*>(a.next_x) + vx * a.Cin_pad;
        const __nv_bfloat16* cond = reinterpret_cast<const __nv_bfloat16*>(a.cond);
        for (int c0 = 8 * sub; c0 < a.Cin_pad; c0 += 32) {
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float f[2];
#pragma unroll
                for (int z = 0; z < 2; ++z) {
                    const int c = c0 + 2 * e + z;
                    float v = 0.f;
                    if (c < C) v = (c == bi) ? 1.f : 0.f;
                    else if (c < C + a.n_cond && cond != nullptr) v = __bfloat162float(cond[vx * a.n_cond + (c - C)]);
                    f[z] = v;
                }
                w[e] = pack_bf16(f[0], f[1]);
            }
            *reinterpret_cast<uint4*>(row + c0) = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
}

// Production form of the same step (noise from the in-kernel generator, no probs_out): HBM-bound.
//  * a quad of lanes owns FOUR consecutive voxels: one Philox4x32-10 call yields their four uniforms, the
//    uint8 labels move as one 32-bit word, index math / coefficient loads are amortised 4x;
//  * x_t is one-hot, so the posterior has only two distinct denominators per voxel (class == label or not):
//    two reciprocals replace C divisions, and the softmax normaliser cancels in the draw;
//  * the draw is an inverse-CDF categorical sample with ONE uniform per voxel (scan over the class axis by
//    warp shuffles) instead of the C-variate exponential race torch.multinomial uses -- same distribution,
//    1/12 of the random numbers and no logarithms.  (The injected-noise path above keeps the race so that
//    labels can be compared with the reference draw for draw.)
__global__ void __launch_bounds__(256) cat_step_cl_fast_kernel(const gg_cat_step_cl_args a) {
    const int b = blockIdx.y;
    const int sub = threadIdx.x & 3;
    const int64_t v0 = ((int64_t)blockIdx.x * 64 + (threadIdx.x >> 2)) * 4;     // first of this quad's 4 voxels
    const bool live = v0 < a.V;
    const int64_t vb = (int64_t)b * a.V + (live ? v0 : 0);
    const int C = a.C;
    float4 lg[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) lg[j] = ldg_nc_f4(a.logits + (vb + j) * 16 + 4 * sub);
    const uint32_t labs = __ldg(reinterpret_cast<const uint32_t*>(a.labels_in + vb));
    const float al = __ldg(a.coef + 2 * b), g = __ldg(a.coef + 2 * b + 1);
    const float k = (1.0f - al) / (float)C, h = (1.0f - g) / (float)C;
    const float hU = h * (al + (float)C * k);
    const float inv_lab = __fdividef(1.0f, fmaf(g, al + k, hU)), inv_oth = __fdividef(1.0f, fmaf(g, k, hU));
    const uint64_t gq = ((uint64_t)vb + (uint64_t)a.vox_base) >> 2;             // global index of this group of 4 voxels
    const uint4 rr = philox4x32_10(make_uint4((uint32_t)gq, (uint32_t)(gq >> 32), (uint32_t)a.offset, (uint32_t)(a.offset >> 32)),
                                   make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32)));
    const uint32_t rbits[4] = {rr.x, rr.y, rr.z, rr.w};
    const bool pad[4] = {4 * sub + 0 >= C, 4 * sub + 1 >= C, 4 * sub + 2 >= C, 4 * sub + 3 >= C};
    uint32_t lab_out = 0;
    const uint16_t* cond = reinterpret_cast<const uint16_t*>(a.cond);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float l[4] = {lg[j].x, lg[j].y, lg[j].z, lg[j].w};
        float m = -INFINITY;
#pragma unroll
        for (int e = 0; e < 4; ++e) { if (pad[e]) l[e] = -INFINITY; m = fmaxf(m, l[e]); }
        m = quad_max(m);
        const int lab = (int)((labs >> (8 * j)) & 255u);
        float ex[4], r[4], u[4], se = 0.f, R = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const bool is_lab = (4 * sub + e) == lab;
            ex[e] = __expf(l[e] - m);                       // 0 for padding classes
            se += ex[e];
            u[e] = (is_lab ? al : 0.f) + k;
            r[e] = ex[e] * (is_lab ? inv_lab : inv_oth);    // un-normalised softmax: 1/sum(ex) is a common factor
            R += r[e];
        }
        se = quad_sum(se);
        R = quad_sum(R);
        const float hR = h * R, floor_p = a.clamp_min * se; // clamp(p / se, min) == max(p, min * se) / se
        float cum[4], run = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float pe = pad[e] ? 0.f : fmaxf(u[e] * fmaf(g, r[e], hR), floor_p);
            run += pe;
            cum[e] = run;
        }
        // inclusive scan of the lane totals over the quad, then the voxel total
        float inc = run;
        float t = __shfl_up_sync(0xffffffffu, inc, 1, 4); if (sub >= 1) inc += t;
        t = __shfl_up_sync(0xffffffffu, inc, 2, 4); if (sub >= 2) inc += t;
        const float excl = inc - run;
        const float P = __shfl_sync(0xffffffffu, inc, 3, 4);
        const float target = ((float)(rbits[j] >> 8) * (1.0f / 16777216.0f) + (0.5f / 16777216.0f)) * P;
        int cand = 99;
#pragma unroll
        for (int e = 3; e >= 0; --e)
            if (!pad[e] && excl + cum[e] >= target) cand = 4 * sub + e;
        cand = min(cand, __shfl_xor_sync(0xffffffffu, cand, 1));
        cand = min(cand, __shfl_xor_sync(0xffffffffu, cand, 2));
        const int bi = cand == 99 ? C - 1 : cand;           // target rounded above the last partial sum
        lab_out |= (uint32_t)bi << (8 * j);
    }
    if (live && a.next_x != nullptr) {
        // next UNet input rows: one-hot(C) | cond | zero pad (bf16).  Lane `sub` writes the whole row of voxel
        // v0 + sub (every lane of the quad holds all four labels), 16-byte stores, 128 contiguous bytes per quad
        const int bi = (int)((lab_out >> (8 * sub)) & 255u);
        uint16_t* row = reinterpret_cast<uint16_t*>(a.next_x) + (vb + sub) * a.Cin_pad;
        for (int c0 = 0; c0 < a.Cin_pad; c0 += 8) {
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int c = c0 + 2 * e;
                uint32_t bits = (c == (bi & ~1)) ? ((bi & 1) ? 0x3F800000u : 0x3F80u) : 0u;
                if (c + 1 >= C && cond != nullptr) {       // pair touches the condition channels
#pragma unroll
                    for (int z = 0; z < 2; ++z) {
                        const int cc = c + z - C;
                        if (cc >= 0 && cc < a.n_cond) bits |= (uint32_t)cond[(vb + sub) * a.n_cond + cc] << (16 * z);
                    }
                }
                w[e] = bits;
            }
            *reinterpret_cast<uint4*>(row + c0) = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
    if (live && sub == 0) *reinterpret_cast<uint32_t*>(a.labels_out + vb) = lab_out;
}

// ------------------------------------------------------------------------------------------
// K14  DDIM update
// ------------------------------------------------------------------------------------------
template <bool VEC>
__global__ void __launch_bounds__(256) ddim_kernel(const gg_ddim_args a) {
    const float a_t = __ldg(a.coef), a_prev = __ldg(a.coef + 1), sigma = __ldg(a.coef + 2), s1m = __ldg(a.coef + 3);
    const float sq_at = __fsqrt_rn(a_t), sq_ap = __fsqrt_rn(a_prev);
    const float c2 = __fsqrt_rn(__fsub_rn(__fsub_rn(1.0f, a_prev), __fmul_rn(sigma, sigma)));
    const float temp = a.temperature;
    auto one = [&](float x, float e, float n, float& xp, float& p0) {
        p0 = __fdiv_rn(__fsub_rn(x, __fmul_rn(s1m, e)), sq_at);
        const float dir = __fmul_rn(c2, e);
        const float nz = __fmul_rn(__fmul_rn(sigma, n), temp);
        xp = __fadd_rn(__fadd_rn(__fmul_rn(sq_ap, p0), dir), nz);
    };
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (VEC) {
        const int64_t i4 = i * 4;
        if (i4 >= a.n) return;
        float4 x = ldg_nc_f4(a.x + i4), e = ldg_nc_f4(a.e_t + i4);
        if (a.e_uncond) {
            const float4 u = ldg_nc_f4(a.e_uncond + i4);
            const float s = a.guidance_scale;
            e.x = __fadd_rn(u.x, __fmul_rn(s, __fsub_rn(e.x, u.x))); e.y = __fadd_rn(u.y, __fmul_rn(s, __fsub_rn(e.y, u.y)));
            e.z = __fadd_rn(u.z, __fmul_rn(s, __fsub_rn(e.z, u.z))); e.w = __fadd_rn(u.w, __fmul_rn(s, __fsub_rn(e.w, u.w)));
        }
        float4 n = a.noise ? ldg_nc_f4(a.noise + i4) : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 xp, p0;
        one(x.x, e.x, n.x, xp.x, p0.x); one(x.y, e.y, n.y, xp.y, p0.y);
        one(x.z, e.z, n.z, xp.z, p0.z); one(x.w, e.w, n.w, xp.w, p0.w);
        stg_na_f4(a.x_prev + i4, xp);
        if (a.pred_x0) stg_na_f4(a.pred_x0 + i4, p0);
    } else {
        if (i >= a.n) return;
        float xp, p0;
        float e = a.e_t[i];
        if (a.e_uncond) e = __fadd_rn(a.e_uncond[i], __fmul_rn(a.guidance_scale, __fsub_rn(e, a.e_uncond[i])));
        one(a.x[i], e, a.noise ? a.noise[i] : 0.f, xp, p0);
        a.x_prev[i] = xp;
        if (a.pred_x0) a.pred_x0[i] = p0;
    }
}

// ------------------------------------------------------------------------------------------
// PLMS: Adams-Bashforth combination of noise predictions (ldm/models/diffusion/plms.py:219-230), torch's fp32
// evaluation order with one rounding per operation
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) plms_eps_kernel(const gg_plms_args a) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    float e = a.e_t[i];
    if (a.e_uncond) e = __fadd_rn(a.e_uncond[i], __fmul_rn(a.guidance_scale, __fsub_rn(e, a.e_uncond[i])));
    if (a.e_cur) a.e_cur[i] = e;
    float r;
    if (a.order == 0) {
        r = __fdiv_rn(__fadd_rn(e, a.old1[i]), 2.0f);
    } else if (a.order == 1) {
        r = __fdiv_rn(__fsub_rn(__fmul_rn(3.0f, e), a.old1[i]), 2.0f);
    } else if (a.order == 2) {
        r = __fdiv_rn(__fadd_rn(__fsub_rn(__fmul_rn(23.0f, e), __fmul_rn(16.0f, a.old1[i])), __fmul_rn(5.0f, a.old2[i])), 12.0f);
    } else {
        r = __fsub_rn(__fmul_rn(55.0f, e), __fmul_rn(59.0f, a.old1[i]));
        r = __fadd_rn(r, __fmul_rn(37.0f, a.old2[i]));
        r = __fdiv_rn(__fsub_rn(r, __fmul_rn(9.0f, a.old3[i])), 24.0f);
    }
    a.e_prime[i] = r;
}

// ------------------------------------------------------------------------------------------
// ancestral DDPM update (ldm/models/diffusion/ddpm.py:1060-1120 p_mean_variance + p_sample)
//   x0 = sr x - srm1 e ; clip ; mean = c1 x0 + c2 x ; x_prev = mean + nz * exp(0.5 logvar) * (noise * T)
// coef[b] = (sr, srm1, c1, c2, logvar, nz) gathered per sample by the host wrapper
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ddpm_kernel(const gg_ddpm_args a) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.per_sample) return;
    const int b = blockIdx.y;
    const float* co = a.coef + 6 * b;
    const float sr = __ldg(co), srm1 = __ldg(co + 1), c1 = __ldg(co + 2), c2 = __ldg(co + 3), lv = __ldg(co + 4), nz = __ldg(co + 5);
    const float sd = __fmul_rn(nz, expf(__fmul_rn(0.5f, lv)));
    const int64_t j = (int64_t)b * a.per_sample + i;
    const float x = a.x[j], e = a.e_t[j];
    float x0 = __fsub_rn(__fmul_rn(sr, x), __fmul_rn(srm1, e));
    if (a.clip_denoised) x0 = fminf(fmaxf(x0, -1.0f), 1.0f);
    const float mean = __fadd_rn(__fmul_rn(c1, x0), __fmul_rn(c2, x));
    const float n = a.noise ? __fmul_rn(a.noise[j], a.temperature) : 0.f;
    a.x_prev[j] = __fadd_rn(mean, __fmul_rn(sd, n));
    if (a.x0_out) a.x0_out[j] = x0;
}

// ------------------------------------------------------------------------------------------
// stage bridge of the autoregressive CT generator (sample_diffusion.py:196-224)
//   labels -> nearest-neighbour zoomed float mask (label / 255);  per-slice-batch min-max normalise
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) labels_to_mask_kernel(const uint8_t* __restrict__ lab, float* __restrict__ out, int D, int H,
                                                             int W, int fh, int fw, float divisor) {
    const int Ho = H * fh, Wo = W * fw;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)D * Ho * Wo) return;
    const int w = (int)(i % Wo), h = (int)((i / Wo) % Ho), d = (int)(i / ((int64_t)Wo * Ho));
    out[i] = __fdiv_rn((float)lab[((int64_t)d * H + h / fh) * W + w / fw], divisor);
}

// general nearest-neighbour resampling: out[d, h, w] = lab[id[d], ih[h], iw[w]] / divisor with host-computed index tables
// (the tables carry scipy.ndimage.zoom(order=0)'s coordinate rule, see ops.zoom_index)
__global__ void __launch_bounds__(256) labels_gather_kernel(const uint8_t* __restrict__ lab, float* __restrict__ out, int H, int W, int Do,
                                                            int Ho, int Wo, const int32_t* __restrict__ id, const int32_t* __restrict__ ih,
                                                            const int32_t* __restrict__ iw, float divisor) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)Do * Ho * Wo) return;
    const int w = (int)(i % Wo), h = (int)((i / Wo) % Ho), d = (int)(i / ((int64_t)Wo * Ho));
    out[i] = __fdiv_rn((float)lab[((int64_t)__ldg(id + d) * H + __ldg(ih + h)) * W + __ldg(iw + w)], divisor);
}

__device__ __forceinline__ void block_minmax(float& mn, float& mx) {
    __shared__ float smn[8], smx[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) { smn[threadIdx.x >> 5] = mn; smx[threadIdx.x >> 5] = mx; }
    __syncthreads();
    mn = smn[0]; mx = smx[0];
#pragma unroll
    for (int k = 1; k < 8; ++k) { mn = fminf(mn, smn[k]); mx = fmaxf(mx, smx[k]); }
    __syncthreads();
}

__global__ void __launch_bounds__(256) minmax_partial_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ part) {
    float mn = INFINITY, mx = -INFINITY;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = __ldg(x + i);
        mn = fminf(mn, v); mx = fmaxf(mx, v);
    }
    block_minmax(mn, mx);
    if (threadIdx.x == 0) { part[2 * blockIdx.x] = mn; part[2 * blockIdx.x + 1] = mx; }
}

// y[b, :] (row stride y_bs) = (x[b, :] - min) / (max - min) with min / max over ALL of x (ds.min(), ds.max())
__global__ void __launch_bounds__(256) minmax_normalize_kernel(const float* __restrict__ x, const float* __restrict__ part, int nparts,
                                                               float* __restrict__ y, int64_t per, int64_t y_bs, int B) {
    float mn = INFINITY, mx = -INFINITY;
    for (int k = threadIdx.x; k < nparts; k += blockDim.x) { mn = fminf(mn, part[2 * k]); mx = fmaxf(mx, part[2 * k + 1]); }
    block_minmax(mn, mx);
    const float den = __fsub_rn(mx, mn);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < per * B; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / per, r = i - b * per;
        y[b * y_bs + r] = __fdiv_rn(__fsub_rn(__ldg(x + i), mn), den);
    }
}

// ------------------------------------------------------------------------------------------
// layout bridges
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) nchw_to_cl_kernel(const float* __restrict__ x1, int C1, const float* __restrict__ x2,
                                                         int C2, __nv_bfloat16* __restrict__ y, int Cpad, int N, int64_t V) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    const int n = blockIdx.y;
    const int64_t i = (int64_t)n * V + v;
    __nv_bfloat16* row = y + i * Cpad;
    for (int c0 = 0; c0 < Cpad; c0 += 8) {
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float f[2];
#pragma unroll
            for (int z = 0; z < 2; ++z) {
                const int c = c0 + 2 * e + z;
                float t = 0.f;
                if (c < C1) t = __ldg(x1 + ((int64_t)n * C1 + c) * V + v);
                else if (c < C1 + C2) t = __ldg(x2 + ((int64_t)n * C2 + (c - C1)) * V + v);
                f[z] = t;
            }
            w[e] = pack_bf16(f[0], f[1]);
        }
        *reinterpret_cast<uint4*>(row + c0) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

template <bool F32>
__global__ void __launch_bounds__(256) cl_to_nchw_kernel(const void* __restrict__ x, int Cs, float* __restrict__ y, int C,
                                                         int N, int64_t V, int softmax) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    const int n = blockIdx.y;
    const int64_t i = (int64_t)n * V + v;
    auto ld = [&](int c) -> float {
        if (F32) return __ldg(reinterpret_cast<const float*>(x) + i * Cs + c);
        return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(x)[i * Cs + c]);
    };
    if (softmax) {
        float m = -INFINITY;
        for (int c = 0; c < C; ++c) m = fmaxf(m, ld(c));
        float s = 0.f;
        for (int c = 0; c < C; ++c) s += expf(ld(c) - m);
        const float inv = 1.0f / s;
        for (int c = 0; c < C; ++c) y[((int64_t)n * C + c) * V + v] = expf(ld(c) - m) * inv;
    } else {
        for (int c = 0; c < C; ++c) y[((int64_t)n * C + c) * V + v] = ld(c);
    }
}

}  // namespace gg

using namespace gg;

extern "C" int gg_cat_posterior_sample(const gg_cat_args* a, gg_stream_t stream) {
    GG_REQUIRE(a != nullptr && a->x0 != nullptr, GG_ERR_BAD_ARG);
    GG_REQUIRE(a->B > 0 && a->V > 0, GG_ERR_BAD_ARG);
    GG_REQUIRE(a->mode >= GG_CAT_POSTERIOR && a->mode <= GG_CAT_PROBS_GIVEN, GG_ERR_BAD_ARG);
    const bool given = a->mode >= GG_CAT_SAMPLE_GIVEN;
    if (!given) GG_REQUIRE(a->xt != nullptr && a->coef != nullptr, GG_ERR_BAD_ARG);
    if (a->mode == GG_CAT_POSTERIOR || a->mode == GG_CAT_PROBS || a->mode == GG_CAT_PROBS_GIVEN) GG_REQUIRE(a->out != nullptr, GG_ERR_BAD_ARG);
    else GG_REQUIRE(a->out != nullptr || a->out_i64 != nullptr || a->labels != nullptr, GG_ERR_BAD_ARG);
    cudaStream_t s = as_stream(stream);
    switch (a->C) {
#define GG_CASE(c) case c: return launch_cat<c>(*a, s);
        GG_CASE(2) GG_CASE(3) GG_CASE(4) GG_CASE(5) GG_CASE(6) GG_CASE(7) GG_CASE(8) GG_CASE(9) GG_CASE(10)
        GG_CASE(11) GG_CASE(12) GG_CASE(13) GG_CASE(14) GG_CASE(15) GG_CASE(16) GG_CASE(17) GG_CASE(18)
        GG_CASE(19) GG_CASE(20) GG_CASE(21) GG_CASE(22) GG_CASE(23) GG_CASE(24)
#undef GG_CASE
        default: return GG_ERR_UNSUPPORTED;
    }
}

extern "C" int gg_cat_step_cl(const gg_cat_step_cl_args* a, gg_stream_t stream) {
    GG_REQUIRE(a != nullptr && a->logits && a->labels_in && a->labels_out && a->coef, GG_ERR_BAD_ARG);
    GG_REQUIRE(a->B > 0 && a->V > 0 && a->C >= 2, GG_ERR_BAD_ARG);
    GG_REQUIRE(a->Cpad == 16 && a->C <= 16, GG_ERR_UNSUPPORTED);
    GG_REQUIRE(a->mode == GG_CAT_SAMPLE || a->mode == GG_CAT_ARGMAX, GG_ERR_BAD_ARG);
    GG_REQUIRE(aligned(a->logits, 16), GG_ERR_ALIGNMENT);
    if (a->next_x) {
        GG_REQUIRE(a->Cin_pad % 8 == 0 && a->Cin_pad >= a->C + a->n_cond, GG_ERR_BAD_ARG);
        GG_REQUIRE(aligned(a->next_x, 16), GG_ERR_ALIGNMENT);
    }
    GG_REQUIRE(a->B <= 65535, GG_ERR_UNSUPPORTED);
    if (a->q == nullptr && a->probs_out == nullptr && a->mode == GG_CAT_SAMPLE && a->V % 4 == 0 && aligned(a->labels_in, 4) &&
        aligned(a->labels_out, 4)) {
        const dim3 fblocks((unsigned)((a->V / 4 + 63) / 64), (unsigned)a->B);
        cat_step_cl_fast_kernel<<<fblocks, 256, 0, as_stream(stream)>>>(*a);
        return launch_result();
    }
    const dim3 blocks((unsigned)((a->V + 63) / 64), (unsigned)a->B);
    cat_step_cl_kernel<<<blocks, 256, 0, as_stream(stream)>>>(*a);
    return launch_result();
}

extern "C" int gg_ddim_update(const gg_ddim_args* a, gg_stream_t stream) {
    GG_REQUIRE(a != nullptr && a->x && a->e_t && a->coef && a->x_prev && a->n > 0, GG_ERR_BAD_ARG);
    const bool vec = (a->n % 4 == 0) && aligned(a->x, 16) && aligned(a->e_t, 16) && aligned(a->x_prev, 16) &&
                     (!a->noise || aligned(a->noise, 16)) && (!a->pred_x0 || aligned(a->pred_x0, 16)) &&
                     (!a->e_uncond || aligned(a->e_uncond, 16));
    if (vec) {
        const unsigned blocks = (unsigned)((a->n / 4 + 255) / 256);
        ddim_kernel<true><<<blocks, 256, 0, as_stream(stream)>>>(*a);
    } else {
        const unsigned blocks = (unsigned)((a->n + 255) / 256);
        ddim_kernel<false><<<blocks, 256, 0, as_stream(stream)>>>(*a);
    }
    return launch_result();
}

extern "C" int gg_plms_eps(const gg_plms_args* a, gg_stream_t stream) {
    GG_REQUIRE(a != nullptr && a->e_t && a->e_prime && a->n > 0 && a->order >= 0 && a->order <= 3, GG_ERR_BAD_ARG);
    GG_REQUIRE(a->old1 != nullptr && (a->order < 2 || a->old2 != nullptr) && (a->order < 3 || a->old3 != nullptr), GG_ERR_BAD_ARG);
    const unsigned blocks = (unsigned)((a->n + 255) / 256);
    plms_eps_kernel<<<blocks, 256, 0, as_stream(stream)>>>(*a);
    return launch_result();
}

extern "C" int gg_ddpm_update(const gg_ddpm_args* a, gg_stream_t stream) {
    GG_REQUIRE(a != nullptr && a->x && a->e_t && a->coef && a->x_prev && a->B > 0 && a->per_sample > 0, GG_ERR_BAD_ARG);
    GG_REQUIRE(a->B <= 65535, GG_ERR_UNSUPPORTED);
    const dim3 grid((unsigned)((a->per_sample + 255) / 256), (unsigned)a->B);
    ddpm_kernel<<<grid, 256, 0, as_stream(stream)>>>(*a);
    return launch_result();
}

extern "C" int gg_labels_to_mask(const uint8_t* labels, float* mask, int32_t D, int32_t H, int32_t W, int32_t fh, int32_t fw,
                                 float divisor, gg_stream_t stream) {
    GG_REQUIRE(labels && mask && D > 0 && H > 0 && W > 0 && fh > 0 && fw > 0 && divisor != 0.f, GG_ERR_BAD_ARG);
    const int64_t n = (int64_t)D * H * fh * W * fw;
    labels_to_mask_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(labels, mask, D, H, W, fh, fw, divisor);
    return launch_result();
}

extern "C" int gg_labels_gather(const uint8_t* labels, float* mask, int32_t D, int32_t H, int32_t W, int32_t Do, int32_t Ho, int32_t Wo,
                                const int32_t* idx_d, const int32_t* idx_h, const int32_t* idx_w, float divisor, gg_stream_t stream) {
    GG_REQUIRE(labels && mask && idx_d && idx_h && idx_w && D > 0 && H > 0 && W > 0 && Do > 0 && Ho > 0 && Wo > 0 && divisor != 0.f, GG_ERR_BAD_ARG);
    const int64_t n = (int64_t)Do * Ho * Wo;
    labels_gather_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(labels, mask, H, W, Do, Ho, Wo, idx_d, idx_h, idx_w, divisor);
    return launch_result();
}

extern "C" int gg_minmax_normalize(const float* x, float* y, float* scratch, int32_t B, int64_t per_sample, int64_t y_batch_stride,
                                   gg_stream_t stream) {
    GG_REQUIRE(x && y && scratch && B > 0 && per_sample > 0 && y_batch_stride >= per_sample, GG_ERR_BAD_ARG);
    const int64_t n = (int64_t)B * per_sample;
    const int nparts = (int)std::min<int64_t>(512, (n + 255) / 256);
    minmax_partial_kernel<<<nparts, 256, 0, as_stream(stream)>>>(x, n, scratch);
    int st = launch_result();
    if (st != GG_OK) return st;
    const unsigned blocks = (unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)num_sms() * 8);
    minmax_normalize_kernel<<<blocks, 256, 0, as_stream(stream)>>>(x, scratch, nparts, y, per_sample, y_batch_stride, B);
    return launch_result();
}

extern "C" int gg_nchw_to_cl(const float* x1, int32_t C1, const float* x2, int32_t C2, void* y_cl, int32_t Cpad, int32_t N,
                             int64_t V, gg_stream_t stream) {
    GG_REQUIRE(x1 && y_cl && C1 > 0 && C2 >= 0 && (C2 == 0 || x2) && N > 0 && V > 0, GG_ERR_BAD_ARG);
    GG_REQUIRE(Cpad % 8 == 0 && Cpad >= C1 + C2, GG_ERR_BAD_ARG);
    GG_REQUIRE(aligned(y_cl, 16), GG_ERR_ALIGNMENT);
    GG_REQUIRE(N <= 65535, GG_ERR_UNSUPPORTED);
    const dim3 blocks((unsigned)((V + 255) / 256), (unsigned)N);
    nchw_to_cl_kernel<<<blocks, 256, 0, as_stream(stream)>>>(x1, C1, x2, C2, reinterpret_cast<__nv_bfloat16*>(y_cl), Cpad, N, V);
    return launch_result();
}

extern "C" int gg_cl_to_nchw(const void* x_cl, int32_t Cstride, int32_t src_is_f32, float* y, int32_t C, int32_t N, int64_t V,
                             int32_t softmax, gg_stream_t stream) {
    GG_REQUIRE(x_cl && y && C > 0 && Cstride >= C && N > 0 && V > 0, GG_ERR_BAD_ARG);
    GG_REQUIRE(N <= 65535, GG_ERR_UNSUPPORTED);
    const dim3 blocks((unsigned)((V + 255) / 256), (unsigned)N);
    if (src_is_f32) cl_to_nchw_kernel<true><<<blocks, 256, 0, as_stream(stream)>>>(x_cl, Cstride, y, C, N, V, softmax);
    else cl_to_nchw_kernel<false><<<blocks, 256, 0, as_stream(stream)>>>(x_cl, Cstride, y, C, N, V, softmax);
    return launch_result();
}
