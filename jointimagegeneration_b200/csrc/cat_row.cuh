// cat_row.cuh -- the per-voxel categorical reverse step for ONE voxel held by ONE thread
//
//   softmax(head logits)  ->  theta_post(x_{t-1} | x_t, p(x0))  ->  clamp  ->  categorical draw
//   (ccdm/ddpm/models/diffusion_denoising.py:203-224, one_hot_categorical.py:25-50, unet.py:720)
//
// Used by the sampler epilogue of the head convolution (conv_roll.cu): there a thread owns an accumulator row, i.e.
// all class logits of its voxel.  The arithmetic mirrors cat_step_cl_fast_kernel (pervoxel.cu) OPERATION FOR
// OPERATION -- that kernel spreads the classes of a voxel over a quad of lanes (4 classes each) and reduces with
// butterflies; here the same partial sums are formed in the same association order -- so both give the same label
// for the same logits, label and random word (checked bit for bit in tests/test_gpu_kernels.py).
#pragma once
#include "common.cuh"

namespace gg {

struct CatRowCoef { float al, g, k, h, hU, inv_lab, inv_oth; };

__device__ __forceinline__ CatRowCoef cat_row_coef(float al, float g, int C) {
    CatRowCoef c;
    c.al = al; c.g = g;
    c.k = (1.0f - al) / (float)C; c.h = (1.0f - g) / (float)C;
    c.hU = c.h * (al + (float)C * c.k);
    c.inv_lab = __fdividef(1.0f, fmaf(g, al + c.k, c.hU));
    c.inv_oth = __fdividef(1.0f, fmaf(g, c.k, c.hU));
    return c;
}

// lg: 16 logits (classes >= C ignored); lab: x_t's class; rbits: the voxel's 32 random bits.  Returns the drawn class.
__device__ __forceinline__ int cat_row_draw(const float (&lg)[16], int C, int lab, const CatRowCoef& cf, float clamp_min, uint32_t rbits) {
    float l[16];
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < 16; ++c) { l[c] = c >= C ? -INFINITY : lg[c]; m = fmaxf(m, l[c]); }
    float ex[16], r[16], u[16], se4[4], R4[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        float se = 0.f, R = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int c = 4 * s + e;
            const bool is_lab = c == lab;
            ex[c] = __expf(l[c] - m);                        // 0 for padding classes
            se += ex[c];
            u[c] = (is_lab ? cf.al : 0.f) + cf.k;
            r[c] = ex[c] * (is_lab ? cf.inv_lab : cf.inv_oth);
            R += r[c];
        }
        se4[s] = se; R4[s] = R;
    }
    const float se = (se4[0] + se4[1]) + (se4[2] + se4[3]);          // the quad butterfly: xor 1, then xor 2
    const float R = (R4[0] + R4[1]) + (R4[2] + R4[3]);
    const float hR = cf.h * R, floor_p = clamp_min * se;
    float cum[16], run4[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        float run = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int c = 4 * s + e;
            const float pe = c >= C ? 0.f : fmaxf(u[c] * fmaf(cf.g, r[c], hR), floor_p);
            run += pe;
            cum[c] = run;
        }
        run4[s] = run;
    }
    // the quad's inclusive scan (shfl_up by 1, then by 2) in its association order
    float inc[4];
    inc[0] = run4[0];
    inc[1] = run4[1] + run4[0];
    inc[2] = (run4[2] + run4[1]) + run4[0];
    inc[3] = (run4[3] + run4[2]) + (run4[1] + run4[0]);
    const float P = inc[3];
    const float target = ((float)(rbits >> 8) * (1.0f / 16777216.0f) + (0.5f / 16777216.0f)) * P;
    int cand = 99;
#pragma unroll
    for (int s = 3; s >= 0; --s) {
        const float excl = inc[s] - run4[s];
#pragma unroll
        for (int e = 3; e >= 0; --e) {
            const int c = 4 * s + e;
            if (c < C && excl + cum[c] >= target) cand = c;
        }
    }
    return cand == 99 ? C - 1 : cand;                         // target rounded above the last partial sum
}

}  // namespace gg
