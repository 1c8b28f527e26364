/*
 * guidegen_sm100.h -- C ABI of libguidegen_sm100.so
 *
 * B200 (sm_100a) kernels for the reverse-diffusion denoising step of GuideGen
 * (OvO1111/JointImageGeneration).  The reference has no FFI/plugin layer: its "backend"
 * is the set of torch library calls made by the Python classes on the sampling path
 * (SURVEY.md section 2.3, K1..K16).  Each entry point below replaces one of those call
 * sites; the reference file:line it stands in for is cited on the declaration
 * (paths relative to the reference root).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless stated;
 *   - the library never allocates, frees or retains device memory;
 *   - every call only ENQUEUES work on `stream` (a cudaStream_t); no host sync, no
 *     allocation -> all calls are legal inside CUDA-graph stream capture;
 *   - return value: 0 = ok, <0 = gg_status (bad argument / unsupported shape / alignment),
 *     >0 = a cudaError_t raised by the launch.  There is NO fallback path of any kind.
 *   - "CL" = channels-last bf16 activation [N, D, H, W, C] (2-D data: D = 1), C % 8 == 0.
 */
#ifndef GUIDEGEN_SM100_H
#define GUIDEGEN_SM100_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* gg_stream_t; /* cudaStream_t */

enum gg_status {
    GG_OK = 0,
    GG_ERR_BAD_ARG = -1,
    GG_ERR_UNSUPPORTED = -2,
    GG_ERR_ALIGNMENT = -3,
    GG_ERR_NO_DEVICE = -4,
    GG_ERR_DRIVER = -5
};

/* library version (major*100+minor) and a static description of status codes */
int gg_version(void);
const char* gg_status_string(int status);
/* 0 when the current device is compute capability 10.x, GG_ERR_NO_DEVICE otherwise */
int gg_device_check(void);
/* sizeof() of the argument structs, in header order (gg_cat_args, gg_cat_step_cl_args, gg_ddim_args, gg_plms_args,
 * gg_ddpm_args, gg_gn_finalize_args, gg_conv_src, gg_conv_args, gg_attn_args, gg_cat_epilogue): lets a binding in another language
 * check its mirror of this header without a GPU.  Writes min(n, 9) entries, returns 9. */
int gg_abi_sizes(int32_t* out, int n);
/* number of kernels launched by this library since load / since last reset (host counter) */
uint64_t gg_launch_count(void);
void gg_launch_count_reset(void);

/* ------------------------------------------------------------------------------------------
 * K11-K13  categorical posterior + clamp + categorical draw
 *   ccdm/ddpm/models/diffusion_denoising.py:105-139  DiffusionModel.theta_post_prob
 *   ccdm/ddpm/models/diffusion_denoising.py:216-224  clamp(1e-12), sample / max_prob / prob
 *   ccdm/ddpm/models/one_hot_categorical.py:25-54    OneHotCategoricalBCHW
 * Layout is the reference's: fp32 [B, C, V] (V = D*H*W, class axis second).
 * ---------------------------------------------------------------------------------------- */
enum gg_cat_mode {
    GG_CAT_POSTERIOR = 0,   /* out = theta_post_prob(xt, x0, t)                       (fp32 probs) */
    GG_CAT_SAMPLE = 1,      /* posterior -> clamp -> normalise -> argmax(p/q) -> one-hot fp32     */
    GG_CAT_ARGMAX = 2,      /* posterior -> clamp -> normalise -> argmax -> one-hot (fp32|int64)  */
    GG_CAT_PROBS = 3,       /* posterior -> clamp -> normalise -> probs                            */
    GG_CAT_SAMPLE_GIVEN = 4,/* x0 IS the distribution: normalise -> argmax(p/q) -> one-hot
                               (OneHotCategoricalBCHW(probs).sample(), one_hot_categorical.py:30) */
    GG_CAT_ARGMAX_GIVEN = 5,/* x0 IS the distribution: normalise -> argmax -> one-hot             */
    GG_CAT_PROBS_GIVEN = 6  /* x0 IS the distribution: normalise -> probs (prob_sample, :48-50)   */
};

typedef struct {
    const float* x0;        /* [B, C, V] predicted x0 probabilities (or the distribution itself) */
    const float* xt;        /* [B, C, V] current state (one-hot or soft); unused for *_GIVEN      */
    const float* q;         /* [B*V, C]  Exp(1) noise rows (channels-last, the block
                               torch.multinomial draws); NULL -> in-kernel Philox(seed, offset)   */
    const float* coef;      /* [B, 2]    (alpha_t, cumalpha_{t-1}) per sample; unused for *_GIVEN */
    float* out;             /* [B, C, V] fp32 result, may be NULL when only labels are wanted     */
    int64_t* out_i64;       /* [B, C, V] int64 one-hot (F.one_hot dtype) or NULL                  */
    uint8_t* labels;        /* [B, V]    drawn class index or NULL                                */
    int32_t B, C;           /* 2 <= C <= 32                                                        */
    int64_t V;
    float clamp_min;        /* 1e-12 in the reference; <= 0 disables                               */
    int32_t mode;           /* gg_cat_mode                                                         */
    uint64_t seed, offset;  /* Philox key / counter base when q == NULL                            */
} gg_cat_args;

int gg_cat_posterior_sample(const gg_cat_args* a, gg_stream_t stream);

/* Fused sampler-loop form of the same step (channels-last, stays on the device between steps):
 * reads the head conv's fp32 logits [V_total, Cpad] (softmax fused, unet.py:720), the current
 * labels uint8 [V_total], draws the next labels and writes the next UNet input: CL bf16
 * [V_total, Cin_pad] = one-hot(C) | condition channel(s) | zero padding  (unet.py:775). */
typedef struct {
    const float* logits;    /* [Vt, Cpad] fp32 */
    const uint8_t* labels_in;   /* [Vt] */
    const float* q;         /* [Vt, C] Exp(1) or NULL -> Philox */
    const float* coef;      /* [B, 2] */
    const void* cond;       /* bf16 [Vt, n_cond] condition channels or NULL (zeros)               */
    uint8_t* labels_out;    /* [Vt] */
    void* next_x;           /* bf16 [Vt, Cin_pad] or NULL */
    float* probs_out;       /* optional fp32 [B, C, V] posterior probabilities (debug/confidence) */
    int32_t B, C, Cpad, n_cond, Cin_pad;
    int64_t V;              /* voxels per sample; Vt = B*V */
    float clamp_min;
    int32_t mode;           /* GG_CAT_SAMPLE or GG_CAT_ARGMAX */
    uint64_t seed, offset;
    int64_t vox_base;       /* global index of this call's first voxel (depth-slab shards): the Philox
                               counter is keyed on the GLOBAL voxel index, so draws do not depend on the
                               number of slabs / ranks                                              */
} gg_cat_step_cl_args;

int gg_cat_step_cl(const gg_cat_step_cl_args* a, gg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K14  DDIM update        latentdiffusion/ldm/models/diffusion/ddim.py:190-205 (p_sample_ddim)
 *   pred_x0 = (x - sqrt(1-a_t) e) / sqrt(a_t);  x_prev = sqrt(a_prev) pred_x0
 *             + sqrt(1 - a_prev - sigma^2) e + sigma * noise * temperature      (fp32, no FMA)
 * coef (device) = [a_t, a_prev, sigma_t, sqrt_one_minus_a_t] as the f32 values torch.full makes.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    const float* x;
    const float* e_t;
    const float* noise;     /* NULL: term omitted (sigma == 0, value identical to the reference) */
    const float* coef;      /* [4] device */
    float* x_prev;
    float* pred_x0;         /* may be NULL */
    int64_t n;              /* elements */
    float temperature;
    const float* e_uncond;  /* NULL, or unconditional eps: e = e_u + guidance_scale * (e_t - e_u)
                               (classifier-free guidance, ddim.py:175-179) before the update   */
    float guidance_scale;
} gg_ddim_args;

int gg_ddim_update(const gg_ddim_args* a, gg_stream_t stream);

/* PLMS noise-prediction combination   latentdiffusion/ldm/models/diffusion/plms.py:178-233 (p_sample_plms).
 * The pseudo linear multistep sampler replaces e_t in the DDIM update (eta = 0) by an Adams-Bashforth
 * combination of the current and up to three previous predictions (:219-230), evaluated in fp32 in the
 * reference's order, one rounding per operation:
 *   order 0 (first step, old1 = the prediction at x_prev, t_next):  (e + old1) / 2
 *   order 1: (3 e - old1) / 2      order 2: (23 e - 16 old1 + 5 old2) / 12
 *   order 3: (55 e - 59 old1 + 37 old2 - 9 old3) / 24
 * With e_uncond the classifier-free-guidance mix (:183-188) is applied to e_t first; e_cur receives that
 * guided prediction (what the sampler appends to old_eps), e_prime the combination (input of gg_ddim_update). */
typedef struct {
    const float* e_t;
    const float* e_uncond;  /* optional */
    const float* old1;      /* newest previous prediction (order >= 1), or e_t_next (order 0) */
    const float* old2;
    const float* old3;
    float* e_cur;           /* optional; may alias e_t */
    float* e_prime;
    int64_t n;
    int32_t order;          /* 0..3 */
    float guidance_scale;
} gg_plms_args;

int gg_plms_eps(const gg_plms_args* a, gg_stream_t stream);

/* Ancestral DDPM update   latentdiffusion/ldm/models/diffusion/ddpm.py:1060-1120 (p_mean_variance, p_sample),
 * :215-230 (predict_start_from_noise, q_posterior); used by sample_diffusion.py --vanilla_sample.
 * coef fp32 [B, 6] = (sqrt_recip_acp[t], sqrt_recipm1_acp[t], posterior_mean_coef1[t], posterior_mean_coef2[t],
 * posterior_log_variance_clipped[t], nonzero_mask) per sample. */
typedef struct {
    const float* x;
    const float* e_t;
    const float* noise;     /* NULL: no noise term */
    const float* coef;      /* [B, 6] device */
    float* x_prev;
    float* x0_out;          /* optional predicted x0 (after clipping) */
    int32_t B;
    int64_t per_sample;     /* elements per sample */
    float temperature;
    int32_t clip_denoised;
} gg_ddpm_args;

int gg_ddpm_update(const gg_ddpm_args* a, gg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Stage bridge of the autoregressive CT generator     latentdiffusion/sample_diffusion.py:196-224
 *   :199-201  label volume -> nearest-neighbour zoom (order 0) -> / 255  (whole-mask conditioning)
 *   :221-222  samples[:, :, m] = (ds - ds.min()) / (ds.max() - ds.min())   (min / max over the whole slice batch)
 * ---------------------------------------------------------------------------------------- */
/* mask[d, h, w] = labels[d, h / fh, w / fw] / divisor   (uint8 [D, H, W] -> fp32 [D, H*fh, W*fw]) */
int gg_labels_to_mask(const uint8_t* labels, float* mask, int32_t D, int32_t H, int32_t W, int32_t fh, int32_t fw, float divisor,
                      gg_stream_t stream);
/* mask[d, h, w] = labels[idx_d[d], idx_h[h], idx_w[w]] / divisor  (uint8 [D, H, W] -> fp32 [Do, Ho, Wo]); the int32 device
 * index tables carry the resampling rule -- scipy.ndimage.zoom(order=0) as sample_diffusion.py:200 calls it maps output
 * index o to input floor(o * (n_in - 1) / (n_out - 1) + 0.5), which is NOT block replication */
int gg_labels_gather(const uint8_t* labels, float* mask, int32_t D, int32_t H, int32_t W, int32_t Do, int32_t Ho, int32_t Wo,
                     const int32_t* idx_d, const int32_t* idx_h, const int32_t* idx_w, float divisor, gg_stream_t stream);
/* y[b, 0:per_sample] = (x[b] - min(x)) / (max(x) - min(x)); x dense fp32 [B, per_sample]; row b of y starts at
 * y + b * y_batch_stride (a slice of a [B, 1, D, H, W] volume); scratch: fp32 [1024] device workspace */
int gg_minmax_normalize(const float* x, float* y, float* scratch, int32_t B, int64_t per_sample, int64_t y_batch_stride,
                        gg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Layout bridges at the drop-in boundary (fp32 NC* <-> CL bf16)
 *   unet.py:775 th.cat([x, input_condition], 1); ddpm.py:1419 torch.cat([x] + c_concat, 1)
 * ---------------------------------------------------------------------------------------- */
/* y[n, v, 0:C1+C2] = cat(x1[n, :, v], x2[n, :, v]); channels [C1+C2, Cpad) zero-filled */
int gg_nchw_to_cl(const float* x1, int32_t C1, const float* x2, int32_t C2, void* y_cl, int32_t Cpad,
                  int32_t N, int64_t V, gg_stream_t stream);
/* y[n, c, v] = x_cl[n, v, c] for c < C  (bf16 or fp32 source selected by src_is_f32) */
int gg_cl_to_nchw(const void* x_cl, int32_t Cstride, int32_t src_is_f32, float* y, int32_t C, int32_t N, int64_t V,
                  int32_t softmax, gg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K6  GroupNorm32(32, C) (+SiLU)       ccdm .../unet_openai/nn.py:17-19,100; unet.py:189-190
 *     fp32 statistics, eps given by caller (1e-5; SpatialTransformer's Normalize uses 1e-6).
 *     Two-source form covers GroupNorm over th.cat([h, skip], 1) (unet.py:812 + :189).
 * Step 1: per-(n, channel) partial sums over spatial chunks (deterministic, no atomics)
 * Step 2: finalize -> per-(n, channel) scale/shift          (the slab all-reduce plugs in here)
 * Step 3: y = act(x * scale + shift) written as ONE CL tensor with C1+C2 channels
 * ---------------------------------------------------------------------------------------- */
/* number of spatial chunks gg_gn_partial will use for S positions of C channels */
int32_t gg_gn_num_chunks(int64_t S, int32_t C);
/* partial: fp32 [N, nchunks, C, 2] (sum, sum of squares) */
int gg_gn_partial(const void* x_cl, int32_t N, int64_t S, int32_t C, float* partial, gg_stream_t stream);
typedef struct {
    const float* partial1; int32_t C1; int32_t nchunks1;
    const float* partial2; int32_t C2; int32_t nchunks2;   /* NULL / 0 when single source */
    const float* gamma; const float* beta;                 /* [C1+C2] */
    float* scale_shift;                                    /* out fp32 [N, C1+C2, 2]       */
    int32_t N; int32_t groups; int64_t S; float eps;
    /* Depth-slab mode over NVLink peer memory (slab_world > 1; see gg_peer_exchange): the partial rows are this RANK's,
     * S is the whole volume's position count.  Phase 1 stores this rank's (sum, sum sq) per (sample, group) -- fp64,
     * 16 bytes each -- into entry [slab_rank] of every rank's table slab_tables[q] ([world][N][groups][2] fp64, inside
     * the peer arenas) and raises *slab_epoch in slab_flag_out[q]; phase 2 waits for slab_flag_in[q] and combines the
     * entries in rank order.  slab_phase = 3: both (two launches). */
    int32_t slab_world, slab_rank, slab_phase;
    void* slab_tables[8];
    uint32_t* slab_flag_out[8];
    const uint32_t* slab_flag_in[8];
    const uint32_t* slab_epoch;
    unsigned int* slab_done_counter;
} gg_gn_finalize_args;
int gg_gn_finalize(const gg_gn_finalize_args* a, gg_stream_t stream);
int gg_gn_apply(const void* x1_cl, int32_t C1, const void* x2_cl, int32_t C2, const float* scale_shift,
                void* y_cl, int32_t N, int64_t S, int32_t silu, gg_stream_t stream);
/* The same GroupNorm (+SiLU) in ONE launch (F.group_norm + F.silu of nn.py:17-19 / util.py:214-216 over th.cat([x1, x2], 1)): a
 * cluster of 1..8 CTAs per sample, channel sums exchanged through distributed shared memory, fp64 combine in rank order.  When
 * a sample fits the shared memory of the cluster (gg_gn_fused_resident() > 0) every CTA loads its slice once with bulk async
 * copies and both passes read shared memory; otherwise the slice is streamed from L2 twice.  C1 + C2 <= 2048.  Same apply
 * formula and rounding points as gg_gn_apply; the statistics differ from gg_gn_partial + gg_gn_finalize only in summation order. */
int gg_gn_fused(const void* x1_cl, int32_t C1, const void* x2_cl, int32_t C2, const float* gamma, const float* beta, void* y_cl,
                int32_t N, int64_t S, int32_t groups, float eps, int32_t silu, gg_stream_t stream);
/* Cluster size (1..8) gg_gn_fused uses for a sample of S positions x C channels when the sample fits the shared memory of one
 * cluster (the slice of every CTA is loaded once by bulk async copies; statistics and normalisation read shared memory), or 0
 * when it does not fit and gg_gn_fused streams the slice from L2 twice instead (slower than the three launches: callers
 * check this before choosing the one-launch form). */
int32_t gg_gn_fused_resident(int64_t S, int32_t C);


/* ------------------------------------------------------------------------------------------
 * K1-K5,K15  implicit-GEMM convolution on tcgen05 tensor cores (bf16 x bf16 -> fp32 in TMEM)
 *   nn.Conv3d / nn.Conv2d / nn.Conv1d / nn.Linear call sites:
 *   unet.py:191,217 (ResBlock 3^d), :135-139 (Downsample stride 2), :104 (Upsample conv),
 *   :228 (1x1 skip), :292,300 (qkv / proj_out 1x1), :522 (input conv), :719 (output conv);
 *   openaimodel.py:207,233,107,154,244,522,688; ldm/modules/attention.py:162-169,239-259
 * A operand = CL activations read by TMA, one 5-D box per (filter tap, 64-channel chunk): no
 * im2col buffer; zero padding and channel padding = TMA out-of-bounds fill.  Up to 4 sources
 * are summed into one accumulator: sources with centre_only = 0 see the full kd x kh x kw
 * filter (two of them = conv over th.cat([h, skip], 1), K15), sources with centre_only = 1
 * contribute a 1x1 term (the ResBlock's skip_connection fused into its second conv).
 * Tap (a, b, c) reads input position  stride * o + (od + a, oh + b, ow + c).
 * Packed weight matrix: bf16 [Cout, Ktot] row-major, Ktot = gg_conv_packed_k(); columns are
 * ordered  source -> tap (a, b, c row-major) -> 64-channel chunk -> channel, channels of the
 * last chunk of a source zero-padded to 64.
 * Epilogue: + bias[c] + emb[n, c] + residual[n, pos, c]; bf16 or fp32 CL output through
 * explicit element strides (so a strided view, e.g. one parity class of an upsampled
 * grid, can be written).  bias / emb rows must be readable up to Cout rounded up to 8, and
 * the output row must hold that many channels (padding channels receive bias only).
 * ---------------------------------------------------------------------------------------- */
/* Sampler epilogue of the output head conv (gg_conv_args.cat; depth-rolling kernel, Cout <= 16): the accumulator row of
 * a voxel holds all its class logits, so the whole reverse step of ccdm/ddpm/models/diffusion_denoising.py:203-224 --
 * Softmax (unet.py:720) -> theta_post_prob (:105-139) -> clamp (:216) -> OneHotCategoricalBCHW.sample()
 * (one_hot_categorical.py:25-31) -> next network input th.cat([x, condition]) (unet.py:775) -- runs in the conv's epilogue:
 * no logits tensor is written or read back and no separate per-voxel kernel is launched.  Same arithmetic and the same
 * Philox stream as gg_cat_step_cl's production path (labels are identical for identical accumulators). */
typedef struct {
    const uint8_t* labels_in;   /* [N*V] x_t as class indices                                          */
    uint8_t* labels_out;        /* [N*V] drawn x_{t-1} (8-byte aligned)                                */
    void* next_x;               /* bf16 [N*V, Cin_pad] = one-hot(C) | cond | zero padding, or NULL     */
    const void* cond;           /* bf16 [N*V, n_cond] condition channels or NULL (zeros)               */
    const float* coef;          /* [N, 2] (alpha_t, cumalpha_{t-1}) per sample                         */
    int32_t C, n_cond, Cin_pad;
    int32_t mode;               /* GG_CAT_SAMPLE                                                       */
    float clamp_min;
    uint64_t seed, offset;      /* Philox key / counter base                                           */
    int64_t vox_base;           /* global index of this launch's first voxel (multiple of 4)           */
} gg_cat_epilogue;

typedef struct {
    const void* x;          /* CL bf16 [N, D, H, W, C] */
    int32_t C;              /* multiple of 8 */
    int32_t centre_only;    /* 1: contributes only a 1x1 (centre) term */
    int32_t d_shift;        /* added to the depth coordinate of every read of this source: a depth slab stored
                               with leading halo planes is read at  d_out + tap + d_shift              */
    int32_t reserved;
} gg_conv_src;

typedef struct {
    gg_conv_src src[4];
    int32_t nsrc;
    int32_t N, D, H, W;               /* input extents (D = 1 for 2-D data, D = H = 1 for tokens)   */
    int32_t dims;                     /* 1, 2 or 3: spatial dims that are strided when stride = 2   */
    int32_t kd, kh, kw;               /* taps per dim (1..3)                                         */
    int32_t od, oh, ow;               /* input offset of tap 0 (-1 for a padded 3-tap filter)        */
    int32_t stride;                   /* 1 or 2 (stride 2 needs 3 taps, offset -1 in strided dims)   */
    int32_t Do, Ho, Wo;               /* output extents                                              */
    const void* w_packed;             /* bf16 [Cout, Ktot]                                           */
    const float* bias;                /* [>= Cout8] or NULL                                          */
    const float* emb;                 /* fp32 [N, emb_stride] per-sample additive term or NULL       */
    int32_t emb_stride;
    const void* residual;             /* CL bf16 [N, Do, Ho, Wo, res_stride] or NULL                 */
    int32_t res_stride;
    void* y;                          /* output                                                      */
    int64_t y_sn, y_sd, y_sh, y_sw;   /* element strides of y for n, d, h, w (multiples of 8)        */
    int32_t y_is_f32;                 /* 0: bf16, 1: fp32                                            */
    int32_t Cout;
    int32_t block_n;                  /* 0 = gg_conv_pick_block_n(Cout); else multiple of 16 <= 256  */
    int32_t brick[4];                 /* 0s = auto; else (bn, bd, bh, bw) with product 128           */
    /* Fused GroupNorm statistics (nn.py:17-19 on the conv's OUTPUT): when gn_partial != NULL the epilogue also
     * writes per-channel (sum, sum of squares) of the stored (bf16-rounded) outputs, one row per (M tile, epilogue
     * warp): fp32 [N, gn_nchunks_total, Cout8, 2], rows [gn_chunk_base, gn_chunk_base + gg_conv_stats_chunks()).
     * Same layout gg_gn_partial produces, so gg_gn_finalize consumes it directly.  Output planes d < stats_d_min
     * are excluded (halo plane of a depth-slab strided conv).  Unsupported (status -2) when an M tile spans
     * samples, i.e. when gg_conv_stats_chunks() returns 0. */
    float* gn_partial;
    int32_t gn_chunk_base, gn_nchunks_total, stats_d_min;
    /* 0: one A tile per (tap, chunk) (any stride / shape).  1: "halo brick" kernel for stride-1 filters: the
     * 16 x 8 output brick's input window is loaded ONCE per 64-channel chunk and every tap reads it in place
     * (shifted UMMA descriptors) -- 6x fewer A-operand bytes through shared memory.  Packed-weight K order for
     * algo 1 is  source -> 64-channel chunk -> tap -> channel.  Algo 1 runs as CTA PAIRS (cta_group::2, one
     * M = 256 MMA stream over two w-adjacent bricks, each CTA holding half of the weight rows) when the brick
     * count along w is even; 2 = halo kernel, single CTAs only; 3 = halo kernel, pairs forced (tests).
     * 4 = "depth-rolling" kernel for 3x3x3 stride-1 filters with 3 * Cout <= 256 (conv_roll.cu): walks input
     * depth planes and stacks the three depth taps along N, so each plane is loaded once and each A slice feeds
     * 3 Cout accumulator columns (same packed-weight K order as algo 1; always CTA pairs).
     * With algo 1..4 (bf16 outputs) gn_partial holds one row per (CTA, epilogue warp): per-tile shuffle-reduced column sums. */
    int32_t algo;
    /* Split-K (algo 0): layers whose output has fewer tiles than the GPU has SMs but a long reduction (the
     * low-resolution 3^d convs) are cut into split_k K ranges; each (tile, range) work item writes raw fp32
     * partial sums into workspace [split_k, N*Do*Ho*Wo, Cout8] (caller-owned) and a second launch sums them
     * and applies bias / emb / residual.  split_k <= 1: off. */
    int32_t split_k;
    float* workspace;
    /* Fused GroupNorm + SiLU on the INPUT (algo 1..4): when src_ss[i] is not NULL, source i is the raw,
     * un-normalised activation and the kernel applies  silu?(x * scale + shift)  to each halo plane in shared
     * memory before the MMAs read it (the separate gg_gn_apply pass and its tensor disappear).  src_ss[i] points
     * at the (scale, shift) pair of the source's first channel for sample 0 -- gg_gn_finalize's output, offset by
     * the source's position in a concatenated norm -- and ss_stride is the number of floats between samples.
     * Positions outside the tensor stay zero (the reference zero-pads the normalised tensor): in h / w by bounds,
     * in depth for local planes z outside [xf_z_lo, xf_z_hi) (0..D unsplit; a depth slab widens it by the halo
     * planes that hold a neighbour's data). */
    const float* src_ss[4];
    int32_t ss_stride, xf_silu, xf_z_lo, xf_z_hi;
    /* NULL, or (algo 4, Cout <= 16): replace the output store by the sampler epilogue above; y is not written. */
    const gg_cat_epilogue* cat;
    /* split_k > 1: NULL = the partial tiles are summed by a second (reduce) launch; else gg_conv_num_tiles() device words,
     * ZERO before the first launch and owned by this conv (not reused by other launches): the split of a tile that finishes
     * last sums the partials in split order and writes the output itself -- one launch, same result. */
    uint32_t* split_counters;
} gg_conv_args;

/* N tile (accumulator columns) the kernel uses for a given Cout */
int32_t gg_conv_pick_block_n(int32_t Cout);
/* output tiles (128 positions x N tile) algo 0 will process for `a`: what split-K decisions are based on */
int32_t gg_conv_num_tiles(const gg_conv_args* a);
/* rows of gn_partial one launch of `a` fills per sample (algo 0: 4 per M tile; algo 1..4: one per (CTA, epilogue
 * warp), accumulated in a fixed order into rows the launch zeroes first), 0 if fused statistics are unsupported */
int32_t gg_conv_stats_chunks(const gg_conv_args* a);
/* K extent (columns) of the packed weight matrix for the sources / taps in `a` */
int64_t gg_conv_packed_k(const gg_conv_args* a);
int gg_conv_fwd(const gg_conv_args* a, gg_stream_t stream);

/* K5  nearest x2 upsample in every spatial dim (unet.py:108-113), CL bf16 */
int gg_upsample2x(const void* x_cl, void* y_cl, int32_t N, int32_t D, int32_t H, int32_t W, int32_t C, int32_t dims,
                  gg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K8/K9  attention, flash style (no T x T buffer), head dim 32 or 64, fp32 softmax
 *   unet.py:343-360 QKVAttentionLegacy (scale ch^-1/4 on q and on k)
 *   ldm/modules/attention.py:170-193 CrossAttention (scale d^-1/2 on the product)
 * Generic strided heads: element (b, t, h, i) of q is q[b*q_bs + t*q_rs + h*q_hs + i].
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    const void* q; const void* k; const void* v; void* o;      /* bf16 */
    int64_t q_bs, k_bs, v_bs, o_bs;                            /* batch strides (elements)  */
    int32_t q_rs, k_rs, v_rs, o_rs;                            /* row (token) strides       */
    int32_t q_hs, k_hs, v_hs, o_hs;                            /* head strides              */
    int32_t B, H, Tq, Tk, d;
    float scale;                                               /* applied to q.k before softmax */
    /* Device scratch of >= gg_attention_workspace_bytes() bytes (128-byte aligned) or NULL.  With it, shapes the
     * tensor-core kernel takes (d 32 / 64, Tq, Tk >= 64: attention_tc.cu -- tcgen05.mma for Q K^T and P V with TMEM
     * accumulators, TMA operands, V^T [B, H, d, Tk] staged in the scratch) run there; without it, or for other
     * shapes, the mma.sync kernel of attention.cu runs.  Same arguments, same result up to bf16 rounding of P. */
    void* workspace;
    int64_t workspace_bytes;
} gg_attn_args;
int64_t gg_attention_workspace_bytes(const gg_attn_args* a);      /* 0: the tensor-core kernel does not take this shape */
int gg_attention_fwd(const gg_attn_args* a, gg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K7/K10  small dense pieces
 *   nn.py:103-121 timestep_embedding; unet.py:511-515 time_embed; unet.py:205-211 emb_layers
 *   ldm/modules/attention.py:37-46 GEGLU; :204-206 nn.LayerNorm
 * ---------------------------------------------------------------------------------------- */
/* emb[b, :] = [cos(t_b f_i) | sin(t_b f_i)], f_i = exp(-ln(max_period) i / half); t on device */
int gg_timestep_embedding(const float* t, float* emb, int32_t B, int32_t dim, float max_period, gg_stream_t stream);
/* y[m, n] = sum_k act(x[m, k]) * w[n, k] + b[n]   fp32, M small (<= 64); act_in: 0 none, 1 SiLU;
 * act_out: 0 none, 1 SiLU */
int gg_small_linear(const float* x, const float* w, const float* b, float* y, int32_t M, int32_t N, int32_t K,
                    int32_t act_in, int32_t act_out, gg_stream_t stream);
/* LayerNorm over the last axis, bf16 in/out, fp32 statistics */
int gg_layernorm(const void* x, const float* gamma, const float* beta, void* y, int64_t rows, int32_t C, float eps,
                 gg_stream_t stream);
/* y[r, j] = x[r, j] * gelu(x[r, inner + j])  (exact erf GELU), bf16 */
int gg_geglu(const void* x, void* y, int64_t rows, int32_t inner, gg_stream_t stream);

/* Single-head full-channel attention of the VAE (latentdiffusion/ldm/modules/diffusionmodules/model.py:237-261
 * AttnBlock2d.forward): the score matrix comes from gg_conv_fwd (queries as the activation, keys as the "weights",
 * fp32 out), then  P = softmax(scale * S) over each row (:250-251), bf16 out, and the values are transposed so that
 * the second product is again a gg_conv_fwd (P as activation, V^T as weights, :254-256).
 * gg_softmax_rows: x fp32 [rows, n] -> y bf16 [rows, n].   gg_transpose_bf16: x [R, C] -> y [C, R], bf16. */
int gg_softmax_rows(const float* x, void* y, int64_t rows, int32_t n, float scale, gg_stream_t stream);
int gg_transpose_bf16(const void* x, void* y, int32_t R, int32_t C, gg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Depth-slab collectives over NVLink peer memory (BASELINE config 5).  The reference has no
 * multi-GPU inference path; the semantics are those of its single-device forward (unet.py:758-823): every conv sees
 * its neighbours' boundary planes, every GroupNorm the statistics of the whole volume, every attention site all keys.
 * Each rank owns a peer-visible arena (gg_peer_alloc; exported with gg_peer_export, mapped by the peers with
 * gg_peer_open -- cudaIpc handles exchanged by the host language's own means, e.g. torch.distributed) laid out
 * identically on every rank.  gg_peer_exchange is ONE kernel per collective site:
 *   phase 1: copy src[i] -> dst[i] (dst = addresses inside PEER arenas), system fence, then store *epoch into
 *            flag_out[i] (addresses inside peer arenas);
 *   phase 2: wait until every flag_in[i] (my arena) has reached *epoch, then csrc[i] -> cdst[i] local copies
 *            (staging slot -> halo plane) and zero fills (halo planes at the ends of the volume).
 * phase = 3 does both (the multi-process case); 1 / 2 let a single process emulate R ranks on one device by running
 * phase 1 of every rank before phase 2 of any (tests).  *epoch is bumped once per forward by gg_peer_epoch_inc.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int32_t nsend;
    const void* src[8]; void* dst[8]; int64_t bytes[8];
    int32_t nflag_out; uint32_t* flag_out[8];
    int32_t nflag_in; const uint32_t* flag_in[8];
    int32_t ncopy; const void* csrc[4]; void* cdst[4]; int64_t cbytes[4];
    int32_t nzero; void* zdst[2]; int64_t zbytes[2];
    const uint32_t* epoch;          /* device counter of this rank                                              */
    unsigned int* done_counter;     /* device word, 0 between launches (last-CTA detection)                      */
    int32_t phase;                  /* 1, 2 or 3                                                                 */
    int32_t ctas;                   /* 0 = by payload size                                                        */
} gg_peer_xchg_args;
int gg_peer_alloc(int64_t bytes, void** out);                 /* cudaMalloc'd, zero-filled                       */
int gg_peer_free(void* p);
int gg_peer_export(void* p, uint8_t* handle64);                /* 64-byte cudaIpcMemHandle_t                      */
int gg_peer_open(const uint8_t* handle64, void** out);
int gg_peer_close(void* p);
int gg_peer_epoch_inc(uint32_t* epoch, gg_stream_t stream);
int gg_peer_exchange(const gg_peer_xchg_args* a, gg_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GUIDEGEN_SM100_H */
