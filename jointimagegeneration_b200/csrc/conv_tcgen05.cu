// conv_tcgen05.cu -- implicit-GEMM convolution / linear layer on the 5th-gen tensor cores
//
//   D[pos, co] = sum over (source, tap, ci)  A_src[pos + tap, ci] * W[co, (source, tap, ci)]
//
// * activations are channels-last bf16 [N, D, H, W, C]; a tile of BM = 128 output positions is
//   a 4-D brick (bn, bd, bh, bw) so that ONE 5-D TMA box per (tap, 64-channel chunk) lands in
//   shared memory already in the K-major SWIZZLE_128B layout tcgen05.mma reads: no im2col
//   buffer, zero padding = TMA out-of-bounds fill (also pads C up to the 64-channel chunk);
// * channel concat (th.cat skip connections), the fused 1x1 skip convolution of a ResBlock
//   and stride-2 convolutions are just more K blocks read through other tensor maps
//   (stride 2: one strided map per input parity class, tap -> (parity map, offset));
// * accumulators live in TMEM (2 x 256 columns, double buffered), fp32;
// * persistent, warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer (one elected
//   lane), warps 2..5 = epilogue (tcgen05.ld -> +bias +emb +residual -> bf16/fp32 stores),
//   which overlaps the next tile's main loop.
#include "tc_common.cuh"

namespace gg {

constexpr int A_BYTES = BM * BK * 2;
constexpr int MAX_STAGES = 8;
constexpr int MAX_MAPS = 8;
constexpr int MAX_SEGS = 4;
constexpr int SMEM_BUDGET = 227 * 1024;
constexpr int NUM_THREADS = 192;
constexpr int ACC_COLS = 256;  // TMEM columns per accumulator buffer

struct ConvSeg {
    int map0;      // first tensor map of this segment
    int nchunks;   // 64-channel chunks
    int kd, kh, kw;  // tap counts
    int od, oh, ow;  // offset of tap 0 (input coordinate = output coordinate + o + tap)
    int stride2;   // 1: tap -> parity map (map0 + parity code) and offset floor((k + off) / 2)
    int s2d, s2h, s2w;  // which dims are strided (dims < 3 leave d (and h) unstrided)
    int dshift;    // extra depth offset of this source (halo-padded depth slabs)
};

struct alignas(64) ConvParams {
    CUtensorMap amap[MAX_MAPS];
    CUtensorMap wmap;
    ConvSeg seg[MAX_SEGS];
    int nseg, num_kb, stages, BN;
    int No, Do, Ho, Wo;
    int bn, bd, bh, bw;
    int tn, td, th, tw;
    int n_tiles_n, total_tiles;
    int Cout8;
    const float* bias;
    const float* emb;
    int emb_stride;
    const __nv_bfloat16* residual;
    int res_stride;
    void* y;
    long long y_sn, y_sd, y_sh, y_sw;
    int y_is_f32;
    // fused GroupNorm statistics of the (bf16-rounded) output: per (sample, M tile, epilogue warp) partial sums
    float* gn_partial;     // [N, gn_nchunks_total, Cout8, 2] or null
    int gn_chunk_base, gn_nchunks_total, stats_d_min;
    // split-K: work item = (tile, split); each split accumulates a K range and writes raw fp32 partials
    int split_k;
    float* workspace;      // [split_k, No*Do*Ho*Wo, Cout8]
    long long ws_split_stride;
    unsigned int* split_counters;   // [total_tiles], zero between launches: the LAST split of a tile to finish reduces it in-kernel
};

// ------------------------------------------------------------------------------------ kernel
__global__ void __launch_bounds__(NUM_THREADS, 1) conv_tcgen05_kernel(const __grid_constant__ ConvParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    const int stages = p.stages;
    const int BN = p.BN;
    const uint32_t stage_bytes = A_BYTES + (uint32_t)BN * 128u;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)stages * stage_bytes);
    uint64_t* empty = full + MAX_STAGES;
    uint64_t* tfull = empty + MAX_STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    volatile uint32_t* split_last = tmem_slot + 2;      // split-K fix-up: "this CTA finished the tile last" (epilogue warps)
    float* bvec = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full) + 256);     // [BN] bias (+ emb[n] when a tile = one sample)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    pdl_launch_dependents();        // the next kernel's CTAs may become resident (and run their prologue) as this one's exit
    if (warp == 0 && lane == 0) {
        for (int i = 0; i < MAX_MAPS; ++i) prefetch_tmap(&p.amap[i]);
        prefetch_tmap(&p.wmap);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int i = 0; i < stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
            for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();                     // everything above touched only this CTA's shared memory / TMEM and the kernel parameters

    if (warp == 0) {
        // ================================================================ TMA producer
        // the whole warp runs the loop (uniform control flow); one elected lane issues the copies
        int stage = 0;
        uint32_t phase = 0;
        for (int item = blockIdx.x; item < p.total_tiles * p.split_k; item += gridDim.x) {
            const int tile = item / p.split_k, split = item - tile * p.split_k;
            const int kb_lo = (int)((long long)p.num_kb * split / p.split_k), kb_hi = (int)((long long)p.num_kb * (split + 1) / p.split_k);
            const int nt = tile % p.n_tiles_n;
            int mt = tile / p.n_tiles_n;
            const int iw = mt % p.tw; mt /= p.tw;
            const int ih = mt % p.th; mt /= p.th;
            const int id = mt % p.td; mt /= p.td;
            const int n0 = mt * p.bn, d0 = id * p.bd, h0 = ih * p.bh, w0 = iw * p.bw;
            int kb = 0;
            for (int s = 0; s < p.nseg; ++s) {
                const ConvSeg sg = p.seg[s];
                for (int a = 0; a < sg.kd; ++a)
                    for (int b = 0; b < sg.kh; ++b)
                        for (int c = 0; c < sg.kw; ++c) {
                            int mi = sg.map0, od, oh, ow;
                            if (sg.stride2) {
                                // input = 2*o + (k + off), off = -1 (pad 1) or 0 (pad (0, 1)): parity grid (k + off) & 1 at
                                // o + floor((k + off) / 2); e.g. off -1: k=0 -> odd grid at o-1, k=1 -> even at o, k=2 -> odd at o
                                const int pd = sg.s2d ? ((a + sg.od) & 1) : 0, ph = sg.s2h ? ((b + sg.oh) & 1) : 0,
                                          pw = sg.s2w ? ((c + sg.ow) & 1) : 0;
                                mi += pd * 4 + ph * 2 + pw;
                                od = sg.s2d ? ((a + sg.od) >> 1) : sg.od + a;
                                oh = sg.s2h ? ((b + sg.oh) >> 1) : sg.oh + b;
                                ow = sg.s2w ? ((c + sg.ow) >> 1) : sg.ow + c;
                            } else {
                                od = sg.od + a; oh = sg.oh + b; ow = sg.ow + c;
                            }
                            for (int j = 0; j < sg.nchunks; ++j) {
                                if (kb < kb_lo || kb >= kb_hi) { ++kb; continue; }      // another split's K block
                                mbar_wait(&empty[stage], phase ^ 1u);
                                if (elect_one()) {
                                    mbar_expect_tx(&full[stage], stage_bytes);
                                    uint8_t* sa = smem + (size_t)stage * stage_bytes;
                                    tma_load_5d(sa, &p.amap[mi], &full[stage], j * BK, w0 + ow, h0 + oh, d0 + od + sg.dshift, n0);
                                    tma_load_2d(sa + A_BYTES, &p.wmap, &full[stage], kb * BK, nt * BN);
                                }
                                __syncwarp();
                                ++kb;
                                if (++stage == stages) { stage = 0; phase ^= 1u; }
                            }
                        }
            }
        }
    } else if (warp == 1) {
        // ================================================================ MMA issuer
        // whole warp loops (uniform control flow, no per-instruction election); one elected lane
        // issues the four K=16 MMAs of a stage and the commit that frees it
        // instruction descriptor: D=f32, A=B=bf16, both K-major, N = BN, M = 128
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        const uint32_t smem_base = smem_u32(smem);
        int stage = 0;
        uint32_t phase = 0, acc = 0, acc_phase = 0;
        for (int item = blockIdx.x; item < p.total_tiles * p.split_k; item += gridDim.x) {
            const int split = item % p.split_k;
            const int kb_lo = (int)((long long)p.num_kb * split / p.split_k), kb_hi = (int)((long long)p.num_kb * (split + 1) / p.split_k);
            mbar_wait(&tempty[acc], acc_phase ^ 1u);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
            for (int kb = kb_lo; kb < kb_hi; ++kb) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t sa = smem_base + (uint32_t)stage * stage_bytes;
                    const uint64_t adesc = make_sw128_desc(sa), bdesc = make_sw128_desc(sa + A_BYTES);
                    // +32 bytes per K=16 step inside the 128-byte swizzle atom (encoded >> 4)
                    umma_bf16(d_tmem, adesc, bdesc, idesc, kb != kb_lo ? 1u : 0u);
                    umma_bf16(d_tmem, adesc + 2, bdesc + 2, idesc, 1u);
                    umma_bf16(d_tmem, adesc + 4, bdesc + 4, idesc, 1u);
                    umma_bf16(d_tmem, adesc + 6, bdesc + 6, idesc, 1u);
                    umma_commit(&empty[stage]);
                    if (kb == kb_hi - 1) umma_commit(&tfull[acc]);
                }
                __syncwarp();
                if (++stage == stages) { stage = 0; phase ^= 1u; }
            }
            acc ^= 1u;
            if (acc == 0) acc_phase ^= 1u;
        }
    } else {
        // ================================================================ epilogue (warps 2..5)
        const int q = warp & 3;               // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;
        uint32_t acc = 0, acc_phase = 0;
        int cur_n = -1, cur_nt = -1;
        const int bvol = p.bd * p.bh * p.bw, bhw = p.bh * p.bw;
        for (int item = blockIdx.x; item < p.total_tiles * p.split_k; item += gridDim.x) {
            const int tile = item / p.split_k, split = item - tile * p.split_k;
            const int nt = tile % p.n_tiles_n;
            int mt = tile / p.n_tiles_n;
            const int iw = mt % p.tw; mt /= p.tw;
            const int ih = mt % p.th; mt /= p.th;
            const int id = mt % p.td; mt /= p.td;
            const int rn = row / bvol, r1 = row - rn * bvol;
            const int rd = r1 / bhw, r2 = r1 - rd * bhw;
            const int rh = r2 / p.bw, rw = r2 - rh * p.bw;
            const int n = mt * p.bn + rn, d = id * p.bd + rd, h = ih * p.bh + rh, w = iw * p.bw + rw;
            const bool valid = n < p.No && d < p.Do && h < p.Ho && w < p.Wo;
            const long long yoff = (long long)n * p.y_sn + (long long)d * p.y_sd + (long long)h * p.y_sh + (long long)w * p.y_sw;
            const long long lin = (((long long)n * p.Do + d) * p.Ho + h) * p.Wo + w;
            const float* embp = p.emb ? p.emb + (long long)n * p.emb_stride : nullptr;

            if (p.split_k > 1) {
                // raw fp32 partial sums of this K range -> workspace[split][position][channel]; bias / emb / residual
                // are applied by splitk_reduce_kernel
                if (cur_nt != -3) {
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    for (int c = (int)threadIdx.x - 64; c < BN; c += 128) bvec[c] = 0.f;
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    cur_nt = -3;
                }
                const int ncols = min(BN, p.Cout8 - nt * BN);
                float* w_row = p.workspace + (long long)split * p.ws_split_stride + lin * p.Cout8 + nt * BN;
                mbar_wait(&tfull[acc], acc_phase);
                tc_fence_after();
                epilogue_row(tmem_base + acc * ACC_COLS + ((uint32_t)(q * 32) << 16), BN, ncols, bvec, nullptr, w_row, 1, valid);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[acc]);
                acc ^= 1u;
                if (acc == 0) acc_phase ^= 1u;
                if (p.split_counters != nullptr) {
                    // In-kernel fix-up (no separate reduce launch): whichever split of this tile finishes LAST sums the partial
                    // tiles in split order 0, 1, ... -- the same order whoever is last, so the result is reproducible -- and runs
                    // the real epilogue (bias, embedding, residual, rounding, store).  atomicInc wraps the counter back to 0.
                    __threadfence();
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    if (threadIdx.x == 64) *split_last = atomicInc(p.split_counters + tile, (unsigned int)p.split_k - 1u) == (unsigned int)p.split_k - 1u;
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    if (*split_last && valid) {
                        __threadfence();
                        const float* wbase = p.workspace + lin * p.Cout8 + nt * BN;
                        for (int c = 0; c < ncols; c += 8) {
                            float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                            for (int s2 = 0; s2 < p.split_k; ++s2) {
                                const float* wp = wbase + (long long)s2 * p.ws_split_stride + c;
                                const float4 a4 = __ldcg(reinterpret_cast<const float4*>(wp)), b4 = __ldcg(reinterpret_cast<const float4*>(wp + 4));
                                v[0] += a4.x; v[1] += a4.y; v[2] += a4.z; v[3] += a4.w; v[4] += b4.x; v[5] += b4.y; v[6] += b4.z; v[7] += b4.w;
                            }
                            const int cg = nt * BN + c;
                            if (p.bias) {
#pragma unroll
                                for (int e = 0; e < 8; ++e) v[e] += __ldg(p.bias + cg + e);
                            }
                            if (embp) {
#pragma unroll
                                for (int e = 0; e < 8; ++e) v[e] += __ldg(embp + cg + e);
                            }
                            if (p.residual) {
                                const uint4 rr = ldg_nc_u4(p.residual + lin * p.res_stride + cg);
                                v[0] += bf16_lo(rr.x); v[1] += bf16_hi(rr.x); v[2] += bf16_lo(rr.y); v[3] += bf16_hi(rr.y);
                                v[4] += bf16_lo(rr.z); v[5] += bf16_hi(rr.z); v[6] += bf16_lo(rr.w); v[7] += bf16_hi(rr.w);
                            }
                            if (p.y_is_f32) {
                                float* yp = reinterpret_cast<float*>(p.y) + yoff + cg;
                                *reinterpret_cast<float4*>(yp) = make_float4(v[0], v[1], v[2], v[3]);
                                *reinterpret_cast<float4*>(yp + 4) = make_float4(v[4], v[5], v[6], v[7]);
                            } else {
                                *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.y) + yoff + cg) =
                                    make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
                            }
                        }
                    }
                }
                continue;
            }
            if (p.gn_partial == nullptr) {
                // fast path: additive vector staged in smem, residual prefetched, paired TMEM loads
                const int vn = (p.bn == 1) ? mt : -2;              // sample whose emb is folded into bvec (-2: none)
                if (vn != cur_n || nt != cur_nt) {                 // uniform over the four epilogue warps
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    for (int c = (int)threadIdx.x - 64; c < BN; c += 128) {
                        const int chn = nt * BN + c;
                        float bv = 0.f;
                        if (chn < p.Cout8) {
                            if (p.bias) bv += __ldg(p.bias + chn);
                            if (p.emb && vn >= 0) bv += __ldg(p.emb + (long long)vn * p.emb_stride + chn);
                        }
                        bvec[c] = bv;
                    }
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    cur_n = vn; cur_nt = nt;
                }
                const int ncols = min(BN, p.Cout8 - nt * BN);
                const __nv_bfloat16* res_row = p.residual ? p.residual + lin * p.res_stride + nt * BN : nullptr;
                void* y_row = p.y_is_f32 ? static_cast<void*>(reinterpret_cast<float*>(p.y) + yoff + nt * BN)
                                         : static_cast<void*>(reinterpret_cast<__nv_bfloat16*>(p.y) + yoff + nt * BN);
                const float* emb_row = (p.emb && vn < 0 && valid) ? embp + nt * BN : nullptr;
                mbar_wait(&tfull[acc], acc_phase);
                tc_fence_after();
                epilogue_row(tmem_base + acc * ACC_COLS + ((uint32_t)(q * 32) << 16), BN, ncols, bvec, res_row, y_row, p.y_is_f32,
                             valid, emb_row);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[acc]);
                acc ^= 1u;
                if (acc == 0) acc_phase ^= 1u;
                continue;
            }
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + acc * ACC_COLS + ((uint32_t)(q * 32) << 16);
            const bool stat_row = valid && d >= p.stats_d_min;
            float* gnp = nullptr;
            if (p.gn_partial) {     // bn == 1 (checked on the host): the whole tile belongs to sample `mt`
                const long long chunk = p.gn_chunk_base + (long long)((id * p.th + ih) * p.tw + iw) * 4 + q;
                gnp = p.gn_partial + (((long long)mt * p.gn_nchunks_total + chunk) * p.Cout8) * 2;
            }
            for (int c0 = 0; c0 < BN; c0 += 16) {
                uint32_t r[16];
                tmem_ld16(t_addr + c0, r);
                tmem_ld_wait();
                const int ch = nt * BN + c0;
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
                if (valid && ch < p.Cout8) {
#pragma unroll
                    for (int g = 0; g < 2; ++g) {
                        const int cg = ch + 8 * g;
                        if (cg >= p.Cout8) break;
                        float* vv = v + 8 * g;
                        if (p.bias) {
                            const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + cg));
                            const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + cg + 4));
                            vv[0] += b0.x; vv[1] += b0.y; vv[2] += b0.z; vv[3] += b0.w;
                            vv[4] += b1.x; vv[5] += b1.y; vv[6] += b1.z; vv[7] += b1.w;
                        }
                        if (embp) {
                            const float4 b0 = __ldg(reinterpret_cast<const float4*>(embp + cg));
                            const float4 b1 = __ldg(reinterpret_cast<const float4*>(embp + cg + 4));
                            vv[0] += b0.x; vv[1] += b0.y; vv[2] += b0.z; vv[3] += b0.w;
                            vv[4] += b1.x; vv[5] += b1.y; vv[6] += b1.z; vv[7] += b1.w;
                        }
                        if (p.residual) {
                            const uint4 rr = ldg_nc_u4(p.residual + lin * p.res_stride + cg);
                            vv[0] += bf16_lo(rr.x); vv[1] += bf16_hi(rr.x); vv[2] += bf16_lo(rr.y); vv[3] += bf16_hi(rr.y);
                            vv[4] += bf16_lo(rr.z); vv[5] += bf16_hi(rr.z); vv[6] += bf16_lo(rr.w); vv[7] += bf16_hi(rr.w);
                        }
                        if (p.y_is_f32) {
                            float* yp = reinterpret_cast<float*>(p.y) + yoff + cg;
                            *reinterpret_cast<float4*>(yp) = make_float4(vv[0], vv[1], vv[2], vv[3]);
                            *reinterpret_cast<float4*>(yp + 4) = make_float4(vv[4], vv[5], vv[6], vv[7]);
                        } else {
                            __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(p.y) + yoff + cg;
                            *reinterpret_cast<uint4*>(yp) = make_uint4(pack_bf16(vv[0], vv[1]), pack_bf16(vv[2], vv[3]),
                                                                       pack_bf16(vv[4], vv[5]), pack_bf16(vv[6], vv[7]));
                        }
                    }
                }
                if (gnp != nullptr && ch < p.Cout8) {      // warp-uniform
                    // per-column (sum, sum of squares) over this warp's 32 rows of the values AS STORED (bf16):
                    // butterfly reduce-scatter, 16 shuffles per quantity; lane l ends with column
                    // 8*b4 + 4*b3 + 2*b2 + b1 (b_k = bit k of l), duplicated in lanes l and l^1
                    float s1[16], s2[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float rv = (stat_row && ch + j < p.Cout8) ? (p.y_is_f32 ? v[j] : __bfloat162float(__float2bfloat16_rn(v[j]))) : 0.f;
                        s1[j] = rv; s2[j] = rv * rv;
                    }
#pragma unroll
                    for (int half = 8; half >= 1; half >>= 1) {
                        const int m = half * 2;
                        const bool up = (lane & m) != 0;
#pragma unroll
                        for (int i = 0; i < half; ++i) {
                            const float k1 = up ? s1[i + half] : s1[i], x1 = up ? s1[i] : s1[i + half];
                            const float k2 = up ? s2[i + half] : s2[i], x2 = up ? s2[i] : s2[i + half];
                            s1[i] = k1 + __shfl_xor_sync(0xffffffffu, x1, m);
                            s2[i] = k2 + __shfl_xor_sync(0xffffffffu, x2, m);
                        }
                    }
                    s1[0] += __shfl_xor_sync(0xffffffffu, s1[0], 1);
                    s2[0] += __shfl_xor_sync(0xffffffffu, s2[0], 1);
                    const int col = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
                    if ((lane & 1) == 0 && ch + col < p.Cout8)
                        *reinterpret_cast<float2*>(gnp + 2 * (ch + col)) = make_float2(s1[0], s2[0]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            acc ^= 1u;
            if (acc == 0) acc_phase ^= 1u;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// sum of the split-K partials + bias + emb + residual -> output (8 channels per thread)
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const ConvParams p, long long npos) {
    pdl_launch_dependents();
    pdl_wait();
    const int P8 = p.Cout8 >> 3;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npos * P8) return;
    const long long lin = i / P8;
    const int cg = (int)(i - lin * P8) * 8;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int s = 0; s < p.split_k; ++s) {
        const float* wp = p.workspace + (long long)s * p.ws_split_stride + lin * p.Cout8 + cg;
        const float4 a = __ldg(reinterpret_cast<const float4*>(wp)), b = __ldg(reinterpret_cast<const float4*>(wp + 4));
        v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
    }
    long long t = lin;
    const int w = (int)(t % p.Wo); t /= p.Wo;
    const int h = (int)(t % p.Ho); t /= p.Ho;
    const int d = (int)(t % p.Do); t /= p.Do;
    const int n = (int)t;
    if (p.bias) {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] += __ldg(p.bias + cg + e);
    }
    if (p.emb) {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] += __ldg(p.emb + (long long)n * p.emb_stride + cg + e);
    }
    if (p.residual) {
        const uint4 rr = ldg_nc_u4(p.residual + lin * p.res_stride + cg);
        v[0] += bf16_lo(rr.x); v[1] += bf16_hi(rr.x); v[2] += bf16_lo(rr.y); v[3] += bf16_hi(rr.y);
        v[4] += bf16_lo(rr.z); v[5] += bf16_hi(rr.z); v[6] += bf16_lo(rr.w); v[7] += bf16_hi(rr.w);
    }
    const long long yoff = (long long)n * p.y_sn + (long long)d * p.y_sd + (long long)h * p.y_sh + (long long)w * p.y_sw + cg;
    if (p.y_is_f32) {
        float* yp = reinterpret_cast<float*>(p.y) + yoff;
        *reinterpret_cast<float4*>(yp) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(yp + 4) = make_float4(v[4], v[5], v[6], v[7]);
    } else {
        *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.y) + yoff) =
            make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    }
}

// -------------------------------------------------------------------------------- host side
EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<EncodeTiledFn>(f);
    }();
    return fn;
}

// 5-D activation map over a (possibly strided) sub-grid of a CL tensor
bool encode_act_map(CUtensorMap* m, const void* base, int C, const int64_t dim[4] /*W,H,D,N extents*/,
                           const int64_t stride_el[4] /*element strides of W,H,D,N*/, const int box[4] /*bw,bh,bd,bn*/) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t gdim[5] = {(cuuint64_t)C, (cuuint64_t)dim[0], (cuuint64_t)dim[1], (cuuint64_t)dim[2], (cuuint64_t)dim[3]};
    cuuint64_t gstr[4] = {(cuuint64_t)stride_el[0] * 2, (cuuint64_t)stride_el[1] * 2, (cuuint64_t)stride_el[2] * 2,
                          (cuuint64_t)stride_el[3] * 2};
    cuuint32_t bx[5] = {(cuuint32_t)BK, (cuuint32_t)box[0], (cuuint32_t)box[1], (cuuint32_t)box[2], (cuuint32_t)box[3]};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), gdim, gstr, bx, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

bool encode_w_map(CUtensorMap* m, const void* base, int64_t Ktot, int rows, int BN) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t gdim[2] = {(cuuint64_t)Ktot, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)Ktot * 2};
    cuuint32_t bx[2] = {(cuuint32_t)BK, (cuuint32_t)BN};
    cuuint32_t es[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, bx, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

static int pow2_ceil(int v) { int r = 1; while (r < v) r *= 2; return r; }

// brick of 128 output positions: as much W as possible, then H, D, N (powers of two)
static void pick_brick(int No, int Do, int Ho, int Wo, int out[4] /*bn,bd,bh,bw*/) {
    int rem = BM;
    int bw = std::min(rem, pow2_ceil(Wo)); rem /= bw;
    int bh = std::min(rem, pow2_ceil(Ho)); rem /= bh;
    int bd = std::min(rem, pow2_ceil(Do)); rem /= bd;
    int bn = rem;  // may exceed No: rows beyond are masked
    (void)No;
    out[0] = bn; out[1] = bd; out[2] = bh; out[3] = bw;
}

}  // namespace gg

using namespace gg;

extern "C" int32_t gg_conv_pick_block_n(int32_t Cout) {
    if (Cout <= 0) return 0;
    const int c16 = (Cout + 15) / 16 * 16;
    if (c16 <= 256) return c16;
    // largest multiple of 16 (<= 256) whose tiling wastes <= 6 % of the columns; else the least waste
    int best = 256, best_waste = 1 << 30;
    for (int bn = 256; bn >= 64; bn -= 16) {
        const int tiles = (Cout + bn - 1) / bn;
        const int waste = tiles * bn - Cout;
        if (waste * 100 <= 6 * Cout) return bn;
        if (waste < best_waste) { best_waste = waste; best = bn; }
    }
    return best;
}

extern "C" int32_t gg_conv_num_tiles(const gg_conv_args* a) {
    if (!a || a->Do <= 0 || a->Ho <= 0 || a->Wo <= 0 || a->Cout <= 0) return 0;
    int brick[4];
    if (a->brick[0] > 0) { for (int i = 0; i < 4; ++i) brick[i] = a->brick[i]; }
    else pick_brick(a->N, a->Do, a->Ho, a->Wo, brick);
    const int BN = a->block_n > 0 ? a->block_n : gg_conv_pick_block_n(a->Cout);
    const int64_t t = (int64_t)((a->N + brick[0] - 1) / brick[0]) * ((a->Do + brick[1] - 1) / brick[1]) * ((a->Ho + brick[2] - 1) / brick[2]) *
                      ((a->Wo + brick[3] - 1) / brick[3]) * ((a->Cout + BN - 1) / BN);
    return (int32_t)std::min<int64_t>(t, 1 << 30);
}

extern "C" int32_t gg_conv_stats_chunks(const gg_conv_args* a) {
    if (!a || a->Do <= 0 || a->Ho <= 0 || a->Wo <= 0) return 0;
    if (a->algo >= 1 && a->algo <= 3) {
        // halo kernel: per-tile shuffle-reduced column sums, one row per (CTA, epilogue warp); bf16 outputs
        if (a->y_is_f32) return 0;
        return conv_halo_grid(a, nullptr) * 4;
    }
    if (a->algo == 4) {
        if ((a->Cout + 7) / 8 * 8 != 64 || a->y_is_f32) return 0;
        return conv_roll_grid(a) * 8;      // one row per (CTA, epilogue warp): two groups of four warps
    }
    int brick[4];
    if (a->brick[0] > 0) { for (int i = 0; i < 4; ++i) brick[i] = a->brick[i]; }
    else pick_brick(a->N, a->Do, a->Ho, a->Wo, brick);
    if (brick[0] != 1) return 0;
    const int td = (a->Do + brick[1] - 1) / brick[1], th = (a->Ho + brick[2] - 1) / brick[2], tw = (a->Wo + brick[3] - 1) / brick[3];
    return td * th * tw * 4;
}

extern "C" int64_t gg_conv_packed_k(const gg_conv_args* a) {
    if (!a || a->nsrc < 1 || a->nsrc > 4) return -1;
    int64_t kb = 0;
    for (int s = 0; s < a->nsrc; ++s) {
        const int nch = (a->src[s].C + BK - 1) / BK;
        const int taps = a->src[s].centre_only ? 1 : a->kd * a->kh * a->kw;
        kb += (int64_t)taps * nch;
    }
    return kb * BK;
}

extern "C" int gg_conv_fwd(const gg_conv_args* a, gg_stream_t stream) {
    GG_REQUIRE(a != nullptr && a->nsrc >= 1 && a->nsrc <= MAX_SEGS, GG_ERR_BAD_ARG);
    GG_REQUIRE(a->N > 0 && a->D > 0 && a->H > 0 && a->W > 0 && a->Cout > 0 && a->y && a->w_packed, GG_ERR_BAD_ARG);
    GG_REQUIRE(a->kd >= 1 && a->kh >= 1 && a->kw >= 1 && a->kd <= 3 && a->kh <= 3 && a->kw <= 3, GG_ERR_UNSUPPORTED);
    GG_REQUIRE(a->stride == 1 || a->stride == 2, GG_ERR_UNSUPPORTED);
    GG_REQUIRE(a->Do > 0 && a->Ho > 0 && a->Wo > 0, GG_ERR_BAD_ARG);
    GG_REQUIRE(aligned(a->y, 16) && aligned(a->w_packed, 16), GG_ERR_ALIGNMENT);
    GG_REQUIRE(a->y_sw % 8 == 0 && a->y_sh % 8 == 0 && a->y_sd % 8 == 0 && a->y_sn % 8 == 0, GG_ERR_ALIGNMENT);
    if (a->residual) GG_REQUIRE(aligned(a->residual, 16) && a->res_stride % 8 == 0, GG_ERR_ALIGNMENT);
    if (a->bias) GG_REQUIRE(aligned(a->bias, 16), GG_ERR_ALIGNMENT);
    if (a->emb) GG_REQUIRE(aligned(a->emb, 16) && a->emb_stride % 4 == 0, GG_ERR_ALIGNMENT);
    if (!encode_fn()) return GG_ERR_DRIVER;
    if (a->cat != nullptr) GG_REQUIRE(a->algo == 4, GG_ERR_UNSUPPORTED);      // sampler epilogue: depth-rolling kernel only
    if (a->algo >= 1 && a->algo <= 3) return conv_halo_fwd(a, as_stream(stream));
    if (a->algo == 4) return conv_roll_fwd(a, as_stream(stream));

    ConvParams p;
    memset(&p, 0, sizeof(p));
    const int BN = a->block_n > 0 ? a->block_n : gg_conv_pick_block_n(a->Cout);
    GG_REQUIRE(BN % 16 == 0 && BN >= 16 && BN <= 256, GG_ERR_UNSUPPORTED);
    p.BN = BN;
    int brick[4];
    if (a->brick[0] > 0) {
        for (int i = 0; i < 4; ++i) brick[i] = a->brick[i];
        GG_REQUIRE(brick[0] * brick[1] * brick[2] * brick[3] == BM, GG_ERR_BAD_ARG);
    } else {
        pick_brick(a->N, a->Do, a->Ho, a->Wo, brick);
    }
    for (int i = 0; i < 4; ++i) GG_REQUIRE(brick[i] >= 1 && brick[i] <= 256, GG_ERR_BAD_ARG);
    p.bn = brick[0]; p.bd = brick[1]; p.bh = brick[2]; p.bw = brick[3];
    p.No = a->N; p.Do = a->Do; p.Ho = a->Ho; p.Wo = a->Wo;
    p.tn = (a->N + p.bn - 1) / p.bn; p.td = (a->Do + p.bd - 1) / p.bd;
    p.th = (a->Ho + p.bh - 1) / p.bh; p.tw = (a->Wo + p.bw - 1) / p.bw;
    p.n_tiles_n = (a->Cout + BN - 1) / BN;
    const int64_t total = (int64_t)p.tn * p.td * p.th * p.tw * p.n_tiles_n;
    GG_REQUIRE(total < (1ll << 31), GG_ERR_UNSUPPORTED);
    p.total_tiles = (int)total;
    p.Cout8 = (a->Cout + 7) / 8 * 8;

    // segments and tensor maps
    const int box[4] = {p.bw, p.bh, p.bd, p.bn};
    int nmaps = 0, num_kb = 0;
    const int64_t W = a->W, H = a->H, D = a->D, N = a->N;
    for (int s = 0; s < a->nsrc; ++s) {
        const gg_conv_src& src = a->src[s];
        GG_REQUIRE(src.x != nullptr && src.C > 0 && src.C % 8 == 0, GG_ERR_BAD_ARG);
        GG_REQUIRE(aligned(src.x, 16), GG_ERR_ALIGNMENT);
        ConvSeg& sg = p.seg[s];
        sg.map0 = nmaps;
        sg.nchunks = (src.C + BK - 1) / BK;
        sg.dshift = src.d_shift;
        const int64_t C = src.C;
        if (src.centre_only || a->stride == 1) {
            GG_REQUIRE(nmaps + 1 <= MAX_MAPS, GG_ERR_UNSUPPORTED);
            if (src.centre_only) {
                sg.kd = sg.kh = sg.kw = 1;
                sg.od = sg.oh = sg.ow = 0;
                // centre tap of a strided conv reads the even sub-grid; of a stride-1 conv the tensor itself
                const int64_t st = a->stride;
                const int64_t dim[4] = {(W + st - 1) / st, a->dims >= 2 ? (H + st - 1) / st : H, a->dims >= 3 ? (D + st - 1) / st : D, N};
                const int64_t str[4] = {C * st, W * C * (a->dims >= 2 ? st : 1), H * W * C * (a->dims >= 3 ? st : 1), D * H * W * C};
                if (!encode_act_map(&p.amap[nmaps], src.x, src.C, dim, str, box)) return GG_ERR_DRIVER;
            } else {
                sg.kd = a->kd; sg.kh = a->kh; sg.kw = a->kw;
                sg.od = a->od; sg.oh = a->oh; sg.ow = a->ow;
                const int64_t dim[4] = {W, H, D, N};
                const int64_t str[4] = {C, W * C, H * W * C, D * H * W * C};
                if (!encode_act_map(&p.amap[nmaps], src.x, src.C, dim, str, box)) return GG_ERR_DRIVER;
            }
            nmaps += 1;
        } else {
            // stride 2, 3-tap in every strided dim: 8 parity sub-grids
            GG_REQUIRE(nmaps + 8 <= MAX_MAPS, GG_ERR_UNSUPPORTED);
            sg.stride2 = 1;
            sg.s2w = 1; sg.s2h = a->dims >= 2; sg.s2d = a->dims >= 3;
            // tap offset -1 = symmetric pad 1 (unet.py:135-139); 0 = the VAE's pad (0, 1) before a pad-0 conv (model.py:75-78)
            GG_REQUIRE(a->kw == 3 && (a->ow == -1 || a->ow == 0), GG_ERR_UNSUPPORTED);
            if (sg.s2h) GG_REQUIRE(a->kh == 3 && (a->oh == -1 || a->oh == 0), GG_ERR_UNSUPPORTED);
            if (sg.s2d) GG_REQUIRE(a->kd == 3 && (a->od == -1 || a->od == 0), GG_ERR_UNSUPPORTED);
            sg.kd = a->kd; sg.kh = a->kh; sg.kw = a->kw;
            sg.od = a->od; sg.oh = a->oh; sg.ow = a->ow;
            for (int code = 0; code < 8; ++code) {
                const int pd = (code >> 2) & 1, ph = (code >> 1) & 1, pw = code & 1;
                const bool used = (sg.s2d || pd == 0) && (sg.s2h || ph == 0);
                const int64_t sd = sg.s2d ? 2 : 1, sh = sg.s2h ? 2 : 1;
                int64_t dim[4] = {(W - pw + 1) / 2, sg.s2h ? (H - ph + 1) / 2 : H, sg.s2d ? (D - pd + 1) / 2 : D, N};
                const int64_t str[4] = {2 * C, sh * W * C, sd * H * W * C, D * H * W * C};
                const char* base = reinterpret_cast<const char*>(src.x) + 2 * (((int64_t)pd * H + ph) * W + pw) * C;
                if (!used || dim[0] <= 0 || dim[1] <= 0 || dim[2] <= 0) {
                    // never addressed by the producer (or empty grid: extent-1 dims only read as zero)
                    base = reinterpret_cast<const char*>(src.x);
                    for (int i = 0; i < 3; ++i) if (dim[i] <= 0) dim[i] = 1;
                    if (used) return GG_ERR_UNSUPPORTED;  // odd-parity grid empty => extent 1 in a strided dim
                }
                if (!encode_act_map(&p.amap[nmaps + code], base, src.C, dim, str, box)) return GG_ERR_DRIVER;
            }
            nmaps += 8;
        }
        num_kb += sg.kd * sg.kh * sg.kw * sg.nchunks;
    }
    // unused map slots: copy map 0 so that prefetch.tensormap never sees garbage
    for (int i = nmaps; i < MAX_MAPS; ++i) p.amap[i] = p.amap[0];
    p.nseg = a->nsrc;
    p.num_kb = num_kb;
    const int64_t Ktot = (int64_t)num_kb * BK;
    if (!encode_w_map(&p.wmap, a->w_packed, Ktot, a->Cout, BN)) return GG_ERR_DRIVER;

    const int stage_bytes = A_BYTES + BN * 128;
    const int bar_bytes = 256 + 1024;      // barriers + the epilogue's [BN] additive vector
    int stages = (SMEM_BUDGET - 1024 - bar_bytes) / stage_bytes;
    stages = std::min(stages, MAX_STAGES);
    GG_REQUIRE(stages >= 2, GG_ERR_UNSUPPORTED);
    p.stages = stages;
    const size_t smem = (size_t)stages * stage_bytes + bar_bytes + 1024;

    p.bias = a->bias; p.emb = a->emb; p.emb_stride = a->emb_stride;
    p.residual = reinterpret_cast<const __nv_bfloat16*>(a->residual); p.res_stride = a->res_stride;
    p.y = a->y; p.y_sn = a->y_sn; p.y_sd = a->y_sd; p.y_sh = a->y_sh; p.y_sw = a->y_sw; p.y_is_f32 = a->y_is_f32;
    if (a->gn_partial) {
        GG_REQUIRE(p.bn == 1, GG_ERR_UNSUPPORTED);      // a tile must not span samples (gg_conv_stats_chunks() == 0)
        GG_REQUIRE(aligned(a->gn_partial, 8), GG_ERR_ALIGNMENT);
        GG_REQUIRE(a->gn_chunk_base >= 0 && a->gn_chunk_base + p.td * p.th * p.tw * 4 <= a->gn_nchunks_total, GG_ERR_BAD_ARG);
        p.gn_partial = a->gn_partial; p.gn_chunk_base = a->gn_chunk_base; p.gn_nchunks_total = a->gn_nchunks_total;
        p.stats_d_min = a->stats_d_min;
    }

    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BUDGET);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    p.split_k = 1;
    if (a->split_k > 1) {
        GG_REQUIRE(a->workspace != nullptr && aligned(a->workspace, 16) && a->gn_partial == nullptr, GG_ERR_BAD_ARG);
        GG_REQUIRE(a->split_k <= num_kb, GG_ERR_BAD_ARG);
        p.split_k = a->split_k;
        p.workspace = a->workspace;
        p.ws_split_stride = (long long)a->N * a->Do * a->Ho * a->Wo * p.Cout8;
        p.split_counters = a->split_counters;
    }
    const int grid = (int)std::min<int64_t>((int64_t)p.total_tiles * p.split_k, num_sms());
    cudaError_t le = launch_k(conv_tcgen05_kernel, dim3(grid), dim3(NUM_THREADS), smem, as_stream(stream), p);
    if (le != cudaSuccess) return (int)le;
    int st = launch_result();
    if (st != GG_OK || p.split_k == 1 || p.split_counters != nullptr) return st;
    const long long npos = (long long)a->N * a->Do * a->Ho * a->Wo;
    const long long nthr = npos * (p.Cout8 / 8);
    le = launch_k(splitk_reduce_kernel, dim3((unsigned)((nthr + 255) / 256)), dim3(256), 0, as_stream(stream), p, npos);
    if (le != cudaSuccess) return (int)le;
    return launch_result();
}
