// im2col-free implicit-GEMM convolution on Blackwell tcgen05 tensor cores (sm_100a).
//
//   fprop / dgrad : D[128 pixels, BN channels] += A[128 pixels, 64 ch] * B[BN, 64 ch]^T per (tap, channel chunk)
//                   A = activation patch fetched by a 4-D TMA box {64ch, TW, TH, 1} (zero-filled halo = padding),
//                   B = packed weights [K][R*S*C] by a 2-D TMA box; both K-major, 128B-swizzled smem.
//   wgrad         : D[128 out-ch, 64*nb in-ch] += dY^T[128, 64 px] * X_shift[64 px, 64*nb]; both operands are
//                   MN-major (channels contiguous) straight out of NHWC memory, K = pixels.
//
// Warp roles (256 threads): warp0 = TMA producer, warp1 = MMA issuer (one thread), warp2 = TMEM allocator,
// warps4-7 = epilogue (tcgen05.ld → registers → global).  fp32 accumulators live in TMEM.
//
// Replaces the cuDNN lowering of nn.Conv2d at /root/reference/furnace/seg_opr/seg_oprs.py:29-31,
// /root/reference/furnace/base_model/resnet.py:11-14,126 and the heads of
// /root/reference/model/bisenet/cityscapes.bisenet.R18/network.py:145-161 (SURVEY.md §8 a1,a2).
#include "conv_common.cuh"
#include "conv_v2.cuh"
#include "sm100_ptx.cuh"

using namespace sm100;
using namespace convhost;

namespace {

// ------------------------------------------------------------------------------------------------
// parameter blocks (passed as __grid_constant__)
// ------------------------------------------------------------------------------------------------

struct alignas(64) KmParams {
    CUtensorMap mapA[4];
    CUtensorMap mapB;
    Tap taps[kMaxTaps];
    int ntaps, kchunks;
    int tiles_w, tiles_h;
    int tw_shift, TW, TH;
    int P, Q;  // extents of the launch's output grid (masking)
    long long out_n_stride, out_p_stride, out_q_stride, out_base;  // in elements
    int k_real;   // real number of output channels (bias / stats bound)
    int k_store;  // channels [0,k_store) are written (multiple of 8)
    int out_f32, accumulate;
    const float* bias;
    float* sum;
    float* sumsq;
    void* out;
};

struct alignas(64) WgParams {
    CUtensorMap mapA;     // dY
    CUtensorMap mapB[4];  // X views
    Tap taps[kMaxTaps];   // bk unused
    int ntaps, cchunks, nb, nboxes;
    int tiles_w, tiles_h, total_tiles, tiles_per_split;
    int tw_shift, TW, TH;
    int K;                    // real out channels
    int C;                    // in channels per tap in dw (row segment length)
    long long dw_row_stride;  // elements between consecutive out channels in dw
    float* dw;
};

constexpr int kThreadsConv = 256;
constexpr int kABytes = 128 * 128;  // 128 rows x 128 B

// ------------------------------------------------------------------------------------------------
// K-major implicit GEMM (fprop, dgrad, stem fprop)
// ------------------------------------------------------------------------------------------------
template <int BN, int STAGES>
struct KmSmem {
    static constexpr int kBBytes = BN * 128;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kBarOff = STAGES * kStageBytes;
    static constexpr int kStatsOff = kBarOff + 256;
    static constexpr int kTotal = kStatsOff + 2 * BN * 4 + 1024 /* alignment slack */;
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(kThreadsConv, (BN <= 128 ? 2 : 1))
igemm_kmajor_kernel(const __grid_constant__ KmParams p) {
    using L = KmSmem<BN, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full = empty_bar + STAGES;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full + 1);
    float* s_stats = reinterpret_cast<float*>(smem + L::kStatsOff);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x;
    const int tw_i = tile % p.tiles_w;
    const int th_i = (tile / p.tiles_w) % p.tiles_h;
    const int n_img = tile / (p.tiles_w * p.tiles_h);
    const int q0 = tw_i * p.TW, p0 = th_i * p.TH;
    const int n0 = blockIdx.y * BN;
    const int total_iters = p.ntaps * p.kchunks;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.mapB);
        tma_prefetch_desc(&p.mapA[0]);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    } else if (warp == 2) {
        tmem_alloc<BN>(tmem_ptr);
    } else if (warp == 3) {
        for (int i = lane; i < 2 * BN; i += 32) s_stats[i] = 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            for (int it = 0; it < total_iters; ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
                mbar_wait(&empty_bar[s], ph ^ 1u);
                const int tap = it / p.kchunks, kc = it - tap * p.kchunks;
                const Tap T = p.taps[tap];
                uint8_t* a_s = smem + s * L::kStageBytes;
                uint8_t* b_s = a_s + kABytes;
                mbar_arrive_expect_tx(&full_bar[s], (uint32_t)L::kStageBytes);
                tma_load_4d(a_s, &p.mapA[T.map], &full_bar[s], kc * 64, q0 + T.dw, p0 + T.dh, n_img);
                tma_load_2d(b_s, &p.mapB, &full_bar[s], T.bk + kc * 64, n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(128, BN, false, false);
            for (int it = 0; it < total_iters; ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem + s * L::kStageBytes);
                const uint32_t b_addr = a_addr + kABytes;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    // K-major, SWIZZLE_128B: 8-row groups are 1024 B apart; advance 16 bf16 (32 B) inside the atom
                    uint64_t da = make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
                    uint64_t db = make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
                    umma_bf16(tmem_base, da, db, idesc, (it | k) != 0 ? 1u : 0u);
                }
                umma_commit(&empty_bar[s]);  // frees the smem slot when these MMAs retire
            }
            if (total_iters > 0) umma_commit(tmem_full);
        }
    } else if (warp >= 4) {
        const int ew = warp - 4;  // == warp % 4 → TMEM lane quarter
        const int m = ew * 32 + lane;
        const int pp = p0 + (m >> p.tw_shift), qq = q0 + (m & (p.TW - 1));
        const bool valid = (pp < p.P) && (qq < p.Q);
        const long long pix_off = p.out_base + (long long)n_img * p.out_n_stride + (long long)pp * p.out_p_stride +
                                  (long long)qq * p.out_q_stride;
        if (total_iters > 0) {
            mbar_wait(tmem_full, 0);
            tc_fence_after();
        }
        const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16);
        const bool do_stats = (p.sum != nullptr);
#pragma unroll 1
        for (int j = 0; j < BN / 32; ++j) {
            uint32_t r[32];
            if (total_iters > 0) {
                tmem_ld_32x32b_x32(taddr + j * 32, r);
                tmem_ld_wait();
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) r[i] = 0u;
            }
            const int col0 = n0 + j * 32;
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
            if (p.bias != nullptr) {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (col0 + i < p.k_real) v[i] += __ldg(p.bias + col0 + i);
            }
            if (p.out_f32) {
                float* o = reinterpret_cast<float*>(p.out) + pix_off + col0;
                if (valid) {
#pragma unroll
                    for (int g = 0; g < 8; ++g)
                        if (col0 + g * 4 < p.k_store)
                            *reinterpret_cast<float4*>(o + g * 4) = make_float4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
                }
            } else {
                __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + pix_off + col0;
                if (p.accumulate && valid) {
#pragma unroll
                    for (int g = 0; g < 4; ++g)
                        if (col0 + g * 8 < p.k_store) {
                            float old[8];
                            uint4 u = *reinterpret_cast<const uint4*>(o + g * 8);
                            unpack8(u, old);
#pragma unroll
                            for (int i = 0; i < 8; ++i) v[g * 8 + i] += old[i];
                        }
                }
                // round once to bf16; statistics are taken of the rounded values the next layer will see
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = bf16_round(v[i]);
                if (valid) {
#pragma unroll
                    for (int g = 0; g < 4; ++g)
                        if (col0 + g * 8 < p.k_store) *reinterpret_cast<uint4*>(o + g * 8) = pack8(&v[g * 8]);
                }
            }
            if (do_stats) {
                float s1[32], s2[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) { s1[i] = valid ? v[i] : 0.f; s2[i] = s1[i] * s1[i]; }
                // warp reduce-scatter: after 5 rounds lane L holds the column-L total over the warp's 32 pixels
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) {
                    const bool upper = (lane & off) != 0;
#pragma unroll
                    for (int i = 0; i < off; ++i) {
                        float send1 = upper ? s1[i] : s1[i + off];
                        float keep1 = upper ? s1[i + off] : s1[i];
                        float send2 = upper ? s2[i] : s2[i + off];
                        float keep2 = upper ? s2[i + off] : s2[i];
                        s1[i] = keep1 + __shfl_xor_sync(0xffffffffu, send1, off);
                        s2[i] = keep2 + __shfl_xor_sync(0xffffffffu, send2, off);
                    }
                }
                atomicAdd(&s_stats[j * 32 + lane], s1[0]);
                atomicAdd(&s_stats[BN + j * 32 + lane], s2[0]);
            }
        }
        if (do_stats) {
            asm volatile("bar.sync 1, 128;" ::: "memory");  // the four epilogue warps only
            for (int i = threadIdx.x - 128; i < BN; i += 128) {
                if (n0 + i < p.k_real) {
                    atomicAdd(p.sum + n0 + i, s_stats[i]);
                    atomicAdd(p.sumsq + n0 + i, s_stats[BN + i]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<BN>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------------
// MN-major GEMM for weight gradients: dW[co, (tap, ci)] += Σ_pixels dY[pix, co] * X[pix + tap, ci]
// ------------------------------------------------------------------------------------------------
constexpr int kWgStages = 4;
constexpr int kWgBoxBytes = 64 * 128;                            // 64 pixels x 64 channels bf16
constexpr int kWgStageBytes = 2 * kWgBoxBytes + 4 * kWgBoxBytes;  // A (2 boxes) + up to 4 B boxes
constexpr int kWgBarOff = kWgStages * kWgStageBytes;
constexpr int kWgSmemTotal = kWgBarOff + 256 + 1024;
constexpr int kWgTmemCols = 256;

__global__ void __launch_bounds__(kThreadsConv, 1) wgrad_mnmajor_kernel(const __grid_constant__ WgParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kWgBarOff);
    uint64_t* empty_bar = full_bar + kWgStages;
    uint64_t* tmem_full = empty_bar + kWgStages;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int split = blockIdx.x;
    const int co0 = blockIdx.y * 128;
    const int box0 = blockIdx.z * p.nb;
    const int nb = min(p.nb, p.nboxes - box0);
    const int t_begin = split * p.tiles_per_split;
    const int t_end = min(p.total_tiles, t_begin + p.tiles_per_split);
    const int total_iters = t_end - t_begin;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.mapA);
        tma_prefetch_desc(&p.mapB[0]);
        for (int s = 0; s < kWgStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    } else if (warp == 2) {
        tmem_alloc<kWgTmemCols>(tmem_ptr);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            for (int it = 0; it < total_iters; ++it) {
                const int s = it % kWgStages;
                const uint32_t ph = (uint32_t)(it / kWgStages) & 1u;
                mbar_wait(&empty_bar[s], ph ^ 1u);
                const int t = t_begin + it;
                const int tw_i = t % p.tiles_w;
                const int th_i = (t / p.tiles_w) % p.tiles_h;
                const int n_img = t / (p.tiles_w * p.tiles_h);
                const int q0 = tw_i * p.TW, p0 = th_i * p.TH;
                uint8_t* a_s = smem + s * kWgStageBytes;
                uint8_t* b_s = a_s + 2 * kWgBoxBytes;
                mbar_arrive_expect_tx(&full_bar[s], (uint32_t)((2 + nb) * kWgBoxBytes));
                tma_load_4d(a_s, &p.mapA, &full_bar[s], co0, q0, p0, n_img);
                tma_load_4d(a_s + kWgBoxBytes, &p.mapA, &full_bar[s], co0 + 64, q0, p0, n_img);  // OOB → zeros
                for (int j = 0; j < nb; ++j) {
                    const int b = box0 + j;
                    const int tap = b / p.cchunks, cc = b - tap * p.cchunks;
                    const Tap T = p.taps[tap];
                    tma_load_4d(b_s + j * kWgBoxBytes, &p.mapB[T.map], &full_bar[s], cc * 64, q0 + T.dw, p0 + T.dh, n_img);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(128, 64 * nb, true, true);
            // MN-major, SWIZZLE_128B: 64-channel groups are LBO = 8192 B apart (one TMA box each), 8-pixel groups
            // SBO = 1024 B apart; one MMA consumes 16 pixels = 2048 B = 128 descriptor units
            const uint64_t d_hi = make_smem_desc_sw128(0, kWgBoxBytes, 1024);
            const uint32_t base_lo = smem_u32(smem) >> 4;
            int s = 0;
            uint32_t ph = 0, acc = 0;
            for (int it = 0; it < total_iters; ++it) {
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                const uint32_t a_lo = base_lo + (uint32_t)(s * (kWgStageBytes >> 4));
                const uint64_t da = d_hi | (uint64_t)a_lo;
                const uint64_t db = d_hi | (uint64_t)(a_lo + ((2 * kWgBoxBytes) >> 4));
                umma_bf16(tmem_base, da, db, idesc, acc);
                umma_bf16(tmem_base, da + 128, db + 128, idesc, 1u);
                umma_bf16(tmem_base, da + 256, db + 256, idesc, 1u);
                umma_bf16(tmem_base, da + 384, db + 384, idesc, 1u);
                acc = 1u;
                umma_commit(&empty_bar[s]);
                if (++s == kWgStages) { s = 0; ph ^= 1u; }
            }
            if (total_iters > 0) umma_commit(tmem_full);
        }
    } else if (warp >= 4 && total_iters > 0) {
        const int ew = warp - 4;
        const int co = co0 + ew * 32 + lane;
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16);
        for (int j = 0; j < nb; ++j) {
            const int b = box0 + j;
            const int tap = b / p.cchunks, cc = b - tap * p.cchunks;
            float* dst = p.dw + (long long)co * p.dw_row_stride + (long long)tap * p.C + cc * 64;
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                uint32_t r[32];
                tmem_ld_32x32b_x32(taddr + j * 64 + half * 32, r);
                tmem_ld_wait();
                if (co < p.K) {
#pragma unroll
                    for (int g = 0; g < 8; ++g)
                        red_add_v4(dst + half * 32 + g * 4, __uint_as_float(r[g * 4]), __uint_as_float(r[g * 4 + 1]),
                                   __uint_as_float(r[g * 4 + 2]), __uint_as_float(r[g * 4 + 3]));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<kWgTmemCols>(tmem_base);
    }
}

template <int BN, int STAGES>
int launch_km(const KmParams& prm, int tiles, int n_tiles, cudaStream_t st) {
    using L = KmSmem<BN, STAGES>;
    { int rc_ = tsb_ensure_dyn_smem(reinterpret_cast<const void*>(igemm_kmajor_kernel<BN, STAGES>), L::kTotal); if (rc_) return rc_; }
    igemm_kmajor_kernel<BN, STAGES><<<dim3(tiles, n_tiles), kThreadsConv, L::kTotal, st>>>(prm);
    TSB_CUDA_CHECK_LAUNCH("igemm_kmajor");
    return TSB_OK;
}

int launch_km_auto(KmParams& prm, const void* w, int w_rows, long long w_k, int ncols, int tiles, cudaStream_t st) {
    // B map box depends on BN
    int BN = (ncols <= 64) ? 64 : 128;
    int rc = encode_2d(&prm.mapB, w, (uint64_t)w_k, (uint64_t)w_rows, (uint64_t)w_k * 2, 64, BN);
    if (rc) return rc;
    int n_tiles = (ncols + BN - 1) / BN;
    if (BN == 64) return launch_km<64, 4>(prm, tiles, n_tiles, st);
    return launch_km<128, 3>(prm, tiles, n_tiles, st);
}

int launch_wg(const WgParams& prm, int splits, int co_tiles, int groups, cudaStream_t st) {
    { int rc_ = tsb_ensure_dyn_smem(reinterpret_cast<const void*>(wgrad_mnmajor_kernel), kWgSmemTotal); if (rc_) return rc_; }
    wgrad_mnmajor_kernel<<<dim3(splits, co_tiles, groups), kThreadsConv, kWgSmemTotal, st>>>(prm);
    TSB_CUDA_CHECK_LAUNCH("wgrad_mnmajor");
    return TSB_OK;
}

}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================
static int fprop_impl(const tsb_conv_shape* s, const void* x, int xcs, const void* w, const float* bias, void* y, int ydtype,
                      int ycs, float* sum, float* sumsq, const void* res, int relu, tsb_stream_t stream) {
    int rc = check_conv_shape(s, "tsb_conv2d_fprop");
    if (rc) return rc;
    TSB_REQUIRE(x && w && y, "tsb_conv2d_fprop: null pointer");
    TSB_REQUIRE(s->C % 64 == 0, "tsb_conv2d_fprop: C must be a multiple of 64 (got %d)", s->C);
    TSB_REQUIRE(xcs % 8 == 0 && xcs >= s->C && tsb_aligned16(x) && tsb_aligned16(w), "tsb_conv2d_fprop: x alignment");
    TSB_REQUIRE(ydtype == TSB_BF16 || ydtype == TSB_F32, "tsb_conv2d_fprop: bad ydtype");
    TSB_REQUIRE(ycs % 8 == 0 && tsb_aligned16(y), "tsb_conv2d_fprop: y channel stride must be a multiple of 8");
    TSB_REQUIRE((sum == nullptr) == (sumsq == nullptr), "tsb_conv2d_fprop: sum and sumsq go together");
    TSB_REQUIRE(!(sum && ydtype != TSB_BF16), "tsb_conv2d_fprop: statistics need bf16 output");
    cudaStream_t st = (cudaStream_t)stream;

    KmParams prm;
    memset(&prm, 0, sizeof(prm));
    int N = s->N, H = s->H, W = s->W, P = s->P, Q = s->Q;
    const bool flat = (s->R == 1 && s->stride == 1 && s->pad == 0);
    if (flat) {  // 1x1: pure GEMM over all pixels
        long long npix = (long long)N * H * W;
        TSB_REQUIRE(npix < (1ll << 31), "tsb_conv2d_fprop: too many pixels");
        W = (int)npix; H = 1; N = 1; P = 1; Q = W;
    }
    int TW, TH;
    if (flat) { TW = 128; TH = 1; } else pick_tile(Q, 128, &TW, &TH);
    rc = encode_act_maps(prm.mapA, x, N, H, W, s->C, xcs, s->stride, TW, TH);
    if (rc) return rc;
    int nt = 0;
    for (int r = 0; r < s->R; ++r)
        for (int ss = 0; ss < s->S; ++ss) {
            Tap& T = prm.taps[nt++];
            int th = r * s->dil - s->pad, tw = ss * s->dil - s->pad;
            if (s->stride == 1) { T.dh = th; T.dw = tw; T.map = 0; }
            else {
                int ph = posmod(th, 2), pw = posmod(tw, 2);
                T.dh = (th - ph) / 2; T.dw = (tw - pw) / 2; T.map = ph * 2 + pw;
            }
            T.bk = (r * s->S + ss) * s->C;
        }
    prm.ntaps = nt;
    prm.kchunks = s->C / 64;
    prm.tiles_w = (Q + TW - 1) / TW;
    prm.tiles_h = (P + TH - 1) / TH;
    prm.TW = TW; prm.TH = TH; prm.tw_shift = ilog2(TW);
    prm.P = P; prm.Q = Q;
    prm.out_n_stride = (long long)P * Q * ycs;
    prm.out_p_stride = (long long)Q * ycs;
    prm.out_q_stride = ycs;
    prm.out_base = 0;
    prm.k_real = s->K;
    int kst = (s->K + 7) / 8 * 8;
    if (kst > ycs) kst = ycs;
    prm.k_store = kst;
    prm.out_f32 = (ydtype == TSB_F32);
    prm.accumulate = 0;
    prm.bias = bias; prm.sum = sum; prm.sumsq = sumsq; prm.out = y;
    if (convv2::g_enabled) {
        convv2::Desc d;
        memset(&d, 0, sizeof(d));
        d.act = x; d.aN = N; d.aH = H; d.aW = W; d.aC = s->C; d.acs = xcs; d.act_stride = s->stride; d.flat = flat;
        for (int t = 0; t < nt; ++t) d.taps[t] = prm.taps[t];
        d.ntaps = nt; d.kchunks = s->C / 64;
        d.w = w; d.w_rows = s->K; d.w_k = (long long)s->R * s->S * s->C; d.ncols = kst;
        d.Nl = N; d.Pl = P; d.Ql = Q;
        d.out_n_stride = prm.out_n_stride; d.out_p_stride = prm.out_p_stride; d.out_q_stride = prm.out_q_stride; d.out_base = 0;
        d.k_real = s->K; d.k_store = kst; d.out_f32 = prm.out_f32; d.accumulate = 0;
        d.bias = bias; d.sum = sum; d.sumsq = sumsq; d.out = y;
        d.relu = relu; d.res = res;
        return convv2::launch(d, st);
    }
    TSB_REQUIRE(res == nullptr && relu == 0, "tsb_conv2d_fprop_fused: the fused inference epilogue needs the persistent kernel (tsb_debug_set 4)");
    int tiles = prm.tiles_w * prm.tiles_h * N;
    return launch_km_auto(prm, w, s->K, (long long)s->R * s->S * s->C, kst, tiles, st);
}

extern "C" int tsb_conv2d_fprop(const tsb_conv_shape* s, const void* x, int xcs, const void* w, const float* bias,
                                void* y, int ydtype, int ycs, float* sum, float* sumsq, tsb_stream_t stream) {
    return fprop_impl(s, x, xcs, w, bias, y, ydtype, ycs, sum, sumsq, nullptr, 0, stream);
}

extern "C" int tsb_conv2d_fprop_fused(const tsb_conv_shape* s, const void* x, int xcs, const void* w, const float* bias,
                                      const void* res, int rescs, int relu, void* y, int ydtype, int ycs,
                                      tsb_stream_t stream) {
    TSB_REQUIRE(res == nullptr || (rescs == ycs && ydtype == TSB_BF16 && tsb_aligned16(res)),
                "tsb_conv2d_fprop_fused: the residual must be a bf16 tensor laid out exactly like the output (rescs == ycs)");
    return fprop_impl(s, x, xcs, w, bias, y, ydtype, ycs, nullptr, nullptr, res, relu != 0, stream);
}

extern "C" int tsb_conv2d_dgrad(const tsb_conv_shape* s, const void* dy, int dycs, const void* wt, void* dx, int dxcs,
                                int accumulate, tsb_stream_t stream) {
    int rc = check_conv_shape(s, "tsb_conv2d_dgrad");
    if (rc) return rc;
    TSB_REQUIRE(dy && wt && dx, "tsb_conv2d_dgrad: null pointer");
    TSB_REQUIRE(s->K % 64 == 0, "tsb_conv2d_dgrad: K must be a multiple of 64 (pad the classifier to 64), got %d", s->K);
    TSB_REQUIRE(s->C % 8 == 0 && dycs % 8 == 0 && dycs >= s->K && dxcs % 8 == 0 && dxcs >= s->C && tsb_aligned16(dy) && tsb_aligned16(dx) && tsb_aligned16(wt),
                "tsb_conv2d_dgrad: alignment");
    cudaStream_t st = (cudaStream_t)stream;
    const int classes = (s->stride == 1) ? 1 : 4;
    for (int cls = 0; cls < classes; ++cls) {
        const int ph = cls >> 1, pw = cls & 1;
        KmParams prm;
        memset(&prm, 0, sizeof(prm));
        int N = s->N, P = s->P, Q = s->Q;          // dY geometry (the A operand)
        int Hl, Wl;                                 // launch output grid
        const bool flat = (s->R == 1 && s->stride == 1 && s->pad == 0);
        if (s->stride == 1) { Hl = s->H; Wl = s->W; }
        else { Hl = (s->H - ph + 1) / 2; Wl = (s->W - pw + 1) / 2; }
        if (Hl <= 0 || Wl <= 0) continue;
        if (flat) {
            long long npix = (long long)N * P * Q;
            TSB_REQUIRE(npix < (1ll << 31), "tsb_conv2d_dgrad: too many pixels");
            Q = (int)npix; P = 1; N = 1; Hl = 1; Wl = Q;
        }
        int TW, TH;
        if (flat) { TW = 128; TH = 1; } else pick_tile(Wl, 128, &TW, &TH);
        rc = encode_act_maps(prm.mapA, dy, N, P, Q, s->K, dycs, 1, TW, TH);
        if (rc) return rc;
        int nt = 0;
        for (int r = 0; r < s->R; ++r) {
            int th = (s->stride == 1 ? 0 : ph) + s->pad - r * s->dil;
            if (s->stride == 2 && posmod(th, 2) != 0) continue;
            for (int ss = 0; ss < s->S; ++ss) {
                int tw = (s->stride == 1 ? 0 : pw) + s->pad - ss * s->dil;
                if (s->stride == 2 && posmod(tw, 2) != 0) continue;
                Tap& T = prm.taps[nt++];
                T.dh = (s->stride == 1) ? th : th / 2;   // th even here, exact
                T.dw = (s->stride == 1) ? tw : tw / 2;
                T.map = 0;
                // wt is [C][R][S][K] with flipped taps: index (R-1-r, S-1-ss)
                T.bk = ((s->R - 1 - r) * s->S + (s->S - 1 - ss)) * s->K;
            }
        }
        prm.ntaps = nt;
        prm.kchunks = s->K / 64;
        prm.tiles_w = (Wl + TW - 1) / TW;
        prm.tiles_h = (Hl + TH - 1) / TH;
        prm.TW = TW; prm.TH = TH; prm.tw_shift = ilog2(TW);
        prm.P = Hl; prm.Q = Wl;
        if (flat) {
            prm.out_n_stride = 0; prm.out_p_stride = 0; prm.out_q_stride = dxcs; prm.out_base = 0;
        } else if (s->stride == 1) {
            prm.out_n_stride = (long long)s->H * s->W * dxcs;
            prm.out_p_stride = (long long)s->W * dxcs;
            prm.out_q_stride = dxcs;
            prm.out_base = 0;
        } else {
            prm.out_n_stride = (long long)s->H * s->W * dxcs;
            prm.out_p_stride = (long long)2 * s->W * dxcs;
            prm.out_q_stride = (long long)2 * dxcs;
            prm.out_base = ((long long)ph * s->W + pw) * dxcs;
        }
        prm.k_real = s->C;
        prm.k_store = s->C;
        prm.out_f32 = 0;
        prm.accumulate = accumulate;
        prm.out = dx;
        if (convv2::g_enabled) {
            convv2::Desc d;
            memset(&d, 0, sizeof(d));
            d.act = dy; d.aN = N; d.aH = P; d.aW = Q; d.aC = s->K; d.acs = dycs; d.act_stride = 1; d.flat = flat;
            for (int t = 0; t < nt; ++t) d.taps[t] = prm.taps[t];
            d.ntaps = nt; d.kchunks = s->K / 64;
            d.w = wt; d.w_rows = s->C; d.w_k = (long long)s->R * s->S * s->K; d.ncols = s->C;
            d.Nl = N; d.Pl = Hl; d.Ql = Wl;
            d.out_n_stride = prm.out_n_stride; d.out_p_stride = prm.out_p_stride; d.out_q_stride = prm.out_q_stride;
            d.out_base = prm.out_base;
            d.k_real = s->C; d.k_store = s->C; d.out_f32 = 0; d.accumulate = accumulate;
            d.out = dx;
            rc = convv2::launch(d, st);
            if (rc) return rc;
            continue;
        }
        int tiles = prm.tiles_w * prm.tiles_h * N;
        rc = launch_km_auto(prm, wt, s->C, (long long)s->R * s->S * s->K, s->C, tiles, st);
        if (rc) return rc;
    }
    return TSB_OK;
}

namespace {
int pick_nb(int nboxes) {
    if (nboxes % 4 == 0) return 4;
    if (nboxes % 3 == 0) return 3;
    if (nboxes % 2 == 0) return 2;
    return 1;
}
int wgrad_common(WgParams& prm, int N, int P, int Q, int K, cudaStream_t st) {
    prm.tiles_w = (Q + prm.TW - 1) / prm.TW;
    prm.tiles_h = (P + prm.TH - 1) / prm.TH;
    prm.total_tiles = prm.tiles_w * prm.tiles_h * N;
    prm.nboxes = prm.ntaps * prm.cchunks;
    prm.nb = pick_nb(prm.nboxes);
    const int groups = prm.nboxes / prm.nb;
    const int co_tiles = (K + 127) / 128;
    // split the pixel range so the grid covers ~2 waves of SMs, at least 8 tiles per CTA
    int base = groups * co_tiles;
    int want = (convv2::g_wgrad_waves_x * tsb_num_sms() + base - 1) / base;
    int max_splits = (prm.total_tiles + 7) / 8;
    if (want > max_splits) want = max_splits;
    if (want < 1) want = 1;
    prm.tiles_per_split = (prm.total_tiles + want - 1) / want;
    int splits = (prm.total_tiles + prm.tiles_per_split - 1) / prm.tiles_per_split;
    return launch_wg(prm, splits, co_tiles, groups, st);
}
}  // namespace

extern "C" int tsb_conv2d_wgrad(const tsb_conv_shape* s, const void* x, int xcs, const void* dy, int dycs, float* dw,
                                tsb_stream_t stream) {
    int rc = check_conv_shape(s, "tsb_conv2d_wgrad");
    if (rc) return rc;
    TSB_REQUIRE(x && dy && dw, "tsb_conv2d_wgrad: null pointer");
    TSB_REQUIRE(s->C % 64 == 0, "tsb_conv2d_wgrad: C must be a multiple of 64 (got %d)", s->C);
    TSB_REQUIRE(xcs % 8 == 0 && dycs % 8 == 0 && dycs >= s->K && tsb_aligned16(x) && tsb_aligned16(dy) && tsb_aligned16(dw),
                "tsb_conv2d_wgrad: alignment");
    cudaStream_t st = (cudaStream_t)stream;
    WgParams prm;
    memset(&prm, 0, sizeof(prm));
    int N = s->N, H = s->H, W = s->W, P = s->P, Q = s->Q;
    const bool flat = (s->R == 1 && s->stride == 1 && s->pad == 0);
    if (flat) {
        long long npix = (long long)N * H * W;
        TSB_REQUIRE(npix < (1ll << 31), "tsb_conv2d_wgrad: too many pixels");
        W = (int)npix; H = 1; N = 1; P = 1; Q = W;
    }
    int TW, TH;
    if (flat) { TW = 64; TH = 1; } else pick_tile(Q, 64, &TW, &TH);
    prm.TW = TW; prm.TH = TH; prm.tw_shift = ilog2(TW);
    rc = encode_4d(&prm.mapA, dy, s->K, Q, P, N, (uint64_t)dycs * 2, (uint64_t)Q * dycs * 2, (uint64_t)P * Q * dycs * 2, 64, TW, TH, 1);
    if (rc) return rc;
    rc = encode_act_maps(prm.mapB, x, N, H, W, s->C, xcs, s->stride, TW, TH);
    if (rc) return rc;
    int nt = 0;
    for (int r = 0; r < s->R; ++r)
        for (int ss = 0; ss < s->S; ++ss) {
            Tap& T = prm.taps[nt++];
            int th = r * s->dil - s->pad, tw = ss * s->dil - s->pad;
            if (s->stride == 1) { T.dh = th; T.dw = tw; T.map = 0; }
            else {
                int ph = posmod(th, 2), pw = posmod(tw, 2);
                T.dh = (th - ph) / 2; T.dw = (tw - pw) / 2; T.map = ph * 2 + pw;
            }
            T.bk = 0;
        }
    prm.ntaps = nt;
    prm.cchunks = s->C / 64;
    prm.K = s->K;
    prm.C = s->C;
    prm.dw_row_stride = (long long)s->R * s->S * s->C;
    prm.dw = dw;
    if (convv2::g_enabled && !flat && s->R > 1) {
        convv2::WgradDesc d;
        memset(&d, 0, sizeof(d));
        d.x = x; d.dy = dy; d.N = N; d.H = H; d.W = W; d.C = s->C; d.xcs = xcs; d.P = P; d.Q = Q; d.K = s->K; d.dycs = dycs;
        d.stride = s->stride;
        for (int t = 0; t < nt; ++t) d.taps[t] = prm.taps[t];
        d.ntaps = nt; d.dw_row_stride = prm.dw_row_stride; d.dw = dw;
        rc = convv2::launch_wgrad_taps(d, s->R, s->pad, s->dil, st);   // 3x3 stride 1: all nine taps in one CTA
        if (rc != TSB_ERR_UNSUPPORTED) return rc;
        rc = convv2::launch_wgrad_rows(d, st);
        if (rc != TSB_ERR_UNSUPPORTED) return rc;
    }
    return wgrad_common(prm, N, P, Q, s->K, st);
}

// ---------------------------------------------------------------------------------- 7x7/2 stem
namespace {
// overlapped-window view of the s2d image: coordinate (e, ow, oh, n) → xs2d[n][oh][ow .. ow+3][0..15]
int encode_stem_map(CUtensorMap* m, const void* xs2d, int N, int H2, int W2, int TW, int TH) {
    const uint64_t row = (uint64_t)(W2 + 4) * 16 * 2;
    return encode_4d(m, xs2d, 64, (uint64_t)W2, (uint64_t)H2, (uint64_t)N, 32, row, row * H2, 64, TW, TH, 1);
}
}  // namespace

static int stem_fprop_impl(const void* xs2d, int N, int H, int W, const void* wp, int K, void* y, int ycs, float* sum,
                           float* sumsq, const float* bias, int relu, tsb_stream_t stream) {
    TSB_REQUIRE(xs2d && wp && y && N > 0 && H > 0 && W > 0 && K > 0, "tsb_conv_stem_fprop: bad args");
    TSB_REQUIRE(H % 2 == 0 && W % 2 == 0 && ycs % 8 == 0 && K % 8 == 0, "tsb_conv_stem_fprop: H,W even; K, ycs multiples of 8");
    TSB_REQUIRE((sum == nullptr) == (sumsq == nullptr), "tsb_conv_stem_fprop: sum and sumsq go together");
    const int H2 = H / 2, W2 = W / 2;
    KmParams prm;
    memset(&prm, 0, sizeof(prm));
    int TW, TH;
    pick_tile(W2, 128, &TW, &TH);
    int rc = encode_stem_map(&prm.mapA[0], xs2d, N, H2, W2, TW, TH);
    if (rc) return rc;
    for (int a = 0; a < 4; ++a) { prm.taps[a].dh = a - 2; prm.taps[a].dw = 0; prm.taps[a].map = 0; prm.taps[a].bk = a * 64; }
    prm.ntaps = 4;
    prm.kchunks = 1;
    prm.tiles_w = (W2 + TW - 1) / TW;
    prm.tiles_h = (H2 + TH - 1) / TH;
    prm.TW = TW; prm.TH = TH; prm.tw_shift = ilog2(TW);
    prm.P = H2; prm.Q = W2;
    prm.out_n_stride = (long long)H2 * W2 * ycs;
    prm.out_p_stride = (long long)W2 * ycs;
    prm.out_q_stride = ycs;
    prm.k_real = K; prm.k_store = K;
    prm.out_f32 = 0; prm.accumulate = 0;
    prm.sum = sum; prm.sumsq = sumsq; prm.out = y;
    if (convv2::g_enabled) {
        convv2::Desc d;
        memset(&d, 0, sizeof(d));
        d.act = xs2d; d.aN = N; d.aH = H2; d.aW = W2; d.aC = 64; d.acs = 16; d.act_stride = 1; d.stem = 1;
        for (int a = 0; a < 4; ++a) d.taps[a] = prm.taps[a];
        d.ntaps = 4; d.kchunks = 1;
        d.w = wp; d.w_rows = K; d.w_k = 256; d.ncols = K;
        d.Nl = N; d.Pl = H2; d.Ql = W2;
        d.out_n_stride = prm.out_n_stride; d.out_p_stride = prm.out_p_stride; d.out_q_stride = prm.out_q_stride;
        d.k_real = K; d.k_store = K;
        d.sum = sum; d.sumsq = sumsq; d.out = y;
        d.bias = bias; d.relu = relu;
        return convv2::launch(d, (cudaStream_t)stream);
    }
    TSB_REQUIRE(bias == nullptr && relu == 0, "tsb_conv_stem_fprop_fused: the fused inference epilogue needs the persistent kernel");
    int tiles = prm.tiles_w * prm.tiles_h * N;
    return launch_km_auto(prm, wp, K, 256, K, tiles, (cudaStream_t)stream);
}

extern "C" int tsb_conv_stem_fprop(const void* xs2d, int N, int H, int W, const void* wp, int K, void* y, int ycs,
                                   float* sum, float* sumsq, tsb_stream_t stream) {
    return stem_fprop_impl(xs2d, N, H, W, wp, K, y, ycs, sum, sumsq, nullptr, 0, stream);
}

extern "C" int tsb_conv_stem_fprop_fused(const void* xs2d, int N, int H, int W, const void* wp, int K, const float* bias,
                                         int relu, void* y, int ycs, tsb_stream_t stream) {
    return stem_fprop_impl(xs2d, N, H, W, wp, K, y, ycs, nullptr, nullptr, bias, relu != 0, stream);
}

extern "C" int tsb_conv_stem_wgrad(const void* xs2d, int N, int H, int W, const void* dy, int dycs, int K, float* dwp,
                                   tsb_stream_t stream) {
    TSB_REQUIRE(xs2d && dy && dwp && N > 0 && H > 0 && W > 0 && K > 0, "tsb_conv_stem_wgrad: bad args");
    TSB_REQUIRE(H % 2 == 0 && W % 2 == 0 && dycs % 8 == 0 && K % 8 == 0, "tsb_conv_stem_wgrad: H,W even; K, dycs multiples of 8");
    const int H2 = H / 2, W2 = W / 2;
    WgParams prm;
    memset(&prm, 0, sizeof(prm));
    int TW, TH;
    pick_tile(W2, 64, &TW, &TH);
    prm.TW = TW; prm.TH = TH; prm.tw_shift = ilog2(TW);
    int rc = encode_4d(&prm.mapA, dy, K, W2, H2, N, (uint64_t)dycs * 2, (uint64_t)W2 * dycs * 2, (uint64_t)H2 * W2 * dycs * 2, 64, TW, TH, 1);
    if (rc) return rc;
    rc = encode_stem_map(&prm.mapB[0], xs2d, N, H2, W2, TW, TH);
    if (rc) return rc;
    for (int a = 0; a < 4; ++a) { prm.taps[a].dh = a - 2; prm.taps[a].dw = 0; prm.taps[a].map = 0; prm.taps[a].bk = 0; }
    prm.ntaps = 4;
    prm.cchunks = 1;
    prm.K = K;
    prm.C = 64;             // each tap row contributes 64 packed elements
    prm.dw_row_stride = 256;  // dwp [K][4][64]
    prm.dw = dwp;
    return wgrad_common(prm, N, H2, W2, K, (cudaStream_t)stream);
}
