// wgrad of 3x3 stride-1 convolutions with ALL NINE TAPS IN ONE CTA ("taps" kernel).
//
//   dW[co, (r,s), ci] = Σ_px dY[px, co] · X[px + ((r-1)·dil, (s-1)·dil), ci]
//
// GEMM roles (swapped with respect to wgrad_rows_kernel): the M side carries (tap, ci), the N side carries co:
//   A = X   : MN-major straight out of NHWC shared-memory rows (64 input channels = 128 B per pixel); ONE UMMA
//             descriptor covers TWO taps — its two 64-row M-groups are the same pixel rows shifted by the tap distance
//             (leading-dimension byte offset = dil·128 B for two taps of a filter row, = the row-slot pitch for the
//             third taps of two filter rows);
//   B = dY  : MN-major, N = 64 output channels, K = 64 pixels per stage;
//   D_j[128 lanes = (tap pair j, ci), 64 columns = co], j = 0..4 → 320 TMEM columns.
// Tap pairs: j=0..2: (r=j; s=0,1), j=3: (r=0,s=2 | r=1,s=2), j=4: (r=2,s=2 | junk: never stored).
//
// Why: the row kernel maps co to M. M = 64 costs a tcgen05.mma as many cycles as M = 128, so the K = 64 layers (layer1, the
// stems' neighbours — the biggest wgrads of the step) ran half-empty MMAs; it also runs one CTA per filter row, reading dY
// (and X) three times. Here every MMA is a full 128 x 64 x 16 (4.5 useful of 5 per K-step), dY is staged once for all nine
// taps and X once per filter row.
//
// Per stage (64 pixels = TH rows of TW): dY box {64 co, TW, TH} (8 KB) + three X boxes {64 ci, TW + 2·dil, TH} (one per
// filter row, at rows p0 + (r-1)·dil; the zero-filled out-of-bounds halo IS the padding). 4 K-steps x 5 MMAs (N = 64: 32
// tensor-pipe cycles each) = 640 cycles per stage against ~780 cycles of L2→SM ingest for the 33 KB at ~43 B/clk/SM.
#include "conv_common.cuh"
#include "conv_v2.cuh"
#include "sm100_ptx.cuh"

using namespace sm100;
using namespace convhost;

namespace {

constexpr int kTpThreads = 256;   // warp 0: TMA, warp 1: MMA, warp 2: TMEM alloc, warps 4-7: epilogue
constexpr int kTpMaxStages = 6;
constexpr int kTpDyBytes = 64 * 128;

struct alignas(64) TpParams {
    CUtensorMap mapDy;   // {K, Q, P, N}, box {64, TW, TH, 1}
    CUtensorMap mapX;    // {C, W, H, N}, box {64, TW + 2·dil, TH, 1}
    int TW, TH, dil;
    int tiles_w, tiles_h, total_tiles, tiles_per_split;
    int slot_bytes;      // one X row-slot: (TW + 2·dil)·TH·128
    int stage_bytes, nstages;
    int koff[4];         // byte offset of K-step k inside an X slot: pixel (k·16) of the tile in the (TW+2·dil)-wide rows
    int kdy[4];          // ... and inside the dY tile (always k·2048)
    int K, C;
    long long dw_row_stride;   // R·S·C
    float* dw;
};

__global__ void __launch_bounds__(kTpThreads, 1) wgrad_taps_kernel(const __grid_constant__ TpParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* tail = smem + p.nstages * p.stage_bytes + 1024;   // 1 KB slack: the junk half of D_4 reads past the last slot
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
    uint64_t* empty_bar = full_bar + kTpMaxStages;
    uint64_t* tmem_full = empty_bar + kTpMaxStages;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int co0 = blockIdx.y * 64, ci0 = blockIdx.z * 64;
    const int t_begin = blockIdx.x * p.tiles_per_split;
    const int t_end = min(p.total_tiles, t_begin + p.tiles_per_split);
    const int iters = t_end - t_begin;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.mapDy);
        tma_prefetch_desc(&p.mapX);
        for (int s = 0; s < p.nstages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    } else if (warp == 2) {
        tmem_alloc<512>(tmem_ptr);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            const uint32_t bytes = (uint32_t)(kTpDyBytes + 3 * p.slot_bytes);
            for (int it = 0; it < iters; ++it) {
                mbar_wait(&empty_bar[s], ph ^ 1u);
                const int t = t_begin + it;
                const int q0 = (t % p.tiles_w) * p.TW;
                const int p0 = ((t / p.tiles_w) % p.tiles_h) * p.TH;
                const int n_img = t / (p.tiles_w * p.tiles_h);
                uint8_t* st = smem + s * p.stage_bytes;
                mbar_arrive_expect_tx(&full_bar[s], bytes);
                tma_load_4d(st, &p.mapDy, &full_bar[s], co0, q0, p0, n_img);
#pragma unroll
                for (int r = 0; r < 3; ++r)
                    tma_load_4d(st + kTpDyBytes + r * p.slot_bytes, &p.mapX, &full_bar[s], ci0, q0 - p.dil,
                                p0 + (r - 1) * p.dil, n_img);
                if (++s == p.nstages) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && iters > 0) {
            constexpr uint32_t idesc = make_idesc_bf16(128, 64, true, true);
            const uint32_t tapb = (uint32_t)p.dil * 128u;            // bytes between horizontally adjacent taps
            // descriptor high parts (start address 0): SBO = 1024 B between 8-pixel groups; LBO = M-group distance
            const uint64_t hi_row = make_smem_desc_sw128(0, tapb, 1024);                  // taps (r,0),(r,1)
            const uint64_t hi_col = make_smem_desc_sw128(0, (uint32_t)p.slot_bytes, 1024);  // taps (0,2),(1,2)
            const uint64_t hi_dy = make_smem_desc_sw128(0, 64 * 128, 1024);               // single N-group
            const uint32_t base_lo = smem_u32(smem) >> 4;
            const uint32_t slot16 = (uint32_t)p.slot_bytes >> 4;
            const uint32_t tap2 = (2u * tapb) >> 4;
            uint32_t ka[4], kb[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) { ka[k] = (uint32_t)p.koff[k] >> 4; kb[k] = (uint32_t)p.kdy[k] >> 4; }
            int s = 0;
            uint32_t ph = 0, acc = 0;
            for (int it = 0; it < iters; ++it) {
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                const uint32_t st_lo = base_lo + (uint32_t)(s * (p.stage_bytes >> 4));
                const uint32_t x_lo = st_lo + (kTpDyBytes >> 4);
                const uint64_t db0 = hi_dy | (uint64_t)st_lo;
                const uint64_t a0 = hi_row | (uint64_t)x_lo;
                const uint64_t a1 = hi_row | (uint64_t)(x_lo + slot16);
                const uint64_t a2 = hi_row | (uint64_t)(x_lo + 2 * slot16);
                const uint64_t a3 = hi_col | (uint64_t)(x_lo + tap2);
                const uint64_t a4 = hi_row | (uint64_t)(x_lo + 2 * slot16 + tap2);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t db = db0 + kb[k];
                    umma_bf16(tmem_base + 0, a0 + ka[k], db, idesc, acc);
                    umma_bf16(tmem_base + 64, a1 + ka[k], db, idesc, acc);
                    umma_bf16(tmem_base + 128, a2 + ka[k], db, idesc, acc);
                    umma_bf16(tmem_base + 192, a3 + ka[k], db, idesc, acc);
                    umma_bf16(tmem_base + 256, a4 + ka[k], db, idesc, acc);
                    acc = 1u;
                }
                umma_commit(&empty_bar[s]);
                if (++s == p.nstages) { s = 0; ph ^= 1u; }
            }
            umma_commit(tmem_full);
        }
    } else if (warp >= 4 && iters > 0) {
        // lanes 0-63 ↔ first tap of the pair, 64-127 ↔ second; column c ↔ output channel co0 + c.
        // dW is [K][3][3][C] fp32: a warp's 32 lanes add to 32 consecutive input channels (one 128-byte line) per column.
        const int ew = warp - 4;                 // TMEM lane quarter (== warp % 4)
        const int half = ew >> 1;
        const int ci = ci0 + (ew & 1) * 32 + lane;
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16);
        const bool ci_ok = ci < p.C;
#pragma unroll 1
        for (int j = 0; j < 5; ++j) {
            int tap;
            if (j < 3) tap = j * 3 + half;
            else if (j == 3) tap = half * 3 + 2;
            else tap = half ? -1 : 8;
            if (tap < 0) continue;               // junk half of D_4 (warp-uniform)
            float* dst = p.dw + (long long)tap * p.C + ci;
#pragma unroll 1
            for (int hc = 0; hc < 2; ++hc) {
                uint32_t r[32];
                tmem_ld_32x32b_x32(taddr + j * 64 + hc * 32, r);
                tmem_ld_wait();
                if (ci_ok) {
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const int co = co0 + hc * 32 + c;
                        if (co < p.K) atomicAdd(dst + (long long)co * p.dw_row_stride, __uint_as_float(r[c]));
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

}  // namespace

namespace convv2 {

int g_allow_taps = 1;   // tsb_debug_set key 8

// returns TSB_ERR_UNSUPPORTED when the shape is not eligible (caller falls back to the row / patch kernels)
int launch_wgrad_taps(const WgradDesc& d, int R, int pad, int dil, cudaStream_t st) {
    if (!g_allow_taps || R != 3 || d.stride != 1 || pad != dil || dil < 1 || dil > 8) return TSB_ERR_UNSUPPORTED;
    if (d.C % 64 != 0 || d.K % 8 != 0 || d.P != d.H || d.Q != d.W) return TSB_ERR_UNSUPPORTED;
    TpParams prm;
    memset(&prm, 0, sizeof(prm));
    int TW = 64;
    while (TW > 8 && TW / 2 >= d.Q) TW /= 2;      // widest power of two that is not more than ~2x the row
    const int TH = 64 / TW;
    prm.TW = TW; prm.TH = TH; prm.dil = dil;
    const int XW = TW + 2 * dil;
    prm.slot_bytes = XW * TH * 128;
    for (int k = 0; k < 4; ++k) {
        const int px = k * 16;
        prm.koff[k] = ((px / TW) * XW + (px % TW)) * 128;
        prm.kdy[k] = px * 128;
    }
    if (TW < 16) return TSB_ERR_UNSUPPORTED;       // a 16-pixel K-step must stay inside one tile row
    prm.stage_bytes = (kTpDyBytes + 3 * prm.slot_bytes + 1023) / 1024 * 1024;
    int ns = (227 * 1024 - 1024 /*align*/ - 1024 /*slack*/ - 256 /*barriers*/) / prm.stage_bytes;
    if (ns > kTpMaxStages) ns = kTpMaxStages;
    if (ns < 2) return TSB_ERR_UNSUPPORTED;
    prm.nstages = ns;
    int rc = encode_4d(&prm.mapDy, d.dy, d.K, d.Q, d.P, d.N, (uint64_t)d.dycs * 2, (uint64_t)d.Q * d.dycs * 2,
                       (uint64_t)d.P * d.Q * d.dycs * 2, 64, TW, TH, 1);
    if (rc) return rc;
    rc = encode_4d(&prm.mapX, d.x, d.C, d.W, d.H, d.N, (uint64_t)d.xcs * 2, (uint64_t)d.W * d.xcs * 2,
                   (uint64_t)d.H * d.W * d.xcs * 2, 64, XW, TH, 1);
    if (rc) return rc;
    prm.tiles_w = (d.Q + TW - 1) / TW;
    prm.tiles_h = (d.P + TH - 1) / TH;
    prm.total_tiles = prm.tiles_w * prm.tiles_h * d.N;
    prm.K = d.K; prm.C = d.C; prm.dw_row_stride = d.dw_row_stride; prm.dw = d.dw;
    const int co_tiles = (d.K + 63) / 64, ci_tiles = d.C / 64;
    const int base = co_tiles * ci_tiles;
    // one CTA per SM (persistent over a contiguous pixel range); at least 8 tiles per CTA so the 320-column drain
    // (288 fp32 reductions per thread) stays small next to the main loop
    // ... and the split count that minimises waves x tiles-per-CTA (wave quantisation: 192 CTAs cost two full waves)
    const int sms = tsb_num_sms();
    int max_splits = (prm.total_tiles + 7) / 8;
    if (max_splits < 1) max_splits = 1;
    if (max_splits > 2 * sms) max_splits = 2 * sms;
    int want = 1;
    long long best = -1;
    for (int sp = 1; sp <= max_splits; ++sp) {
        const long long per = (prm.total_tiles + sp - 1) / sp;
        const long long waves = ((long long)base * sp + sms - 1) / sms;
        const long long cost = waves * (per + 6);     // + ~6 stage-times of prologue / 320-column drain per CTA
        if (best < 0 || cost < best) { best = cost; want = sp; }
    }
    prm.tiles_per_split = (prm.total_tiles + want - 1) / want;
    const int splits = (prm.total_tiles + prm.tiles_per_split - 1) / prm.tiles_per_split;
    const size_t smem = 1024 + (size_t)ns * prm.stage_bytes + 1024 + 256;
    rc = tsb_ensure_dyn_smem(reinterpret_cast<const void*>(wgrad_taps_kernel), smem);
    if (rc) return rc;
    wgrad_taps_kernel<<<dim3(splits, co_tiles, ci_tiles), kTpThreads, smem, st>>>(prm);
    TSB_CUDA_CHECK_LAUNCH("wgrad_taps");
    return TSB_OK;
}

}  // namespace convv2
