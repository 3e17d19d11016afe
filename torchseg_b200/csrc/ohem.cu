// OHEM cross-entropy for sm_100a — replaces ProbOhemCrossEntropy2d.forward
// (/root/reference/furnace/seg_opr/loss_opr.py:68-98).  HBM-bound byte work: coalesced loads,
// warp-shuffle / shared-memory histogram reductions, an exact radix *select* instead of the full sort.
#include "tsb_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int ST_HIST = 0;  // 4096 bins
constexpr int ST_NUM_VALID = TSB_OHEM_ST_NUM_VALID;
constexpr int ST_COUNT_LE = TSB_OHEM_ST_COUNT_LE;
constexpr int ST_ACTIVE = TSB_OHEM_ST_ACTIVE;
constexpr int ST_THRESH = TSB_OHEM_ST_THRESH;
constexpr int ST_KEPT = TSB_OHEM_ST_KEPT;
constexpr int ST_LOSS = TSB_OHEM_ST_LOSS;
constexpr int ST_INVDEN = TSB_OHEM_ST_INVDEN;
constexpr int ST_DONE = 4103;
constexpr int ST_PREFIX = 4104;
constexpr int ST_KREM = 4105;
constexpr int ST_LOSS_SUM = 4108;  // double
constexpr int ST_W_SUM = 4110;     // double
constexpr int ST_KEPT_ACC = 4112;

// bilinear source coordinate, align_corners=True, exactly ATen's float recipe:
// scale = float(in-1)/float(out-1) (0 if out==1); src = scale*dst; i0 = (int)src; lambda = src - i0
struct Lerp {
    int i0, i1;
    float l0, l1;
};
__device__ __forceinline__ Lerp make_lerp(float scale, int dst, int in_size) {
    Lerp r;
    float src = __fmul_rn(scale, (float)dst);
    r.i0 = (int)src;
    if (r.i0 > in_size - 1) r.i0 = in_size - 1;
    r.i1 = r.i0 + ((r.i0 < in_size - 1) ? 1 : 0);
    r.l1 = __fsub_rn(src, (float)r.i0);
    r.l0 = __fsub_rn(1.0f, r.l1);
    return r;
}
__device__ __forceinline__ float lerp4(const Lerp& ly, const Lerp& lx, float a, float b, float c, float d) {
    // no FMA contraction: identical to the oracle's separate multiply/add sequence
    float t0 = __fadd_rn(__fmul_rn(lx.l0, a), __fmul_rn(lx.l1, b));
    float t1 = __fadd_rn(__fmul_rn(lx.l0, c), __fmul_rn(lx.l1, d));
    return __fadd_rn(__fmul_rn(ly.l0, t0), __fmul_rn(ly.l1, t1));
}
__host__ __device__ __forceinline__ float area_scale(int in_size, int out_size) {
    return out_size > 1 ? (float)(in_size - 1) / (float)(out_size - 1) : 0.0f;
}

// p_target + nll from the C logits of one pixel held in registers (C <= CMAX).
template <int CMAX>
__device__ __forceinline__ void softmax_target(const float (&x)[CMAX], int C, int t, float& p_t, float& nll) {
    float m = x[0];
#pragma unroll
    for (int c = 1; c < CMAX; ++c)
        if (c < C) m = fmaxf(m, x[c]);
    float s = 0.f, et = 0.f, xt = 0.f;
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
        if (c < C) {
            float e = tsb_exp_det(__fsub_rn(x[c], m));
            s = __fadd_rn(s, e);
            if (c == t) { et = e; xt = x[c]; }
        }
    }
    p_t = __fdiv_rn(et, s);
    nll = logf(s) - (xt - m);
}

// shared tail of the two p_target kernels: writes p/nll, feeds histogram + counters
struct PtAccum {
    unsigned int num_valid = 0, count_le = 0;
};
__device__ __forceinline__ void pt_emit(long long i, bool valid, float p_t, float nll_v, float thresh, float* p,
                                        float* nll, unsigned int* s_hist, PtAccum& acc) {
    float pv = valid ? p_t : 1.0f;  // loss_opr.py:81 masked_fill_(~valid, 1) then gather of class 0
    p[i] = pv;
    nll[i] = valid ? nll_v : 0.f;
    acc.num_valid += valid ? 1u : 0u;
    acc.count_le += (pv <= thresh) ? 1u : 0u;
    unsigned int bin = __float_as_uint(pv) >> 20;
    atomicAdd(&s_hist[bin > 4095u ? 4095u : bin], 1u);
}
__device__ __forceinline__ void pt_flush(unsigned int* s_hist, const PtAccum& acc, uint32_t* state, float* s_red) {
    __syncthreads();
    for (int b = threadIdx.x; b < 4096; b += kThreads) {
        unsigned int v = s_hist[b];
        if (v) atomicAdd(&state[ST_HIST + b], v);
    }
    float nv = block_sum<kThreads>((float)acc.num_valid, s_red);
    float cl = block_sum<kThreads>((float)acc.count_le, s_red);
    if (threadIdx.x == 0) {
        // per-block partial counts are < 2^24, exact in float
        atomicAdd(&state[ST_NUM_VALID], (unsigned int)nv);
        atomicAdd(&state[ST_COUNT_LE], (unsigned int)cl);
    }
}

template <typename T, int CMAX>
__global__ void __launch_bounds__(kThreads)
ohem_ptarget_kernel(const T* __restrict__ logits, long long sn, long long sc, long long sy, long long sx,
                    const int64_t* __restrict__ labels, int N, int C, int H, int W, int ignore_label, float thresh,
                    float* __restrict__ p, float* __restrict__ nll, uint32_t* state) {
    __shared__ unsigned int s_hist[4096];
    __shared__ float s_red[33];
    for (int b = threadIdx.x; b < 4096; b += kThreads) s_hist[b] = 0;
    __syncthreads();
    PtAccum acc;
    const long long total = (long long)N * H * W;
    const long long hw = (long long)H * W;
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
        int n = (int)(i / hw);
        long long r = i - (long long)n * hw;
        int y = (int)(r / W), x = (int)(r - (long long)y * W);
        long long lab = labels[i];
        bool valid = lab != (long long)ignore_label;
        int t = valid ? (int)lab : 0;
        const T* base = logits + n * sn + y * sy + x * sx;
        float v[CMAX];
#pragma unroll
        for (int c = 0; c < CMAX; ++c)
            if (c < C) v[c] = ld_as_float<T>(base + c * sc);
        float p_t, nl;
        softmax_target<CMAX>(v, C, t, p_t, nl);
        pt_emit(i, valid, p_t, nl, thresh, p, nll, s_hist, acc);
    }
    pt_flush(s_hist, acc, state, s_red);
}

// generic-C variant (C > 32, e.g. ADE 150 classes): two passes over the pixel's logits
template <typename T>
__global__ void __launch_bounds__(kThreads)
ohem_ptarget_kernel_anyc(const T* __restrict__ logits, long long sn, long long sc, long long sy, long long sx,
                         const int64_t* __restrict__ labels, int N, int C, int H, int W, int ignore_label,
                         float thresh, float* __restrict__ p, float* __restrict__ nll, uint32_t* state) {
    __shared__ unsigned int s_hist[4096];
    __shared__ float s_red[33];
    for (int b = threadIdx.x; b < 4096; b += kThreads) s_hist[b] = 0;
    __syncthreads();
    PtAccum acc;
    const long long total = (long long)N * H * W;
    const long long hw = (long long)H * W;
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
        int n = (int)(i / hw);
        long long r = i - (long long)n * hw;
        int y = (int)(r / W), x = (int)(r - (long long)y * W);
        long long lab = labels[i];
        bool valid = lab != (long long)ignore_label;
        int t = valid ? (int)lab : 0;
        const T* base = logits + n * sn + y * sy + x * sx;
        float m = ld_as_float<T>(base);
        for (int c = 1; c < C; ++c) m = fmaxf(m, ld_as_float<T>(base + c * sc));
        float s = 0.f, et = 0.f, xt = 0.f;
        for (int c = 0; c < C; ++c) {
            float xv = ld_as_float<T>(base + c * sc);
            float e = tsb_exp_det(__fsub_rn(xv, m));
            s = __fadd_rn(s, e);
            if (c == t) { et = e; xt = xv; }
        }
        float p_t = __fdiv_rn(et, s);
        float nl = logf(s) - (xt - m);
        pt_emit(i, valid, p_t, nl, thresh, p, nll, s_hist, acc);
    }
    pt_flush(s_hist, acc, state, s_red);
}

// fused bilinear-upsample variant: low-res fp32 NHWC logits [N,h,w,cs]
template <int CMAX>
__global__ void __launch_bounds__(kThreads)
ohem_ptarget_up_kernel(const float* __restrict__ lo, int cs, int h, int w, const int64_t* __restrict__ labels, int N,
                       int C, int H, int W, int ignore_label, float thresh, float* __restrict__ p,
                       float* __restrict__ nll, uint32_t* state) {
    __shared__ unsigned int s_hist[4096];
    __shared__ float s_red[33];
    for (int b = threadIdx.x; b < 4096; b += kThreads) s_hist[b] = 0;
    __syncthreads();
    PtAccum acc;
    const float ry = area_scale(h, H), rx = area_scale(w, W);
    const long long total = (long long)N * H * W;
    const long long hw = (long long)H * W;
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
        int n = (int)(i / hw);
        long long r = i - (long long)n * hw;
        int y = (int)(r / W), x = (int)(r - (long long)y * W);
        long long lab = labels[i];
        bool valid = lab != (long long)ignore_label;
        int t = valid ? (int)lab : 0;
        Lerp ly = make_lerp(ry, y, h), lx = make_lerp(rx, x, w);
        const float* b00 = lo + (((long long)n * h + ly.i0) * w + lx.i0) * cs;
        const float* b01 = lo + (((long long)n * h + ly.i0) * w + lx.i1) * cs;
        const float* b10 = lo + (((long long)n * h + ly.i1) * w + lx.i0) * cs;
        const float* b11 = lo + (((long long)n * h + ly.i1) * w + lx.i1) * cs;
        float v[CMAX];
#pragma unroll
        for (int c = 0; c < CMAX; ++c)
            if (c < C) v[c] = lerp4(ly, lx, __ldg(b00 + c), __ldg(b01 + c), __ldg(b10 + c), __ldg(b11 + c));
        float p_t, nl;
        softmax_target<CMAX>(v, C, t, p_t, nl);
        pt_emit(i, valid, p_t, nl, thresh, p, nll, s_hist, acc);
    }
    pt_flush(s_hist, acc, state, s_red);
}

// ---------------------------------------------------------------------------------------------
// radix select: decide kernel (1 block of 1024 threads) + histogram kernels for levels 1 and 2
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) ohem_decide_kernel(uint32_t* state, long long n, long long min_kept,
                                                            float thresh, int level) {
    __shared__ unsigned long long s_warp[32];
    __shared__ int s_bin;
    __shared__ unsigned long long s_before;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (level == 0) {
        if (tid == 0) {
            unsigned int num_valid = state[ST_NUM_VALID];
            unsigned int count_le = state[ST_COUNT_LE];
            // loss_opr.py:78-85: skip when min_kept > num_valid; kept mask only if num_valid>0 and min_kept>0
            bool active = !((long long)min_kept > (long long)num_valid) && num_valid > 0 && min_kept > 0;
            long long k = min_kept < n ? min_kept : n;
            unsigned int done = 1;
            float T = thresh;
            if (active && (long long)count_le < k) done = 0;  // k-th smallest p exceeds thresh → need it
            state[ST_ACTIVE] = active ? 1u : 0u;
            state[ST_THRESH] = __float_as_uint(T);
            state[ST_DONE] = done;
            state[ST_PREFIX] = 0;
            state[ST_KREM] = (unsigned int)k;
        }
        __syncthreads();
    }
    const bool done = state[ST_DONE] != 0;
    if (!done) {
        const unsigned int krem = state[ST_KREM];
        // each thread owns 4 consecutive bins
        unsigned int h0 = state[ST_HIST + tid * 4 + 0], h1 = state[ST_HIST + tid * 4 + 1];
        unsigned int h2 = state[ST_HIST + tid * 4 + 2], h3 = state[ST_HIST + tid * 4 + 3];
        unsigned long long mine = (unsigned long long)h0 + h1 + h2 + h3;
        unsigned long long incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) s_warp[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            unsigned long long wv = s_warp[lane];
            unsigned long long wi = wv;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                unsigned long long v = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += v;
            }
            s_warp[lane] = wi - wv;  // exclusive prefix of warp totals
        }
        __syncthreads();
        unsigned long long excl = s_warp[wid] + incl - mine;  // elements strictly before my 4 bins
        if (excl < krem && excl + mine >= krem) {
            unsigned long long c = excl;
            int b = tid * 4;
            if (c + h0 >= krem) { b += 0; }
            else { c += h0; if (c + h1 >= krem) { b += 1; } else { c += h1; if (c + h2 >= krem) { b += 2; } else { c += h2; b += 3; } } }
            s_bin = b;
            s_before = c;
        }
        __syncthreads();
        if (tid == 0) {
            unsigned int prefix = state[ST_PREFIX];
            int b = s_bin;
            unsigned int newk = krem - (unsigned int)s_before;
            if (level < 2) {
                state[ST_PREFIX] = (prefix << 12) | (unsigned int)b;
                state[ST_KREM] = newk;
            } else {
                unsigned int bits = (prefix << 8) | (unsigned int)b;
                state[ST_THRESH] = bits;  // T = k-th smallest p (> thresh, loss_opr.py:88-89)
                state[ST_DONE] = 1;
            }
        }
    }
    __syncthreads();
    // leave a clean histogram for the next level / next use
    for (int b = tid; b < 4096; b += 1024) state[ST_HIST + b] = 0;
}

__global__ void __launch_bounds__(kThreads) ohem_hist_kernel(const float* __restrict__ p, long long n, uint32_t* state,
                                                              int level) {
    if (state[ST_DONE] != 0) return;
    __shared__ unsigned int s_hist[4096];
    for (int b = threadIdx.x; b < 4096; b += kThreads) s_hist[b] = 0;
    __syncthreads();
    const unsigned int prefix = state[ST_PREFIX];
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kThreads) {
        unsigned int u = __float_as_uint(p[i]);
        if (level == 1) {
            if ((u >> 20) == prefix) atomicAdd(&s_hist[(u >> 8) & 0xfffu], 1u);
        } else {
            if ((u >> 8) == prefix) atomicAdd(&s_hist[u & 0xffu], 1u);
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < 4096; b += kThreads) {
        unsigned int v = s_hist[b];
        if (v) atomicAdd(&state[ST_HIST + b], v);
    }
}

// ---------------------------------------------------------------------------------------------
// loss over kept pixels
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) ohem_loss_kernel(const float* __restrict__ p, const float* __restrict__ nll,
                                                              const int64_t* __restrict__ labels, long long n,
                                                              int ignore_label, const float* __restrict__ cw,
                                                              uint32_t* state) {
    __shared__ float s_red[33];
    const bool active = state[ST_ACTIVE] != 0;
    const float T = __uint_as_float(state[ST_THRESH]);
    float ls = 0.f, ws = 0.f, kc = 0.f;
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kThreads) {
        long long lab = labels[i];
        bool kept = (lab != (long long)ignore_label) && (!active || p[i] <= T);
        if (kept) {
            float wt = cw ? __ldg(cw + (int)lab) : 1.f;
            ls += wt * nll[i];
            ws += wt;
            kc += 1.f;
        }
    }
    ls = block_sum<kThreads>(ls, s_red);
    ws = block_sum<kThreads>(ws, s_red);
    kc = block_sum<kThreads>(kc, s_red);
    if (threadIdx.x == 0) {
        atomicAdd(reinterpret_cast<double*>(state + ST_LOSS_SUM), (double)ls);
        atomicAdd(reinterpret_cast<double*>(state + ST_W_SUM), (double)ws);
        atomicAdd(&state[ST_KEPT_ACC], (unsigned int)kc);
    }
}
__global__ void ohem_finalize_kernel(uint32_t* state, float* loss_out) {
    double ls = *reinterpret_cast<double*>(state + ST_LOSS_SUM);
    double ws = *reinterpret_cast<double*>(state + ST_W_SUM);
    float loss = (float)(ls / ws);  // 0/0 → NaN, as CrossEntropyLoss over zero pixels
    state[ST_KEPT] = state[ST_KEPT_ACC];
    state[ST_LOSS] = __float_as_uint(loss);
    state[ST_INVDEN] = __float_as_uint((float)(1.0 / ws));
    if (loss_out) *loss_out = loss;
}

// ---------------------------------------------------------------------------------------------
// gradient, materialised logits
// ---------------------------------------------------------------------------------------------
template <typename T, int CMAX>
__global__ void __launch_bounds__(kThreads)
ohem_grad_kernel(const T* __restrict__ logits, long long sn, long long sc, long long sy, long long sx,
                 const int64_t* __restrict__ labels, const float* __restrict__ p, int N, int C, int H, int W,
                 int ignore_label, const float* __restrict__ cw, const uint32_t* __restrict__ state,
                 const float* __restrict__ gscale, T* __restrict__ dlogits) {
    const bool active = state[ST_ACTIVE] != 0;
    const float Tth = __uint_as_float(state[ST_THRESH]);
    const float scale = __uint_as_float(state[ST_INVDEN]) * (gscale ? *gscale : 1.f);
    const long long total = (long long)N * H * W;
    const long long hw = (long long)H * W;
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
        int n = (int)(i / hw);
        long long r = i - (long long)n * hw;
        int y = (int)(r / W), x = (int)(r - (long long)y * W);
        long long lab = labels[i];
        bool kept = (lab != (long long)ignore_label) && (!active || p[i] <= Tth);
        long long off = n * sn + y * sy + x * sx;
        if (!kept) {
            for (int c = 0; c < C; ++c) st_from_float<T>(dlogits + off + c * sc, 0.f);
            continue;
        }
        int t = (int)lab;
        float wt = (cw ? __ldg(cw + t) : 1.f) * scale;
        if (CMAX > 0) {
            float v[CMAX > 0 ? CMAX : 1];
            float m = -INFINITY;
#pragma unroll
            for (int c = 0; c < CMAX; ++c)
                if (c < C) { v[c] = ld_as_float<T>(logits + off + c * sc); m = fmaxf(m, v[c]); }
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < CMAX; ++c)
                if (c < C) { v[c] = __expf(v[c] - m); s += v[c]; }
            float inv = 1.f / s;
#pragma unroll
            for (int c = 0; c < CMAX; ++c)
                if (c < C) st_from_float<T>(dlogits + off + c * sc, (v[c] * inv - (c == t ? 1.f : 0.f)) * wt);
        } else {
            float m = -INFINITY;
            for (int c = 0; c < C; ++c) m = fmaxf(m, ld_as_float<T>(logits + off + c * sc));
            float s = 0.f;
            for (int c = 0; c < C; ++c) s += __expf(ld_as_float<T>(logits + off + c * sc) - m);
            float inv = 1.f / s;
            for (int c = 0; c < C; ++c) {
                float e = __expf(ld_as_float<T>(logits + off + c * sc) - m);
                st_from_float<T>(dlogits + off + c * sc, (e * inv - (c == t ? 1.f : 0.f)) * wt);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// gradient, fused-upsample form.  One warp per low-res cell (n,i,j): the hi-res pixels whose bilinear
// stencil has (i,j) as top-left corner contribute only to the cell's 4 corners, so the warp keeps
// 4 x C register accumulators, shuffle-reduces them and issues 4*C red.global.add.f32.
// ---------------------------------------------------------------------------------------------
template <int CMAX>
__global__ void __launch_bounds__(kThreads)
ohem_grad_up_kernel(const float* __restrict__ lo, int cs, int h, int w, const int64_t* __restrict__ labels,
                    const float* __restrict__ p, int N, int C, int H, int W, int ignore_label,
                    const float* __restrict__ cw, const uint32_t* __restrict__ state,
                    const float* __restrict__ gscale, float* __restrict__ dlo) {
    __shared__ float s_corner[kThreads / 32][4][CMAX];
    __shared__ float s_out[kThreads / 32][4][CMAX];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const bool active = state[ST_ACTIVE] != 0;
    const float Tth = __uint_as_float(state[ST_THRESH]);
    const float scale = __uint_as_float(state[ST_INVDEN]) * (gscale ? *gscale : 1.f);
    const float ry = area_scale(h, H), rx = area_scale(w, W);
    const long long ncells = (long long)N * h * w;
    for (long long cell = (long long)blockIdx.x * (kThreads / 32) + wib; cell < ncells;
         cell += (long long)gridDim.x * (kThreads / 32)) {
        int n = (int)(cell / ((long long)h * w));
        int rem = (int)(cell - (long long)n * h * w);
        int ci = rem / w, cj = rem - ci * w;
        int i1 = ci + (ci < h - 1 ? 1 : 0), j1 = cj + (cj < w - 1 ? 1 : 0);
        const long long o00 = (((long long)n * h + ci) * w + cj) * cs, o01 = (((long long)n * h + ci) * w + j1) * cs;
        const long long o10 = (((long long)n * h + i1) * w + cj) * cs, o11 = (((long long)n * h + i1) * w + j1) * cs;
        __syncwarp();
        if (lane < C) {
            s_corner[wib][0][lane] = __ldg(lo + o00 + lane);
            s_corner[wib][1][lane] = __ldg(lo + o01 + lane);
            s_corner[wib][2][lane] = __ldg(lo + o10 + lane);
            s_corner[wib][3][lane] = __ldg(lo + o11 + lane);
        }
        __syncwarp();
        // candidate hi-res box (one pixel of slack each side; membership is re-checked exactly)
        int y_lo, y_hi, x_lo, x_hi;
        if (ry > 0.f) {
            y_lo = max(0, (int)floorf((float)ci / ry) - 1);
            y_hi = min(H - 1, (int)ceilf((float)(ci + 1) / ry) + 1);
        } else { y_lo = 0; y_hi = H - 1; }
        if (rx > 0.f) {
            x_lo = max(0, (int)floorf((float)cj / rx) - 1);
            x_hi = min(W - 1, (int)ceilf((float)(cj + 1) / rx) + 1);
        } else { x_lo = 0; x_hi = W - 1; }
        const int bw = x_hi - x_lo + 1, bh = y_hi - y_lo + 1;
        float acc[4][CMAX];
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int c = 0; c < CMAX; ++c) acc[k][c] = 0.f;
        for (int q = lane; q < bw * bh; q += 32) {
            int yy = y_lo + q / bw, xx = x_lo + q % bw;
            Lerp ly = make_lerp(ry, yy, h), lx = make_lerp(rx, xx, w);
            if (ly.i0 != ci || lx.i0 != cj) continue;
            long long pi = ((long long)n * H + yy) * W + xx;
            long long lab = labels[pi];
            bool kept = (lab != (long long)ignore_label) && (!active || p[pi] <= Tth);
            if (!kept) continue;
            int t = (int)lab;
            float wt = (cw ? __ldg(cw + t) : 1.f) * scale;
            float v[CMAX];
            float m = -INFINITY;
#pragma unroll
            for (int c = 0; c < CMAX; ++c)
                if (c < C) {
                    v[c] = lerp4(ly, lx, s_corner[wib][0][c], s_corner[wib][1][c], s_corner[wib][2][c],
                                 s_corner[wib][3][c]);
                    m = fmaxf(m, v[c]);
                }
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < CMAX; ++c)
                if (c < C) { v[c] = __expf(v[c] - m); s += v[c]; }
            float inv = wt / s;
            const float w00 = ly.l0 * lx.l0, w01 = ly.l0 * lx.l1, w10 = ly.l1 * lx.l0, w11 = ly.l1 * lx.l1;
#pragma unroll
            for (int c = 0; c < CMAX; ++c)
                if (c < C) {
                    float g = v[c] * inv - (c == t ? wt : 0.f);
                    acc[0][c] += w00 * g; acc[1][c] += w01 * g; acc[2][c] += w10 * g; acc[3][c] += w11 * g;
                }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int c = 0; c < CMAX; ++c)
                if (c < C) {
                    float r = warp_sum(acc[k][c]);
                    if (lane == 0) s_out[wib][k][c] = r;
                }
        __syncwarp();
        if (lane < C) {
            float a0 = s_out[wib][0][lane], a1 = s_out[wib][1][lane], a2 = s_out[wib][2][lane], a3 = s_out[wib][3][lane];
            if (a0 != 0.f) atomicAdd(dlo + o00 + lane, a0);
            if (a1 != 0.f) atomicAdd(dlo + o01 + lane, a1);
            if (a2 != 0.f) atomicAdd(dlo + o10 + lane, a2);
            if (a3 != 0.f) atomicAdd(dlo + o11 + lane, a3);
        }
    }
}

}  // namespace

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" int tsb_ohem_begin(uint32_t* state, tsb_stream_t stream) {
    TSB_REQUIRE(state != nullptr, "tsb_ohem_begin: null state");
    TSB_CUDA_CALL(cudaMemsetAsync(state, 0, sizeof(uint32_t) * TSB_OHEM_STATE_WORDS, (cudaStream_t)stream));
    return TSB_OK;
}

extern "C" int tsb_ohem_ptarget(const void* logits, int dtype, long long sn, long long sc, long long sy, long long sx,
                                const int64_t* labels, int N, int C, int H, int W, int ignore_label, float thresh,
                                float* p, float* nll, uint32_t* state, tsb_stream_t stream) {
    TSB_REQUIRE(logits && labels && p && nll && state, "tsb_ohem_ptarget: null pointer");
    TSB_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0, "tsb_ohem_ptarget: bad shape");
    TSB_REQUIRE(dtype == TSB_F32 || dtype == TSB_BF16, "tsb_ohem_ptarget: bad dtype");
    long long total = (long long)N * H * W;
    int grid = tsb_grid_for(total, kThreads, 8);
    cudaStream_t st = (cudaStream_t)stream;
    if (C <= 32) {
        if (dtype == TSB_F32)
            ohem_ptarget_kernel<float, 32><<<grid, kThreads, 0, st>>>((const float*)logits, sn, sc, sy, sx, labels, N, C, H, W, ignore_label, thresh, p, nll, state);
        else
            ohem_ptarget_kernel<__nv_bfloat16, 32><<<grid, kThreads, 0, st>>>((const __nv_bfloat16*)logits, sn, sc, sy, sx, labels, N, C, H, W, ignore_label, thresh, p, nll, state);
    } else {
        if (dtype == TSB_F32)
            ohem_ptarget_kernel_anyc<float><<<grid, kThreads, 0, st>>>((const float*)logits, sn, sc, sy, sx, labels, N, C, H, W, ignore_label, thresh, p, nll, state);
        else
            ohem_ptarget_kernel_anyc<__nv_bfloat16><<<grid, kThreads, 0, st>>>((const __nv_bfloat16*)logits, sn, sc, sy, sx, labels, N, C, H, W, ignore_label, thresh, p, nll, state);
    }
    TSB_CUDA_CHECK_LAUNCH("ohem_ptarget");
    return TSB_OK;
}

extern "C" int tsb_ohem_ptarget_up(const float* logits_lo, int cs, int h, int w, const int64_t* labels, int N, int C,
                                   int H, int W, int ignore_label, float thresh, float* p, float* nll, uint32_t* state,
                                   tsb_stream_t stream) {
    TSB_REQUIRE(logits_lo && labels && p && nll && state, "tsb_ohem_ptarget_up: null pointer");
    TSB_REQUIRE(N > 0 && C > 0 && C <= 32 && cs >= C && H > 0 && W > 0 && h > 0 && w > 0,
                "tsb_ohem_ptarget_up: bad shape (C must be <= 32)");
    long long total = (long long)N * H * W;
    int grid = tsb_grid_for(total, kThreads, 8);
    ohem_ptarget_up_kernel<32><<<grid, kThreads, 0, (cudaStream_t)stream>>>(logits_lo, cs, h, w, labels, N, C, H, W, ignore_label, thresh, p, nll, state);
    TSB_CUDA_CHECK_LAUNCH("ohem_ptarget_up");
    return TSB_OK;
}

extern "C" int tsb_ohem_select(const float* p, long long n, long long min_kept, float thresh, uint32_t* state,
                               tsb_stream_t stream) {
    TSB_REQUIRE(p && state && n > 0, "tsb_ohem_select: bad args");
    cudaStream_t st = (cudaStream_t)stream;
    int grid = tsb_grid_for(n, kThreads, 8);
    ohem_decide_kernel<<<1, 1024, 0, st>>>(state, n, min_kept, thresh, 0);
    TSB_CUDA_CHECK_LAUNCH("ohem_decide0");
    ohem_hist_kernel<<<grid, kThreads, 0, st>>>(p, n, state, 1);
    TSB_CUDA_CHECK_LAUNCH("ohem_hist1");
    ohem_decide_kernel<<<1, 1024, 0, st>>>(state, n, min_kept, thresh, 1);
    TSB_CUDA_CHECK_LAUNCH("ohem_decide1");
    ohem_hist_kernel<<<grid, kThreads, 0, st>>>(p, n, state, 2);
    TSB_CUDA_CHECK_LAUNCH("ohem_hist2");
    ohem_decide_kernel<<<1, 1024, 0, st>>>(state, n, min_kept, thresh, 2);
    TSB_CUDA_CHECK_LAUNCH("ohem_decide2");
    return TSB_OK;
}

extern "C" int tsb_ohem_loss(const float* p, const float* nll, const int64_t* labels, long long n, int ignore_label,
                             const float* class_weight, uint32_t* state, float* loss_out, tsb_stream_t stream) {
    TSB_REQUIRE(p && nll && labels && state && n > 0, "tsb_ohem_loss: bad args");
    cudaStream_t st = (cudaStream_t)stream;
    int grid = tsb_grid_for(n, kThreads, 8);
    ohem_loss_kernel<<<grid, kThreads, 0, st>>>(p, nll, labels, n, ignore_label, class_weight, state);
    TSB_CUDA_CHECK_LAUNCH("ohem_loss");
    ohem_finalize_kernel<<<1, 1, 0, st>>>(state, loss_out);
    TSB_CUDA_CHECK_LAUNCH("ohem_finalize");
    return TSB_OK;
}

extern "C" int tsb_ohem_grad(const void* logits, int dtype, long long sn, long long sc, long long sy, long long sx,
                             const int64_t* labels, const float* p, int N, int C, int H, int W, int ignore_label,
                             const float* class_weight, const uint32_t* state, const float* gscale, void* dlogits,
                             tsb_stream_t stream) {
    TSB_REQUIRE(logits && labels && p && state && dlogits, "tsb_ohem_grad: null pointer");
    TSB_REQUIRE(dtype == TSB_F32 || dtype == TSB_BF16, "tsb_ohem_grad: bad dtype");
    long long total = (long long)N * H * W;
    int grid = tsb_grid_for(total, kThreads, 8);
    cudaStream_t st = (cudaStream_t)stream;
#define TSB_LAUNCH_GRAD(T, CM)                                                                                   \
    ohem_grad_kernel<T, CM><<<grid, kThreads, 0, st>>>((const T*)logits, sn, sc, sy, sx, labels, p, N, C, H, W,  \
                                                       ignore_label, class_weight, state, gscale, (T*)dlogits)
    if (C <= 32) {
        if (dtype == TSB_F32) TSB_LAUNCH_GRAD(float, 32); else TSB_LAUNCH_GRAD(__nv_bfloat16, 32);
    } else {
        if (dtype == TSB_F32) TSB_LAUNCH_GRAD(float, 0); else TSB_LAUNCH_GRAD(__nv_bfloat16, 0);
    }
#undef TSB_LAUNCH_GRAD
    TSB_CUDA_CHECK_LAUNCH("ohem_grad");
    return TSB_OK;
}

extern "C" int tsb_ohem_grad_up(const float* logits_lo, int cs, int h, int w, const int64_t* labels, const float* p,
                                int N, int C, int H, int W, int ignore_label, const float* class_weight,
                                const uint32_t* state, const float* gscale, float* dlogits_lo, tsb_stream_t stream) {
    TSB_REQUIRE(logits_lo && labels && p && state && dlogits_lo, "tsb_ohem_grad_up: null pointer");
    TSB_REQUIRE(C > 0 && C <= 32 && cs >= C, "tsb_ohem_grad_up: C must be <= 32");
    long long ncells = (long long)N * h * w;
    int grid = tsb_grid_for(ncells, kThreads / 32, 8);
    ohem_grad_up_kernel<32><<<grid, kThreads, 0, (cudaStream_t)stream>>>(logits_lo, cs, h, w, labels, p, N, C, H, W, ignore_label, class_weight, state, gscale, dlogits_lo);
    TSB_CUDA_CHECK_LAUNCH("ohem_grad_up");
    return TSB_OK;
}
