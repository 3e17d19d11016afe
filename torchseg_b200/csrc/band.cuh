// Bilinear (align_corners=True) interpolation coefficients and the BAND decomposition shared by the fused
// upsample + loss kernels (ohem.cu, ce_up.cu): CTA = (strip of hi-res columns, low-res row i, image n); all hi-res rows y with
// floor(ry*y) == i share the two low-res source rows (i, i1).
#pragma once
#include "tsb_common.cuh"

namespace band {

// bilinear source coordinate, align_corners=True, exactly ATen's float recipe:
// scale = float(in-1)/float(out-1) (0 if out==1); src = scale*dst; i0 = (int)src; lambda = src - i0
struct Lerp {
    int i0, i1;
    float l0, l1;
};
static __device__ __forceinline__ Lerp make_lerp(float scale, int dst, int in_size) {
    Lerp r;
    float src = __fmul_rn(scale, (float)dst);
    r.i0 = (int)src;
    if (r.i0 > in_size - 1) r.i0 = in_size - 1;
    r.i1 = r.i0 + ((r.i0 < in_size - 1) ? 1 : 0);
    r.l1 = __fsub_rn(src, (float)r.i0);
    r.l0 = __fsub_rn(1.0f, r.l1);
    return r;
}
// 4-tap bilinear blend in the oracle's operation order (no FMA contraction: separate multiply / add roundings),
// split at its own intermediate values: t = horizontal blend of one source row (depends on the column only, so a band
// kernel forms it ONCE per thread), v = vertical blend per hi-res row.
static __device__ __forceinline__ float lerp_h(const Lerp& lx, float a, float b) {
    return __fadd_rn(__fmul_rn(lx.l0, a), __fmul_rn(lx.l1, b));
}
static __device__ __forceinline__ float lerp_v(const Lerp& ly, float t0, float t1) {
    return __fadd_rn(__fmul_rn(ly.l0, t0), __fmul_rn(ly.l1, t1));
}
static __host__ __device__ __forceinline__ float area_scale(int in_size, int out_size) {
    return out_size > 1 ? (float)(in_size - 1) / (float)(out_size - 1) : 0.0f;
}

struct BandGeom {
    int y_lo, y_hi;   // candidate hi-res rows (membership re-checked exactly)
    int jbase, ncols; // low-res columns [jbase, jbase+ncols) touched by this strip
};
static __device__ __forceinline__ BandGeom band_geom(float ry, float rx, int ci, int x0, int x1 /*exclusive*/, int H, int w) {
    BandGeom g;
    if (ry > 0.f) {
        g.y_lo = max(0, (int)floorf((float)ci / ry) - 1);
        g.y_hi = min(H - 1, (int)ceilf((float)(ci + 1) / ry) + 1);
    } else { g.y_lo = 0; g.y_hi = H - 1; }
    Lerp a = make_lerp(rx, x0, w), b = make_lerp(rx, x1 - 1, w);
    g.jbase = a.i0;
    g.ncols = b.i1 - a.i0 + 1;
    return g;
}


}  // namespace band
