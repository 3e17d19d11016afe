// Training-time input pipeline on the device: random mirror → random scale (cv2.resize INTER_LINEAR for the
// image, INTER_NEAREST for the label) → normalize → random crop + centred constant pad → HWC→CHW, fused into ONE pass
// from the decoded uint8 image to the fp32 NCHW batch and the int64 label batch the train loop consumes.
// Replaces TrainPre.__call__ (/root/reference/model/bisenet/cityscapes.bisenet.R18/dataloader.py:16-33) and the
// img_utils functions it calls (/root/reference/furnace/utils/img_utils.py:24-78,118-125,140-145,181-187); the random draws
// stay on the host (torchseg_b200/utils/gpu_pipeline.py) in the reference's order.
//
// Bit-exact by construction: OpenCV's 8-bit INTER_LINEAR is FIXED POINT (11-bit coefficients, int32 horizontal pass,
// (((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2 vertically — imgproc/src/resize.cpp); the coefficient arithmetic below
// uses explicit IEEE double / float operations (no FMA contraction), and normalisation is a 3 x 256 table built on the host
// with the reference's own float32 / float64 sequence. HBM-bound byte work: one thread per output pixel, coalesced
// plane-wise stores (the 6 MB uint8 source of a Cityscapes frame replaces 20 MB of fp32 image + int64 label on the H2D path).
#include "tsb_common.cuh"

namespace {

struct PreSample {           // mirror of the 10 x int64 descriptor row (include/tsb.h)
    long long img, gt, H, W, flip, sh, sw, pos_h, pos_w, reserved;
};

struct Coef {
    int i0, i1, w0, w1;
};
// resize.cpp: fx = (float)((dx + 0.5) * scale - 0.5); sx = cvFloor(fx); fx -= sx; [x only: clamp at the borders with fx = 0]
__device__ __forceinline__ Coef linear_coef(int d, int src, int dst, bool border_zero) {
    const double scale = __ddiv_rn(1.0, __ddiv_rn((double)dst, (double)src));
    float f = (float)__dadd_rn(__dmul_rn(__dadd_rn((double)d, 0.5), scale), -0.5);
    int s = (int)floorf(f);
    f = __fsub_rn(f, (float)s);
    if (border_zero) {
        if (s < 0) { f = 0.f; s = 0; }
        if (s >= src - 1) { f = 0.f; s = src - 1; }
    }
    Coef c;
    c.w0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
    c.w1 = __float2int_rn(__fmul_rn(f, 2048.f));
    c.i0 = min(max(s, 0), src - 1);
    c.i1 = min(max(s + 1, 0), src - 1);
    return c;
}
__device__ __forceinline__ int nearest_index(int d, int src, int dst) {   // x_ofs = min(cvFloor(x * ifx), size - 1)
    const double scale = __ddiv_rn(1.0, __ddiv_rn((double)dst, (double)src));
    const int s = (int)floor(__dmul_rn((double)d, scale));
    return min(s, src - 1);
}

__global__ void __launch_bounds__(256)
train_preprocess_kernel(const PreSample* __restrict__ samples, int crop_h, int crop_w, int reverse_channels,
                        const float* __restrict__ lut, float img_pad, int gt_pad, float* __restrict__ out_img,
                        long long* __restrict__ out_gt) {
    const int n = blockIdx.z;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= crop_w) return;
    const PreSample S = samples[n];
    const int H = (int)S.H, W = (int)S.W, sh = (int)S.sh, sw = (int)S.sw;
    // random_crop_pad_to_shape: the crop [pos, pos + crop) ∩ image, centred in the target, constant border
    const int c_h = min(crop_h, sh - (int)S.pos_h), c_w = min(crop_w, sw - (int)S.pos_w);
    const int top = (crop_h - c_h) / 2, left = (crop_w - c_w) / 2;
    const int yy = y - top, xx = x - left;
    const size_t plane = (size_t)crop_h * crop_w;
    float* o = out_img + (size_t)n * 3 * plane + (size_t)y * crop_w + x;
    long long* og = out_gt + (size_t)n * plane + (size_t)y * crop_w + x;
    if (yy < 0 || yy >= c_h || xx < 0 || xx >= c_w) {
        o[0] = img_pad; o[plane] = img_pad; o[2 * plane] = img_pad;
        *og = (long long)gt_pad;
        return;
    }
    const int ys = (int)S.pos_h + yy, xs = (int)S.pos_w + xx;      // coordinates in the (mirrored, scaled) image
    const bool flip = S.flip != 0;
    // label: INTER_NEAREST
    {
        const int gy = nearest_index(ys, H, sh), gx = nearest_index(xs, W, sw);
        const uint8_t* g = reinterpret_cast<const uint8_t*>(S.gt);
        *og = (long long)g[(size_t)gy * W + (flip ? W - 1 - gx : gx)];
    }
    // image: INTER_LINEAR, fixed point
    const Coef cx = linear_coef(xs, W, sw, true), cy = linear_coef(ys, H, sh, false);
    const int xa = flip ? W - 1 - cx.i0 : cx.i0, xb = flip ? W - 1 - cx.i1 : cx.i1;
    const uint8_t* p = reinterpret_cast<const uint8_t*>(S.img);
    const uint8_t* r0a = p + ((size_t)cy.i0 * W + xa) * 3;
    const uint8_t* r0b = p + ((size_t)cy.i0 * W + xb) * 3;
    const uint8_t* r1a = p + ((size_t)cy.i1 * W + xa) * 3;
    const uint8_t* r1b = p + ((size_t)cy.i1 * W + xb) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int cs = reverse_channels ? 2 - c : c;
        const int s0 = (int)r0a[cs] * cx.w0 + (int)r0b[cs] * cx.w1;
        const int s1 = (int)r1a[cs] * cx.w0 + (int)r1b[cs] * cx.w1;
        int v = (((cy.w0 * (s0 >> 4)) >> 16) + ((cy.w1 * (s1 >> 4)) >> 16) + 2) >> 2;
        v = min(max(v, 0), 255);
        o[(size_t)c * plane] = __ldg(lut + c * 256 + v);
    }
}

}  // namespace

extern "C" int tsb_train_preprocess(const long long* samples_dev, int n, int crop_h, int crop_w, int reverse_channels,
                                    const float* lut, float img_pad, int gt_pad, float* out_img, long long* out_gt,
                                    tsb_stream_t stream) {
    TSB_REQUIRE(samples_dev && lut && out_img && out_gt, "tsb_train_preprocess: null pointer");
    TSB_REQUIRE(n > 0 && n <= 65535 && crop_h > 0 && crop_h <= 65535 && crop_w > 0, "tsb_train_preprocess: bad sizes");
    dim3 grid((crop_w + 255) / 256, crop_h, n);
    train_preprocess_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const PreSample*>(samples_dev), crop_h, crop_w,
                                                                    reverse_channels, lut, img_pad, gt_pad, out_img, out_gt);
    TSB_CUDA_CHECK_LAUNCH("train_preprocess");
    return TSB_OK;
}
