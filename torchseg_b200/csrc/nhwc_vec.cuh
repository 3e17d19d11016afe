// 8-channel (16-byte for bf16) vector access helpers for NHWC tensors with a channel stride.
#pragma once
#include "tsb_common.cuh"

#ifdef __CUDACC__
template <typename T> struct Vec8;

template <> struct Vec8<__nv_bfloat16> {
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float f[8]) {
        uint4 u = *reinterpret_cast<const uint4*>(p);
        unpack8(u, f);
    }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, const float f[8]) {
        *reinterpret_cast<uint4*>(p) = pack8(f);
    }
};
template <> struct Vec8<float> {
    static __device__ __forceinline__ void load(const float* p, float f[8]) {
        float4 a = *reinterpret_cast<const float4*>(p);
        float4 b = *reinterpret_cast<const float4*>(p + 4);
        f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
        f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    }
    static __device__ __forceinline__ void store(float* p, const float f[8]) {
        *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
        *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
    }
};

__device__ __forceinline__ void ldg8f(const float* p, float f[8]) {
    float4 a = __ldg(reinterpret_cast<const float4*>(p));
    float4 b = __ldg(reinterpret_cast<const float4*>(p + 4));
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
    f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// Grid-stride walk over (pixel, 8-channel group) pairs: i = pixel * C8 + c8 advances by `stride` per trip. The naive
// `i % C8`, `i / C8` costs a 64-bit integer division per 16-byte vector (~100 issue slots — half the issue bandwidth of
// a streaming kernel, ncu: 50 % issue-active on bn_apply); the walker divides once per thread.
struct VecWalk {
    long long p, dp;
    int c8, dc, C8;
    __device__ __forceinline__ VecWalk(int C8_, long long start, long long stride) : C8(C8_) {
        p = start / C8_;
        c8 = (int)(start - p * C8_);
        dp = stride / C8_;
        dc = (int)(stride - dp * C8_);
    }
    __device__ __forceinline__ void next() {
        p += dp;
        c8 += dc;
        if (c8 >= C8) { c8 -= C8; ++p; }
    }
};
#endif

