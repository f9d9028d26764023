// Tap generation shared by the blur kernels (reference pytorch/gaussianblur.py:9-14 and
// gaussian_blur.py:83-88), in the reference's own fp32 expression order, normalised to sum 1.
#pragma once

#include "common.cuh"

namespace bgan {

constexpr float kSqrt2Pi = 2.5066282746310002f;

__device__ __forceinline__ float raw_tap(int rule, float s, int i) {
    const float x = static_cast<float>(i);
    if (rule == BGAN_RULE_TORCH) {
        // g = exp(-(x^2 / (2 s^2)) / (sqrt(2 pi) s)) -- the normaliser sits inside the exp
        const float num = (x * x) / (2.0f * (s * s));
        const float den = kSqrt2Pi * s;
        return expf(-(num / den));
    }
    // g = exp(-(x^2 / (2 s^2))) / (sqrt(2 pi) s)
    return expf(-((x * x) / (2.0f * (s * s)))) / (kSqrt2Pi * s);
}

// Fills s_g[0 .. 2r] with normalised taps; s_red is one double of scratch.  All threads call it.
__device__ __forceinline__ void normalised_taps(int rule, float s, int r, const float* __restrict__ taps_dev, float* s_g, double* s_red) {
    const int n = 2 * r + 1;
    for (int t = threadIdx.x; t < n; t += blockDim.x)
        s_g[t] = (rule == BGAN_RULE_EXPLICIT) ? taps_dev[t] : raw_tap(rule, s, t - r);
    __syncthreads();
    if (rule != BGAN_RULE_EXPLICIT) {
        if (threadIdx.x < 32) {
            double acc = 0.0;
            for (int t = threadIdx.x; t < n; t += 32) acc += static_cast<double>(s_g[t]);
            acc = warp_sum(acc);
            if (threadIdx.x == 0) *s_red = acc;
        }
        __syncthreads();
        const float total = static_cast<float>(*s_red);
        for (int t = threadIdx.x; t < n; t += blockDim.x) s_g[t] = s_g[t] / total;
        __syncthreads();
    }
}


}  // namespace bgan
