// One LPIPS feature tap in one pass (spi/criteria/lpips/lpips.py:50-71 with utils.normalize_activation):
//     xn = x / (sqrt(sum_c x^2) + 1e-10);   tap = sum_n mean_{h,w} sum_c lin_c (xn_c - yn_c)^2
// x: raw (post-ReLU) VGG features of the generated image, channels-last fp32 [N, H, W, C]; yn: the target's features, ALREADY
// unit-normalised (constant: cached or computed under no_grad), batch N or 1; lin: the 1x1 "lin" layer weights [C].
// Replaces ~12 ATen passes per tap and direction (pow, sum, sqrt, add, div, sub, pow, 1x1 conv, mean + their adjoints) with one
// streaming read of x and yn (forward) and one read + one write (backward).  A warp owns a pixel: lanes stride over the C/4
// channel groups with 128-bit loads, two shuffle reductions per pixel.
#include "common.cuh"

namespace {

constexpr int LT_MAXG = 4;      // float4 groups per lane: C <= 512

template <int NG, bool BWD>
__global__ void __launch_bounds__(256) lpips_tap_kernel(const float* __restrict__ x, const float* __restrict__ yn, const float* __restrict__ lin,
                                                        int n, int hw, int C, int ny, float inv_hw, const float* __restrict__ gout,
                                                        float* __restrict__ out, float* __restrict__ dx, const float* __restrict__ sample_w) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int groups = C >> 2;
    float4 wl[NG];
#pragma unroll
    for (int g = 0; g < NG; g++) {
        const int gi = lane + 32 * g;
        wl[g] = gi < groups ? *reinterpret_cast<const float4*>(lin + 4 * gi) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float go0 = BWD ? gout[0] * inv_hw : 0.f;
    const long long pixels = (long long)n * hw;
    float part = 0.f;
    for (long long p = (long long)blockIdx.x * 8 + warp; p < pixels; p += (long long)gridDim.x * 8) {
        const float sw = sample_w ? __ldg(sample_w + p / hw) : 1.f;           // per-sample weight of the sum over n (NULL: 1)
        const float go = go0 * sw;
        const float4* xr = reinterpret_cast<const float4*>(x + p * C);
        const long long py = ny == n ? p : p % hw;
        const float4* yr = reinterpret_cast<const float4*>(yn + py * C);
        float4 xv[NG], yv[NG];
        float ss = 0.f;
#pragma unroll
        for (int g = 0; g < NG; g++) {
            const int gi = lane + 32 * g;
            const bool in = gi < groups;
            xv[g] = in ? ldg_stream(xr + gi) : make_float4(0.f, 0.f, 0.f, 0.f);
            yv[g] = in ? __ldg(yr + gi) : make_float4(0.f, 0.f, 0.f, 0.f);
            ss += xv[g].x * xv[g].x + xv[g].y * xv[g].y + xv[g].z * xv[g].z + xv[g].w * xv[g].w;
        }
        ss = warp_sum(ss);
        const float nrm = sqrtf(ss);
        const float inv = 1.f / (nrm + 1e-10f);
        float d = 0.f, s = 0.f;
        float4 q[NG];
#pragma unroll
        for (int g = 0; g < NG; g++) {
            const float ex = xv[g].x * inv - yv[g].x, ey = xv[g].y * inv - yv[g].y, ez = xv[g].z * inv - yv[g].z, ew = xv[g].w * inv - yv[g].w;
            d += wl[g].x * ex * ex + wl[g].y * ey * ey + wl[g].z * ez * ez + wl[g].w * ew * ew;
            if (BWD) {
                q[g] = make_float4(2.f * wl[g].x * ex, 2.f * wl[g].y * ey, 2.f * wl[g].z * ez, 2.f * wl[g].w * ew);
                s += q[g].x * xv[g].x + q[g].y * xv[g].y + q[g].z * xv[g].z + q[g].w * xv[g].w;
            }
        }
        if (!BWD) part += d * sw;
        else {
            s = warp_sum(s);
            // d tap / d x_k = q_k / (n + eps) - (sum_c q_c x_c) x_k / (n (n + eps)^2); an all-zero feature vector gets a zero gradient
            const float k1 = go * inv, k2 = nrm > 0.f ? go * s * inv * inv / nrm : 0.f;
            float4* dr = reinterpret_cast<float4*>(dx + p * C);
#pragma unroll
            for (int g = 0; g < NG; g++) {
                const int gi = lane + 32 * g;
                if (gi < groups)
                    dr[gi] = make_float4(k1 * q[g].x - k2 * xv[g].x, k1 * q[g].y - k2 * xv[g].y, k1 * q[g].z - k2 * xv[g].z, k1 * q[g].w - k2 * xv[g].w);
            }
        }
    }
    if (!BWD) {
        part = warp_sum(part);
        __shared__ float red[8];
        if (lane == 0) red[warp] = part;
        __syncthreads();
        if (threadIdx.x == 0) {
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < 8; w++) t += red[w];
            atomicAdd(out, t * inv_hw);
        }
    }
}

template <bool BWD>
int launch_tap(const float* x, const float* yn, const float* lin, int n, int hw, int c, int ny, const float* gout, float* out, float* dx,
               const float* sample_w, cudaStream_t stream) {
    const long long pixels = (long long)n * hw;
    long long want = (pixels + 7) / 8, cap = (long long)spi_num_sms() * 8;
    const int grid = (int)(want < cap ? want : cap);
    const float inv_hw = 1.f / (float)hw;
    const int ng = (c + 127) / 128;
    if (ng <= 1) lpips_tap_kernel<1, BWD><<<grid, 256, 0, stream>>>(x, yn, lin, n, hw, c, ny, inv_hw, gout, out, dx, sample_w);
    else if (ng <= 2) lpips_tap_kernel<2, BWD><<<grid, 256, 0, stream>>>(x, yn, lin, n, hw, c, ny, inv_hw, gout, out, dx, sample_w);
    else lpips_tap_kernel<4, BWD><<<grid, 256, 0, stream>>>(x, yn, lin, n, hw, c, ny, inv_hw, gout, out, dx, sample_w);
    SPI_COUNT_LAUNCH(1);
    SPI_LAUNCH_CHECK("lpips_tap");
    return SPI_OK;
}

int check_tap(const float* x, const float* yn, const float* lin, int n, int hw, int c, int ny) {
    SPI_CHECK_ARG(x && yn && lin && n >= 1 && hw >= 1, "lpips_tap: null pointer / empty tensor");
    SPI_CHECK_ARG(c >= 4 && c % 4 == 0 && c <= 128 * LT_MAXG, "lpips_tap: C must be a multiple of 4, <= 512");
    SPI_CHECK_ARG(ny == n || ny == 1, "lpips_tap: target batch must be 1 or match");
    SPI_CHECK_ARG((((uintptr_t)x | (uintptr_t)yn | (uintptr_t)lin) & 15) == 0, "lpips_tap: tensors must be 16-byte aligned");
    return SPI_OK;
}

}  // namespace

/* out[0] += sum_n sw_n mean_hw sum_c lin_c (x_c / (|x| + 1e-10) - yn_c)^2   (out is accumulated: the caller zeroes it once per LPIPS call;
 * sample_weight [n] or NULL = all ones: lets several (image, target) pairs with different loss weights share one pass of the trunk) */
extern "C" int spi_lpips_tap_forward(const float* x, const float* yn, const float* lin, int n, int hw, int c, int ny, float* out,
                                     const float* sample_weight, cudaStream_t stream) {
    if (int rc = check_tap(x, yn, lin, n, hw, c, ny)) return rc;
    SPI_CHECK_ARG(out, "lpips_tap_forward: null output");
    return launch_tap<false>(x, yn, lin, n, hw, c, ny, nullptr, out, nullptr, sample_weight, stream);
}

/* dx = gout[0] * d tap / d x   (written) */
extern "C" int spi_lpips_tap_backward(const float* x, const float* yn, const float* lin, int n, int hw, int c, int ny, const float* gout,
                                      float* dx, const float* sample_weight, cudaStream_t stream) {
    if (int rc = check_tap(x, yn, lin, n, hw, c, ny)) return rc;
    SPI_CHECK_ARG(gout && dx && (((uintptr_t)dx) & 15) == 0, "lpips_tap_backward: null / misaligned pointer");
    return launch_tap<true>(x, yn, lin, n, hw, c, ny, gout, nullptr, dx, sample_weight, stream);
}
