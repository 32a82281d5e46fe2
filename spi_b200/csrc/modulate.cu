// Weight modulation / demodulation for StyleGAN2 modulated convolutions (sm_100a).
//
// Replaces the ~8 ATen launches per layer (and ~20 in backward) of the fused branch of `modulated_conv2d`
// (eg3d/training/networks_stylegan2.py:58-68):
//     w'[n,o,i,k] = W[o,i,k] * s[n,i];   d[n,o] = rsqrt(sum_{i,k} w'^2 + 1e-8);   w'' = w' * d      (demodulate)
// Forward: one CTA per (o, n) row of I*KK weights (warp-shuffle + smem reduction), writing w'' directly in the layout
// the conv engine consumes.  Backward (given g = dL/dw''):
//     A[n,o]   = sum_{i,k} g * W * s
//     t        = d * (g - d^2 * A * W * s)          (t = g when demodulate is off, with d = 1)
//     dW[o,i,k] = sum_n s[n,i] * t;    ds[n,i] = sum_{o,k} W[o,i,k] * t      (ds accumulated with atomics over o)
// Tensors are tiny (<= 4 x 2.4 M floats); the point is launch count and fusing the reductions, not bandwidth.
#include "common.cuh"

namespace {

// memory layout of the per-sample weight tensor (logical [n][o][i][k]):
//   0: O I K (contiguous OIHW)   1: O K I (channels-last OHWI, what an NHWC conv consumes)   2: I K O (channels-last of the
//   transposed [I,O,kh,kw] weight a conv_transpose2d consumes)
__device__ __forceinline__ size_t widx(int layout, int n, int o, int i, int k, int O, int I, int KK) {
    if (layout == 0) return (((size_t)n * O + o) * I + i) * KK + k;
    if (layout == 1) return (((size_t)n * O + o) * KK + k) * I + i;
    return (((size_t)n * I + i) * KK + k) * O + o;
}

__device__ __forceinline__ float block_sum(float v, float* sh) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    float t = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.f;
    if (warp == 0) t = warp_sum(t);
    if (threadIdx.x == 0) sh[0] = t;
    __syncthreads();
    return sh[0];
}

// grid (O, N), block 256: demodulation coefficients d[n,o] = rsqrt(sum_{i,k} (W s)^2 + 1e-8)
__global__ void __launch_bounds__(256) demod_coef_kernel(const float* __restrict__ W, const float* __restrict__ s, float* __restrict__ dcoef,
                                                         int O, int I, int KK) {
    __shared__ float sh[32];
    const int o = blockIdx.x, n = blockIdx.y;
    const int len = I * KK;
    const float* w = W + (size_t)o * len;
    const float* sn = s + (size_t)n * I;
    float acc = 0.f;
    for (int e = threadIdx.x; e < len; e += blockDim.x) { float v = w[e] * sn[e / KK]; acc += v * v; }
    float tot = block_sum(acc, sh);
    if (threadIdx.x == 0) dcoef[(size_t)n * O + o] = rsqrtf(tot + 1e-8f);
}

// elementwise in OUTPUT memory order (coalesced stores for every layout): out = W * s * d
__global__ void __launch_bounds__(256) modulate_apply_kernel(const float* __restrict__ W, const float* __restrict__ s,
                                                             const float* __restrict__ dcoef, float* __restrict__ out, int N, int O,
                                                             int I, int KK, int layout_flags) {
    const int layout = layout_flags & 3;
    const bool flip = (layout_flags & 4) != 0;      // taps written in reverse order: out[.., k] = W[.., KK-1-k] (= w.flip([3, 4]))
    const long long total = (long long)N * O * I * KK;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
        int n, o, i, k;
        long long r = q;
        if (layout == 0) { k = (int)(r % KK); r /= KK; i = (int)(r % I); r /= I; o = (int)(r % O); n = (int)(r / O); }
        else if (layout == 1) { i = (int)(r % I); r /= I; k = (int)(r % KK); r /= KK; o = (int)(r % O); n = (int)(r / O); }
        else { o = (int)(r % O); r /= O; k = (int)(r % KK); r /= KK; i = (int)(r % I); n = (int)(r / I); }
        float v = W[((size_t)o * I + i) * KK + (flip ? KK - 1 - k : k)] * s[(size_t)n * I + i];
        if (dcoef) v *= dcoef[(size_t)n * O + o];
        out[q] = v;
    }
}

// grid (O), block 256: loops over n; dW written (no atomics), ds accumulated with atomics across o
__global__ void __launch_bounds__(256) modulate_bwd_kernel(const float* __restrict__ W, const float* __restrict__ s, const float* __restrict__ dcoef,
                                                           const float* __restrict__ g, float* __restrict__ dW, float* __restrict__ ds,
                                                           int N, int O, int I, int KK, int demod, int layout_flags) {
    __shared__ float sh[32];
    const int layout = layout_flags & 3;
    const bool flip = (layout_flags & 4) != 0;      // g holds the gradient of the tap-reversed weights
    const int o = blockIdx.x;
    const int len = I * KK;
    const float* w = W + (size_t)o * len;
    float* dw = dW ? dW + (size_t)o * len : nullptr;
    for (int n = 0; n < N; n++) {
        const float* sn = s + (size_t)n * I;
        float d = 1.f, A = 0.f;
        if (demod) {
            d = dcoef[(size_t)n * O + o];
            float acc = 0.f;
            for (int q = threadIdx.x; q < len; q += blockDim.x) {      // q walks g's memory order for layouts 0 and 1
                const int i = (layout == 1) ? q % I : q / KK, k = (layout == 1) ? q / I : q % KK;
                acc += g[widx(layout, n, o, i, k, O, I, KK)] * w[i * KK + (flip ? KK - 1 - k : k)] * sn[i];
            }
            A = block_sum(acc, sh);
        }
        const float dA = d * d * A;
        for (int i = threadIdx.x; i < I; i += blockDim.x) {       // thread owns channel i: one atomic per (n, i) per CTA
            const float si = sn[i];
            float dsi = 0.f;
            for (int k = 0; k < KK; k++) {
                const int e = i * KK + k;
                const float ws = w[e] * si;
                const float ge = g[widx(layout, n, o, i, flip ? KK - 1 - k : k, O, I, KK)];
                const float t = demod ? d * (ge - dA * ws) : ge;
                if (dw) dw[e] = (n == 0 ? 0.f : dw[e]) + si * t;
                dsi += w[e] * t;
            }
            if (ds) atomicAdd(ds + (size_t)n * I + i, dsi);
        }
    }
}

}  // namespace

extern "C" int spi_modulate_weights(const float* weight, const float* styles, float* out, float* dcoef, int n, int o, int i, int kk,
                                    int demodulate, int layout, cudaStream_t stream) {
    SPI_CHECK_ARG(layout >= 0 && (layout & 3) <= 2 && layout < 8, "modulate_weights: layout must be 0 (OIK), 1 (OKI) or 2 (IKO), +4 to reverse the taps");
    SPI_CHECK_ARG(weight && styles && out, "modulate_weights: null pointer");
    SPI_CHECK_ARG(n >= 1 && o >= 1 && i >= 1 && kk >= 1 && n <= 65535, "modulate_weights: bad shape");
    SPI_CHECK_ARG(!demodulate || dcoef, "modulate_weights: dcoef buffer required when demodulating");
    if (demodulate) demod_coef_kernel<<<dim3(o, n), 256, 0, stream>>>(weight, styles, dcoef, o, i, kk);
    long long total = (long long)n * o * i * kk;
    long long blocks = (total + 255) / 256, cap = (long long)spi_num_sms() * 8;
    modulate_apply_kernel<<<(int)(blocks > cap ? cap : blocks), 256, 0, stream>>>(weight, styles, demodulate ? dcoef : nullptr, out, n, o, i, kk, layout);
    SPI_COUNT_LAUNCH(demodulate ? 2 : 1);
    SPI_LAUNCH_CHECK("modulate_weights");
    return SPI_OK;
}

extern "C" int spi_modulate_weights_backward(const float* weight, const float* styles, const float* dcoef, const float* grad_out,
                                             float* grad_weight, float* grad_styles, int n, int o, int i, int kk, int demodulate,
                                             int layout, cudaStream_t stream) {
    SPI_CHECK_ARG(weight && styles && grad_out, "modulate_weights_backward: null pointer");
    SPI_CHECK_ARG(!demodulate || dcoef, "modulate_weights_backward: dcoef required when demodulating");
    if (grad_styles) cudaMemsetAsync(grad_styles, 0, sizeof(float) * (size_t)n * i, stream);
    modulate_bwd_kernel<<<o, 256, 0, stream>>>(weight, styles, dcoef, grad_out, grad_weight, grad_styles, n, o, i, kk, demodulate, layout);
    SPI_COUNT_LAUNCH(1);
    SPI_LAUNCH_CHECK("modulate_weights_backward");
    return SPI_OK;
}
