// 1x1 convolution onto a handful of output channels (the RGB heads: ToRGBLayer of the super-resolution blocks, Cout = 3,
// eg3d/training/networks_stylegan2.py:503-518 via superresolution.py:279-290) -- forward, data gradient and weight gradient.
// With 3 output channels the contraction has an arithmetic intensity of 1.5 FLOP/byte: it is a streaming op over x (134 MB for
// 128 channels at 512^2), not tensor-core work, so these are plain coalesced kernels: a warp owns one pixel at a time, a lane owns
// 4 consecutive input channels (float4) of every 128-channel slab, the per-pixel dot products are finished with shuffles.
//   forward   y[n,p,o]  = sum_i x[n,p,i] * w[g,o,i]
//   dgrad     dx[n,p,i] = sum_o dy[n,p,o] * w[g,o,i]
//   wgrad     dw[g,o,i] = sum_{n in g, p} dy[n,p,o] * x[n,p,i]
// x [N][P][Ci] channels-last fp32 (Ci a multiple of 128, at most 512), y / dy [N][P][Co] (Co <= 4), w [G][Co][Ci], G in {1, N}.
#include "common.cuh"

namespace {

constexpr int MAXCO = 4;
constexpr int MAXSLAB = 4;        // Ci <= 512

template <int CO, int SLABS>
__global__ void __launch_bounds__(256) rgb_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ y, long long pixels_per_img,
                                                      int n, int per_sample) {
    const int ci = SLABS * 128;
    const int lane = threadIdx.x & 31;
    const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const long long total = pixels_per_img * n;
    int cur_g = -1;
    float4 wr[CO][SLABS];
    for (long long p = warp; p < total; p += nwarps) {
        const int g = per_sample ? (int)(p / pixels_per_img) : 0;
        if (g != cur_g) {
            cur_g = g;
#pragma unroll
            for (int o = 0; o < CO; o++)
#pragma unroll
                for (int s = 0; s < SLABS; s++) wr[o][s] = __ldg(reinterpret_cast<const float4*>(w + ((size_t)g * CO + o) * ci + s * 128) + lane);
        }
        const float4* xp = reinterpret_cast<const float4*>(x + (size_t)p * ci);
        float acc[CO];
#pragma unroll
        for (int o = 0; o < CO; o++) acc[o] = 0.f;
#pragma unroll
        for (int s = 0; s < SLABS; s++) {
            const float4 v = ldg_stream(xp + s * 32 + lane);
#pragma unroll
            for (int o = 0; o < CO; o++) acc[o] += v.x * wr[o][s].x + v.y * wr[o][s].y + v.z * wr[o][s].z + v.w * wr[o][s].w;
        }
#pragma unroll
        for (int o = 0; o < CO; o++) acc[o] = warp_sum(acc[o]);
        if (lane < CO) {
            float r = acc[0];
#pragma unroll
            for (int o = 1; o < CO; o++) r = (lane == o) ? acc[o] : r;
            y[(size_t)p * CO + lane] = r;
        }
    }
}

template <int CO, int SLABS>
__global__ void __launch_bounds__(256) rgb_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w, float* __restrict__ dx,
                                                        long long pixels_per_img, int n, int per_sample) {
    const int ci = SLABS * 128;
    const int lane = threadIdx.x & 31;
    const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const long long total = pixels_per_img * n;
    int cur_g = -1;
    float4 wr[CO][SLABS];
    for (long long p = warp; p < total; p += nwarps) {
        const int g = per_sample ? (int)(p / pixels_per_img) : 0;
        if (g != cur_g) {
            cur_g = g;
#pragma unroll
            for (int o = 0; o < CO; o++)
#pragma unroll
                for (int s = 0; s < SLABS; s++) wr[o][s] = __ldg(reinterpret_cast<const float4*>(w + ((size_t)g * CO + o) * ci + s * 128) + lane);
        }
        float d[CO];
#pragma unroll
        for (int o = 0; o < CO; o++) d[o] = __ldg(dy + (size_t)p * CO + o);
        float4* xp = reinterpret_cast<float4*>(dx + (size_t)p * ci);
#pragma unroll
        for (int s = 0; s < SLABS; s++) {
            float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int o = 0; o < CO; o++) {
                r.x += d[o] * wr[o][s].x; r.y += d[o] * wr[o][s].y; r.z += d[o] * wr[o][s].z; r.w += d[o] * wr[o][s].w;
            }
            stg_stream(xp + s * 32 + lane, r);
        }
    }
}

// one block per (group, slice of its pixels): per-lane partial sums over the block's pixels, combined across the block's warps in
// shared memory, one atomicAdd per weight and block
template <int CO, int SLABS>
__global__ void __launch_bounds__(128) rgb_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw,
                                                        long long pixels_per_group, int blocks_per_group) {
    const int ci = SLABS * 128;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int g = blockIdx.x / blocks_per_group, b = blockIdx.x % blocks_per_group;
    const long long per_block = (pixels_per_group + blocks_per_group - 1) / blocks_per_group;
    const long long p0 = (long long)b * per_block, p1 = min(pixels_per_group, p0 + per_block);
    const float* xg = x + (size_t)g * pixels_per_group * ci;
    const float* dg = dy + (size_t)g * pixels_per_group * CO;
    float4 acc[CO][SLABS];
#pragma unroll
    for (int o = 0; o < CO; o++)
#pragma unroll
        for (int s = 0; s < SLABS; s++) acc[o][s] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (long long p = p0 + wid; p < p1; p += nw) {
        float d[CO];
#pragma unroll
        for (int o = 0; o < CO; o++) d[o] = __ldg(dg + (size_t)p * CO + o);
        const float4* xp = reinterpret_cast<const float4*>(xg + (size_t)p * ci);
#pragma unroll
        for (int s = 0; s < SLABS; s++) {
            const float4 v = ldg_stream(xp + s * 32 + lane);
#pragma unroll
            for (int o = 0; o < CO; o++) {
                acc[o][s].x += d[o] * v.x; acc[o][s].y += d[o] * v.y; acc[o][s].z += d[o] * v.z; acc[o][s].w += d[o] * v.w;
            }
        }
    }
    __shared__ float red[4][CO * SLABS * 128];
#pragma unroll
    for (int o = 0; o < CO; o++)
#pragma unroll
        for (int s = 0; s < SLABS; s++) *reinterpret_cast<float4*>(&red[wid][(o * SLABS + s) * 128 + lane * 4]) = acc[o][s];
    __syncthreads();
    for (int e = threadIdx.x; e < CO * ci; e += blockDim.x) {
        float sum = 0.f;
        for (int k = 0; k < nw; k++) sum += red[k][e];
        atomicAdd(dw + (size_t)g * CO * ci + e, sum);        // e = o * ci + (s * 128 + c): the [Co][Ci] order of dw
    }
}

template <int CO>
int dispatch(int which, const float* a, const float* b, float* c, long long pixels, int n, int ci, int per_sample, cudaStream_t stream) {
    const int slabs = ci / 128;
    const int sms = spi_num_sms();
    const int grid = sms * 8;
    const int groups = per_sample ? n : 1;
    const long long ppg = pixels * (n / groups);
    const int bpg = max(1, (sms * 8) / groups);
#define RGB_CASE(S)                                                                                                               \
    case S:                                                                                                                       \
        if (which == 0) rgb_fwd_kernel<CO, S><<<grid, 256, 0, stream>>>(a, b, c, pixels, n, per_sample);                           \
        else if (which == 1) rgb_dgrad_kernel<CO, S><<<grid, 256, 0, stream>>>(a, b, c, pixels, n, per_sample);                    \
        else rgb_wgrad_kernel<CO, S><<<groups * bpg, 128, 0, stream>>>(a, b, c, ppg, bpg);                                         \
        break;
    switch (slabs) {
        RGB_CASE(1) RGB_CASE(2) RGB_CASE(3) RGB_CASE(4)
        default: return SPI_ERR_ARG;
    }
#undef RGB_CASE
    return SPI_OK;
}

}  // namespace

extern "C" int spi_conv1x1_rgb_supported(int ci, int co) { return (ci % 128 == 0 && ci >= 128 && ci <= 512 && co >= 1 && co <= MAXCO) ? 1 : 0; }

// which: 0 forward (a = x, b = w, c = y), 1 data gradient (a = dy, b = w, c = dx), 2 weight gradient (a = x, b = dy, c = dw, overwritten;
// 2 + 4: dw is already zero on entry, the fill is skipped)
extern "C" int spi_conv1x1_rgb(int which, const float* a, const float* b, float* c, long long pixels, int n, int ci, int co, int per_sample,
                               cudaStream_t stream) {
    const bool prezeroed = (which & 4) != 0;
    which &= 3;
    SPI_CHECK_ARG(a && b && c && which >= 0 && which <= 2, "spi_conv1x1_rgb: bad arguments");
    SPI_CHECK_ARG(spi_conv1x1_rgb_supported(ci, co), "spi_conv1x1_rgb: unsupported shape ci=%d co=%d", ci, co);
    SPI_CHECK_ARG((((uintptr_t)a | (uintptr_t)b | (uintptr_t)c) & 15) == 0 || co != 4, "spi_conv1x1_rgb: tensors must be 16-byte aligned");
    if (which == 2 && !prezeroed) cudaMemsetAsync(c, 0, (size_t)(per_sample ? n : 1) * co * ci * 4, stream);
    int rc;
    switch (co) {
        case 1: rc = dispatch<1>(which, a, b, c, pixels, n, ci, per_sample, stream); break;
        case 2: rc = dispatch<2>(which, a, b, c, pixels, n, ci, per_sample, stream); break;
        case 3: rc = dispatch<3>(which, a, b, c, pixels, n, ci, per_sample, stream); break;
        default: rc = dispatch<4>(which, a, b, c, pixels, n, ci, per_sample, stream); break;
    }
    if (rc != SPI_OK) { spi_set_error("spi_conv1x1_rgb: unsupported channel count %d", ci); return rc; }
    SPI_COUNT_LAUNCH(1);
    SPI_LAUNCH_CHECK("spi_conv1x1_rgb");
    return SPI_OK;
}
