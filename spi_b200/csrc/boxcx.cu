// Kernels of the mirror-view contextual loss (spi/criteria/bbox_cx_loss.py):
//   * roi_align 80x80 crops of the landmark boxes (:41-59, torchvision.ops.roi_align semantics: aligned = False, sampling_ratio = -1)
//   * the row-wise chain on the [B, M, N] cosine-similarity matrix (:103-131, 175-178):
//       d = 1 - S;  dmin_i = min_j d;  dt = clamp(d / (dmin_i + 1e-5), -10, 10);  w = exp((1 - dt) / h);  cx = w / sum_j w;
//       colmax_j = max_i cx[i, j]                                  (then mean_j, -log: [B, N] work left to the caller)
//     as ONE pass over S forward (the reference's ~10 ATen passes over [4,1600,1600]) and one pass backward.
#include "common.cuh"

namespace {

// ---------------------------------------------------------------------------------------------------------------- roi_align
struct RoiParams {
    const float* in; const float* rois; float* out;        // in [N,C,H,W] (strides s*), rois [K,5] = (batch, x1, y1, x2, y2), out [K,C,P,P]
    int c, h, w, k, p;
    long long sn, sc, sh, sw;
};

__device__ __forceinline__ void bilinear_setup(float y, float x, int h, int w, int& yl, int& xl, int& yh, int& xh, float& w1, float& w2, float& w3,
                                               float& w4, bool& valid) {
    valid = !(y < -1.f || y > (float)h || x < -1.f || x > (float)w);
    if (y <= 0.f) y = 0.f;
    if (x <= 0.f) x = 0.f;
    yl = (int)y; xl = (int)x;
    if (yl >= h - 1) { yh = yl = h - 1; y = (float)yl; } else yh = yl + 1;
    if (xl >= w - 1) { xh = xl = w - 1; x = (float)xl; } else xh = xl + 1;
    const float ly = y - yl, lx = x - xl, hy = 1.f - ly, hx = 1.f - lx;
    w1 = hy * hx; w2 = hy * lx; w3 = ly * hx; w4 = ly * lx;
}

// one thread per output element; BWD: scatters g[out] * weight / count into gin with atomics
template <bool BWD>
__global__ void roi_align_kernel(RoiParams q, const float* __restrict__ gout, float* __restrict__ gin) {
    const long long total = (long long)q.k * q.c * q.p * q.p;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int pw = (int)(idx % q.p), ph = (int)((idx / q.p) % q.p), c = (int)((idx / ((long long)q.p * q.p)) % q.c), k = (int)(idx / ((long long)q.p * q.p * q.c));
        const float* r = q.rois + (size_t)k * 5;
        const int b = (int)r[0];
        const float sw0 = r[1], sh0 = r[2];
        const float rw = fmaxf(r[3] - sw0, 1.f), rh = fmaxf(r[4] - sh0, 1.f);
        const float bh = rh / q.p, bw = rw / q.p;
        const int gh = (int)ceilf(rh / q.p), gw = (int)ceilf(rw / q.p);
        const float count = fmaxf((float)(gh * gw), 1.f);
        const long long base = (long long)b * q.sn + (long long)c * q.sc;
        float acc = 0.f;
        const float g = BWD ? gout[idx] / count : 0.f;
        for (int iy = 0; iy < gh; iy++) {
            const float y = sh0 + ph * bh + (iy + .5f) * bh / gh;
            for (int ix = 0; ix < gw; ix++) {
                const float x = sw0 + pw * bw + (ix + .5f) * bw / gw;
                int yl, xl, yh, xh; float w1, w2, w3, w4; bool valid;
                bilinear_setup(y, x, q.h, q.w, yl, xl, yh, xh, w1, w2, w3, w4, valid);
                if (!valid) continue;
                if (BWD) {
                    atomicAdd(gin + base + yl * q.sh + xl * q.sw, g * w1);
                    atomicAdd(gin + base + yl * q.sh + xh * q.sw, g * w2);
                    atomicAdd(gin + base + yh * q.sh + xl * q.sw, g * w3);
                    atomicAdd(gin + base + yh * q.sh + xh * q.sw, g * w4);
                } else {
                    acc += w1 * q.in[base + yl * q.sh + xl * q.sw] + w2 * q.in[base + yl * q.sh + xh * q.sw] + w3 * q.in[base + yh * q.sh + xl * q.sw] +
                           w4 * q.in[base + yh * q.sh + xh * q.sw];
                }
            }
        }
        if (!BWD) q.out[idx] = acc / count;
    }
}

// ---------------------------------------------------------------------------------------------------------------- CX rows
constexpr int CX_THREADS = 256;
constexpr int CX_MAXN = 2048;          // columns per row held in registers: CX_MAXN / CX_THREADS = 8 per thread

__device__ __forceinline__ float block_reduce(float v, float* red, bool is_min) {
    v = is_min ? warp_min(v) : warp_sum(v);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    float r = red[0];
    for (int k = 1; k < CX_THREADS / 32; k++) r = is_min ? fminf(r, red[k]) : r + red[k];
    return r;
}

// forward: one block per row (b, i).  stats[b, i] = {dmin, rowsum}; colmax[b, j] = max_i cx (atomicMax on the bits: cx > 0)
__global__ void __launch_bounds__(CX_THREADS) cx_rows_fwd_kernel(const float* __restrict__ S, int m, int n, float inv_h, float* __restrict__ stats,
                                                                 unsigned int* __restrict__ colmax) {
    __shared__ float red[CX_THREADS / 32];
    const long long row = blockIdx.x;                 // b * m + i
    const int b = (int)(row / m);
    const float* s = S + row * n;
    float d[CX_MAXN / CX_THREADS];
    float dmin = 3.0e38f;
#pragma unroll
    for (int k = 0; k < CX_MAXN / CX_THREADS; k++) {
        const int j = threadIdx.x + k * CX_THREADS;
        d[k] = j < n ? 1.f - __ldg(s + j) : 3.0e38f;
        dmin = fminf(dmin, d[k]);
    }
    dmin = block_reduce(dmin, red, true);
    const float inv = 1.f / (dmin + 1e-5f);
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < CX_MAXN / CX_THREADS; k++) {
        const int j = threadIdx.x + k * CX_THREADS;
        float w = 0.f;
        if (j < n) {
            const float dt = fminf(fmaxf(d[k] * inv, -10.f), 10.f);
            w = expf((1.f - dt) * inv_h);
        }
        d[k] = w;
        sum += w;
    }
    sum = block_reduce(sum, red, false);
    if (threadIdx.x == 0) { stats[row * 2] = dmin; stats[row * 2 + 1] = sum; }
#pragma unroll
    for (int k = 0; k < CX_MAXN / CX_THREADS; k++) {
        const int j = threadIdx.x + k * CX_THREADS;
        if (j < n) atomicMax(colmax + (size_t)b * n + j, __float_as_uint(d[k] / sum));
    }
}

// backward: one block per row.  gcol[b, j] = dL/d colmax[b, j].  The row that holds the column maximum (recomputed bit-identically)
// receives it; dS[b, i, :] follows through the normalisation, the exponential, the clamp, the relative distance and its row minimum.
__global__ void __launch_bounds__(CX_THREADS) cx_rows_bwd_kernel(const float* __restrict__ S, int m, int n, float inv_h, const float* __restrict__ stats,
                                                                 const unsigned int* __restrict__ colmax, const float* __restrict__ gcol, float* __restrict__ dS) {
    __shared__ float red[CX_THREADS / 32];
    __shared__ int jmin_s;
    const long long row = blockIdx.x;
    const int b = (int)(row / m);
    const float* s = S + row * n;
    const float dmin = stats[row * 2], sum = stats[row * 2 + 1];
    const float inv = 1.f / (dmin + 1e-5f);
    float d[CX_MAXN / CX_THREADS], cx[CX_MAXN / CX_THREADS], g[CX_MAXN / CX_THREADS];
    float inner = 0.f;
    if (threadIdx.x == 0) jmin_s = 0x7fffffff;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < CX_MAXN / CX_THREADS; k++) {
        const int j = threadIdx.x + k * CX_THREADS;
        d[k] = 0.f; cx[k] = 0.f; g[k] = 0.f;
        if (j < n) {
            d[k] = 1.f - __ldg(s + j);
            const float dt = fminf(fmaxf(d[k] * inv, -10.f), 10.f);
            cx[k] = expf((1.f - dt) * inv_h) / sum;
            if (__float_as_uint(cx[k]) == colmax[(size_t)b * n + j]) g[k] = gcol[(size_t)b * n + j];
            inner += g[k] * cx[k];
            if (d[k] == dmin) atomicMin(&jmin_s, j);          // first index of the row minimum (torch.min's gradient goes to one index)
        }
    }
    inner = block_reduce(inner, red, false);
    // d cx / d w: dw = (g - inner) / sum ; w = exp((1 - dt) / h): ddt = -w / h * dw = -cx * (g - inner) / h
    float ddmin = 0.f;
#pragma unroll
    for (int k = 0; k < CX_MAXN / CX_THREADS; k++) {
        const int j = threadIdx.x + k * CX_THREADS;
        float dd = 0.f;
        if (j < n) {
            const float ddt = -cx[k] * (g[k] - inner) * inv_h;
            const float raw = d[k] * inv;
            if (raw >= -10.f && raw <= 10.f) {              // clamp passes the gradient inside [min, max] (torch.clamp)
                dd = ddt * inv;
                ddmin -= ddt * d[k] * inv * inv;
            }
        }
        g[k] = dd;
    }
    ddmin = block_reduce(ddmin, red, false);
    const int jmin = jmin_s;
#pragma unroll
    for (int k = 0; k < CX_MAXN / CX_THREADS; k++) {
        const int j = threadIdx.x + k * CX_THREADS;
        if (j < n) dS[row * n + j] = -(g[k] + (j == jmin ? ddmin : 0.f));          // d = 1 - S
    }
}

}  // namespace

extern "C" int spi_roi_align(const float* in, const float* rois, float* out, int n, int c, int h, int w, const long long* strides, int k, int pooled,
                             cudaStream_t stream) {
    SPI_CHECK_ARG(in && rois && out && k > 0 && pooled > 0 && n > 0, "spi_roi_align: bad arguments");
    RoiParams q{in, rois, out, c, h, w, k, pooled, strides[0], strides[1], strides[2], strides[3]};
    const long long total = (long long)k * c * pooled * pooled;
    roi_align_kernel<false><<<cdiv(total, 256), 256, 0, stream>>>(q, nullptr, nullptr);
    SPI_COUNT_LAUNCH(1);
    SPI_LAUNCH_CHECK("spi_roi_align");
    return SPI_OK;
}

// gin [N,C,H,W] with the given strides is ACCUMULATED into (the caller zeroes it)
extern "C" int spi_roi_align_backward(const float* gout, const float* rois, float* gin, int n, int c, int h, int w, const long long* strides, int k,
                                      int pooled, cudaStream_t stream) {
    SPI_CHECK_ARG(gout && rois && gin && k > 0 && pooled > 0 && n > 0, "spi_roi_align_backward: bad arguments");
    RoiParams q{nullptr, rois, nullptr, c, h, w, k, pooled, strides[0], strides[1], strides[2], strides[3]};
    const long long total = (long long)k * c * pooled * pooled;
    roi_align_kernel<true><<<cdiv(total, 256), 256, 0, stream>>>(q, gout, gin);
    SPI_COUNT_LAUNCH(1);
    SPI_LAUNCH_CHECK("spi_roi_align_backward");
    return SPI_OK;
}

// S [B, M, N] cosine similarities -> stats [B, M, 2] (row minimum of 1 - S, row sum of w), colmax [B, N] = max_i cx (zeroed here)
extern "C" int spi_cx_rows_forward(const float* S, int b, int m, int n, float band_width, float* stats, float* colmax, cudaStream_t stream) {
    SPI_CHECK_ARG(S && stats && colmax && b > 0 && m > 0 && n > 0 && n <= CX_MAXN, "spi_cx_rows_forward: bad arguments (n must be <= %d)", CX_MAXN);
    cudaMemsetAsync(colmax, 0, (size_t)b * n * 4, stream);
    cx_rows_fwd_kernel<<<b * m, CX_THREADS, 0, stream>>>(S, m, n, 1.f / band_width, stats, reinterpret_cast<unsigned int*>(colmax));
    SPI_COUNT_LAUNCH(1);
    SPI_LAUNCH_CHECK("spi_cx_rows_forward");
    return SPI_OK;
}

extern "C" int spi_cx_rows_backward(const float* S, int b, int m, int n, float band_width, const float* stats, const float* colmax, const float* gcol,
                                    float* dS, cudaStream_t stream) {
    SPI_CHECK_ARG(S && stats && colmax && gcol && dS && b > 0 && m > 0 && n > 0 && n <= CX_MAXN, "spi_cx_rows_backward: bad arguments");
    cx_rows_bwd_kernel<<<b * m, CX_THREADS, 0, stream>>>(S, m, n, 1.f / band_width, stats, reinterpret_cast<const unsigned int*>(colmax), gcol, dS);
    SPI_COUNT_LAUNCH(1);
    SPI_LAUNCH_CHECK("spi_cx_rows_backward");
    return SPI_OK;
}
