// Noise-buffer regulariser and re-normalisation of the stage-1 projectors, one launch each for ALL buffers.
//
// Replaces the ~150 ATen launches per step of spi/training/projectors/mirror_projector.py:107-115 (and the identical loops of
// w_projector.py / w_plus_projector.py):
//     for v in noise_bufs:  noise = v[None, None]
//         while True:
//             reg += (noise * roll(noise, 1, dims=3)).mean()**2 + (noise * roll(noise, 1, dims=2)).mean()**2
//             if noise.shape[2] <= 8: break
//             noise = avg_pool2d(noise, 2)
// and of the post-step re-normalisation :128-131   buf -= buf.mean(); buf *= rsqrt(mean(buf^2)).
//
// One CTA per buffer (13 buffers, 4^2 .. 256^2 = 0.7 MB in total: latency-bound work, not bandwidth-bound).  The CTA builds the
// average-pool pyramid of its buffer in shared memory (level 1 from global, deeper levels from shared memory; <= 85 KB), reduces
// the two shifted auto-correlations of every level, and in the backward pass evaluates, for every full-resolution element, the
// sum over levels of the local gradient of its ancestor cell:
//     d/dx_l[i,j] (S/N)^2 = 2 S / N^2 * (x_l[i,j-1] + x_l[i,j+1])      (circular; likewise for the row shift)
//     d x_l[i>>l, j>>l] / d x_0[i,j] = 4^-l
#include "common.cuh"

namespace {

constexpr int NR_THREADS = 1024;
constexpr int NR_MAXLEV = 8;

struct NoiseBuf { float* x; long long out_off; int size; int pad; };     // square [size, size] fp32; gradient at out_base + out_off

__device__ __forceinline__ float block_sum(float v, float* red) {      // red: 32 floats of shared memory
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (warp == 0) {
        float t = lane < (blockDim.x >> 5) ? red[lane] : 0.f;
        t = warp_sum(t);
        if (lane == 0) red[0] = t;
    }
    __syncthreads();
    return red[0];
}

// levels 1.. of the pyramid into shared memory; returns the number of levels (level 0 = the buffer itself, in global memory)
__device__ __forceinline__ int build_pyramid(const float* __restrict__ x, int size, float* pyr, int* off) {
    int levels = 1, s = size, o = 0;
    while (s > 8) {
        const int h = s >> 1;
        off[levels] = o;
        const float* src = levels == 1 ? x : pyr + off[levels - 1];
        float* dst = pyr + o;
        for (int i = threadIdx.x; i < h * h; i += blockDim.x) {
            const int r = i / h, c = i - r * h;
            const float* p = src + (2 * r) * s + 2 * c;
            dst[i] = (p[0] + p[1] + p[s] + p[s + 1]) * 0.25f;
        }
        __syncthreads();
        o += h * h;
        s = h;
        levels++;
    }
    return levels;
}

// S_w = sum x[i,j] x[i,j-1], S_h = sum x[i,j] x[i-1,j] (circular) of one level
__device__ __forceinline__ void level_stats(const float* __restrict__ v, int s, float* red, float& sw, float& sh) {
    float a = 0.f, b = 0.f;
    for (int i = threadIdx.x; i < s * s; i += blockDim.x) {
        const int r = i / s, c = i - r * s;
        const float x = v[i];
        a = fmaf(x, v[r * s + (c == 0 ? s - 1 : c - 1)], a);
        b = fmaf(x, v[(r == 0 ? s - 1 : r - 1) * s + c], b);
    }
    sw = block_sum(a, red);
    sh = block_sum(b, red);
}

__global__ void __launch_bounds__(NR_THREADS) noise_reg_fwd_kernel(const NoiseBuf* __restrict__ bufs, float* __restrict__ partial,
                                                                   float* __restrict__ stats) {
    extern __shared__ float pyr[];
    __shared__ float red[32];
    __shared__ int off[NR_MAXLEV];
    const NoiseBuf nb = bufs[blockIdx.x];
    const int levels = build_pyramid(nb.x, nb.size, pyr, off);
    float loss = 0.f;
    int s = nb.size;
    for (int l = 0; l < levels; l++, s >>= 1) {
        float sw, sh;
        level_stats(l == 0 ? nb.x : pyr + off[l], s, red, sw, sh);
        const float inv = 1.f / (float)(s * s);
        const float mw = sw * inv, mh = sh * inv;
        loss += mw * mw;
        loss += mh * mh;
        if (threadIdx.x == 0) { stats[(blockIdx.x * NR_MAXLEV + l) * 2] = mw; stats[(blockIdx.x * NR_MAXLEV + l) * 2 + 1] = mh; }
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = loss;
}

// grad[i,j] = gout * sum_l 4^-l * 2/N_l * ( mw_l (x_l[.,j-1] + x_l[.,j+1]) + mh_l (x_l[i-1,.] + x_l[i+1,.]) ) at the ancestor cell
__global__ void __launch_bounds__(NR_THREADS) noise_reg_bwd_kernel(const NoiseBuf* __restrict__ bufs, const float* __restrict__ stats,
                                                                   const float* __restrict__ gout, float* __restrict__ out_base) {
    extern __shared__ float pyr[];
    __shared__ int off[NR_MAXLEV];
    __shared__ float cw[NR_MAXLEV], ch[NR_MAXLEV];
    const NoiseBuf nb = bufs[blockIdx.x];
    const int levels = build_pyramid(nb.x, nb.size, pyr, off);
    if (threadIdx.x < levels) {
        const int l = threadIdx.x, s = nb.size >> l;
        const float k = gout[0] * 2.f / ((float)(s * s)) / (float)(1 << (2 * l));
        cw[l] = k * stats[(blockIdx.x * NR_MAXLEV + l) * 2];
        ch[l] = k * stats[(blockIdx.x * NR_MAXLEV + l) * 2 + 1];
    }
    __syncthreads();
    const int size = nb.size;
    float* out = out_base + nb.out_off;
    for (int i = threadIdx.x; i < size * size; i += blockDim.x) {
        const int r0 = i / size, c0 = i - r0 * size;
        float g = 0.f;
        int s = size;
        for (int l = 0; l < levels; l++, s >>= 1) {
            const float* v = l == 0 ? nb.x : pyr + off[l];
            const int r = r0 >> l, c = c0 >> l;
            const int cl = c == 0 ? s - 1 : c - 1, cr = c == s - 1 ? 0 : c + 1;
            const int ru = r == 0 ? s - 1 : r - 1, rd = r == s - 1 ? 0 : r + 1;
            g = fmaf(cw[l], v[r * s + cl] + v[r * s + cr], g);
            g = fmaf(ch[l], v[ru * s + c] + v[rd * s + c], g);
        }
        out[i] = g;
    }
}

// buf -= mean(buf); buf *= rsqrt(mean(buf^2))   (in place; the second mean is taken after the shift, as the reference does)
__global__ void __launch_bounds__(NR_THREADS) noise_renorm_kernel(const NoiseBuf* __restrict__ bufs) {
    __shared__ float red[32];
    const NoiseBuf nb = bufs[blockIdx.x];
    float* x = nb.x;
    const int n = nb.size * nb.size;
    float a = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) a += x[i];
    const float mean = block_sum(a, red) / (float)n;
    float q = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) { const float d = x[i] - mean; q = fmaf(d, d, q); }
    const float scale = rsqrtf(block_sum(q, red) / (float)n);
    for (int i = threadIdx.x; i < n; i += blockDim.x) x[i] = (x[i] - mean) * scale;
}

size_t pyramid_bytes(int max_size) {
    size_t f = 0;
    for (int s = max_size; s > 8; s >>= 1) f += (size_t)(s / 2) * (s / 2);
    return f * sizeof(float);
}

int configure(const void* k, size_t smem, const char* name) {
    if (smem > 48 * 1024 && cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        spi_set_error("%s: cannot reserve %zu bytes of shared memory", name, smem);
        return SPI_ERR_CUDA;
    }
    return SPI_OK;
}

}  // namespace

/* table: device array of `count` records {float* x; long long out_off; int size; int pad} (24 bytes each, 8-byte aligned).
 * forward: partial[count] = per-buffer loss, stats[count*8*2] = per-level means of the two shifted products (kept for backward). */
extern "C" int spi_noise_reg_forward(const void* table, int count, int max_size, float* partial, float* stats, cudaStream_t stream) {
    SPI_CHECK_ARG(table && partial && stats && count > 0, "noise_reg_forward: null argument");
    SPI_CHECK_ARG(max_size >= 1 && max_size <= 256 && (max_size & (max_size - 1)) == 0, "noise_reg_forward: buffer sizes must be powers of two <= 256");
    const size_t smem = pyramid_bytes(max_size);
    if (int rc = configure((const void*)noise_reg_fwd_kernel, smem, "noise_reg_forward")) return rc;
    noise_reg_fwd_kernel<<<count, NR_THREADS, smem, stream>>>((const NoiseBuf*)table, partial, stats);
    SPI_COUNT_LAUNCH(1);
    SPI_LAUNCH_CHECK("noise_reg_forward");
    return SPI_OK;
}

/* backward: (out_base + out_off_b)[i,j] = gout[0] * d loss / d x_b[i,j] (written, not accumulated). */
extern "C" int spi_noise_reg_backward(const void* table, int count, int max_size, const float* stats, const float* gout, float* out_base,
                                      cudaStream_t stream) {
    SPI_CHECK_ARG(table && stats && gout && out_base && count > 0, "noise_reg_backward: null argument");
    SPI_CHECK_ARG(max_size >= 1 && max_size <= 256 && (max_size & (max_size - 1)) == 0, "noise_reg_backward: buffer sizes must be powers of two <= 256");
    const size_t smem = pyramid_bytes(max_size);
    if (int rc = configure((const void*)noise_reg_bwd_kernel, smem, "noise_reg_backward")) return rc;
    noise_reg_bwd_kernel<<<count, NR_THREADS, smem, stream>>>((const NoiseBuf*)table, stats, gout, out_base);
    SPI_COUNT_LAUNCH(1);
    SPI_LAUNCH_CHECK("noise_reg_backward");
    return SPI_OK;
}

/* in place on table[b].x (mirror_projector.py:128-131). */
extern "C" int spi_noise_renorm(const void* table, int count, cudaStream_t stream) {
    SPI_CHECK_ARG(table && count > 0, "noise_renorm: null argument");
    noise_renorm_kernel<<<count, NR_THREADS, 0, stream>>>((const NoiseBuf*)table);
    SPI_COUNT_LAUNCH(1);
    SPI_LAUNCH_CHECK("noise_renorm");
    return SPI_OK;
}
