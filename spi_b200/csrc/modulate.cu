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
// Tensors are small (<= 4 x 2.4 M floats) but there are 29 layers per pass: rows are staged in shared memory so that W and the
// gradient are read once, coalesced (the first, unstaged kernels are kept for rows that do not fit).
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


// ---- row-staged kernels (the per-layer weight tensors are small, but there are 29 layers per generator pass and the
// three launches per layer added up to 4.5 % of an iteration: these variants read W once, coalesced, into shared memory)

// Forward, layouts 0 (OIK) and 1 (OKI): grid (O, N), one CTA per output row; the modulated row W*s is staged in shared
// memory, reduced for the demodulation coefficient, and written out in the destination order with coalesced stores.
__device__ __forceinline__ void modulate_row_body(const float* __restrict__ W, const float* __restrict__ s, float* __restrict__ dcoef,
                                                  float* __restrict__ out, int O, int I, int KK, int layout_flags, int demod, int o, int n,
                                                  float* row, float* sh) {
    const int layout = layout_flags & 3;
    const bool flip = (layout_flags & 4) != 0;
    const int len = I * KK;
    const float* w = W + (size_t)o * len;
    const float* sn = s + (size_t)n * I;
    float acc = 0.f;
    for (int e = threadIdx.x; e < len; e += blockDim.x) { const float v = w[e] * sn[e / KK]; row[e] = v; acc += v * v; }
    float d = 1.f;
    if (demod) {
        d = rsqrtf(block_sum(acc, sh) + 1e-8f);
        if (threadIdx.x == 0) dcoef[(size_t)n * O + o] = d;
    } else {
        __syncthreads();
    }
    float* dst = out + ((size_t)n * O + o) * len;
    if (layout == 0) {
        for (int e = threadIdx.x; e < len; e += blockDim.x) {
            int src = e;
            if (flip) { const int k = e % KK; src = e - k + (KK - 1 - k); }
            const float v = row[src];
            dst[e] = demod ? v * d : v;
        }
    } else {
        for (int k = 0; k < KK; k++) {
            const int kk = flip ? KK - 1 - k : k;
            for (int i = threadIdx.x; i < I; i += blockDim.x) {       // shared-memory stride KK (odd): conflict-free
                const float v = row[i * KK + kk];
                dst[k * I + i] = demod ? v * d : v;
            }
        }
    }
}

__global__ void __launch_bounds__(256) modulate_row_kernel(const float* __restrict__ W, const float* __restrict__ s, float* __restrict__ dcoef,
                                                           float* __restrict__ out, int O, int I, int KK, int layout_flags, int demod) {
    extern __shared__ float row[];
    __shared__ float sh[32];
    modulate_row_body(W, s, dcoef, out, O, I, KK, layout_flags, demod, blockIdx.x, blockIdx.y, row, sh);
}

// Forward, layout 2 (IKO, the transposed-convolution weights): 32 x 32 (o, e) tiles transposed through shared memory so that
// both the W reads (contiguous in e = i*KK + k) and the stores (contiguous in o) are coalesced.  grid (ceil(len/32), ceil(O/32), N).
__global__ void __launch_bounds__(256) modulate_apply_iko_kernel(const float* __restrict__ W, const float* __restrict__ s,
                                                                 const float* __restrict__ dcoef, float* __restrict__ out, int O, int I,
                                                                 int KK, int flip) {
    __shared__ float tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int e0 = blockIdx.x * 32, o0 = blockIdx.y * 32, n = blockIdx.z, len = I * KK;
    {
        const int e = e0 + tx;
        const float se = e < len ? s[(size_t)n * I + e / KK] : 0.f;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int o = o0 + ty + 8 * j;
            if (o < O && e < len) {
                float v = W[(size_t)o * len + e] * se;
                if (dcoef) v *= dcoef[(size_t)n * O + o];
                tile[ty + 8 * j][tx] = v;
            }
        }
    }
    __syncthreads();
    const int o = o0 + tx;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int el = ty + 8 * j, e = e0 + el;
        if (e < len && o < O) {
            int e2 = e;
            if (flip) { const int k = e % KK; e2 = e - k + (KK - 1 - k); }
            out[((size_t)n * len + e2) * O + o] = tile[tx][el];
        }
    }
}

// Backward, every layout: grid (O); the W row and the incoming gradient row are staged in shared memory in W's own (i, k)
// order (coalesced loads for layouts 0 and 1), then each thread owns whole channels i: A, t, dW and ds come from shared memory,
// dW leaves through a coalesced store.  Shared memory: 2 * I * KK floats.
__device__ __forceinline__ void modulate_bwd_rows_body(const float* __restrict__ W, const float* __restrict__ s, const float* __restrict__ dcoef,
                                                       const float* __restrict__ g, float* __restrict__ dW, float* __restrict__ ds, int N, int O,
                                                       int I, int KK, int demod, int layout_flags, int o, float* sm, float* sh) {
    const int layout = layout_flags & 3;
    const bool flip = (layout_flags & 4) != 0;
    const int len = I * KK;
    float* wrow = sm;
    float* grow = sm + len;
    const float* w = W + (size_t)o * len;
    for (int e = threadIdx.x; e < len; e += blockDim.x) wrow[e] = w[e];
    for (int n = 0; n < N; n++) {
        __syncthreads();                                   // previous sample's grow fully consumed; wrow visible
        if (layout == 0) {
            const float* gr = g + ((size_t)n * O + o) * len;
            for (int e = threadIdx.x; e < len; e += blockDim.x) {
                int dst = e;
                if (flip) { const int k = e % KK; dst = e - k + (KK - 1 - k); }
                grow[dst] = gr[e];
            }
        } else if (layout == 1) {
            const float* gr = g + ((size_t)n * O + o) * len;
            for (int k = 0; k < KK; k++) {
                const int kk = flip ? KK - 1 - k : k;
                for (int i = threadIdx.x; i < I; i += blockDim.x) grow[i * KK + kk] = gr[k * I + i];
            }
        } else {
            const float* gr = g + (size_t)n * len * O + o;
            for (int e = threadIdx.x; e < len; e += blockDim.x) {
                int dst = e;
                if (flip) { const int k = e % KK; dst = e - k + (KK - 1 - k); }
                grow[dst] = gr[(size_t)e * O];
            }
        }
        __syncthreads();
        const float* sn = s + (size_t)n * I;
        float d = 1.f, dA = 0.f;
        if (demod) {
            d = dcoef[(size_t)n * O + o];
            float acc = 0.f;
            for (int i = threadIdx.x; i < I; i += blockDim.x) {
                const float si = sn[i];
                float a = 0.f;
                for (int k = 0; k < KK; k++) a = fmaf(grow[i * KK + k], wrow[i * KK + k], a);
                acc = fmaf(a, si, acc);
            }
            dA = d * d * block_sum(acc, sh);
        }
        for (int i = threadIdx.x; i < I; i += blockDim.x) {       // thread owns channel i: one atomic per (n, i) per CTA
            const float si = sn[i];
            float dsi = 0.f;
            for (int k = 0; k < KK; k++) {
                const int e = i * KK + k;
                const float we = wrow[e];
                const float ge = grow[e];
                const float t = demod ? d * (ge - dA * we * si) : ge;
                grow[e] = si * t;                                  // this sample's contribution to dW[o, i, k]
                dsi = fmaf(we, t, dsi);
            }
            if (ds) atomicAdd(ds + (size_t)n * I + i, dsi);
        }
        if (dW) {
            __syncthreads();
            float* dw = dW + (size_t)o * len;
            if (n == 0) { for (int e = threadIdx.x; e < len; e += blockDim.x) dw[e] = grow[e]; }
            else { for (int e = threadIdx.x; e < len; e += blockDim.x) dw[e] += grow[e]; }
        }
    }
}

__global__ void __launch_bounds__(256) modulate_bwd_rows_kernel(const float* __restrict__ W, const float* __restrict__ s,
                                                                const float* __restrict__ dcoef, const float* __restrict__ g,
                                                                float* __restrict__ dW, float* __restrict__ ds, int N, int O, int I, int KK,
                                                                int demod, int layout_flags) {
    extern __shared__ float sm[];
    __shared__ float sh[32];
    modulate_bwd_rows_body(W, s, dcoef, g, dW, ds, N, O, I, KK, demod, layout_flags, blockIdx.x, sm, sh);
}

// ---- the same two kernels over EVERY modulated convolution of a synthesis network in one launch each (the per-layer launches are small
// -- 20 MB of traffic, 8 / 12 us -- and latency-bound; grouped, the 26 layers of the generator stream their 120 MB of weights once).
constexpr int MOD_MAX = 32;
struct ModLayer {
    const float* W; const float* s; float* dcoef; float* out;      // forward
    const float* g; float* dW; float* ds;                          // backward
    int O, I, KK, n, layout_flags, demod, row0;
};
struct ModArgs {
    ModLayer layer[MOD_MAX];
    int layers, rows;
};

__global__ void __launch_bounds__(256) modulate_rows_many_kernel(const __grid_constant__ ModArgs a) {
    extern __shared__ float row[];
    __shared__ float sh[32];
    int l = 0;
    while (l + 1 < a.layers && a.layer[l + 1].row0 <= (int)blockIdx.x) l++;
    const ModLayer& L = a.layer[l];
    const int r = blockIdx.x - L.row0;
    modulate_row_body(L.W, L.s, L.dcoef, L.out, L.O, L.I, L.KK, L.layout_flags, L.demod, r % L.O, r / L.O, row, sh);
}

__global__ void __launch_bounds__(256) modulate_bwd_rows_many_kernel(const __grid_constant__ ModArgs a) {
    extern __shared__ float sm[];
    __shared__ float sh[32];
    int l = 0;
    while (l + 1 < a.layers && a.layer[l + 1].row0 <= (int)blockIdx.x) l++;
    const ModLayer& L = a.layer[l];
    modulate_bwd_rows_body(L.W, L.s, L.dcoef, L.g, L.dW, L.ds, L.n, L.O, L.I, L.KK, L.demod, L.layout_flags, blockIdx.x - L.row0, sm, sh);
}

}  // namespace

extern "C" int spi_modulate_weights(const float* weight, const float* styles, float* out, float* dcoef, int n, int o, int i, int kk,
                                    int demodulate, int layout, cudaStream_t stream) {
    SPI_CHECK_ARG(layout >= 0 && (layout & 3) <= 2 && layout < 8, "modulate_weights: layout must be 0 (OIK), 1 (OKI) or 2 (IKO), +4 to reverse the taps");
    SPI_CHECK_ARG(weight && styles && out, "modulate_weights: null pointer");
    SPI_CHECK_ARG(n >= 1 && o >= 1 && i >= 1 && kk >= 1 && n <= 65535, "modulate_weights: bad shape");
    SPI_CHECK_ARG(!demodulate || dcoef, "modulate_weights: dcoef buffer required when demodulating");
    const size_t row_bytes = sizeof(float) * (size_t)i * kk;
    if ((layout & 3) != 2 && row_bytes <= 40 * 1024) {
        // one launch: the row is staged in shared memory between the reduction and the store
        modulate_row_kernel<<<dim3(o, n), 256, row_bytes, stream>>>(weight, styles, dcoef, out, o, i, kk, layout, demodulate);
        SPI_COUNT_LAUNCH(1);
    } else if ((layout & 3) == 2) {
        if (demodulate) demod_coef_kernel<<<dim3(o, n), 256, 0, stream>>>(weight, styles, dcoef, o, i, kk);
        modulate_apply_iko_kernel<<<dim3((i * kk + 31) / 32, (o + 31) / 32, n), 256, 0, stream>>>(weight, styles, demodulate ? dcoef : nullptr, out,
                                                                                                 o, i, kk, (layout & 4) != 0);
        SPI_COUNT_LAUNCH(demodulate ? 2 : 1);
    } else {
        if (demodulate) demod_coef_kernel<<<dim3(o, n), 256, 0, stream>>>(weight, styles, dcoef, o, i, kk);
        long long total = (long long)n * o * i * kk;
        long long blocks = (total + 255) / 256, cap = (long long)spi_num_sms() * 8;
        modulate_apply_kernel<<<(int)(blocks > cap ? cap : blocks), 256, 0, stream>>>(weight, styles, demodulate ? dcoef : nullptr, out, n, o, i, kk, layout);
        SPI_COUNT_LAUNCH(demodulate ? 2 : 1);
    }
    SPI_LAUNCH_CHECK("modulate_weights");
    return SPI_OK;
}

extern "C" int spi_modulate_weights_backward(const float* weight, const float* styles, const float* dcoef, const float* grad_out,
                                             float* grad_weight, float* grad_styles, int n, int o, int i, int kk, int demodulate,
                                             int layout, cudaStream_t stream) {
    SPI_CHECK_ARG(weight && styles && grad_out, "modulate_weights_backward: null pointer");
    SPI_CHECK_ARG(!demodulate || dcoef, "modulate_weights_backward: dcoef required when demodulating");
    const bool prezeroed = (layout & 8) != 0;           // + 8: grad_styles is already zero on entry (a slice of the caller's zero arena)
    layout &= 7;
    if (grad_styles && !prezeroed) cudaMemsetAsync(grad_styles, 0, sizeof(float) * (size_t)n * i, stream);
    const size_t rows_bytes = 2 * sizeof(float) * (size_t)i * kk;
    if (rows_bytes <= 40 * 1024)
        modulate_bwd_rows_kernel<<<o, 256, rows_bytes, stream>>>(weight, styles, dcoef, grad_out, grad_weight, grad_styles, n, o, i, kk, demodulate, layout);
    else
        modulate_bwd_kernel<<<o, 256, 0, stream>>>(weight, styles, dcoef, grad_out, grad_weight, grad_styles, n, o, i, kk, demodulate, layout);
    SPI_COUNT_LAUNCH(1);
    SPI_LAUNCH_CHECK("modulate_weights_backward");
    return SPI_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Style bank: every affine layer of a synthesis network (FullyConnectedLayer(w_dim, in_channels, bias_init=1) of each
// SynthesisLayer / ToRGBLayer, eg3d/training/networks_stylegan2.py:282,316 and :352,357-358) evaluated by ONE launch, and
// their weight / bias / latent gradients by one more (two when the latents are being optimised).  Per layer the work is a
// [I x 512] matrix-vector product -- a few microseconds of HBM time each, but 26 layers x (addmm, mul, and three backward
// GEMV / outer-product launches) per iteration; grouped, the whole generator's affines cost one read of their 20 MB of weights.
//   styles_l[n,i] = ogain_l * (wgain_l * sum_k ws[n, widx_l, k] * W_l[i,k] + bgain_l * b_l[i])
namespace {

constexpr int BANK_MAX = 32;
struct BankLayer {
    const float* W; const float* b; float* out;            // forward
    const float* ds; float* dW; float* db;                 // backward (ds null: this layer's styles were not used)
    int I, widx, row0;
    float wgain, bgain, ogain;
};
struct BankArgs {
    BankLayer layer[BANK_MAX];
    int layers, n, k, rows;
    long long ws_sn, ws_sl;                                // element strides of ws over samples / latent index (last dim contiguous)
};

__device__ __forceinline__ int bank_find(const BankArgs& a, int row) {
    int l = 0;
    while (l + 1 < a.layers && a.layer[l + 1].row0 <= row) l++;
    return l;
}

// one warp per (layer, output row)
__global__ void __launch_bounds__(256) style_bank_fwd_kernel(const __grid_constant__ BankArgs a, const float* __restrict__ ws) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= a.rows) return;
    const int l = bank_find(a, row);
    const BankLayer& L = a.layer[l];
    const int i = row - L.row0;
    if (i >= L.I) return;                                  // padding rows between layers
    const float4* wr = reinterpret_cast<const float4*>(L.W + (size_t)i * a.k);
    const float bias = L.b ? L.b[i] * L.bgain : 0.f;
    for (int n = 0; n < a.n; n++) {
        const float4* x = reinterpret_cast<const float4*>(ws + n * a.ws_sn + L.widx * a.ws_sl);
        float acc = 0.f;
        for (int c = lane; c < a.k / 4; c += 32) {
            const float4 w4 = __ldg(wr + c), x4 = __ldg(x + c);
            acc = fmaf(w4.x, x4.x, acc); acc = fmaf(w4.y, x4.y, acc); acc = fmaf(w4.z, x4.z, acc); acc = fmaf(w4.w, x4.w, acc);
        }
        acc = warp_sum(acc);
        if (lane == 0) L.out[(size_t)n * L.I + i] = L.ogain * (L.wgain * acc + bias);
    }
}

// one warp per (layer, row): dW_l[i,:] = ogain*wgain * sum_n ds[n,i] * ws[n,widx,:],  db_l[i] = ogain*bgain * sum_n ds[n,i]
__global__ void __launch_bounds__(256) style_bank_bwd_kernel(const __grid_constant__ BankArgs a, const float* __restrict__ ws) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= a.rows) return;
    const int l = bank_find(a, row);
    const BankLayer& L = a.layer[l];
    const int i = row - L.row0;
    if (i >= L.I || (!L.dW && !L.db)) return;
    float4* dw = L.dW ? reinterpret_cast<float4*>(L.dW + (size_t)i * a.k) : nullptr;
    const float gw = L.ogain * L.wgain;
    float dsum = 0.f;
    for (int n = 0; n < a.n; n++) dsum += L.ds ? L.ds[(size_t)n * L.I + i] : 0.f;
    if (L.db && lane == 0) L.db[i] = L.ogain * L.bgain * dsum;
    if (!dw) return;
    for (int c = lane; c < a.k / 4; c += 32) {
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
        if (L.ds)
            for (int n = 0; n < a.n; n++) {
                const float d = L.ds[(size_t)n * L.I + i] * gw;
                const float4 x4 = __ldg(reinterpret_cast<const float4*>(ws + n * a.ws_sn + L.widx * a.ws_sl) + c);
                r.x = fmaf(d, x4.x, r.x); r.y = fmaf(d, x4.y, r.y); r.z = fmaf(d, x4.z, r.z); r.w = fmaf(d, x4.w, r.w);
            }
        dw[c] = r;
    }
}

// latent gradient: dws[n, widx_l, :] += ogain*wgain * sum_i ds_l[n,i] * W_l[i,:]; grid (row chunks of 32 over all layers, n), a thread owns
// 4 consecutive k; dws (contiguous [n][L][k]) zeroed by the caller
__global__ void __launch_bounds__(128) style_bank_dws_kernel(const __grid_constant__ BankArgs a, float* __restrict__ dws, int num_ws) {
    const int row_begin = blockIdx.x * 32, n = blockIdx.y;
    const int l = bank_find(a, row_begin);
    const BankLayer& L = a.layer[l];
    if (!L.ds) return;
    const int i0 = row_begin - L.row0, i1 = min(L.I, i0 + 32);          // row0 of every layer is a multiple of 32 (host pads)
    const float g = L.ogain * L.wgain;
    for (int c = threadIdx.x; c < a.k / 4; c += blockDim.x) {
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i = i0; i < i1; i++) {
            const float d = L.ds[(size_t)n * L.I + i] * g;
            const float4 w4 = __ldg(reinterpret_cast<const float4*>(L.W + (size_t)i * a.k) + c);
            r.x = fmaf(d, w4.x, r.x); r.y = fmaf(d, w4.y, r.y); r.z = fmaf(d, w4.z, r.z); r.w = fmaf(d, w4.w, r.w);
        }
        float* dst = dws + ((size_t)n * num_ws + L.widx) * a.k + c * 4;
        atomicAdd(dst, r.x); atomicAdd(dst + 1, r.y); atomicAdd(dst + 2, r.z); atomicAdd(dst + 3, r.w);
    }
}

int bank_args(BankArgs& a, int layers, int n, int k, const int* I, const int* widx, const float* wgain, const float* bgain, const float* ogain,
              long long ws_sn, long long ws_sl) {
    if (layers < 1 || layers > BANK_MAX || n < 1 || k < 4 || (k & 3)) return SPI_ERR_ARG;
    int rows = 0;
    for (int l = 0; l < layers; l++) {
        if (I[l] < 1 || widx[l] < 0) return SPI_ERR_ARG;
        BankLayer& L = a.layer[l];
        L.I = I[l]; L.widx = widx[l]; L.row0 = rows; L.wgain = wgain[l]; L.bgain = bgain[l]; L.ogain = ogain[l];
        rows += (I[l] + 31) / 32 * 32;                       // rows of a layer start on a multiple of 32 (the dws kernel's chunks)
    }
    a.layers = layers; a.n = n; a.k = k; a.rows = rows; a.ws_sn = ws_sn; a.ws_sl = ws_sl;
    return SPI_OK;
}

}  // namespace

// ws: latents, element strides (ws_sn, ws_sl, 1); per layer l < layers: W[l] [I[l]][k], b[l] [I[l]] (may be null), out[l] [n][I[l]].
// The pointer / scalar tables are HOST arrays (read during the call only).
extern "C" int spi_style_bank_forward(const float* ws, long long ws_sn, long long ws_sl, int n, int k, int layers, const float* const* W,
                                      const float* const* b, float* const* out, const int* I, const int* widx, const float* wgain,
                                      const float* bgain, const float* ogain, cudaStream_t stream) {
    SPI_CHECK_ARG(ws && W && b && out && I && widx && wgain && bgain && ogain, "style_bank_forward: null table");
    BankArgs a{};
    SPI_CHECK_ARG(bank_args(a, layers, n, k, I, widx, wgain, bgain, ogain, ws_sn, ws_sl) == SPI_OK, "style_bank_forward: bad shape (layers=%d n=%d k=%d)", layers, n, k);
    SPI_CHECK_ARG((((uintptr_t)ws | (uintptr_t)(ws_sn * 4) | (uintptr_t)(ws_sl * 4)) & 15) == 0, "style_bank_forward: latents must be 16-byte aligned");
    for (int l = 0; l < layers; l++) {
        SPI_CHECK_ARG(W[l] && out[l] && ((uintptr_t)W[l] & 15) == 0, "style_bank_forward: layer %d: null / misaligned weight", l);
        a.layer[l].W = W[l]; a.layer[l].b = b[l]; a.layer[l].out = out[l];
    }
    style_bank_fwd_kernel<<<(a.rows + 7) / 8, 256, 0, stream>>>(a, ws);
    SPI_COUNT_LAUNCH(1);
    SPI_LAUNCH_CHECK("style_bank_forward");
    return SPI_OK;
}

// ds[l] [n][I[l]] (null: zero); dW[l] / db[l] (null: not wanted) are overwritten; dws (null: not wanted) is contiguous [n][num_ws][k], overwritten.
extern "C" int spi_style_bank_backward(const float* ws, long long ws_sn, long long ws_sl, int n, int k, int layers, const float* const* W,
                                       const float* const* ds, float* const* dW, float* const* db, const int* I, const int* widx,
                                       const float* wgain, const float* bgain, const float* ogain, float* dws, int num_ws, cudaStream_t stream) {
    SPI_CHECK_ARG(ws && W && ds && dW && db && I && widx && wgain && bgain && ogain, "style_bank_backward: null table");
    BankArgs a{};
    SPI_CHECK_ARG(bank_args(a, layers, n, k, I, widx, wgain, bgain, ogain, ws_sn, ws_sl) == SPI_OK, "style_bank_backward: bad shape (layers=%d n=%d k=%d)", layers, n, k);
    SPI_CHECK_ARG((((uintptr_t)ws | (uintptr_t)(ws_sn * 4) | (uintptr_t)(ws_sl * 4)) & 15) == 0, "style_bank_backward: latents must be 16-byte aligned");
    bool any = false;
    for (int l = 0; l < layers; l++) {
        SPI_CHECK_ARG(W[l] && ((uintptr_t)W[l] & 15) == 0 && ((uintptr_t)dW[l] & 15) == 0, "style_bank_backward: layer %d: null / misaligned weight", l);
        SPI_CHECK_ARG(!dws || widx[l] < num_ws, "style_bank_backward: layer %d: latent index %d outside num_ws=%d", l, widx[l], num_ws);
        a.layer[l].W = W[l]; a.layer[l].ds = ds[l]; a.layer[l].dW = dW[l]; a.layer[l].db = db[l];
        any = any || dW[l] || db[l];
    }
    int launches = 0;
    if (any) { style_bank_bwd_kernel<<<(a.rows + 7) / 8, 256, 0, stream>>>(a, ws); launches++; }
    if (dws) {
        cudaMemsetAsync(dws, 0, sizeof(float) * (size_t)n * num_ws * k, stream);
        style_bank_dws_kernel<<<dim3(a.rows / 32, n), 128, 0, stream>>>(a, dws, num_ws);
        launches++;
    }
    SPI_COUNT_LAUNCH(launches);
    SPI_LAUNCH_CHECK("style_bank_backward");
    return SPI_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Grouped entry points: every modulated convolution of a network per launch.  All tables are HOST arrays of `layers` entries (<= 32).
namespace {

// w [G][O][T][I] -> wt [G][I][T'][O] for many tensors in one launch (T' = T-1-t when reverse): 32 x 32 tiles through shared memory
struct TrLayer { const float* w; float* wt; int g, o, taps, i, reverse, tiles_i, tiles_o, tile0; };
struct TrArgs { TrLayer layer[MOD_MAX]; int layers; };

__global__ void __launch_bounds__(256) weight_transpose_many_kernel(const __grid_constant__ TrArgs a) {
    __shared__ float tile[32][33];
    int l = 0;
    while (l + 1 < a.layers && a.layer[l + 1].tile0 <= (int)blockIdx.x) l++;
    const TrLayer& L = a.layer[l];
    int r = blockIdx.x - L.tile0;
    const int ti = r % L.tiles_i; r /= L.tiles_i;
    const int to = r % L.tiles_o; r /= L.tiles_o;
    const int t = r % L.taps, g = r / L.taps;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int o0 = to * 32, i0 = ti * 32;
    const float* src = L.w + ((size_t)g * L.o * L.taps + t) * L.i;
    float* dst = L.wt + ((size_t)g * L.i * L.taps + (L.reverse ? L.taps - 1 - t : t)) * L.o;
    for (int q = ty; q < 32; q += 8) {
        const int oo = o0 + q, ii = i0 + tx;
        tile[q][tx] = (oo < L.o && ii < L.i) ? src[(size_t)oo * L.taps * L.i + ii] : 0.f;
    }
    __syncthreads();
    for (int q = ty; q < 32; q += 8) {
        const int ii = i0 + q, oo = o0 + tx;
        if (ii < L.i && oo < L.o) dst[(size_t)ii * L.taps * L.o + oo] = tile[tx][q];
    }
}

}  // namespace

extern "C" int spi_conv_weight_transpose_many(int layers, const float* const* w, float* const* wt, const int* g, const int* o, const int* taps,
                                              const int* i, const int* reverse, cudaStream_t stream) {
    SPI_CHECK_ARG(layers >= 1 && layers <= MOD_MAX && w && wt && g && o && taps && i && reverse, "conv_weight_transpose_many: bad tables (layers=%d)", layers);
    TrArgs a{};
    long long tiles = 0;
    for (int l = 0; l < layers; l++) {
        SPI_CHECK_ARG(w[l] && wt[l] && g[l] > 0 && o[l] > 0 && taps[l] > 0 && i[l] > 0, "conv_weight_transpose_many: layer %d: bad arguments", l);
        TrLayer& L = a.layer[l];
        L.w = w[l]; L.wt = wt[l]; L.g = g[l]; L.o = o[l]; L.taps = taps[l]; L.i = i[l]; L.reverse = reverse[l] ? 1 : 0;
        L.tiles_i = (i[l] + 31) / 32; L.tiles_o = (o[l] + 31) / 32; L.tile0 = (int)tiles;
        tiles += (long long)L.tiles_i * L.tiles_o * taps[l] * g[l];
        SPI_CHECK_ARG(tiles < (1LL << 30), "conv_weight_transpose_many: too many tiles");
    }
    a.layers = layers;
    weight_transpose_many_kernel<<<(unsigned)tiles, 256, 0, stream>>>(a);
    SPI_COUNT_LAUNCH(1);
    SPI_LAUNCH_CHECK("conv_weight_transpose_many");
    return SPI_OK;
}

// Forward of spi_modulate_weights for `layers` convolutions at once; layout[l] in {0, 1} (+4: taps reversed), rows of at most 40 KB.
extern "C" int spi_modulate_weights_many(int layers, const float* const* weight, const float* const* styles, float* const* out, float* const* dcoef,
                                         const int* n, const int* o, const int* i, const int* kk, const int* demodulate, const int* layout,
                                         cudaStream_t stream) {
    SPI_CHECK_ARG(layers >= 1 && layers <= MOD_MAX && weight && styles && out && dcoef && n && o && i && kk && demodulate && layout,
                  "modulate_weights_many: bad tables (layers=%d)", layers);
    ModArgs a{};
    long long rows = 0;
    size_t smem = 0;
    for (int l = 0; l < layers; l++) {
        SPI_CHECK_ARG(weight[l] && styles[l] && out[l] && n[l] >= 1 && o[l] >= 1 && i[l] >= 1 && kk[l] >= 1, "modulate_weights_many: layer %d: bad arguments", l);
        SPI_CHECK_ARG(layout[l] >= 0 && layout[l] < 8 && (layout[l] & 3) <= 1, "modulate_weights_many: layer %d: layout must be 0 (OIK) or 1 (OKI), +4 to reverse the taps", l);
        SPI_CHECK_ARG(!demodulate[l] || dcoef[l], "modulate_weights_many: layer %d: dcoef buffer required when demodulating", l);
        const size_t row_bytes = sizeof(float) * (size_t)i[l] * kk[l];
        SPI_CHECK_ARG(row_bytes <= 40 * 1024, "modulate_weights_many: layer %d: row of %zu bytes does not fit (use spi_modulate_weights)", l, row_bytes);
        smem = row_bytes > smem ? row_bytes : smem;
        ModLayer& L = a.layer[l];
        L.W = weight[l]; L.s = styles[l]; L.dcoef = dcoef[l]; L.out = out[l];
        L.O = o[l]; L.I = i[l]; L.KK = kk[l]; L.n = n[l]; L.layout_flags = layout[l]; L.demod = demodulate[l] ? 1 : 0; L.row0 = (int)rows;
        rows += (long long)o[l] * n[l];
        SPI_CHECK_ARG(rows < (1LL << 30), "modulate_weights_many: too many rows");
    }
    a.layers = layers; a.rows = (int)rows;
    modulate_rows_many_kernel<<<(unsigned)rows, 256, smem, stream>>>(a);
    SPI_COUNT_LAUNCH(1);
    SPI_LAUNCH_CHECK("modulate_weights_many");
    return SPI_OK;
}

// Backward of the same: grad_out[l] in layout[l]; grad_weight[l] (null: not wanted) overwritten; grad_styles[l] (null: not wanted) must be ZERO
// on entry (the caller packs them into one zero-filled buffer: they are accumulated with atomics over the output channels).
extern "C" int spi_modulate_weights_backward_many(int layers, const float* const* weight, const float* const* styles, const float* const* dcoef,
                                                  const float* const* grad_out, float* const* grad_weight, float* const* grad_styles, const int* n,
                                                  const int* o, const int* i, const int* kk, const int* demodulate, const int* layout,
                                                  cudaStream_t stream) {
    SPI_CHECK_ARG(layers >= 1 && layers <= MOD_MAX && weight && styles && dcoef && grad_out && grad_weight && grad_styles && n && o && i && kk && demodulate && layout,
                  "modulate_weights_backward_many: bad tables (layers=%d)", layers);
    ModArgs a{};
    long long rows = 0;
    size_t smem = 0;
    for (int l = 0; l < layers; l++) {
        SPI_CHECK_ARG(weight[l] && styles[l] && grad_out[l] && n[l] >= 1 && o[l] >= 1 && i[l] >= 1 && kk[l] >= 1, "modulate_weights_backward_many: layer %d: bad arguments", l);
        SPI_CHECK_ARG(layout[l] >= 0 && layout[l] < 8 && (layout[l] & 3) <= 2, "modulate_weights_backward_many: layer %d: bad layout", l);
        SPI_CHECK_ARG(!demodulate[l] || dcoef[l], "modulate_weights_backward_many: layer %d: dcoef required when demodulating", l);
        const size_t rows_bytes = 2 * sizeof(float) * (size_t)i[l] * kk[l];
        SPI_CHECK_ARG(rows_bytes <= 80 * 1024, "modulate_weights_backward_many: layer %d: rows of %zu bytes do not fit (use spi_modulate_weights_backward)", l, rows_bytes);
        smem = rows_bytes > smem ? rows_bytes : smem;
        ModLayer& L = a.layer[l];
        L.W = weight[l]; L.s = styles[l]; L.dcoef = const_cast<float*>(dcoef[l]); L.g = grad_out[l]; L.dW = grad_weight[l]; L.ds = grad_styles[l];
        L.O = o[l]; L.I = i[l]; L.KK = kk[l]; L.n = n[l]; L.layout_flags = layout[l]; L.demod = demodulate[l] ? 1 : 0; L.row0 = (int)rows;
        rows += o[l];
    }
    a.layers = layers; a.rows = (int)rows;
    if (smem > 48 * 1024) {
        static bool configured = false;
        if (!configured) {
            if (cudaFuncSetAttribute(modulate_bwd_rows_many_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024) != cudaSuccess) {
                spi_set_error("modulate_weights_backward_many: cannot reserve 80 KB of shared memory");
                return SPI_ERR_CUDA;
            }
            configured = true;
        }
    }
    modulate_bwd_rows_many_kernel<<<(unsigned)rows, 256, smem, stream>>>(a);
    SPI_COUNT_LAUNCH(1);
    SPI_LAUNCH_CHECK("modulate_weights_backward_many");
    return SPI_OK;
}
