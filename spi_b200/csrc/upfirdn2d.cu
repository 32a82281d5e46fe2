// upfirdn2d for sm_100a: zero-insert upsample -> pad/crop -> 2-D FIR -> decimate, per channel.
//
// Replaces the reference plugin op `upfirdn2d` (eg3d/torch_utils/ops/upfirdn2d.cpp:20-105, kernels
// upfirdn2d.cu:33-204) with the same argument meaning (f is a rank-2 fp32 filter [fh, fw]; flip=false means
// true convolution).  Definition used (equivalent to upfirdn2d.py:169-211):
//   U[t]  = x[(t - pad0) / up]  if (t - pad0) % up == 0 and in range, else 0
//   y[o]  = gain * sum_k U[o*down + k] * w[k],   w[k] = flip ? f[k] : f[fw-1-k]
//
// B200 design: HBM-streaming; the polyphase structure is resolved per thread so only the non-zero taps are
// visited ((fw/up) x (fh/up) loads instead of fw x fh).  Two thread mappings:
//   * x-fastest (contiguous NCHW): each thread produces a strip of 4 outputs along W, re-using loaded inputs in
//     registers; warps read 128-B rows, L1 absorbs the vertical re-use.
//   * channel-fastest (channels_last): each thread produces 4 consecutive channels of one pixel with LDG.128.
// The filter (<= 32x32) is staged in shared memory once per CTA.  Grid is capped to a multiple of the SM count
// and loops (persistent-style) over work items.
#include "common.cuh"

namespace {

struct UpfirdnParams {
    const void* x; const float* f; void* y;
    int upx, upy, downx, downy, padx0, pady0, flip;
    float gain;
    int n, c, inH, inW, outH, outW, fw, fh;
    long long xs_n, xs_c, xs_h, xs_w;   // element strides of x
    long long ys_n, ys_c, ys_h, ys_w;
    // optional SynthesisLayer epilogue fused into the 4x4 blur kernels (spi_blur4_bias_act_noise): bias [C], noise map
    // [outH*outW] * device scalar, act 1 (linear) / 3 (lrelu), gain, clamp (< 0: off)
    const float* eb; const float* enoise; const float* estrength;
    float ealpha, egain, eclamp; int eact;
};

// y = clamp(act(v + (bias + noise)) * gain): the arithmetic of bias_act.cu::act_eval for act 1 / 3, forward
template <bool EPI>
__device__ __forceinline__ float4 blur_epilogue(const UpfirdnParams& p, float4 v, float4 bias, float nz) {
    if (!EPI) return v;
    float r[4] = {v.x + (bias.x + nz), v.y + (bias.y + nz), v.z + (bias.z + nz), v.w + (bias.w + nz)};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        float t = r[k];
        if (p.eact == 3) t = (t > 0.f) ? t : t * p.ealpha;
        t *= p.egain * 1.f;
        if (p.eclamp >= 0.f) t = (t > -p.eclamp && t < p.eclamp) ? t : ((t >= 0.f) ? p.eclamp : -p.eclamp);
        r[k] = t;
    }
    return make_float4(r[0], r[1], r[2], r[3]);
}

__device__ __forceinline__ int floordiv(int a, int b) { int q = a / b; return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q; }

constexpr int MAXF = 32;

// tap range helper: for output coordinate o, returns first tap k0 (>=0) with (o*down + k0 - pad0) % up == 0,
// and the matching input coordinate i0; subsequent taps step k += up, i += 1.
__device__ __forceinline__ void tap_start(int o, int down, int up, int pad0, int& k0, int& i0) {
    int t = o * down - pad0;              // position of tap 0 in (unpadded) upsampled coordinates
    int r = ((t % up) + up) % up;         // t mod up, non-negative
    k0 = (up - r) % up;
    i0 = floordiv(t + k0, up);
}

template <class T>
__global__ void __launch_bounds__(256) upfirdn2d_xfast(UpfirdnParams p) {
    typedef typename Acc<T>::t S;
    __shared__ float sf[MAXF * MAXF];
    for (int i = threadIdx.x; i < p.fw * p.fh; i += blockDim.x) {
        int ky = i / p.fw, kx = i % p.fw;
        sf[i] = p.flip ? p.f[ky * p.fw + kx] : p.f[(p.fh - 1 - ky) * p.fw + (p.fw - 1 - kx)];
    }
    __syncthreads();
    const T* x = (const T*)p.x; T* y = (T*)p.y;
    constexpr int STRIP = 4;
    const int stripsW = (p.outW + STRIP - 1) / STRIP;
    const long long total = (long long)p.n * p.c * p.outH * stripsW;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        int sx = (int)(idx % stripsW); long long r = idx / stripsW;
        int oy = (int)(r % p.outH); r /= p.outH;
        int ch = (int)(r % p.c); int nn = (int)(r / p.c);
        const T* xb = x + nn * p.xs_n + ch * p.xs_c;
        int ky0, iy0; tap_start(oy, p.downy, p.upy, p.pady0, ky0, iy0);
        S acc[STRIP];
#pragma unroll
        for (int s = 0; s < STRIP; s++) acc[s] = 0;
        for (int ky = ky0, iy = iy0; ky < p.fh; ky += p.upy, iy++) {
            if (iy < 0 || iy >= p.inH) continue;
            const T* xr = xb + iy * p.xs_h;
            const float* fr = sf + ky * p.fw;
#pragma unroll
            for (int s = 0; s < STRIP; s++) {
                int ox = sx * STRIP + s;
                if (ox >= p.outW) break;
                int kx0, ix0; tap_start(ox, p.downx, p.upx, p.padx0, kx0, ix0);
                S a = 0;
                for (int kx = kx0, ix = ix0; kx < p.fw; kx += p.upx, ix++)
                    if (ix >= 0 && ix < p.inW) a += (S)xr[ix * p.xs_w] * (S)fr[kx];
                acc[s] += a;
            }
        }
        T* yb = y + nn * p.ys_n + ch * p.ys_c + oy * p.ys_h;
#pragma unroll
        for (int s = 0; s < STRIP; s++) {
            int ox = sx * STRIP + s;
            if (ox < p.outW) yb[ox * p.ys_w] = (T)(acc[s] * (S)p.gain);
        }
    }
}

// channels_last, fp32, C % 4 == 0, 16-B aligned
// IDX: 32-bit index arithmetic when the launch fits (the 64-bit div/mod chain per thread made these kernels instruction-bound);
// UP / DOWN: compile-time resampling factors of the two hot cases (x2 up-sampling, x2 down-sampling), 0 = run-time values.
template <typename IDX, int UP, int DOWN>
__global__ void __launch_bounds__(256) upfirdn2d_cfast_f32(UpfirdnParams p) {
    __shared__ float sf[MAXF * MAXF];
    for (int i = threadIdx.x; i < p.fw * p.fh; i += blockDim.x) {
        int ky = i / p.fw, kx = i % p.fw;
        sf[i] = p.flip ? p.f[ky * p.fw + kx] : p.f[(p.fh - 1 - ky) * p.fw + (p.fw - 1 - kx)];
    }
    __syncthreads();
    const float* x = (const float*)p.x; float* y = (float*)p.y;
    const int upx = UP ? UP : p.upx, upy = UP ? UP : p.upy, downx = DOWN ? DOWN : p.downx, downy = DOWN ? DOWN : p.downy;
    const int c4 = p.c / 4;
    const IDX total = (IDX)p.n * p.outH * p.outW * c4;
    for (IDX idx = (IDX)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (IDX)gridDim.x * blockDim.x) {
        int cv = (int)(idx % (IDX)c4); IDX r = idx / (IDX)c4;
        int ox = (int)(r % (IDX)p.outW); r /= (IDX)p.outW;
        int oy = (int)(r % (IDX)p.outH); int nn = (int)(r / (IDX)p.outH);
        int ky0, iy0, kx0, ix0;
        tap_start(oy, downy, upy, p.pady0, ky0, iy0);
        tap_start(ox, downx, upx, p.padx0, kx0, ix0);
        const float* xb = x + nn * p.xs_n + cv * 4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int ky = ky0, iy = iy0; ky < p.fh; ky += upy, iy++) {
            if (iy < 0 || iy >= p.inH) continue;
            for (int kx = kx0, ix = ix0; kx < p.fw; kx += upx, ix++) {
                if (ix < 0 || ix >= p.inW) continue;
                float w = sf[ky * p.fw + kx];
                float4 v = __ldg((const float4*)(xb + iy * p.xs_h + ix * p.xs_w));
                acc.x += v.x * w; acc.y += v.y * w; acc.z += v.z * w; acc.w += v.w * w;
            }
        }
        acc.x *= p.gain; acc.y *= p.gain; acc.z *= p.gain; acc.w *= p.gain;
        *(float4*)(y + nn * p.ys_n + oy * p.ys_h + ox * p.ys_w + cv * 4) = acc;
    }
}

// Hot specialisation: up = down = 1, 4x4 filter, channels_last fp32 (the blur that follows every stride-2 transposed conv,
// conv2d_resample.py:127-128, and its adjoint).  Each thread produces a 4(y) x 2(x) patch of outputs for 4 channels:
// 35 LDG.128 for 8 outputs (4.4 loads/output instead of 16), so the L1/L2 read amplification drops from 16x to ~4x and the
// kernel becomes HBM-bound.  A warp covers 128 consecutive channels of one patch: every load/store is a 512-byte row.
template <bool EPI>
__global__ void __launch_bounds__(256) upfirdn2d_blur4_cl_f32(UpfirdnParams p) {
    __shared__ float sf[16];
    if (threadIdx.x < 16) {
        int ky = threadIdx.x / 4, kx = threadIdx.x % 4;
        sf[threadIdx.x] = (p.flip ? p.f[ky * 4 + kx] : p.f[(3 - ky) * 4 + (3 - kx)]) * p.gain;
    }
    __syncthreads();
    float fl[16];
#pragma unroll
    for (int i = 0; i < 16; i++) fl[i] = sf[i];
    constexpr int TY = 4, TX = 2;
    const float* x = (const float*)p.x; float* y = (float*)p.y;
    const int c4 = p.c / 4;
    const int tilesX = (p.outW + TX - 1) / TX, tilesY = (p.outH + TY - 1) / TY;
    const long long total = (long long)p.n * tilesY * tilesX * c4;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        int cv = (int)(idx % c4); long long r = idx / c4;
        int tx = (int)(r % tilesX); r /= tilesX;
        int ty = (int)(r % tilesY); int nn = (int)(r / tilesY);
        const int ox0 = tx * TX, oy0 = ty * TY;
        const int ix0 = ox0 - p.padx0, iy0 = oy0 - p.pady0;
        const float* xb = x + nn * p.xs_n + cv * 4;
        float4 acc[TY][TX];
#pragma unroll
        for (int a = 0; a < TY; a++)
#pragma unroll
            for (int b = 0; b < TX; b++) acc[a][b] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int ry = 0; ry < TY + 3; ry++) {
            const int iy = iy0 + ry;
            float4 row[TX + 3];
            const bool yin = iy >= 0 && iy < p.inH;
#pragma unroll
            for (int rx = 0; rx < TX + 3; rx++) {
                const int ix = ix0 + rx;
                row[rx] = (yin && ix >= 0 && ix < p.inW) ? __ldg((const float4*)(xb + iy * p.xs_h + ix * p.xs_w)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int a = 0; a < TY; a++) {
                const int ky = ry - a;
                if (ky < 0 || ky > 3) continue;
#pragma unroll
                for (int b = 0; b < TX; b++)
#pragma unroll
                    for (int kx = 0; kx < 4; kx++) {
                        const float w = fl[ky * 4 + kx];
                        const float4 v = row[b + kx];
                        acc[a][b].x = fmaf(v.x, w, acc[a][b].x); acc[a][b].y = fmaf(v.y, w, acc[a][b].y);
                        acc[a][b].z = fmaf(v.z, w, acc[a][b].z); acc[a][b].w = fmaf(v.w, w, acc[a][b].w);
                    }
            }
        }
#pragma unroll
        for (int a = 0; a < TY; a++)
#pragma unroll
            for (int b = 0; b < TX; b++) {
                const int oy = oy0 + a, ox = ox0 + b;
                if (oy < p.outH && ox < p.outW) {
                    float4 o = acc[a][b];
                    if (EPI) o = blur_epilogue<EPI>(p, o, *(const float4*)(p.eb + cv * 4), p.enoise ? __ldg(p.enoise + oy * p.outW + ox) * p.estrength[0] : 0.f);
                    *(float4*)(y + nn * p.ys_n + oy * p.ys_h + ox * p.ys_w + cv * 4) = o;
                }
            }
    }
}

// Strip variant of the kernel above for the large layers: a thread walks TY output rows of a 4-pixel-wide strip, keeping the
// four partially accumulated output rows in registers, so every input row is loaded once per strip: (TY+3)*7 float4 loads for
// TY*4 outputs (2.1 per output at TY = 16, 2.4 at TY = 8; the 4x2 patch kernel needs 4.4) and a 256-thread CTA covers a
// 32 x TY pixel region at C = 128 (unique L2->SM traffic 1.3x the output instead of 2.1x).  Two CTAs per SM (128 registers, a
// few spilled words) measured faster than one CTA with 255 registers and than 2-pixel-wide strips.
template <int TY, int TX, int MINB, bool EPI>
__global__ void __launch_bounds__(256, MINB) upfirdn2d_blur4_strip_cl_f32(UpfirdnParams p) {
    __shared__ float sf[16];
    if (threadIdx.x < 16) {
        int ky = threadIdx.x / 4, kx = threadIdx.x % 4;
        sf[threadIdx.x] = (p.flip ? p.f[ky * 4 + kx] : p.f[(3 - ky) * 4 + (3 - kx)]) * p.gain;
    }
    __syncthreads();
    float fl[16];
#pragma unroll
    for (int i = 0; i < 16; i++) fl[i] = sf[i];
    const float* x = (const float*)p.x; float* y = (float*)p.y;
    const int c4 = p.c / 4;
    const int tilesX = (p.outW + TX - 1) / TX, tilesY = (p.outH + TY - 1) / TY;
    const long long total = (long long)p.n * tilesY * tilesX * c4;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        int cv = (int)(idx % c4); long long r = idx / c4;
        int tx = (int)(r % tilesX); r /= tilesX;
        int ty = (int)(r % tilesY); int nn = (int)(r / tilesY);
        const int ox0 = tx * TX, oy0 = ty * TY;
        const int ix0 = ox0 - p.padx0, iy0 = oy0 - p.pady0;
        const float* xb = x + nn * p.xs_n + cv * 4;
        float* yb = y + nn * p.ys_n + cv * 4;
        float4 ebias = make_float4(0.f, 0.f, 0.f, 0.f);
        float est = 0.f;
        if (EPI) { ebias = *(const float4*)(p.eb + cv * 4); est = p.enoise ? p.estrength[0] : 0.f; }
        float4 acc[4][TX];
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int b = 0; b < TX; b++) acc[a][b] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int ry = 0; ry < TY + 3; ry++) {
            const int iy = iy0 + ry;
            float4 row[TX + 3];
            const bool yin = iy >= 0 && iy < p.inH;
#pragma unroll
            for (int rx = 0; rx < TX + 3; rx++) {
                const int ix = ix0 + rx;
                row[rx] = (yin && ix >= 0 && ix < p.inW) ? __ldg((const float4*)(xb + iy * p.xs_h + ix * p.xs_w)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int ky = 0; ky < 4; ky++) {
                const int a = ry - ky;                       // output row (within the strip) this input row feeds with filter row ky
                if (a < 0 || a >= TY) continue;
#pragma unroll
                for (int b = 0; b < TX; b++)
#pragma unroll
                    for (int kx = 0; kx < 4; kx++) {
                        const float w = fl[ky * 4 + kx];
                        const float4 v = row[b + kx];
                        float4& t = acc[a & 3][b];
                        t.x = fmaf(v.x, w, t.x); t.y = fmaf(v.y, w, t.y); t.z = fmaf(v.z, w, t.z); t.w = fmaf(v.w, w, t.w);
                    }
            }
            if (ry >= 3) {                                   // output row ry - 3 has seen its four input rows
                const int a = ry - 3, oy = oy0 + a;
#pragma unroll
                for (int b = 0; b < TX; b++) {
                    const int ox = ox0 + b;
                    if (oy < p.outH && ox < p.outW)
                        *(float4*)(yb + oy * p.ys_h + ox * p.ys_w) =
                            blur_epilogue<EPI>(p, acc[a & 3][b], ebias, (EPI && p.enoise) ? __ldg(p.enoise + oy * p.outW + ox) * est : 0.f);
                    acc[a & 3][b] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        }
    }
}

// channel-fastest scalar fallback (any dtype / C)
template <class T, typename IDX, int UP, int DOWN>
__global__ void __launch_bounds__(256) upfirdn2d_cfast(UpfirdnParams p) {
    typedef typename Acc<T>::t S;
    __shared__ float sf[MAXF * MAXF];
    for (int i = threadIdx.x; i < p.fw * p.fh; i += blockDim.x) {
        int ky = i / p.fw, kx = i % p.fw;
        sf[i] = p.flip ? p.f[ky * p.fw + kx] : p.f[(p.fh - 1 - ky) * p.fw + (p.fw - 1 - kx)];
    }
    __syncthreads();
    const T* x = (const T*)p.x; T* y = (T*)p.y;
    const int upx = UP ? UP : p.upx, upy = UP ? UP : p.upy, downx = DOWN ? DOWN : p.downx, downy = DOWN ? DOWN : p.downy;
    const IDX total = (IDX)p.n * p.outH * p.outW * p.c;
    for (IDX idx = (IDX)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (IDX)gridDim.x * blockDim.x) {
        int ch = (int)(idx % (IDX)p.c); IDX r = idx / (IDX)p.c;
        int ox = (int)(r % (IDX)p.outW); r /= (IDX)p.outW;
        int oy = (int)(r % (IDX)p.outH); int nn = (int)(r / (IDX)p.outH);
        int ky0, iy0, kx0, ix0;
        tap_start(oy, downy, upy, p.pady0, ky0, iy0);
        tap_start(ox, downx, upx, p.padx0, kx0, ix0);
        const T* xb = x + nn * p.xs_n + ch * p.xs_c;
        S acc = 0;
        for (int ky = ky0, iy = iy0; ky < p.fh; ky += upy, iy++) {
            if (iy < 0 || iy >= p.inH) continue;
            for (int kx = kx0, ix = ix0; kx < p.fw; kx += upx, ix++)
                if (ix >= 0 && ix < p.inW) acc += (S)xb[iy * p.xs_h + ix * p.xs_w] * (S)sf[ky * p.fw + kx];
        }
        y[nn * p.ys_n + ch * p.ys_c + oy * p.ys_h + ox * p.ys_w] = (T)(acc * (S)p.gain);
    }
}

}  // namespace

namespace {
struct BlurEpilogue { const float* b; const float* noise; const float* strength; float alpha, gain, clamp; int act; };
}

static int upfirdn2d_impl(const void* x, const float* f, void* y, int dtype, int n, int c, int in_h, int in_w,
                          const long long* x_strides, const long long* y_strides, int fh, int fw, int upx, int upy,
                          int downx, int downy, int padx0, int padx1, int pady0, int pady1, int flip, float gain,
                          const BlurEpilogue* epi, cudaStream_t stream) {
    SPI_CHECK_ARG(x && f && y, "upfirdn2d: null pointer");
    SPI_CHECK_ARG(fw >= 1 && fh >= 1, "upfirdn2d: f must be at least 1x1");
    SPI_CHECK_ARG(fw <= MAXF && fh <= MAXF, "upfirdn2d: filter larger than %dx%d is not supported", MAXF, MAXF);
    SPI_CHECK_ARG(upx >= 1 && upy >= 1, "upfirdn2d: upsampling factor must be at least 1");
    SPI_CHECK_ARG(downx >= 1 && downy >= 1, "upfirdn2d: downsampling factor must be at least 1");
    int out_w = (in_w * upx + padx0 + padx1 - fw + downx) / downx;
    int out_h = (in_h * upy + pady0 + pady1 - fh + downy) / downy;
    SPI_CHECK_ARG(out_w >= 1 && out_h >= 1, "upfirdn2d: output must be at least 1x1");
    if (n == 0 || c == 0) return SPI_OK;
    UpfirdnParams p;
    p.x = x; p.f = f; p.y = y;
    p.upx = upx; p.upy = upy; p.downx = downx; p.downy = downy; p.padx0 = padx0; p.pady0 = pady0; p.flip = flip;
    p.gain = gain; p.n = n; p.c = c; p.inH = in_h; p.inW = in_w; p.outH = out_h; p.outW = out_w; p.fw = fw; p.fh = fh;
    p.xs_n = x_strides[0]; p.xs_c = x_strides[1]; p.xs_h = x_strides[2]; p.xs_w = x_strides[3];
    p.ys_n = y_strides[0]; p.ys_c = y_strides[1]; p.ys_h = y_strides[2]; p.ys_w = y_strides[3];
    p.eb = nullptr; p.enoise = nullptr; p.estrength = nullptr; p.ealpha = 0.f; p.egain = 1.f; p.eclamp = -1.f; p.eact = 0;
    if (epi) { p.eb = epi->b; p.enoise = epi->noise; p.estrength = epi->strength; p.ealpha = epi->alpha; p.egain = epi->gain; p.eclamp = epi->clamp; p.eact = epi->act; }
    const bool cfast = (c > 1 && p.xs_c == 1 && p.ys_c == 1);
    const int block = 256;
    const long long cap = (long long)spi_num_sms() * 8;
    auto grid_for = [&](long long items) {
        long long b = (items + block - 1) / block;
        return (int)(b < 1 ? 1 : (b > cap ? cap : b));
    };
    long long outs = (long long)n * c * out_h * out_w;
    const bool small = outs < (1LL << 31) - (1LL << 24);          // 32-bit index arithmetic (grid stride < 2^22 threads)
    const bool up2 = upx == 2 && upy == 2 && downx == 1 && downy == 1, down2 = upx == 1 && upy == 1 && downx == 2 && downy == 2;
    if (cfast) {
        bool v4 = dtype == SPI_DT_F32 && (c % 4 == 0) && (((uintptr_t)x | (uintptr_t)y) % 16 == 0) &&
                  (p.xs_n % 4 == 0) && (p.xs_h % 4 == 0) && (p.xs_w % 4 == 0) && (p.ys_n % 4 == 0) && (p.ys_h % 4 == 0) && (p.ys_w % 4 == 0);
        const bool blur4 = v4 && upx == 1 && upy == 1 && downx == 1 && downy == 1 && fw == 4 && fh == 4;
        auto strips = [&](int ty) { return (long long)n * ((out_h + ty - 1) / ty) * ((out_w + 3) / 4) * (c / 4); };
        if (epi && !blur4) {
            spi_set_error("blur4_bias_act_noise: needs channels-last float32 with C %% 4 == 0, 16-byte aligned, and a 4x4 filter at up = down = 1");
            return SPI_ERR_ARG;
        }
        // large layers: 16-row strips while they still give every SM two CTAs, else 8-row strips (measured on B200: 84 -> 72 us at
        // [1,128,513,513], 307 -> 227 us at [4,128,513,513], 45 -> 41 us at [1,256,257,257]; profiles/r1_bench_stream_kernels.txt)
        const bool big = blur4 && out_h >= 64;
        const long long patches = (long long)n * ((out_h + 3) / 4) * ((out_w + 1) / 2) * (c / 4);
        if (big && !epi && strips(16) >= 148LL * 512) {
            // (with the epilogue the 16-row body schedules badly under the 128-register cap: 186 us vs 70 us at [1,128,513,513];
            // the 8-row strips keep their speed, so the fused variant always uses them)
            upfirdn2d_blur4_strip_cl_f32<16, 4, 2, false><<<grid_for(strips(16)), block, 0, stream>>>(p);
        } else if (big) {
            if (epi) upfirdn2d_blur4_strip_cl_f32<8, 4, 2, true><<<grid_for(strips(8)), block, 0, stream>>>(p);
            else upfirdn2d_blur4_strip_cl_f32<8, 4, 2, false><<<grid_for(strips(8)), block, 0, stream>>>(p);
        } else if (blur4) {
            if (epi) upfirdn2d_blur4_cl_f32<true><<<grid_for(patches), block, 0, stream>>>(p);
            else upfirdn2d_blur4_cl_f32<false><<<grid_for(patches), block, 0, stream>>>(p);
        }
        else if (v4) {
            if (!small) upfirdn2d_cfast_f32<long long, 0, 0><<<grid_for(outs / 4), block, 0, stream>>>(p);
            else if (up2) upfirdn2d_cfast_f32<unsigned, 2, 1><<<grid_for(outs / 4), block, 0, stream>>>(p);
            else if (down2) upfirdn2d_cfast_f32<unsigned, 1, 2><<<grid_for(outs / 4), block, 0, stream>>>(p);
            else upfirdn2d_cfast_f32<unsigned, 0, 0><<<grid_for(outs / 4), block, 0, stream>>>(p);
        } else if (dtype == SPI_DT_F32) {
            if (!small) upfirdn2d_cfast<float, long long, 0, 0><<<grid_for(outs), block, 0, stream>>>(p);
            else if (up2) upfirdn2d_cfast<float, unsigned, 2, 1><<<grid_for(outs), block, 0, stream>>>(p);
            else if (down2) upfirdn2d_cfast<float, unsigned, 1, 2><<<grid_for(outs), block, 0, stream>>>(p);
            else upfirdn2d_cfast<float, unsigned, 0, 0><<<grid_for(outs), block, 0, stream>>>(p);
        }
        else if (dtype == SPI_DT_F16) upfirdn2d_cfast<__half, long long, 0, 0><<<grid_for(outs), block, 0, stream>>>(p);
        else if (dtype == SPI_DT_F64) upfirdn2d_cfast<double, long long, 0, 0><<<grid_for(outs), block, 0, stream>>>(p);
        else { spi_set_error("upfirdn2d: unsupported dtype %d", dtype); return SPI_ERR_ARG; }
    } else {
        if (epi) { spi_set_error("blur4_bias_act_noise: x and y must be channels-last"); return SPI_ERR_ARG; }
        long long items = (long long)n * c * out_h * ((out_w + 3) / 4);
        if (dtype == SPI_DT_F32) upfirdn2d_xfast<float><<<grid_for(items), block, 0, stream>>>(p);
        else if (dtype == SPI_DT_F16) upfirdn2d_xfast<__half><<<grid_for(items), block, 0, stream>>>(p);
        else if (dtype == SPI_DT_F64) upfirdn2d_xfast<double><<<grid_for(items), block, 0, stream>>>(p);
        else { spi_set_error("upfirdn2d: unsupported dtype %d", dtype); return SPI_ERR_ARG; }
    }
    SPI_COUNT_LAUNCH(1);
    SPI_LAUNCH_CHECK("upfirdn2d");
    return SPI_OK;
}

extern "C" int spi_upfirdn2d(const void* x, const float* f, void* y, int dtype, int n, int c, int in_h, int in_w,
                             const long long* x_strides, const long long* y_strides, int fh, int fw, int upx, int upy,
                             int downx, int downy, int padx0, int padx1, int pady0, int pady1, int flip, float gain,
                             cudaStream_t stream) {
    return upfirdn2d_impl(x, f, y, dtype, n, c, in_h, in_w, x_strides, y_strides, fh, fw, upx, upy, downx, downy, padx0, padx1, pady0,
                          pady1, flip, gain, nullptr, stream);
}

// The tail of an up-sampling SynthesisLayer in one pass (networks_stylegan2.py:320-329 after conv2d_resample.py:117-119):
//   y = clamp(act(upfirdn2d(x, f[4x4], pad, fir_gain) + noise[h,w] * noise_strength + b[c]) * gain)
// x, y channels-last float32, C % 4 == 0; act 1 (linear) or 3 (lrelu); noise may be NULL; clamp < 0 = off.  Same arithmetic, in the
// same order, as spi_upfirdn2d followed by spi_bias_act_noise, without the intermediate tensor.
extern "C" int spi_blur4_bias_act_noise(const float* x, const float* f, float* y, const float* b, const float* noise,
                                        const float* noise_strength, int n, int c, int in_h, int in_w, const long long* x_strides,
                                        const long long* y_strides, int padx0, int padx1, int pady0, int pady1, int flip, float fir_gain,
                                        int act, float alpha, float gain, float clamp, cudaStream_t stream) {
    SPI_CHECK_ARG(b != nullptr, "blur4_bias_act_noise: bias required");
    SPI_CHECK_ARG(act == 1 || act == 3, "blur4_bias_act_noise: act must be 1 (linear) or 3 (lrelu)");
    SPI_CHECK_ARG(!noise || noise_strength, "blur4_bias_act_noise: noise_strength required with a noise map");
    BlurEpilogue epi{b, noise, noise_strength, alpha, gain, clamp, act};
    return upfirdn2d_impl(x, f, y, SPI_DT_F32, n, c, in_h, in_w, x_strides, y_strides, 4, 4, 1, 1, 1, 1, padx0, padx1, pady0, pady1, flip,
                          fir_gain, &epi, stream);
}
