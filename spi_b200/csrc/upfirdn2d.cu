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
};

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
__global__ void __launch_bounds__(256) upfirdn2d_cfast_f32(UpfirdnParams p) {
    __shared__ float sf[MAXF * MAXF];
    for (int i = threadIdx.x; i < p.fw * p.fh; i += blockDim.x) {
        int ky = i / p.fw, kx = i % p.fw;
        sf[i] = p.flip ? p.f[ky * p.fw + kx] : p.f[(p.fh - 1 - ky) * p.fw + (p.fw - 1 - kx)];
    }
    __syncthreads();
    const float* x = (const float*)p.x; float* y = (float*)p.y;
    const int c4 = p.c / 4;
    const long long total = (long long)p.n * p.outH * p.outW * c4;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        int cv = (int)(idx % c4); long long r = idx / c4;
        int ox = (int)(r % p.outW); r /= p.outW;
        int oy = (int)(r % p.outH); int nn = (int)(r / p.outH);
        int ky0, iy0, kx0, ix0;
        tap_start(oy, p.downy, p.upy, p.pady0, ky0, iy0);
        tap_start(ox, p.downx, p.upx, p.padx0, kx0, ix0);
        const float* xb = x + nn * p.xs_n + cv * 4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int ky = ky0, iy = iy0; ky < p.fh; ky += p.upy, iy++) {
            if (iy < 0 || iy >= p.inH) continue;
            for (int kx = kx0, ix = ix0; kx < p.fw; kx += p.upx, ix++) {
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
                if (oy < p.outH && ox < p.outW) *(float4*)(y + nn * p.ys_n + oy * p.ys_h + ox * p.ys_w + cv * 4) = acc[a][b];
            }
    }
}

// channel-fastest scalar fallback (any dtype / C)
template <class T>
__global__ void __launch_bounds__(256) upfirdn2d_cfast(UpfirdnParams p) {
    typedef typename Acc<T>::t S;
    __shared__ float sf[MAXF * MAXF];
    for (int i = threadIdx.x; i < p.fw * p.fh; i += blockDim.x) {
        int ky = i / p.fw, kx = i % p.fw;
        sf[i] = p.flip ? p.f[ky * p.fw + kx] : p.f[(p.fh - 1 - ky) * p.fw + (p.fw - 1 - kx)];
    }
    __syncthreads();
    const T* x = (const T*)p.x; T* y = (T*)p.y;
    const long long total = (long long)p.n * p.outH * p.outW * p.c;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        int ch = (int)(idx % p.c); long long r = idx / p.c;
        int ox = (int)(r % p.outW); r /= p.outW;
        int oy = (int)(r % p.outH); int nn = (int)(r / p.outH);
        int ky0, iy0, kx0, ix0;
        tap_start(oy, p.downy, p.upy, p.pady0, ky0, iy0);
        tap_start(ox, p.downx, p.upx, p.padx0, kx0, ix0);
        const T* xb = x + nn * p.xs_n + ch * p.xs_c;
        S acc = 0;
        for (int ky = ky0, iy = iy0; ky < p.fh; ky += p.upy, iy++) {
            if (iy < 0 || iy >= p.inH) continue;
            for (int kx = kx0, ix = ix0; kx < p.fw; kx += p.upx, ix++)
                if (ix >= 0 && ix < p.inW) acc += (S)xb[iy * p.xs_h + ix * p.xs_w] * (S)sf[ky * p.fw + kx];
        }
        y[nn * p.ys_n + ch * p.ys_c + oy * p.ys_h + ox * p.ys_w] = (T)(acc * (S)p.gain);
    }
}

}  // namespace

extern "C" int spi_upfirdn2d(const void* x, const float* f, void* y, int dtype, int n, int c, int in_h, int in_w,
                             const long long* x_strides, const long long* y_strides, int fh, int fw, int upx, int upy,
                             int downx, int downy, int padx0, int padx1, int pady0, int pady1, int flip, float gain,
                             cudaStream_t stream) {
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
    const bool cfast = (c > 1 && p.xs_c == 1 && p.ys_c == 1);
    const int block = 256;
    const long long cap = (long long)spi_num_sms() * 8;
    auto grid_for = [&](long long items) {
        long long b = (items + block - 1) / block;
        return (int)(b < 1 ? 1 : (b > cap ? cap : b));
    };
    long long outs = (long long)n * c * out_h * out_w;
    if (cfast) {
        bool v4 = dtype == SPI_DT_F32 && (c % 4 == 0) && (((uintptr_t)x | (uintptr_t)y) % 16 == 0) &&
                  (p.xs_n % 4 == 0) && (p.xs_h % 4 == 0) && (p.xs_w % 4 == 0) && (p.ys_n % 4 == 0) && (p.ys_h % 4 == 0) && (p.ys_w % 4 == 0);
        const bool blur4 = v4 && upx == 1 && upy == 1 && downx == 1 && downy == 1 && fw == 4 && fh == 4;
        if (blur4) upfirdn2d_blur4_cl_f32<<<grid_for((long long)n * ((out_h + 3) / 4) * ((out_w + 1) / 2) * (c / 4)), block, 0, stream>>>(p);
        else if (v4) upfirdn2d_cfast_f32<<<grid_for(outs / 4), block, 0, stream>>>(p);
        else if (dtype == SPI_DT_F32) upfirdn2d_cfast<float><<<grid_for(outs), block, 0, stream>>>(p);
        else if (dtype == SPI_DT_F16) upfirdn2d_cfast<__half><<<grid_for(outs), block, 0, stream>>>(p);
        else if (dtype == SPI_DT_F64) upfirdn2d_cfast<double><<<grid_for(outs), block, 0, stream>>>(p);
        else { spi_set_error("upfirdn2d: unsupported dtype %d", dtype); return SPI_ERR_ARG; }
    } else {
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
