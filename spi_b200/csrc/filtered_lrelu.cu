// filtered_lrelu for sm_100a: bias -> up-FIR -> gain*lrelu*clamp (+2-bit sign/clamp mask) -> down-FIR, fused.
//
// Replaces the reference plugin ops `filtered_lrelu` / `filtered_lrelu_act_`
// (eg3d/torch_utils/ops/filtered_lrelu.cpp:20-213,217-293; kernels filtered_lrelu.cu:144-1100,1110-1230) with the same
// argument meaning and the same sign-tensor format: uint8 [N, C, sh, sw/4], element (x, y) lives in byte x>>2 at
// bits 2*(x&3); bit0 = negative (slope applied), bit1 = clamped; sh = yh*down-(down-1)+fdh-1, sw = that width
// rounded up to 16 (filtered_lrelu.cpp:78-86).  Read mode applies offsets (sx, sy) and ignores positions that
// fall outside the mask (filtered_lrelu.cu:1178-1192).
//
// B200 design: one CTA = one 16x16 output tile of one (n, c) image.  The bias-added input footprint is staged in
// shared memory, the up-filtered + activated intermediate tile is produced into shared memory (only non-zero
// polyphase taps are visited) and the down filter reads it from there -- one HBM read of x, one write of y, no
// intermediate tensor.  Filters live in shared memory (NOT in a global __constant__ buffer as in the reference, which
// makes that implementation unsafe on concurrent streams -- filtered_lrelu.py:217-218); this one is stream-safe.
// 1-D (separable) filters are expanded to their outer product by the host shim.
#include "common.cuh"

namespace {

struct FlrParams {
    const void* x; void* y; const void* b; unsigned char* s; const float* fu; const float* fd;
    int up, down, px0, py0, fuw, fuh, fdw, fdh, flip;
    float gain, slope, clamp;
    int n, c, xh, xw, ch, cw, yh, yw;            // ch/cw: size of the up-filtered intermediate
    long long xs_n, xs_c, xs_h, xs_w, ys_n, ys_c, ys_h, ys_w;
    int s_w, s_h, sx, sy;                         // sign tensor width (elements) / height, read offsets
    int mode;                                     // 0 none, 1 write signs, 2 read signs
    int tin_w, tin_h, tmid_w, tmid_h;             // shared tile extents
};

constexpr int TO = 16;  // output tile edge

__device__ __forceinline__ int fdiv(int a, int b) { int q = a / b; return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q; }

template <class T>
__global__ void __launch_bounds__(256) filtered_lrelu_kernel(FlrParams p) {
    extern __shared__ float smem[];
    float* s_fu = smem;                                   // fuh*fuw (already flipped for correlation)
    float* s_fd = s_fu + p.fuh * p.fuw;                   // fdh*fdw
    float* s_in = s_fd + p.fdh * p.fdw;                   // tin_h*tin_w
    float* s_mid = s_in + p.tin_h * p.tin_w;              // tmid_h*tmid_w
    unsigned char* s_sg = (unsigned char*)(s_mid + p.tmid_h * p.tmid_w);  // tmid_h*tmid_w sign codes
    const int tid = threadIdx.x;
    const int nc = blockIdx.z;
    const int nn = nc / p.c, cc = nc % p.c;
    const int oy0 = blockIdx.y * TO, ox0 = blockIdx.x * TO;
    for (int i = tid; i < p.fuh * p.fuw; i += blockDim.x) {
        int ky = i / p.fuw, kx = i % p.fuw;
        s_fu[i] = p.flip ? p.fu[i] : p.fu[(p.fuh - 1 - ky) * p.fuw + (p.fuw - 1 - kx)];
    }
    for (int i = tid; i < p.fdh * p.fdw; i += blockDim.x) {
        int ky = i / p.fdw, kx = i % p.fdw;
        s_fd[i] = p.flip ? p.fd[i] : p.fd[(p.fdh - 1 - ky) * p.fdw + (p.fdw - 1 - kx)];
    }
    // intermediate tile origin (in up-filtered coordinates) and the input footprint origin
    const int ty0 = oy0 * p.down, tx0 = ox0 * p.down;
    const int iy0 = fdiv(ty0 - p.py0, p.up), ix0 = fdiv(tx0 - p.px0, p.up);
    const T* xb = (const T*)p.x + nn * p.xs_n + cc * p.xs_c;
    const float bias = p.b ? (float)((const T*)p.b)[cc] : 0.f;
    for (int i = tid; i < p.tin_h * p.tin_w; i += blockDim.x) {
        int iy = iy0 + i / p.tin_w, ix = ix0 + i % p.tin_w;
        float v = 0.f;
        if (iy >= 0 && iy < p.xh && ix >= 0 && ix < p.xw) v = (float)xb[iy * p.xs_h + ix * p.xs_w] + bias;
        s_in[i] = v;
    }
    __syncthreads();
    // up-FIR + activation into the intermediate tile
    const float up2 = (float)(p.up * p.up);
    const size_t s_plane = (size_t)(p.s_w >> 2) * p.s_h;
    for (int i = tid; i < p.tmid_h * p.tmid_w; i += blockDim.x) {
        int my = i / p.tmid_w, mx = i % p.tmid_w;
        int ty = ty0 + my, tx = tx0 + mx;
        float v = 0.f;
        unsigned char code = 0;
        if (ty < p.ch && tx < p.cw) {
            // taps: upsampled coordinate t = ty + ky - py0 must be a multiple of up
            int t = ty - p.py0; int r = ((t % p.up) + p.up) % p.up; int ky0 = (p.up - r) % p.up; int iy = fdiv(t + ky0, p.up) - iy0;
            int u = tx - p.px0; int q = ((u % p.up) + p.up) % p.up; int kx0 = (p.up - q) % p.up; int ixs = fdiv(u + kx0, p.up) - ix0;
            for (int ky = ky0; ky < p.fuh; ky += p.up, iy++) {
                const float* fr = s_fu + ky * p.fuw;
                const float* ir = s_in + iy * p.tin_w;
                int ix = ixs;
                for (int kx = kx0; kx < p.fuw; kx += p.up, ix++) v += ir[ix] * fr[kx];
            }
            v *= up2 * p.gain;
            if (p.mode == 2) {
                unsigned int sx = (unsigned int)(tx + p.sx), sy = (unsigned int)(ty + p.sy);
                if (sx < (unsigned int)p.s_w && sy < (unsigned int)p.s_h) {
                    unsigned char sb = p.s[s_plane * nc + (size_t)(p.s_w >> 2) * sy + (sx >> 2)];
                    sb >>= (sx & 3) << 1;
                    if (sb & 1) v *= p.slope;
                    if (sb & 2) v = 0.f;
                }
            } else {
                if (v < 0.f) { v *= p.slope; code = 1; }
                if (fabsf(v) > p.clamp) { v = v < 0.f ? -p.clamp : p.clamp; code = 2; }
            }
        }
        s_mid[i] = v;
        s_sg[i] = code;
    }
    __syncthreads();
    // sign write: this CTA owns intermediate rows/cols [t0, t0 + TO*down); the last tile also owns the filter tail
    if (p.mode == 1) {
        const int own_h = (blockIdx.y == gridDim.y - 1) ? p.tmid_h : TO * p.down;
        const int own_w = (blockIdx.x == gridDim.x - 1) ? p.tmid_w : TO * p.down;   // TO*down is a multiple of 4
        const int bytes_w = (own_w + 3) >> 2;
        for (int i = tid; i < own_h * bytes_w; i += blockDim.x) {
            int my = i / bytes_w, bx = i % bytes_w;
            int sy = ty0 + my, sx = tx0 + bx * 4;
            if (sy >= p.s_h || sx >= p.s_w) continue;
            unsigned int v = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                int mx = bx * 4 + k;
                if (mx < p.tmid_w) v |= (unsigned int)s_sg[my * p.tmid_w + mx] << (2 * k);
            }
            p.s[s_plane * nc + (size_t)(p.s_w >> 2) * sy + (sx >> 2)] = (unsigned char)v;
        }
    }
    // down-FIR
    T* yb = (T*)p.y + nn * p.ys_n + cc * p.ys_c;
    for (int i = tid; i < TO * TO; i += blockDim.x) {
        int ly = i / TO, lx = i % TO;
        int oy = oy0 + ly, ox = ox0 + lx;
        if (oy >= p.yh || ox >= p.yw) continue;
        float acc = 0.f;
        const float* mr = s_mid + (ly * p.down) * p.tmid_w + lx * p.down;
        for (int ky = 0; ky < p.fdh; ky++)
            for (int kx = 0; kx < p.fdw; kx++) acc += mr[ky * p.tmid_w + kx] * s_fd[ky * p.fdw + kx];
        yb[oy * p.ys_h + ox * p.ys_w] = (T)acc;
    }
}

struct ActParams {
    void* x; unsigned char* s;
    int w, h, c, n;
    long long xs_n, xs_c, xs_h, xs_w;
    int s_w, s_h, sx, sy, mode;
    float gain, slope, clamp;
};

// in-place gain*lrelu*clamp with sign write/read: the generic-fallback op `filtered_lrelu_act_`
template <class T>
__global__ void __launch_bounds__(256) filtered_lrelu_act_kernel(ActParams p) {
    const int ymax = (p.mode == 1) ? p.s_h : p.h;
    const int wq = (p.mode == 1) ? (p.s_w >> 2) : ((p.w + 3) >> 2);   // work item = 4 consecutive x
    const long long total = (long long)p.n * p.c * ymax * wq;
    const size_t s_plane = (size_t)(p.s_w >> 2) * p.s_h;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        int xq = (int)(idx % wq); long long r = idx / wq;
        int y = (int)(r % ymax); int q = (int)(r / ymax);
        int nn = q / p.c, cc = q % p.c;
        T* xb = (T*)p.x + nn * p.xs_n + cc * p.xs_c + (long long)y * p.xs_h;
        unsigned int packed = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            int x = xq * 4 + k;
            if (x >= p.w || y >= p.h) continue;
            float v = (float)xb[x * p.xs_w] * p.gain;
            if (p.mode == 2) {
                unsigned int sx = (unsigned int)(x + p.sx), sy = (unsigned int)(y + p.sy);
                if (sx < (unsigned int)p.s_w && sy < (unsigned int)p.s_h) {
                    unsigned char sb = p.s[s_plane * q + (size_t)(p.s_w >> 2) * sy + (sx >> 2)];
                    sb >>= (sx & 3) << 1;
                    if (sb & 1) v *= p.slope;
                    if (sb & 2) v = 0.f;
                }
            } else {
                unsigned int code = 0;
                if (v < 0.f) { v *= p.slope; code = 1; }
                if (fabsf(v) > p.clamp) { v = v < 0.f ? -p.clamp : p.clamp; code = 2; }
                packed |= code << (2 * k);
            }
            xb[x * p.xs_w] = (T)v;
        }
        if (p.mode == 1) p.s[s_plane * q + (size_t)(p.s_w >> 2) * y + xq] = (unsigned char)packed;
    }
}

}  // namespace

extern "C" int spi_filtered_lrelu_sign_shape(int yh, int yw, int down, int fdh, int fdw, int* sh, int* sw_bytes) {
    int sw_active = yw * down - (down - 1) + (fdw - 1);
    *sh = yh * down - (down - 1) + (fdh - 1);
    *sw_bytes = ((sw_active + 15) & ~15) >> 2;
    return SPI_OK;
}

extern "C" int spi_filtered_lrelu(const void* x, void* y, const void* b, unsigned char* s, const float* fu, const float* fd,
                                  int dtype, int n, int c, int xh, int xw, const long long* x_strides,
                                  const long long* y_strides, int fuh, int fuw, int fdh, int fdw, int up, int down, int px0,
                                  int px1, int py0, int py1, int s_h, int s_w_bytes, int sx, int sy, float gain, float slope,
                                  float clamp, int flip, int sign_mode, cudaStream_t stream) {
    SPI_CHECK_ARG(x && y && fu && fd, "filtered_lrelu: null pointer");
    SPI_CHECK_ARG(dtype == SPI_DT_F32 || dtype == SPI_DT_F16, "filtered_lrelu: x and b must be float16 or float32");
    SPI_CHECK_ARG(up >= 1 && down >= 1, "filtered_lrelu: up and down must be at least 1");
    SPI_CHECK_ARG(n > 0 && c > 0 && xh > 0 && xw > 0, "filtered_lrelu: x is empty");
    int cw = xw * up + (px0 + px1) - (fuw - 1);
    int ch = xh * up + (py0 + py1) - (fuh - 1);
    SPI_CHECK_ARG(cw > fdw - 1 && ch > fdh - 1, "filtered_lrelu: upsampled buffer must be at least the size of downsampling filter");
    int yw = (cw - (fdw - 1) + (down - 1)) / down;
    int yh = (ch - (fdh - 1) + (down - 1)) / down;
    SPI_CHECK_ARG(yw > 0 && yh > 0, "filtered_lrelu: output must be at least 1x1");
    SPI_CHECK_ARG(sign_mode == 0 || s, "filtered_lrelu: sign tensor required");
    FlrParams p;
    p.x = x; p.y = y; p.b = b; p.s = s; p.fu = fu; p.fd = fd;
    p.up = up; p.down = down; p.px0 = px0; p.py0 = py0; p.fuw = fuw; p.fuh = fuh; p.fdw = fdw; p.fdh = fdh; p.flip = flip;
    p.gain = gain; p.slope = slope; p.clamp = clamp;
    p.n = n; p.c = c; p.xh = xh; p.xw = xw; p.ch = ch; p.cw = cw; p.yh = yh; p.yw = yw;
    p.xs_n = x_strides[0]; p.xs_c = x_strides[1]; p.xs_h = x_strides[2]; p.xs_w = x_strides[3];
    p.ys_n = y_strides[0]; p.ys_c = y_strides[1]; p.ys_h = y_strides[2]; p.ys_w = y_strides[3];
    p.s_w = s_w_bytes * 4; p.s_h = s_h; p.sx = sx; p.sy = sy; p.mode = sign_mode;
    p.tmid_h = (TO - 1) * down + fdh; p.tmid_w = (TO - 1) * down + fdw;
    p.tin_h = (p.tmid_h + fuh - 1) / up + 2; p.tin_w = (p.tmid_w + fuw - 1) / up + 2;
    size_t smem = sizeof(float) * ((size_t)fuh * fuw + (size_t)fdh * fdw + (size_t)p.tin_h * p.tin_w + (size_t)p.tmid_h * p.tmid_w) +
                  (size_t)p.tmid_h * p.tmid_w;
    if (smem > 220 * 1024) return SPI_ERR_UNSUPPORTED;   // -> python falls back to upfirdn2d + act (reference return_code -1)
    dim3 grid(cdiv(yw, TO), cdiv(yh, TO), n * c);
    if (grid.z > 65535u * 32u) return SPI_ERR_UNSUPPORTED;
    void* kern = dtype == SPI_DT_F32 ? (void*)filtered_lrelu_kernel<float> : (void*)filtered_lrelu_kernel<__half>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    // gridDim.z limit is 65535: fold the overflow into multiple launches
    for (unsigned int z0 = 0; z0 < grid.z; z0 += 65535u) {
        unsigned int gz = grid.z - z0 < 65535u ? grid.z - z0 : 65535u;
        FlrParams q = p;
        if (z0) {
            // advance base pointers by z0 images: (n, c) index = z0 + blockIdx.z -> handled by offsetting n*c counting
            // (rare path: > 65535 channel images)
            return SPI_ERR_UNSUPPORTED;
        }
        void* args[] = {&q};
        cudaError_t e = cudaLaunchKernel(kern, dim3(grid.x, grid.y, gz), dim3(256), args, smem, stream);
        SPI_COUNT_LAUNCH(1);
        if (e != cudaSuccess) { spi_set_error("filtered_lrelu: %s", cudaGetErrorString(e)); return SPI_ERR_CUDA; }
    }
    return SPI_OK;
}

extern "C" int spi_filtered_lrelu_act(void* x, unsigned char* s, int dtype, int n, int c, int h, int w,
                                      const long long* x_strides, int s_h, int s_w_bytes, int sx, int sy, float gain,
                                      float slope, float clamp, int sign_mode, cudaStream_t stream) {
    SPI_CHECK_ARG(x, "filtered_lrelu_act_: null pointer");
    SPI_CHECK_ARG(dtype == SPI_DT_F32 || dtype == SPI_DT_F16, "filtered_lrelu_act_: x must be float16 or float32");
    SPI_CHECK_ARG(sign_mode == 0 || s, "filtered_lrelu_act_: sign tensor required");
    ActParams p;
    p.x = x; p.s = s; p.w = w; p.h = h; p.c = c; p.n = n;
    p.xs_n = x_strides[0]; p.xs_c = x_strides[1]; p.xs_h = x_strides[2]; p.xs_w = x_strides[3];
    p.s_w = s_w_bytes * 4; p.s_h = s_h; p.sx = sx; p.sy = sy; p.mode = sign_mode;
    p.gain = gain; p.slope = slope; p.clamp = clamp;
    int ymax = sign_mode == 1 ? s_h : h;
    int wq = sign_mode == 1 ? s_w_bytes : (w + 3) / 4;
    long long total = (long long)n * c * ymax * wq;
    if (total == 0) return SPI_OK;
    long long cap = (long long)spi_num_sms() * 8;
    long long blocks = (total + 255) / 256;
    int grid = (int)(blocks > cap ? cap : blocks);
    if (dtype == SPI_DT_F32) filtered_lrelu_act_kernel<float><<<grid, 256, 0, stream>>>(p);
    else filtered_lrelu_act_kernel<__half><<<grid, 256, 0, stream>>>(p);
    SPI_COUNT_LAUNCH(1);
    SPI_LAUNCH_CHECK("filtered_lrelu_act_");
    return SPI_OK;
}
