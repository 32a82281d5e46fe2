// Fused Adam and small streaming helpers for sm_100a.
//
// Adam replaces `torch.optim.Adam(...).step()` as used at spi/training/coaches/base_coach.py:132-135 (all
// G.parameters(), lr 3e-4) and spi/training/projectors/*_projector.py:55-58 ([w_opt] + noise buffers): defaults
// betas=(0.9, 0.999), eps=1e-8, no weight decay, no amsgrad.  Same arithmetic as torch's single-tensor path:
//   m = m + (g - m)*(1-b1);  v = v*b2 + g*g*(1-b2);  p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)
// The host keeps every parameter of an optimiser as a view into one flat fp32 arena (params, grads, m, v), so one
// launch streams 28 B/param (836 MB per G-step) with 128-bit loads/stores instead of ~400 foreach launches.
// `hyper` (device, optional) = {lr, bc1, bc2} lets a captured CUDA graph be replayed with a new learning rate.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, long long n, float lr, float b1, float b2, float eps,
                                                   float bc1, float bc2, const float* __restrict__ hyper, int zero_grad, float* gz,
                                                   const float* __restrict__ cond, float cond_thr) {
    // device-side early exit (rot_bbox_cx_coach.py:148-151 breaks BEFORE optimizer.step()): no update when *cond <= thr
    if (cond && cond[0] <= cond_thr) return;
    if (hyper) { lr = hyper[0]; bc1 = hyper[1]; bc2 = hyper[2]; }
    const float step_size = lr / bc1;
    const float rsbc2 = 1.f / sqrtf(bc2);
    const long long nv = n / 4;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
        float4 pp = ((float4*)p)[i], gg = ldg_stream((const float4*)g + i), mm = ((float4*)m)[i], vv = ((float4*)v)[i];
        float* P = (float*)&pp; float* G = (float*)&gg; float* M = (float*)&mm; float* V = (float*)&vv;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            M[k] = M[k] + (G[k] - M[k]) * (1.f - b1);
            V[k] = V[k] * b2 + G[k] * G[k] * (1.f - b2);
            float denom = sqrtf(V[k]) * rsbc2 + eps;
            P[k] = P[k] - step_size * (M[k] / denom);
        }
        ((float4*)p)[i] = pp; ((float4*)m)[i] = mm; ((float4*)v)[i] = vv;
        if (zero_grad) ((float4*)gz)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (long long i = nv * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float gk = g[i];
        float mk = m[i] + (gk - m[i]) * (1.f - b1);
        float vk = v[i] * b2 + gk * gk * (1.f - b2);
        m[i] = mk; v[i] = vk;
        p[i] = p[i] - step_size * (mk / (sqrtf(vk) * rsbc2 + eps));
        if (zero_grad) gz[i] = 0.f;
    }
}

// y[i] = mean of the 2x2 block (== F.interpolate bilinear/area at exactly 1/2 scale, align_corners=False), NCHW planes
__global__ void __launch_bounds__(256) half_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long long planes, int oh, int ow) {
    const long long total = planes * oh * ow;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int ox = (int)(i % ow); long long r = i / ow; int oy = (int)(r % oh); long long pl = r / oh;
        const float2* r0 = (const float2*)(x + (pl * 2 * oh + 2 * oy) * (2LL * ow)) + ox;
        const float2* r1 = (const float2*)(x + (pl * 2 * oh + 2 * oy + 1) * (2LL * ow)) + ox;
        float2 a = __ldg(r0), b = __ldg(r1);
        // torch's bilinear: (1-l)*((1-l)*a.x + l*a.y) + l*(...) with l = 0.5
        y[i] = 0.5f * (0.5f * a.x + 0.5f * a.y) + 0.5f * (0.5f * b.x + 0.5f * b.y);
    }
}

__global__ void __launch_bounds__(256) half_bwd_kernel(const float* __restrict__ gy, float* __restrict__ gx, long long planes, int oh, int ow) {
    const long long total = planes * oh * ow;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int ox = (int)(i % ow); long long r = i / ow; int oy = (int)(r % oh); long long pl = r / oh;
        float g = 0.25f * gy[i];
        float2* r0 = (float2*)(gx + (pl * 2 * oh + 2 * oy) * (2LL * ow)) + ox;
        float2* r1 = (float2*)(gx + (pl * 2 * oh + 2 * oy + 1) * (2LL * ow)) + ox;
        *r0 = make_float2(g, g); *r1 = make_float2(g, g);
    }
}

}  // namespace

extern "C" int spi_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr, float beta1,
                             float beta2, float eps, int step, const float* hyper, int zero_grad, const float* skip_if_le, float skip_threshold,
                             cudaStream_t stream) {
    SPI_CHECK_ARG(param && grad && exp_avg && exp_avg_sq, "adam_step: null pointer");
    SPI_CHECK_ARG(step >= 1 || hyper, "adam_step: step must be >= 1");
    SPI_CHECK_ARG((((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) == 0, "adam_step: arenas must be 16-byte aligned");
    if (n == 0) return SPI_OK;
    float bc1 = 1.f, bc2 = 1.f;
    if (!hyper) {
        bc1 = (float)(1.0 - pow((double)beta1, (double)step));
        bc2 = (float)(1.0 - pow((double)beta2, (double)step));
    }
    long long blocks = (n / 4 + 255) / 256;
    long long cap = (long long)spi_num_sms() * 8;
    int grid = (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
    adam_kernel<<<grid, 256, 0, stream>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, bc1, bc2, hyper, zero_grad, (float*)grad, skip_if_le, skip_threshold);
    SPI_COUNT_LAUNCH(1);
    SPI_LAUNCH_CHECK("adam_step");
    return SPI_OK;
}

extern "C" int spi_downsample2x(const float* x, float* y, long long planes, int out_h, int out_w, int backward, cudaStream_t stream) {
    SPI_CHECK_ARG(x && y && out_h >= 1 && out_w >= 1, "downsample2x: bad argument");
    SPI_CHECK_ARG((((uintptr_t)x | (uintptr_t)y) & 7) == 0, "downsample2x: tensors must be 8-byte aligned");
    long long total = planes * out_h * out_w;
    if (total == 0) return SPI_OK;
    long long cap = (long long)spi_num_sms() * 8, blocks = (total + 255) / 256;
    int grid = (int)(blocks > cap ? cap : blocks);
    if (!backward) half_fwd_kernel<<<grid, 256, 0, stream>>>(x, y, planes, out_h, out_w);
    else half_bwd_kernel<<<grid, 256, 0, stream>>>(x, y, planes, out_h, out_w);   // x = grad_out [planes,oh,ow], y = grad_in
    SPI_COUNT_LAUNCH(1);
    SPI_LAUNCH_CHECK("downsample2x");
    return SPI_OK;
}

// Column sums of a tall row-major matrix x [rows, cols] (cols % 4 == 0, <= 128): the bias gradients of the OSG decoder,
// db1 = sum_rows dpre, db2 = sum_rows dout (rows = millions of samples).  HBM-streaming: a warp reads whole rows with 128-bit
// loads (32 / (cols/4) rows per instruction), 4 loads in flight per lane; per-lane partials are combined through shared memory
// and one atomicAdd per column per CTA.
namespace {

__global__ void __launch_bounds__(256) column_sums_kernel(const float* __restrict__ x, long long rows, int cols, float* __restrict__ out) {
    __shared__ float sm[128];
    if (threadIdx.x < 128) sm[threadIdx.x] = 0.f;
    __syncthreads();
    const int c4 = cols >> 2;
    const long long nvec = rows * c4;
    const long long T = (long long)gridDim.x * blockDim.x;            // host guarantees T % c4 == 0: a thread's column never changes
    const long long v0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    long long v = v0;
    for (; v + 3 * T < nvec; v += 4 * T) {
        const float4 a = ldg_stream(reinterpret_cast<const float4*>(x) + v), b = ldg_stream(reinterpret_cast<const float4*>(x) + v + T);
        const float4 c = ldg_stream(reinterpret_cast<const float4*>(x) + v + 2 * T), d = ldg_stream(reinterpret_cast<const float4*>(x) + v + 3 * T);
        acc.x += (a.x + b.x) + (c.x + d.x); acc.y += (a.y + b.y) + (c.y + d.y);
        acc.z += (a.z + b.z) + (c.z + d.z); acc.w += (a.w + b.w) + (c.w + d.w);
    }
    for (; v < nvec; v += T) {
        const float4 a = ldg_stream(reinterpret_cast<const float4*>(x) + v);
        acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
    }
    const int col = (int)(v0 % c4) * 4;
    atomicAdd(sm + col, acc.x); atomicAdd(sm + col + 1, acc.y); atomicAdd(sm + col + 2, acc.z); atomicAdd(sm + col + 3, acc.w);
    __syncthreads();
    if (threadIdx.x < cols) atomicAdd(out + threadIdx.x, sm[threadIdx.x]);
}

}  // namespace

extern "C" int spi_column_sums(const float* x, long long rows, int cols, float* out, cudaStream_t stream) {
    SPI_CHECK_ARG(x && out && rows >= 0 && cols >= 4 && cols % 4 == 0 && cols <= 128, "column_sums: cols must be a multiple of 4, <= 128");
    SPI_CHECK_ARG(((uintptr_t)x & 15) == 0, "column_sums: x must be 16-byte aligned");
    cudaMemsetAsync(out, 0, sizeof(float) * cols, stream);
    if (rows == 0) return SPI_OK;
    const int c4 = cols / 4;
    long long nvec = rows * c4, blocks = (nvec + 256 * 16 - 1) / (256 * 16), cap = (long long)spi_num_sms() * 8;
    long long grid = blocks < cap ? blocks : cap;
    int need = c4, r256 = 256;                                   // grid * 256 must be a multiple of c4
    while (need % 2 == 0 && r256 % 2 == 0) { need /= 2; r256 /= 2; }
    grid = (grid + need - 1) / need * need;
    column_sums_kernel<<<(unsigned)grid, 256, 0, stream>>>(x, rows, cols, out);
    SPI_COUNT_LAUNCH(1);
    SPI_LAUNCH_CHECK("column_sums");
    return SPI_OK;
}

// 2x2 / stride-2 max pooling for channels-last fp32 activations (the VGG16 / VGG19 feature extractors of the losses,
// spi/criteria/lpips/networks.py:75-80, bbox_cx_loss.py:79-87).  Forward: one read of x, one write of y (ATen's NHWC kernel also
// writes an int64 index per output).  Backward recomputes the arg-max from x (first maximum in row-major window order, like
// torch) and writes every dx element exactly once: read x + dy/4, write dx.
namespace {

__global__ void __launch_bounds__(256) maxpool2_cl_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ out, int n,
                                                          int h, int w, int c4, int backward) {
    const int oh = h >> 1, ow = w >> 1;
    const long long total = (long long)n * oh * ow * c4;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int cv = (int)(idx % c4); long long r = idx / c4;
        const int ox = (int)(r % ow); r /= ow;
        const int oy = (int)(r % oh); const int nn = (int)(r / oh);
        const float4* xp = reinterpret_cast<const float4*>(x) + (((long long)nn * h + 2 * oy) * w + 2 * ox) * c4 + cv;
        const float4 a = __ldg(xp), b = __ldg(xp + c4), c = __ldg(xp + (long long)w * c4), d = __ldg(xp + (long long)w * c4 + c4);
        if (!backward) {
            float4 m;
            m.x = fmaxf(fmaxf(a.x, b.x), fmaxf(c.x, d.x)); m.y = fmaxf(fmaxf(a.y, b.y), fmaxf(c.y, d.y));
            m.z = fmaxf(fmaxf(a.z, b.z), fmaxf(c.z, d.z)); m.w = fmaxf(fmaxf(a.w, b.w), fmaxf(c.w, d.w));
            reinterpret_cast<float4*>(out)[idx] = m;
        } else {
            const float4 g = __ldg(reinterpret_cast<const float4*>(dy) + idx);
            float4 ga = make_float4(0.f, 0.f, 0.f, 0.f), gb = ga, gc = ga, gd = ga;
#define ROUTE(f)                                                                                   \
            {                                                                                      \
                int k = 0; float m = a.f;                                                          \
                if (b.f > m) { m = b.f; k = 1; }                                                   \
                if (c.f > m) { m = c.f; k = 2; }                                                   \
                if (d.f > m) { k = 3; }                                                            \
                if (k == 0) ga.f = g.f; else if (k == 1) gb.f = g.f; else if (k == 2) gc.f = g.f; else gd.f = g.f; \
            }
            ROUTE(x) ROUTE(y) ROUTE(z) ROUTE(w)
#undef ROUTE
            float4* op = reinterpret_cast<float4*>(out) + (((long long)nn * h + 2 * oy) * w + 2 * ox) * c4 + cv;
            op[0] = ga; op[c4] = gb; op[(long long)w * c4] = gc; op[(long long)w * c4 + c4] = gd;
        }
    }
}

}  // namespace

extern "C" int spi_maxpool2x2(const float* x, const float* dy, float* out, int n, int h, int w, int c, int backward, cudaStream_t stream) {
    SPI_CHECK_ARG(x && out && (!backward || dy), "maxpool2x2: null pointer");
    SPI_CHECK_ARG(n >= 1 && h >= 2 && w >= 2 && h % 2 == 0 && w % 2 == 0 && c >= 4 && c % 4 == 0, "maxpool2x2: even H, W and C %% 4 == 0 required");
    SPI_CHECK_ARG((((uintptr_t)x | (uintptr_t)out | (uintptr_t)dy) & 15) == 0, "maxpool2x2: tensors must be 16-byte aligned");
    const long long total = (long long)n * (h / 2) * (w / 2) * (c / 4);
    long long blocks = (total + 255) / 256, cap = (long long)spi_num_sms() * 16;
    maxpool2_cl_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, stream>>>(x, dy, out, n, h, w, c / 4, backward);
    SPI_COUNT_LAUNCH(1);
    SPI_LAUNCH_CHECK("maxpool2x2");
    return SPI_OK;
}

// Multi-tensor Adam over a device pointer table: rows of 5 x int64 = {param, grad, exp_avg, exp_avg_sq, numel}.  Parameters and
// moments live in the flat arenas of FlatAdam; the gradients are whatever tensors autograd produced (no accumulation into a
// pre-seated gradient arena: that cost one ATen add launch per parameter per iteration).  blockIdx.y = tensor, blockIdx.x strides
// over its elements.  Same arithmetic and same device-side early exit as adam_kernel.
namespace {

__global__ void __launch_bounds__(256) adam_multi_kernel(const long long* __restrict__ table, float lr, float b1, float b2, float eps, float bc1,
                                                         float bc2, const float* __restrict__ hyper, const float* __restrict__ cond, float cond_thr) {
    if (cond && cond[0] <= cond_thr) return;
    if (hyper) { lr = hyper[0]; bc1 = hyper[1]; bc2 = hyper[2]; }
    const long long* row = table + 5 * (long long)blockIdx.y;
    float* p = reinterpret_cast<float*>(row[0]);
    const float* g = reinterpret_cast<const float*>(row[1]);
    float* m = reinterpret_cast<float*>(row[2]);
    float* v = reinterpret_cast<float*>(row[3]);
    const long long n = row[4];
    const float step_size = lr / bc1;
    const float rsbc2 = 1.f / sqrtf(bc2);
    const bool vec = ((row[0] | row[1] | row[2] | row[3]) & 15) == 0;
    const long long nv = vec ? n / 4 : 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
        float4 pp = ((float4*)p)[i], gg = ldg_stream((const float4*)g + i), mm = ((float4*)m)[i], vv = ((float4*)v)[i];
        float* P = (float*)&pp; float* G = (float*)&gg; float* M = (float*)&mm; float* V = (float*)&vv;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            M[k] = M[k] + (G[k] - M[k]) * (1.f - b1);
            V[k] = V[k] * b2 + G[k] * G[k] * (1.f - b2);
            float denom = sqrtf(V[k]) * rsbc2 + eps;
            P[k] = P[k] - step_size * (M[k] / denom);
        }
        ((float4*)p)[i] = pp; ((float4*)m)[i] = mm; ((float4*)v)[i] = vv;
    }
    for (long long i = nv * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float gk = g[i];
        float mk = m[i] + (gk - m[i]) * (1.f - b1);
        float vk = v[i] * b2 + gk * gk * (1.f - b2);
        m[i] = mk; v[i] = vk;
        p[i] = p[i] - step_size * (mk / (sqrtf(vk) * rsbc2 + eps));
    }
}

}  // namespace

extern "C" int spi_adam_step_multi(const void* table, int count, int blocks_per_row, float lr, float beta1, float beta2, float eps, int step,
                                   const float* hyper, const float* skip_if_le, float skip_threshold, cudaStream_t stream) {
    SPI_CHECK_ARG(table && count >= 0 && count <= 65535 && blocks_per_row >= 1 && blocks_per_row <= 1024, "adam_step_multi: bad table");
    SPI_CHECK_ARG(step >= 1 || hyper, "adam_step_multi: step must be >= 1");
    if (count == 0) return SPI_OK;
    float bc1 = 1.f, bc2 = 1.f;
    if (!hyper) {
        bc1 = (float)(1.0 - pow((double)beta1, (double)step));
        bc2 = (float)(1.0 - pow((double)beta2, (double)step));
    }
    adam_multi_kernel<<<dim3(blocks_per_row, count), 256, 0, stream>>>((const long long*)table, lr, beta1, beta2, eps, bc1, bc2, hyper, skip_if_le, skip_threshold);
    SPI_COUNT_LAUNCH(1);
    SPI_LAUNCH_CHECK("adam_step_multi");
    return SPI_OK;
}
