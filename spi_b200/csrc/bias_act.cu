// bias_act for sm_100a: y = clamp(act(x + b) * gain), plus first/second-order gradient modes.
//
// Replaces the reference plugin op `bias_act` (eg3d/torch_utils/ops/bias_act.cpp:36-93,
// kernel bias_act.cu:28-147) with the same argument meaning:
//   grad=0: x is the activation input;            y = clamp(act(x+b)*gain)
//   grad=1: x is the incoming gradient dy;        y = x * act'(.) * gain, masked where the fwd clamp saturated
//   grad=2: x is the gradient of grad=1's output; second-order term (uses `dy`)
// `xref` / `yref` are the saved forward input / output; empty (NULL) means absent.
//
// B200 design: pure HBM streaming.  Each thread moves 16 bytes per tensor per iteration (LDG.128 / STG.128,
// L1 no-allocate), 2 independent iterations in flight, grid sized to a multiple of the SM count.  The bias index is
// resolved per vector: stepB % VEC == 0 (NCHW) -> one scalar bias per vector; stepB == 1 (channels_last) -> a
// vector of consecutive biases.  Anything else takes the scalar tail path.
#include "common.cuh"

namespace {

struct BiasActParams {
    const void* x; const void* b; const void* xref; const void* yref; const void* dy; void* y;
    int grad, act;
    float alpha, gain, clamp;
    long long n;
    int sizeB, stepB, force_scalar;
    const float* noise; const float* noise_strength;   // optional per-pixel noise map [H*W] * device scalar, added with the bias
    int hw, cl_c;                                       // pixels per image; channels if channels-last else 0
};

template <class S, int A>
__device__ __forceinline__ S act_eval(S x, S b, S xref, S yref, S dy, int G, S alpha, S gain, S clamp) {
    const S one = (S)1, two = (S)2, expRange = (S)80, halfExpRange = (S)40;
    const S seluScale = (S)1.0507009873554804934193349852946;
    const S seluAlpha = (S)1.6732632423543772848170429916717;
    S yy = (gain != 0) ? yref / gain : (S)0;
    S y = 0;
    if (G == 0) x += b; else xref += b;
    if (A == 1) { y = x; }
    if (A == 2) { y = (G == 0) ? ((x > 0) ? x : (S)0) : (G == 1 ? ((yy > 0) ? x : (S)0) : (S)0); }
    if (A == 3) { y = (G == 0) ? ((x > 0) ? x : x * alpha) : (G == 1 ? ((yy > 0) ? x : x * alpha) : (S)0); }
    if (A == 4) {
        if (G == 0) { S c = exp(x), d = one / c; y = (x < -expRange) ? -one : (x > expRange) ? one : (c - d) / (c + d); }
        if (G == 1) y = x * (one - yy * yy);
        if (G == 2) y = x * (one - yy * yy) * (-two * yy);
    }
    if (A == 5) {
        if (G == 0) y = (x < -expRange) ? (S)0 : one / (exp(-x) + one);
        if (G == 1) y = x * yy * (one - yy);
        if (G == 2) y = x * yy * (one - yy) * (one - two * yy);
    }
    if (A == 6) {
        if (G == 0) y = (x >= 0) ? x : exp(x) - one;
        if (G == 1) y = (yy >= 0) ? x : x * (yy + one);
        if (G == 2) y = (yy >= 0) ? (S)0 : x * (yy + one);
    }
    if (A == 7) {
        if (G == 0) y = (x >= 0) ? seluScale * x : (seluScale * seluAlpha) * (exp(x) - one);
        if (G == 1) y = (yy >= 0) ? x * seluScale : x * (yy + seluScale * seluAlpha);
        if (G == 2) y = (yy >= 0) ? (S)0 : x * (yy + seluScale * seluAlpha);
    }
    if (A == 8) {
        if (G == 0) y = (x > expRange) ? x : log(exp(x) + one);
        if (G == 1) y = x * (one - exp(-yy));
        if (G == 2) { S c = exp(-yy); y = x * c * (one - c); }
    }
    if (A == 9) {
        if (G == 0) y = (x < -expRange) ? (S)0 : x / (exp(-x) + one);
        else {
            S c = exp(xref), d = c + one;
            if (G == 1) y = (xref > halfExpRange) ? x : x * c * (xref + d) / (d * d);
            else y = (xref > halfExpRange) ? (S)0 : x * c * (xref * (two - d) + two * d) / (d * d * d);
            yref = (xref < -expRange) ? (S)0 : xref / (exp(-xref) + one) * gain;
        }
    }
    y *= gain * dy;
    if (clamp >= 0) {
        if (G == 0) y = (y > -clamp && y < clamp) ? y : ((y >= 0) ? clamp : -clamp);
        else y = (yref > -clamp && yref < clamp) ? y : (S)0;
    }
    return y;
}

template <class T> struct Vec;       // 16-byte vectors
template <> struct Vec<float>  { static constexpr int N = 4; };
template <> struct Vec<__half> { static constexpr int N = 8; };
template <> struct Vec<double> { static constexpr int N = 2; };

template <class T, int N> struct alignas(16) Pack { T v[N]; };

template <class T, int A>
__global__ void __launch_bounds__(256) bias_act_kernel(BiasActParams p) {
    typedef typename Acc<T>::t S;
    constexpr int V = Vec<T>::N;
    typedef Pack<T, V> P;
    const S alpha = (S)p.alpha, gain = (S)p.gain, clamp = (S)p.clamp;
    const int G = p.grad;
    const T* x = (const T*)p.x; const T* b = (const T*)p.b; const T* xr = (const T*)p.xref;
    const T* yr = (const T*)p.yref; const T* dyp = (const T*)p.dy; T* y = (T*)p.y;
    const long long nvec = p.n / V;
    const bool vec_ok = !p.force_scalar && (!p.noise || (p.cl_c ? (p.cl_c % V == 0) : true)) && ((b == nullptr) || (p.stepB % V == 0) || (p.stepB == 1 && p.sizeB % V == 0));
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (vec_ok) {
        for (; i < nvec; i += stride) {
            P vx = ((const P*)x)[i], vxr, vyr, vdy, vb, out;
            if (xr) vxr = ((const P*)xr)[i];
            if (yr) vyr = ((const P*)yr)[i];
            if (dyp) vdy = ((const P*)dyp)[i];
            long long e0 = i * V;
            S bs = 0;
            bool bvec = false;
            S nz[V];
#pragma unroll
            for (int k = 0; k < V; k++) nz[k] = 0;
            if (p.noise) {
                const S st = (S)p.noise_strength[0];
                if (p.cl_c) { S v = (S)p.noise[(e0 / p.cl_c) % p.hw] * st;
#pragma unroll
                    for (int k = 0; k < V; k++) nz[k] = v; }
                else {
#pragma unroll
                    for (int k = 0; k < V; k++) nz[k] = (S)p.noise[(e0 + k) % p.hw] * st; }
            }
            if (b) {
                if (p.stepB == 1) { vb = *(const P*)(b + (e0 % p.sizeB)); bvec = true; }
                else bs = (S)b[(e0 / p.stepB) % p.sizeB];
            }
#pragma unroll
            for (int k = 0; k < V; k++) {
                S bb = (bvec ? (S)vb.v[k] : bs) + nz[k];
                out.v[k] = (T)act_eval<S, A>((S)vx.v[k], bb, xr ? (S)vxr.v[k] : (S)0, yr ? (S)vyr.v[k] : (S)0,
                                             dyp ? (S)vdy.v[k] : (S)1, G, alpha, gain, clamp);
            }
            ((P*)y)[i] = out;
        }
    }
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (long long e = (vec_ok ? nvec * V : 0) + tid; e < p.n; e += stride) {
        S bb = b ? (S)b[(e / p.stepB) % p.sizeB] : (S)0;
        if (p.noise) bb += (S)p.noise[p.cl_c ? (e / p.cl_c) % p.hw : e % p.hw] * (S)p.noise_strength[0];
        y[e] = (T)act_eval<S, A>((S)x[e], bb, xr ? (S)xr[e] : (S)0, yr ? (S)yr[e] : (S)0, dyp ? (S)dyp[e] : (S)1, G,
                                 alpha, gain, clamp);
    }
}


// ---- hot path: fp32, bias along the fastest dimension (channels-last rows of C floats) or no bias at all --------------------
// Every thread keeps ONE column of 4 channels for its whole life (the total thread count is a multiple of the vectors per
// row), so the bias vector sits in registers and the pixel index advances by a constant: no division or modulo in the loop.
// Four independent 16-byte loads per operand are in flight per thread.
template <int A, int G>
__global__ void __launch_bounds__(256) bias_act_rows_kernel(BiasActParams p, unsigned c4, unsigned rows) {
    const float* __restrict__ x = (const float*)p.x;
    const float* __restrict__ yr = (const float*)p.yref;
    float* __restrict__ y = (float*)p.y;
    const unsigned T = gridDim.x * blockDim.x, t = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned col = t % c4, rstep = T / c4;
    unsigned row = t / c4;
    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (G == 0 && p.b) bv = *(const float4*)((const float*)p.b + 4 * col);
    const bool nz = G == 0 && p.noise != nullptr;
    const float st = nz ? p.noise_strength[0] : 0.f;
    const unsigned hw = (unsigned)p.hw;
    unsigned rhw = row % hw;
    const unsigned hstep = rstep % hw;
    constexpr int U = 4;
    const unsigned hstep_u = (U * hstep) % hw;
    while (row < rows) {
        float4 vx[U], vy[U];
        float nv[U];
        unsigned rr[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            rr[u] = row + u * rstep;
            const bool ok = rr[u] < rows;
            const size_t off = ((size_t)rr[u] * c4 + col) * 4;
            vx[u] = ok ? ldg_stream((const float4*)(x + off)) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (G == 1) vy[u] = ok ? ldg_stream((const float4*)(yr + off)) : make_float4(0.f, 0.f, 0.f, 0.f);
            nv[u] = 0.f;
            if (nz) {
                unsigned h = rhw + u * hstep;            // < 4*hw: at most three conditional subtractions, no division
                if (h >= 2 * hw) h -= 2 * hw;
                if (h >= hw) h -= hw;
                if (h >= hw) h -= hw;
                nv[u] = ok ? __ldg(p.noise + h) * st : 0.f;
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (rr[u] < rows) {
                float4 o;
                if (G == 0) {
                    o.x = act_eval<float, A>(vx[u].x, bv.x + nv[u], 0.f, 0.f, 1.f, 0, p.alpha, p.gain, p.clamp);
                    o.y = act_eval<float, A>(vx[u].y, bv.y + nv[u], 0.f, 0.f, 1.f, 0, p.alpha, p.gain, p.clamp);
                    o.z = act_eval<float, A>(vx[u].z, bv.z + nv[u], 0.f, 0.f, 1.f, 0, p.alpha, p.gain, p.clamp);
                    o.w = act_eval<float, A>(vx[u].w, bv.w + nv[u], 0.f, 0.f, 1.f, 0, p.alpha, p.gain, p.clamp);
                } else {
                    o.x = act_eval<float, A>(vx[u].x, 0.f, 0.f, vy[u].x, 1.f, 1, p.alpha, p.gain, p.clamp);
                    o.y = act_eval<float, A>(vx[u].y, 0.f, 0.f, vy[u].y, 1.f, 1, p.alpha, p.gain, p.clamp);
                    o.z = act_eval<float, A>(vx[u].z, 0.f, 0.f, vy[u].z, 1.f, 1, p.alpha, p.gain, p.clamp);
                    o.w = act_eval<float, A>(vx[u].w, 0.f, 0.f, vy[u].w, 1.f, 1, p.alpha, p.gain, p.clamp);
                }
                stg_stream((float4*)(y + ((size_t)rr[u] * c4 + col) * 4), o);
            }
        }
        row += U * rstep;
        if (nz) {
            unsigned h = rhw + hstep_u;
            rhw = h >= hw ? h - hw : h;
        }
    }
}

// Gradient of the activation AND its reductions in one pass (grad = 1 of linear / relu / lrelu; channels-last rows of C floats):
// dx = act'(yref) * dy is written, while every thread -- which keeps one column of 4 channels for its whole life -- adds what it wrote to
// four private sums (the bias gradient db[c] = sum over pixels of dx) and, when a noise map is given, to sum dx * noise[pixel] (the
// gradient of the layer's noise_strength).  Private sums meet in shared memory, then one atomic per channel and CTA.  Replaces the
// epilogue_grad_reduce pass over dx that used to follow bias_act(grad = 1) (one read of dx saved per layer).
template <int A>
__global__ void __launch_bounds__(256) bias_act_grad_reduce_rows_kernel(BiasActParams p, unsigned c4, unsigned rows, float* __restrict__ db,
                                                                        float* __restrict__ dstrength) {
    extern __shared__ float sred[];                       // 4 * c4 channel sums + 1
    const float* __restrict__ x = (const float*)p.x;
    const float* __restrict__ yr = (const float*)p.yref;
    float* __restrict__ y = (float*)p.y;
    const unsigned T = gridDim.x * blockDim.x, t = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned col = t % c4, rstep = T / c4;
    unsigned row = t / c4;
    const bool nz = dstrength != nullptr;
    const unsigned hw = (unsigned)p.hw;
    unsigned rhw = row % hw;
    const unsigned hstep = rstep % hw;
    constexpr int U = 4;
    const unsigned hstep_u = (U * hstep) % hw;
    for (unsigned i = threadIdx.x; i < 4 * c4 + 1; i += blockDim.x) sred[i] = 0.f;
    __syncthreads();
    float4 sb = make_float4(0.f, 0.f, 0.f, 0.f);
    float ss = 0.f;
    while (row < rows) {
        float4 vx[U], vy[U];
        float nv[U];
        unsigned rr[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            rr[u] = row + u * rstep;
            const bool ok = rr[u] < rows;
            const size_t off = ((size_t)rr[u] * c4 + col) * 4;
            vx[u] = ok ? ldg_stream((const float4*)(x + off)) : make_float4(0.f, 0.f, 0.f, 0.f);
            vy[u] = ok ? ldg_stream((const float4*)(yr + off)) : make_float4(0.f, 0.f, 0.f, 0.f);
            nv[u] = 0.f;
            if (nz) {
                unsigned h = rhw + u * hstep;            // < 4*hw: at most three conditional subtractions, no division
                if (h >= 2 * hw) h -= 2 * hw;
                if (h >= hw) h -= hw;
                if (h >= hw) h -= hw;
                nv[u] = ok ? __ldg(p.noise + h) : 0.f;
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (rr[u] < rows) {
                float4 o;
                o.x = act_eval<float, A>(vx[u].x, 0.f, 0.f, vy[u].x, 1.f, 1, p.alpha, p.gain, p.clamp);
                o.y = act_eval<float, A>(vx[u].y, 0.f, 0.f, vy[u].y, 1.f, 1, p.alpha, p.gain, p.clamp);
                o.z = act_eval<float, A>(vx[u].z, 0.f, 0.f, vy[u].z, 1.f, 1, p.alpha, p.gain, p.clamp);
                o.w = act_eval<float, A>(vx[u].w, 0.f, 0.f, vy[u].w, 1.f, 1, p.alpha, p.gain, p.clamp);
                stg_stream((float4*)(y + ((size_t)rr[u] * c4 + col) * 4), o);
                sb.x += o.x; sb.y += o.y; sb.z += o.z; sb.w += o.w;
                ss = fmaf((o.x + o.y) + (o.z + o.w), nv[u], ss);
            }
        }
        row += U * rstep;
        if (nz) {
            unsigned h = rhw + hstep_u;
            rhw = h >= hw ? h - hw : h;
        }
    }
    if (db) {
        atomicAdd(&sred[4 * col + 0], sb.x); atomicAdd(&sred[4 * col + 1], sb.y);
        atomicAdd(&sred[4 * col + 2], sb.z); atomicAdd(&sred[4 * col + 3], sb.w);
    }
    if (nz) {
        ss = warp_sum(ss);
        if ((threadIdx.x & 31) == 0) atomicAdd(&sred[4 * c4], ss);
    }
    __syncthreads();
    if (db)
        for (unsigned i = threadIdx.x; i < 4 * c4; i += blockDim.x) atomicAdd(db + i, sred[i]);
    if (nz && threadIdx.x == 0) atomicAdd(dstrength, sred[4 * c4]);
}

template <int G>
void* pick_rows_kernel(int act) {
    switch (act) {
        case 1: return (void*)bias_act_rows_kernel<1, G>;
        case 2: return (void*)bias_act_rows_kernel<2, G>;
        case 3: return (void*)bias_act_rows_kernel<3, G>;
        case 5: return (void*)bias_act_rows_kernel<5, G>;
    }
    return nullptr;
}

template <class T>
void* pick_kernel(int act) {
    switch (act) {
        case 1: return (void*)bias_act_kernel<T, 1>;
        case 2: return (void*)bias_act_kernel<T, 2>;
        case 3: return (void*)bias_act_kernel<T, 3>;
        case 4: return (void*)bias_act_kernel<T, 4>;
        case 5: return (void*)bias_act_kernel<T, 5>;
        case 6: return (void*)bias_act_kernel<T, 6>;
        case 7: return (void*)bias_act_kernel<T, 7>;
        case 8: return (void*)bias_act_kernel<T, 8>;
        case 9: return (void*)bias_act_kernel<T, 9>;
    }
    return nullptr;
}

}  // namespace

static int bias_act_impl(const void* x, const void* b, const void* xref, const void* yref, const void* dy, void* y,
                         long long numel, int size_b, int step_b, int dtype, int grad, int act, float alpha,
                         float gain, float clamp, const float* noise, const float* noise_strength, int hw, int cl_c,
                         cudaStream_t stream) {
    SPI_CHECK_ARG(x && y, "bias_act: x and y must be non-null");
    SPI_CHECK_ARG(numel >= 0 && numel <= 2147483647LL, "bias_act: x is too large");
    SPI_CHECK_ARG(grad >= 0 && grad <= 2, "bias_act: grad must be 0, 1 or 2");
    SPI_CHECK_ARG(act >= 1 && act <= 9, "bias_act: no CUDA kernel found for the specified activation func");
    SPI_CHECK_ARG(!b || (size_b > 0 && step_b > 0), "bias_act: b has wrong number of elements");
    if (numel == 0) return SPI_OK;
    void* k = nullptr;
    int vec = 4;
    if (dtype == SPI_DT_F32) { k = pick_kernel<float>(act); vec = 4; }
    else if (dtype == SPI_DT_F16) { k = pick_kernel<__half>(act); vec = 8; }
    else if (dtype == SPI_DT_F64) { k = pick_kernel<double>(act); vec = 2; }
    SPI_CHECK_ARG(k, "bias_act: unsupported dtype %d", dtype);
    // The vector path needs 16-byte aligned pointers (torch allocations are; offset views may not be).
    uintptr_t al = (uintptr_t)x | (uintptr_t)y | (uintptr_t)xref | (uintptr_t)yref | (uintptr_t)dy | (uintptr_t)b;
    BiasActParams p{x, b, xref, yref, dy, y, grad, act, alpha, gain, clamp, numel, b ? size_b : 1, b ? step_b : 1,
                    (al & 15) ? 1 : 0, noise, noise_strength, hw > 0 ? hw : 1, cl_c};
    // hot path: fp32 rows (see bias_act_rows_kernel).  grad=1 of relu / lrelu / linear / sigmoid reads only dy and yref.
    {
        // grad = 1 of relu / lrelu / linear / sigmoid depends on dy and yref only (the bias enters through xref, which they ignore)
        const bool g1 = grad == 1 && yref && !xref && !dy && !noise && (act == 1 || act == 2 || act == 3 || act == 5);
        const bool g_ok = (grad == 0 && !xref && !yref && !dy) || g1;
        long long c = 0;
        if (g1) c = (numel % 128 == 0) ? 128 : 0;
        else if (b) { if (step_b == 1 && size_b % 4 == 0 && (!noise || cl_c == size_b)) c = size_b; }
        else c = noise ? ((cl_c > 0 && cl_c % 4 == 0) ? cl_c : 0) : ((numel % 128 == 0) ? 128 : 0);
        void* rk = grad == 0 ? pick_rows_kernel<0>(act) : pick_rows_kernel<1>(act);
        if (dtype == SPI_DT_F32 && g_ok && rk && c > 0 && numel % c == 0 && !(al & 15) && numel >= 4096) {
            const unsigned c4 = (unsigned)(c / 4), rows = (unsigned)(numel / c);
            unsigned need = c4, r256 = 256;                      // grid * 256 must be a multiple of c4
            while (need % 2 == 0 && r256 % 2 == 0) { need /= 2; r256 /= 2; }
            long long want = (numel / 4 + 256 * 4 - 1) / (256 * 4);
            long long capb = (long long)spi_num_sms() * 8;
            long long grid = want < capb ? want : capb;
            grid = (grid + need - 1) / need * need;
            void* args[] = {&p, (void*)&c4, (void*)&rows};
            cudaError_t e = cudaLaunchKernel(rk, dim3((unsigned)grid), dim3(256), args, 0, stream);
            SPI_COUNT_LAUNCH(1);
            if (e != cudaSuccess) { spi_set_error("bias_act: %s", cudaGetErrorString(e)); return SPI_ERR_CUDA; }
            return SPI_OK;
        }
    }
    const int block = 256;
    long long work = p.force_scalar ? numel : (numel + vec - 1) / vec;
    long long blocks = (work + block * 2 - 1) / (block * 2);
    long long cap = (long long)spi_num_sms() * 16;
    int grid = (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
    void* args[] = {&p};
    cudaError_t e = cudaLaunchKernel(k, dim3(grid), dim3(block), args, 0, stream);
    SPI_COUNT_LAUNCH(1);
    if (e != cudaSuccess) { spi_set_error("bias_act: %s", cudaGetErrorString(e)); return SPI_ERR_CUDA; }
    return SPI_OK;
}

extern "C" int spi_bias_act(const void* x, const void* b, const void* xref, const void* yref, const void* dy, void* y,
                            long long numel, int size_b, int step_b, int dtype, int grad, int act, float alpha,
                            float gain, float clamp, cudaStream_t stream) {
    return bias_act_impl(x, b, xref, yref, dy, y, numel, size_b, step_b, dtype, grad, act, alpha, gain, clamp, nullptr, nullptr, 1, 0, stream);
}

// Layer epilogue of SynthesisLayer.forward (networks_stylegan2.py:320-329) in one pass:
//   y = clamp(act(x + noise[h,w] * noise_strength + b[c]) * gain);  x is [N,C,H,W] dense, NCHW (cl_c = 0) or channels-last (cl_c = C).
extern "C" int spi_bias_act_noise(const void* x, const void* b, void* y, const float* noise, const float* noise_strength,
                                  long long numel, int size_b, int step_b, int hw, int channels_last_c, int dtype, int act,
                                  float alpha, float gain, float clamp, cudaStream_t stream) {
    SPI_CHECK_ARG(noise && noise_strength && hw >= 1, "bias_act_noise: noise map required");
    SPI_CHECK_ARG(dtype == SPI_DT_F32, "bias_act_noise: float32 only");
    return bias_act_impl(x, b, nullptr, nullptr, nullptr, y, numel, size_b, step_b, dtype, 0, act, alpha, gain, clamp, noise,
                         noise_strength, hw, channels_last_c, stream);
}

// ---------------------------------------------------------------------------------------------------------------------
// Gradient reductions of the layer epilogue in ONE pass over dx (channels-last fp32, C % 4 == 0):
//   db[c]      = sum_{n,h,w} dx[n,c,h,w]                                   (bias_act.py:166-167)
//   pix[h,w]   = sum_{n,c} dx[n,c,h,w]  ->  dnoise = pix * strength,  dstrength = sum pix * noise   (noise branch)
// A warp owns one pixel at a time (lanes stride over the C/4 channel groups with LDG.128); per-channel partials stay in
// registers for the whole kernel and are combined through shared memory + one atomicAdd per channel per CTA.
namespace {

constexpr int RG_MAXG = 8;     // channel groups per lane: C <= 4*32*8 = 1024
constexpr int RG_U = 4;        // pixels in flight per warp

// A warp owns RG_U pixels (hw positions) at a time and walks the batch for them, so the per-pixel sum over (n, c) is complete
// in registers and is stored once (no atomics, no memset); per-channel partials stay in registers for the whole kernel and are
// combined through shared memory + one atomicAdd per channel per CTA.  NG = ceil(C / 128): float4 groups per lane (32-bit index
// math, no predicates inside the unrolled loads for the common C % 128 == 0 case).
template <int NG>
__global__ void __launch_bounds__(256) epilogue_grad_reduce_kernel(const float* __restrict__ dx, int n, int C, int HW,
                                                                   const float* __restrict__ noise, float* __restrict__ db,
                                                                   float* __restrict__ dpix, float* __restrict__ dstrength) {
    extern __shared__ float sm[];                 // [8 warps][C]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int groups = C >> 2;
    float4 acc[NG];
#pragma unroll
    for (int g = 0; g < NG; g++) acc[g] = make_float4(0.f, 0.f, 0.f, 0.f);
    float ds = 0.f;
    const bool want_pix = (dpix != nullptr) || (dstrength != nullptr);
    const int wstride = gridDim.x * 8 * RG_U;
    for (int hw0 = (blockIdx.x * 8 + warp) * RG_U; hw0 < HW; hw0 += wstride) {
        float ps[RG_U];
#pragma unroll
        for (int u = 0; u < RG_U; u++) ps[u] = 0.f;
        for (int nn = 0; nn < n; nn++) {
            float4 v[RG_U][NG];
#pragma unroll
            for (int u = 0; u < RG_U; u++) {
                const bool ok = hw0 + u < HW;
                const float4* row = (const float4*)(dx + ((size_t)nn * HW + hw0 + u) * C);
#pragma unroll
                for (int g = 0; g < NG; g++) {
                    const int gi = lane + 32 * g;
                    v[u][g] = (ok && gi < groups) ? ldg_stream(row + gi) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
#pragma unroll
            for (int u = 0; u < RG_U; u++)
#pragma unroll
                for (int g = 0; g < NG; g++) {
                    acc[g].x += v[u][g].x; acc[g].y += v[u][g].y; acc[g].z += v[u][g].z; acc[g].w += v[u][g].w;
                    ps[u] += (v[u][g].x + v[u][g].y) + (v[u][g].z + v[u][g].w);
                }
        }
        if (want_pix) {
#pragma unroll
            for (int u = 0; u < RG_U; u++) ps[u] = warp_sum(ps[u]);
            if (lane == 0) {
#pragma unroll
                for (int u = 0; u < RG_U; u++)
                    if (hw0 + u < HW) {
                        if (dpix) dpix[hw0 + u] = ps[u];
                        if (dstrength) ds = fmaf(ps[u], noise[hw0 + u], ds);
                    }
            }
        }
    }
    if (db) {
#pragma unroll
        for (int g = 0; g < NG; g++) {
            const int gi = lane + 32 * g;
            if (gi < groups) *(float4*)(sm + warp * C + gi * 4) = acc[g];
        }
        __syncthreads();
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < 8; w++) t += sm[w * C + c];
            atomicAdd(db + c, t);
        }
    }
    if (dstrength && lane == 0 && ds != 0.f) atomicAdd(dstrength, ds);
}

// db only, for channel counts that are not a multiple of 4 (the RGB outputs of the toRGB layers, C = 3): the flat [pixels*C]
// array is read as float4 vectors; the grid-stride is a multiple of C vectors, so each of a thread's 4 accumulators always sees
// the same channel ((4 v + k) mod C is invariant), and the block combines them through shared memory.
__global__ void __launch_bounds__(256) bias_grad_smallc_kernel(const float* __restrict__ dx, long long nvec, int C, float* __restrict__ db) {
    __shared__ float sm[8];
    if (threadIdx.x < 8) sm[threadIdx.x] = 0.f;
    __syncthreads();
    const long long T = (long long)gridDim.x * blockDim.x;            // host guarantees T % C == 0
    const long long v0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (long long v = v0; v < nvec; v += T) {
        const float4 x = ldg_stream(reinterpret_cast<const float4*>(dx) + v);
        acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
    }
    const int c0 = (int)((4 * v0) % C);
    atomicAdd(sm + c0, acc.x);
    atomicAdd(sm + (c0 + 1) % C, acc.y);
    atomicAdd(sm + (c0 + 2) % C, acc.z);
    atomicAdd(sm + (c0 + 3) % C, acc.w);
    __syncthreads();
    if (threadIdx.x < C) atomicAdd(db + threadIdx.x, sm[threadIdx.x]);
}

}  // namespace

// dx = d act(x + b)/dx * dy from the saved output yref (grad = 1 of act 1 linear / 2 relu / 3 lrelu, with gain and clamp as spi_bias_act),
// db[c] = sum over pixels of dx (NULL: not wanted), dstrength = sum dx * noise[pixel % hw] (NULL: not wanted; noise [hw]) -- one pass.
// dy / yref / dx: channels-last rows of c floats (c % 4 == 0, c <= 1024), numel = pixels * c, 16-byte aligned.  db / dstrength overwritten.
extern "C" int spi_bias_act_grad_reduce(const float* dy, const float* yref, float* dx, long long numel, int c, int hw, int act, float alpha,
                                        float gain, float clamp, const float* noise, float* db, float* dstrength, cudaStream_t stream) {
    SPI_CHECK_ARG(dy && yref && dx && (db || dstrength), "bias_act_grad_reduce: null pointer");
    SPI_CHECK_ARG(act >= 1 && act <= 3, "bias_act_grad_reduce: act must be 1 (linear), 2 (relu) or 3 (lrelu)");
    SPI_CHECK_ARG(c >= 4 && c % 4 == 0 && c <= 1024 && numel > 0 && numel % c == 0 && numel / c <= 4294967295LL, "bias_act_grad_reduce: bad shape (numel=%lld c=%d)", numel, c);
    SPI_CHECK_ARG(!dstrength || noise, "bias_act_grad_reduce: noise map required for dstrength");
    SPI_CHECK_ARG((((uintptr_t)dy | (uintptr_t)yref | (uintptr_t)dx) & 15) == 0, "bias_act_grad_reduce: tensors must be 16-byte aligned");
    if (db && dstrength == db + c) cudaMemsetAsync(db, 0, sizeof(float) * (c + 1), stream);      // one fill when the caller packed them
    else {
        if (db) cudaMemsetAsync(db, 0, sizeof(float) * c, stream);
        if (dstrength) cudaMemsetAsync(dstrength, 0, sizeof(float), stream);
    }
    BiasActParams p{dy, nullptr, nullptr, yref, nullptr, dx, 1, act, alpha, gain, clamp, numel, 1, 1, 0, noise, nullptr, hw > 0 ? hw : 1, c};
    const unsigned c4 = (unsigned)(c / 4), rows = (unsigned)(numel / c);
    unsigned need = c4, r256 = 256;                      // grid * 256 must be a multiple of c4
    while (need % 2 == 0 && r256 % 2 == 0) { need /= 2; r256 /= 2; }
    long long want = (numel / 4 + 256 * 4 - 1) / (256 * 4);
    long long capb = (long long)spi_num_sms() * 8;
    long long grid = want < capb ? want : capb;
    grid = (grid + need - 1) / need * need;
    const size_t smem = sizeof(float) * (4 * c4 + 1);
    void* k = act == 1 ? (void*)bias_act_grad_reduce_rows_kernel<1> : (act == 2 ? (void*)bias_act_grad_reduce_rows_kernel<2> : (void*)bias_act_grad_reduce_rows_kernel<3>);
    void* args[] = {&p, (void*)&c4, (void*)&rows, (void*)&db, (void*)&dstrength};
    cudaError_t e = cudaLaunchKernel(k, dim3((unsigned)grid), dim3(256), args, smem, stream);
    SPI_COUNT_LAUNCH(1);
    if (e != cudaSuccess) { spi_set_error("bias_act_grad_reduce: %s", cudaGetErrorString(e)); return SPI_ERR_CUDA; }
    return SPI_OK;
}

extern "C" int spi_epilogue_grad_reduce(const float* dx, long long pixels, int c, int hw, const float* noise, float* db, float* dpix,
                                        float* dstrength, cudaStream_t stream) {
    SPI_CHECK_ARG(dx && pixels >= 0 && c >= 1, "epilogue_grad_reduce: bad argument");
    SPI_CHECK_ARG(((uintptr_t)dx & 15) == 0, "epilogue_grad_reduce: dx must be 16-byte aligned");
    if (c % 4 != 0) {      // small channel counts (RGB): bias gradient only
        SPI_CHECK_ARG(c <= 8 && db && !dpix && !dstrength && (pixels * c) % 4 == 0, "epilogue_grad_reduce: C %% 4 != 0 supports db only, C <= 8, numel %% 4 == 0");
        cudaMemsetAsync(db, 0, sizeof(float) * c, stream);
        if (pixels == 0) return SPI_OK;
        const long long nvec = pixels * c / 4;
        long long blocks = (nvec + 256 * 8 - 1) / (256 * 8), capb = (long long)spi_num_sms() * 8;
        long long grid = blocks < capb ? blocks : capb;
        grid = (grid + c - 1) / c * c;                          // grid * 256 vectors per sweep: a multiple of C
        bias_grad_smallc_kernel<<<(unsigned)grid, 256, 0, stream>>>(dx, nvec, c, db);
        SPI_COUNT_LAUNCH(1);
        SPI_LAUNCH_CHECK("epilogue_grad_reduce");
        return SPI_OK;
    }
    SPI_CHECK_ARG(c >= 4 && c <= 4 * 32 * RG_MAXG, "epilogue_grad_reduce: C must be <= 1024");
    SPI_CHECK_ARG(!dstrength || noise, "epilogue_grad_reduce: noise map required for dstrength");
    if (hw <= 0) hw = 1;
    SPI_CHECK_ARG(pixels % hw == 0 && pixels / hw <= 2147483647LL && pixels * c <= (1LL << 40), "epilogue_grad_reduce: pixels must be n * hw");
    if (db && dstrength == db + c) cudaMemsetAsync(db, 0, sizeof(float) * (c + 1), stream);      // one fill when the caller packed them
    else {
        if (db) cudaMemsetAsync(db, 0, sizeof(float) * c, stream);
        if (dstrength) cudaMemsetAsync(dstrength, 0, sizeof(float), stream);
    }
    if (pixels == 0) {
        if (dpix) cudaMemsetAsync(dpix, 0, sizeof(float) * hw, stream);
        return SPI_OK;
    }
    const int n = (int)(pixels / hw);
    long long want = ((long long)hw + 8 * RG_U - 1) / (8 * RG_U);
    long long cap = (long long)spi_num_sms() * 8;
    int grid = (int)(want < cap ? want : cap);
    const size_t smem = sizeof(float) * 8 * c;
    const int ng = (c + 127) / 128;
    if (ng <= 1) epilogue_grad_reduce_kernel<1><<<grid, 256, smem, stream>>>(dx, n, c, hw, noise, db, dpix, dstrength);
    else if (ng <= 2) epilogue_grad_reduce_kernel<2><<<grid, 256, smem, stream>>>(dx, n, c, hw, noise, db, dpix, dstrength);
    else if (ng <= 4) epilogue_grad_reduce_kernel<4><<<grid, 256, smem, stream>>>(dx, n, c, hw, noise, db, dpix, dstrength);
    else epilogue_grad_reduce_kernel<8><<<grid, 256, smem, stream>>>(dx, n, c, hw, noise, db, dpix, dstrength);
    SPI_COUNT_LAUNCH(1);
    SPI_LAUNCH_CHECK("epilogue_grad_reduce");
    return SPI_OK;
}
