// Fused tri-plane volume renderer for sm_100a.
//
// Replaces, in ONE forward kernel and ONE backward kernel, the ~60 ATen launches of
//   ImportanceRenderer.forward      eg3d/training/volumetric_rendering/renderer.py:88-140
//   sample_from_planes / grid_sample renderer.py:39-65
//   OSGDecoder.forward              eg3d/training/triplane.py:123-135
//   MipRayMarcher2.run_forward      eg3d/training/volumetric_rendering/ray_marcher.py:25-57
//   sample_stratified / sample_importance / sample_pdf / unify_samples   renderer.py:157-253
//   RaySampler.forward              eg3d/training/volumetric_rendering/ray_sampler.py:24-61
// without materialising any [N, R*D, 32] intermediate (the reference writes ~2.5 GB/img through HBM here).
//
// Data layout in HBM
//   planes   channels-last [N, H, W, 96] fp32: the 3 planes' 32 features of a texel are three consecutive
//            128-byte lines, so one bilinear corner of one plane is exactly one L2 line / one LDG.128 x 8 lanes.
//   rays     computed in-kernel from the 25-float camera (cam2world 4x4 | intrinsics 3x3).
//   jitter   [N, R, Dc], u [N*R, Df]: the two uniform draws of renderer.py:190,237 (injected or device RNG).
//   outputs  feature [N, R, 32], depth [N, R] (unclamped + global min/max), wsum [N, R]; the backward pass re-reads
//            only depths_all [N, R, D] (sorted merged sample depths).
//
// Kernel structure (warp = one ray at a time, lane = one sample; persistent CTAs, decoder weights staged once in smem)
//   pass 1   coarse samples: gather 12 texels -> mean -> FC 32->64 -> softplus -> sigma row of FC 64->33
//   per ray  coarse weights (ray_marcher) -> smoothed pdf/cdf -> inverse-CDF fine depths -> merge by rank
//   pass 2a  fine samples: sigma only
//   per ray  final weights; colour coefficients a_i = (w_{i-1}+w_i)/2; depth, weight sum
//   pass 2b  all merged samples: full decoder, colour accumulated as sum_i a_i * rgb_i in registers, then a
//            shared-memory transpose-reduce across the warp ("compositing without storing colours").
// Compositing is linear in the colours, so recomputing sigma costs +25% decoder FLOPs and zero HBM bytes.
// The backward kernel re-derives everything from depths_all and scatters plane gradients with red.global.add.v4.f32.
#include "common.cuh"
#include "raymarch_mma.cuh"
#include "tc05.cuh"
#include <stdlib.h>

namespace {

constexpr int NF = 32;        // features per plane / decoder input
constexpr int NH = 64;        // decoder hidden
constexpr int NO = 33;        // decoder output (sigma + 32 colour)
constexpr int MAXDC = 128;
constexpr int MAXDF = 128;
constexpr int WARPS = 4;

struct Decoder {               // shared-memory image of the decoder (gains folded in)
    float w1[NH * NF];         // [h][k]
    float b1[NH];
    float w2[NO * NH];         // [o][h]
    float b2[NO + 3];
};

struct RenderParams {
    const float* planes;       // [N, H, W, 96]
    const float* origins;      // [N, R, 3]
    const float* dirs;         // [N, R, 3]
    const float* jitter;       // [N, R, Dc]
    const float* u;            // [N*R, Df]
    const float* w1; const float* b1; const float* w2; const float* b2;   // raw decoder tensors (triplane.py:117-121)
    float w1_gain, w2_gain, b_gain;   // lr_mul/sqrt(32), lr_mul/sqrt(64), lr_mul (FullyConnectedLayer gains)
    float* feat;               // [N, R, 32]
    float* depth;              // [N, R]
    float* wsum;               // [N, R]
    float* depths_all;         // [N, R, D]   (saved for backward / index-parity tests), may be null
    float* sigma_all;          // optional debug output [N, R, D]
    int* minmax;               // 2 ints: ordered-int min / max of all sample depths (ray_marcher.py:50)
    int n, R, H, W, dc, df;    // R = rays per image
    long long plane_bs;        // batch stride of planes in floats (0: one tri-plane set shared by all n views)
    long long gplane_bs;       // batch stride of g_planes (may differ: per-view gradient planes for shared planes avoid RED contention)
    float ray_start, ray_end, box_warp;
    int disparity;
    // backward
    const float* g_feat;       // [N, R, 32]
    const float* g_depth;      // [N, R]
    float* g_planes;           // [N, H, W, 96]  (accumulated)
    float* sc_f; float* sc_hid; float* sc_dpre; float* sc_dout;   // per-sample [S,32] [S,64] [S,64] [S,36] rows for the decoder
                               // weight-gradient GEMMs; null when the decoder is frozen (stage 1)
    // activations kept by the forward pass for the backward pass (tcgen05 kernels only; rows in STORAGE order: coarse sample i ->
    // i, importance sample j -> dc + j): hidden layer [S,64], pre-activation outputs incl. bias [S,36] (32 colours, sigma, 3 pad),
    // gathered features [S,32] (optional), and the storage index of every merged sample [n,R,D]
    float* sv_h; float* sv_o; float* sv_f; unsigned char* sv_src;
};

__device__ __forceinline__ void load_decoder(Decoder* s, const RenderParams& p) {
    for (int i = threadIdx.x; i < NH * NF; i += blockDim.x) s->w1[i] = p.w1[i] * p.w1_gain;
    for (int i = threadIdx.x; i < NH; i += blockDim.x) s->b1[i] = p.b1[i] * p.b_gain;
    for (int i = threadIdx.x; i < NO * NH; i += blockDim.x) s->w2[i] = p.w2[i] * p.w2_gain;
    for (int i = threadIdx.x; i < NO + 3; i += blockDim.x) s->b2[i] = i < NO ? p.b2[i] * p.b_gain : 0.f;
}

// ---------------------------------------------------------------- rays (ray_sampler.py:24-61)
struct Ray { float ox, oy, oz, dx, dy, dz; };

__device__ __forceinline__ Ray make_ray(const float* c, int res, int m) {
    const float fx = c[16], sk = c[17], cx = c[18], fy = c[20], cy = c[21];
    const int i = m / res, j = m % res;
    const float xc = (float)j * (1.f / res) + (0.5f / res);
    const float yc = (float)i * (1.f / res) + (0.5f / res);
    const float xl = (xc - cx + cy * sk / fy - sk * yc / fy) / fx;
    const float yl = (yc - cy) / fy;
    // world = cam2world @ [xl, yl, 1, 1]
    float wx = c[0] * xl + c[1] * yl + c[2] + c[3];
    float wy = c[4] * xl + c[5] * yl + c[6] + c[7];
    float wz = c[8] * xl + c[9] * yl + c[10] + c[11];
    Ray r;
    r.ox = c[3]; r.oy = c[7]; r.oz = c[11];
    float dx = wx - r.ox, dy = wy - r.oy, dz = wz - r.oz;
    float nrm = fmaxf(sqrtf(dx * dx + dy * dy + dz * dz), 1e-12f);   // F.normalize eps
    r.dx = dx / nrm; r.dy = dy / nrm; r.dz = dz / nrm;
    return r;
}

// ---------------------------------------------------------------- tri-plane gather (renderer.py:39-65)
// grid_sample(bilinear, zeros, align_corners=False): pixel = ((g + 1) * size - 1) / 2
struct Corner { int off[4]; float w[4]; };

__device__ __forceinline__ void corners(float gx, float gy, int W, int H, Corner& c) {
    float ix = ((gx + 1.f) * W - 1.f) * 0.5f, iy = ((gy + 1.f) * H - 1.f) * 0.5f;
    float fx0 = floorf(ix), fy0 = floorf(iy);
    int x0 = (int)fx0, y0 = (int)fy0;
    float tx = ix - fx0, ty = iy - fy0;
    float wx[2] = {1.f - tx, tx}, wy[2] = {1.f - ty, ty};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        int xx = x0 + (k & 1), yy = y0 + (k >> 1);
        bool in = (xx >= 0) && (xx < W) && (yy >= 0) && (yy < H);
        c.off[k] = in ? (yy * W + xx) * 96 : -1;
        c.w[k] = wy[k >> 1] * wx[k & 1];
    }
}

__device__ __forceinline__ void plane_coords(float x, float y, float z, float scale, float g[3][2]) {
    x *= scale; y *= scale; z *= scale;
    g[0][0] = x; g[0][1] = y;      // plane 0: (x, y)
    g[1][0] = x; g[1][1] = z;      // plane 1: (x, z)
    g[2][0] = z; g[2][1] = x;      // plane 2: (z, x)
}

__device__ __forceinline__ void gather_features(const float* __restrict__ pl, int W, int H, float x, float y, float z,
                                                float scale, float f[NF]) {
    float g[3][2];
    plane_coords(x, y, z, scale, g);
#pragma unroll
    for (int k = 0; k < NF; k++) f[k] = 0.f;
#pragma unroll
    for (int p = 0; p < 3; p++) {
        Corner c;
        corners(g[p][0], g[p][1], W, H, c);
        float acc[NF];
#pragma unroll
        for (int k = 0; k < NF; k++) acc[k] = 0.f;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (c.off[q] < 0) continue;
            const float4* t = (const float4*)(pl + c.off[q] + p * NF);
            const float w = c.w[q];
#pragma unroll
            for (int v = 0; v < NF / 4; v++) {
                float4 a = __ldg(t + v);
                acc[4 * v + 0] += a.x * w; acc[4 * v + 1] += a.y * w; acc[4 * v + 2] += a.z * w; acc[4 * v + 3] += a.w * w;
            }
        }
#pragma unroll
        for (int k = 0; k < NF; k++) f[k] += acc[k];
    }
#pragma unroll
    for (int k = 0; k < NF; k++) f[k] = f[k] / 3.f;   // .mean(1) over the 3 planes (triplane.py:125)
}

// Warp-cooperative scatter of feature gradients (adjoint of gather_features).  Every lane deposits its sample's
// df[32] and its 12 (texel offset, weight) pairs in per-warp shared memory; the warp then walks the 32x12 tasks with
// 8 lanes per texel, so each texel update is ONE coalesced 128-byte red.global.add.v4.f32 group.
constexpr int SC_DF_STRIDE = 36;
constexpr int SCATTER_FLOATS = 32 * SC_DF_STRIDE + 2 * 32 * 12;     // per-warp scratch, in floats

__device__ __forceinline__ void warp_scatter(float* __restrict__ gp, float* sc, int W, int H, float x, float y, float z,
                                             float scale, const float df[NF], bool valid, int lane) {
    float* s_df = sc;
    int* s_off = (int*)(sc + 32 * SC_DF_STRIDE);
    float* s_w = sc + 32 * SC_DF_STRIDE + 32 * 12;
#pragma unroll
    for (int v = 0; v < NF / 4; v++)
        *(float4*)(s_df + lane * SC_DF_STRIDE + 4 * v) = make_float4(df[4 * v], df[4 * v + 1], df[4 * v + 2], df[4 * v + 3]);
    float g[3][2];
    plane_coords(x, y, z, scale, g);
#pragma unroll
    for (int p = 0; p < 3; p++) {
        Corner c;
        corners(g[p][0], g[p][1], W, H, c);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            s_off[lane * 12 + p * 4 + q] = (valid && c.off[q] >= 0) ? c.off[q] + p * NF : -1;
            s_w[lane * 12 + p * 4 + q] = c.w[q] * (1.f / 3.f);
        }
    }
    __syncwarp();
    const int sub = lane & 7;
    for (int t = lane >> 3; t < 32 * 12; t += 4) {
        const int off = s_off[t];
        if (off < 0) continue;
        const float w = s_w[t];
        const float4 v = *(const float4*)(s_df + (t / 12) * SC_DF_STRIDE + sub * 4);
        red_add_v4(gp + off + sub * 4, make_float4(v.x * w, v.y * w, v.z * w, v.w * w));
    }
    __syncwarp();
}

// ---------------------------------------------------------------- decoder (triplane.py:123-135)
__device__ __forceinline__ void fc1(const Decoder* s, const float f[NF], float hid[NH]) {
#pragma unroll
    for (int h = 0; h < NH; h++) {
        float a = s->b1[h];
        const float4* w = (const float4*)(s->w1 + h * NF);
#pragma unroll
        for (int v = 0; v < NF / 4; v++) {
            float4 ww = w[v];
            a = fmaf(f[4 * v], ww.x, a); a = fmaf(f[4 * v + 1], ww.y, a); a = fmaf(f[4 * v + 2], ww.z, a); a = fmaf(f[4 * v + 3], ww.w, a);
        }
        hid[h] = a;
    }
}

__device__ __forceinline__ float fc2_row(const Decoder* s, const float hid[NH], int o) {
    float a = s->b2[o];
    const float4* w = (const float4*)(s->w2 + o * NH);
#pragma unroll
    for (int v = 0; v < NH / 4; v++) {
        float4 ww = w[v];
        a = fmaf(hid[4 * v], ww.x, a); a = fmaf(hid[4 * v + 1], ww.y, a); a = fmaf(hid[4 * v + 2], ww.z, a); a = fmaf(hid[4 * v + 3], ww.w, a);
    }
    return a;
}

__device__ __forceinline__ float decode_sigma(const Decoder* s, const float f[NF]) {
    float hid[NH];
    fc1(s, f, hid);
#pragma unroll
    for (int h = 0; h < NH; h++) hid[h] = softplus_f(hid[h]);
    return fc2_row(s, hid, 0);
}

__device__ __forceinline__ float rgb_act(float x) { return sigmoid_f(x) * (1.f + 2.f * 0.001f) - 0.001f; }

// ---------------------------------------------------------------- per-ray (warp-cooperative) stages
// All arrays live in per-warp shared memory.  Serial scans are executed redundantly by every lane (no divergence).

__device__ __forceinline__ float coarse_depth(const RenderParams& p, int s, float jit) {
    // sample_stratified (renderer.py:169-192), scalar ray_start/ray_end
    const int dc = p.dc;
    if (p.disparity) {
        float step = 1.f / (float)(dc - 1);
        float t = (s < dc / 2) ? (0.f + step * (float)s) : (1.f - step * (float)(dc - 1 - s));
        t += jit * step;
        return 1.f / (1.f / p.ray_start * (1.f - t) + 1.f / p.ray_end * t);
    }
    float step = (p.ray_end - p.ray_start) / (float)(dc - 1);
    float base = (s < dc / 2) ? (p.ray_start + step * (float)s) : (p.ray_end - step * (float)(dc - 1 - s));   // torch.linspace
    return base + jit * step;
}

// ray_marcher.py:25-45: weights[i] for i in [0, D-1)
__device__ __forceinline__ void warp_weights(const float* d, const float* sig, int D, float* w, int lane) {
    for (int i = lane; i < D - 1; i += 32) {
        float delta = d[i + 1] - d[i];
        float sm = softplus_f((sig[i] + sig[i + 1]) * 0.5f - 1.f);
        w[i] = 1.f - expf(-(sm * delta));            // alpha, turned into weight below
    }
    __syncwarp();
    // exclusive cumprod of (1 - alpha + 1e-10): 32-wide chunks, shuffle scan inside, carry across
    float carry = 1.f;
    for (int base = 0; base < D - 1; base += 32) {
        int i = base + lane;
        float a = (i < D - 1) ? w[i] : 0.f;
        float t = (i < D - 1) ? (1.f - a + 1e-10f) : 1.f;
        float incl = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            float v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl *= v;
        }
        float excl = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) excl = 1.f;
        if (i < D - 1) w[i] = a * (carry * excl);
        carry *= __shfl_sync(0xffffffffu, incl, 31);
    }
    __syncwarp();
}

// sample_importance + sample_pdf (renderer.py:194-253). d[Dc], w[Dc-1] -> fine[Df]; scratch cdf[Dc], bins implicit.
__device__ __forceinline__ void warp_importance(const float* d, const float* w, int Dc, const float* u, int Df, float* cdf,
                                                float* fine, int* inds_out, int lane) {
    const int nw = Dc - 1;        // number of weights
    const int ns = Dc - 3;        // pdf bins ("N_samples_")
    // smoothed weights: max_pool1d(2,1,pad=1) -> Dc values m[j] = max(w[j-1], w[j]); avg_pool1d(2,1) -> Dc-1 values
    // a[j] = (m[j] + m[j+1])/2; +0.01; keep a[1 .. Dc-3]; +1e-5
    float lsum = 0.f;
    for (int k = lane; k < ns; k += 32) {
        int j = k + 1;
        float m0 = fmaxf(w[j - 1], w[j]);
        float m1 = (j + 1 < nw) ? fmaxf(w[j], w[j + 1]) : w[j];
        float a = (m0 + m1) * 0.5f + 0.01f + 1e-5f;
        cdf[k + 1] = a;
        lsum += a;
    }
    float total = warp_sum(lsum);
    __syncwarp();
    // pdf = a / total in parallel, then the SEQUENTIAL cumsum of torch (the order of the additions decides the searchsorted indices)
    for (int k = lane; k < ns; k += 32) cdf[k + 1] = cdf[k + 1] / total;
    __syncwarp();
    if (lane == 0) {
        float run = 0.f;
        cdf[0] = 0.f;
        for (int k = 0; k < ns; k++) { run += cdf[k + 1]; cdf[k + 1] = run; }
    }
    __syncwarp();
    for (int t = lane; t < Df; t += 32) {
        float uu = u[t];
        // searchsorted(cdf[0..ns], uu, right=True): first index with cdf[idx] > uu
        int lo = 0, hi = ns + 1;
        while (lo < hi) { int mid = (lo + hi) >> 1; if (cdf[mid] > uu) hi = mid; else lo = mid + 1; }
        int ind = lo;
        int below = max(ind - 1, 0), above = min(ind, ns);
        float cb = cdf[below], ca = cdf[above];
        float bb = 0.5f * (d[below] + d[below + 1]), ba = 0.5f * (d[above] + d[above + 1]);
        float denom = ca - cb;
        if (denom < 1e-5f) denom = 1.f;
        fine[t] = bb + (uu - cb) / denom * (ba - bb);
        if (inds_out) inds_out[t] = ind;
    }
    __syncwarp();
}

// unify_samples (renderer.py:157-167): stable rank of every sample in cat([coarse, fine]).  pos_c[i], pos_f[j].
__device__ __forceinline__ void warp_merge_ranks(const float* dcs, int Dc, const float* fine, int Df, int* pos_c, int* pos_f, int lane) {
    for (int i = lane; i < Dc; i += 32) {
        float v = dcs[i];
        int r = 0;
        for (int k = 0; k < Dc; k++) r += (dcs[k] < v) || (dcs[k] == v && k < i);
        for (int k = 0; k < Df; k++) r += (fine[k] < v);
        pos_c[i] = r;
    }
    for (int j = lane; j < Df; j += 32) {
        float v = fine[j];
        int r = 0;
        for (int k = 0; k < Dc; k++) r += (dcs[k] <= v);
        for (int k = 0; k < Df; k++) r += (fine[k] < v) || (fine[k] == v && k < j);
        pos_f[j] = r;
    }
    __syncwarp();
}

struct WarpSmem {
    float* dcs;    // [Dc]   coarse depths
    float* sig;    // [D]    sigma: coarse while sampling, then merged order
    float* w;      // [D]    weights / coefficients
    float* cdf;    // [Dc]
    float* fine;   // [Df]
    float* dall;   // [D]    merged sorted depths
    int* pos_c;    // [Dc]
    int* pos_f;    // [Df]
    float* red;    // [32*33] transpose-reduce scratch
};

__device__ __forceinline__ size_t warp_smem_floats(int dc, int df) {
    return (size_t)dc + (dc + df) + (dc + df) + dc + df + (dc + df) + dc + df + 32 * 33;
}

__device__ __forceinline__ WarpSmem carve(float* base, int dc, int df) {
    WarpSmem s;
    s.dcs = base; base += dc;
    s.sig = base; base += dc + df;
    s.w = base; base += dc + df;
    s.cdf = base; base += dc;
    s.fine = base; base += df;
    s.dall = base; base += dc + df;
    s.pos_c = (int*)base; base += dc;
    s.pos_f = (int*)base; base += df;
    s.red = base;
    return s;
}

// final composite scalars (ray_marcher.py:41-55): given w[D-1] -> coefficients a[D] (in place of w), depth, wsum
__device__ __forceinline__ void warp_finalize(const float* dall, float* w, int D, float& depth, float& wsum, int lane) {
    float ws = 0.f, wd = 0.f;
    for (int i = lane; i < D - 1; i += 32) { ws += w[i]; wd += w[i] * (0.5f * (dall[i] + dall[i + 1])); }
    ws = warp_sum(ws); wd = warp_sum(wd);
    wsum = ws;
    depth = wd / ws;                                 // NaN when ws == 0; fixed up by the clamp kernel
    __syncwarp();
    // a_i = (w_{i-1} + w_i) / 2 with w_{-1} = w_{D-1} = 0  (colors_mid = (c_i + c_{i+1})/2)
    float prev_hi = 0.f;                             // carries w[base-1] across chunks
    for (int base = 0; base < D; base += 32) {
        int i = base + lane;
        float wi = (i < D - 1) ? w[i] : 0.f;
        float wl = __shfl_up_sync(0xffffffffu, wi, 1);
        if (lane == 0) wl = prev_hi;
        prev_hi = __shfl_sync(0xffffffffu, wi, 31);
        __syncwarp();
        if (i < D) w[i] = 0.5f * (wl + wi);
    }
    __syncwarp();
}

__device__ __forceinline__ int float_as_ordered(float f) { return __float_as_int(f); }   // depths are positive

// ================================================================= forward kernel
__global__ void __launch_bounds__(WARPS * 32) render_fwd_kernel(RenderParams p) {
    extern __shared__ __align__(16) float smem[];
    Decoder* dec = (Decoder*)smem;
    float* wbase = smem + (sizeof(Decoder) + 15) / 16 * 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int dc = p.dc, df = p.df, D = dc + df;
    WarpSmem s = carve(wbase + warp * warp_smem_floats(dc, df), dc, df);
    load_decoder(dec, p);
    __syncthreads();
    const int R = p.R;
    const long long total = (long long)p.n * R;
    const float scale = 2.f / p.box_warp;
    int lmin = 0x7f800000, lmax = 0;
    for (long long ray = (long long)blockIdx.x * WARPS + warp; ray < total; ray += (long long)gridDim.x * WARPS) {
        const int n = (int)(ray / R);
        const float* pl = p.planes + (size_t)n * p.plane_bs;
        Ray r;
        r.ox = p.origins[ray * 3]; r.oy = p.origins[ray * 3 + 1]; r.oz = p.origins[ray * 3 + 2];
        r.dx = p.dirs[ray * 3]; r.dy = p.dirs[ray * 3 + 1]; r.dz = p.dirs[ray * 3 + 2];
        // ---- pass 1: coarse sigma
        for (int i = lane; i < dc; i += 32) {
            float d = coarse_depth(p, i, p.jitter[ray * dc + i]);
            s.dcs[i] = d;
            float f[NF];
            gather_features(pl, p.W, p.H, r.ox + d * r.dx, r.oy + d * r.dy, r.oz + d * r.dz, scale, f);
            s.sig[i] = decode_sigma(dec, f);
        }
        __syncwarp();
        float depth, wsum;
        if (df > 0) {
            warp_weights(s.dcs, s.sig, dc, s.w, lane);
            warp_importance(s.dcs, s.w, dc, p.u + ray * df, df, s.cdf, s.fine, nullptr, lane);
            warp_merge_ranks(s.dcs, dc, s.fine, df, s.pos_c, s.pos_f, lane);
            // move coarse sigma to merged order (through registers: pos_c is monotone but overlaps)
            float tmp[MAXDC / 32];
#pragma unroll
            for (int k = 0; k < MAXDC / 32; k++) { int i = lane + 32 * k; tmp[k] = (i < dc) ? s.sig[i] : 0.f; }
            __syncwarp();
#pragma unroll
            for (int k = 0; k < MAXDC / 32; k++) { int i = lane + 32 * k; if (i < dc) { s.sig[s.pos_c[i]] = tmp[k]; s.dall[s.pos_c[i]] = s.dcs[i]; } }
            // ---- pass 2a: fine sigma
            for (int j = lane; j < df; j += 32) {
                float d = s.fine[j];
                float f[NF];
                gather_features(pl, p.W, p.H, r.ox + d * r.dx, r.oy + d * r.dy, r.oz + d * r.dz, scale, f);
                int q = s.pos_f[j];
                s.sig[q] = decode_sigma(dec, f);
                s.dall[q] = d;
            }
            __syncwarp();
        } else {
            for (int i = lane; i < dc; i += 32) s.dall[i] = s.dcs[i];
            __syncwarp();
        }
        warp_weights(s.dall, s.sig, D, s.w, lane);
        warp_finalize(s.dall, s.w, D, depth, wsum, lane);
        // ---- pass 2b: colours
        float acc[NF];
#pragma unroll
        for (int k = 0; k < NF; k++) acc[k] = 0.f;
        for (int i = lane; i < D; i += 32) {
            float d = s.dall[i];
            float a = s.w[i];
            float f[NF], hid[NH];
            gather_features(pl, p.W, p.H, r.ox + d * r.dx, r.oy + d * r.dy, r.oz + d * r.dz, scale, f);
            fc1(dec, f, hid);
#pragma unroll
            for (int h = 0; h < NH; h++) hid[h] = softplus_f(hid[h]);
#pragma unroll
            for (int o = 0; o < NF; o++) acc[o] = fmaf(a, rgb_act(fc2_row(dec, hid, o + 1)), acc[o]);
            lmin = min(lmin, float_as_ordered(d)); lmax = max(lmax, float_as_ordered(d));
            if (p.depths_all) p.depths_all[ray * D + i] = d;
            if (p.sigma_all) p.sigma_all[ray * D + i] = s.sig[i];
        }
        // transpose-reduce over lanes through shared memory
#pragma unroll
        for (int k = 0; k < NF; k++) s.red[lane * 33 + k] = acc[k];
        __syncwarp();
        float tot = 0.f;
#pragma unroll
        for (int l = 0; l < 32; l++) tot += s.red[l * 33 + lane];
        p.feat[ray * NF + lane] = tot * 2.f - 1.f;                       // rgb*2-1 (ray_marcher.py:55)
        if (lane == 0) { p.depth[ray] = depth; p.wsum[ray] = wsum; }
        __syncwarp();
    }
    // global min/max of all sample depths (clamp bounds, ray_marcher.py:50)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { lmin = min(lmin, __shfl_xor_sync(0xffffffffu, lmin, o)); lmax = max(lmax, __shfl_xor_sync(0xffffffffu, lmax, o)); }
    if (lane == 0 && lmax != 0) { atomicMin(p.minmax, lmin); atomicMax(p.minmax + 1, lmax); }
}

__global__ void minmax_init_kernel(int* mm) { mm[0] = 0x7f800000; mm[1] = 0; }

// nan_to_num(inf) + clamp(min, max) (ray_marcher.py:49-50); also emits a mask of rays whose depth carries gradient
__global__ void depth_clamp_kernel(float* depth, const int* mm, long long n) {
    const float lo = __int_as_float(mm[0]), hi = __int_as_float(mm[1]);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float d = depth[i];
        if (d != d) d = INFINITY;
        depth[i] = fminf(fmaxf(d, lo), hi);
    }
}

// ================================================================= backward kernel
// Inputs: depths_all (sorted merged depths from the forward), dL/dfeat, dL/ddepth.  Sample positions carry no gradient
// (renderer.py:198,211 no_grad/detach; cameras are constants), so gradients flow only through colours and sigmas of
// the merged samples (coarse samples are re-used by the final composite).
//   B1  recompute sigma_i and p_i = <g_rgb, rgb_i> per sample
//   B2  adjoint of the compositing scan: dL/dsigma_i and colour coefficients a_i
//   B3  per sample: decoder backward -> feature gradient -> red.global.add.v4.f32 into the 12 texels
__global__ void __launch_bounds__(WARPS * 32) render_bwd_kernel(RenderParams p) {
    extern __shared__ __align__(16) float smem[];
    Decoder* dec0 = (Decoder*)smem;
    float* wbase = smem + (sizeof(Decoder) + 15) / 16 * 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int D = p.dc + p.df;
    float* base = wbase + warp * (7 * (size_t)D + 32 + SCATTER_FLOATS);
    float* dall = base; float* sig = dall + D; float* w = sig + D; float* pdot = w + D; float* alpha = pdot + D;
    float* Tarr = alpha + D; float* gmid = Tarr + D; float* gfe = gmid + D; float* scat = gfe + 32;
    load_decoder(dec0, p);
    __syncthreads();
    const int R = p.R;
    const long long total = (long long)p.n * R;
    const float scale = 2.f / p.box_warp;
    const float lo = __int_as_float(p.minmax[0]), hi = __int_as_float(p.minmax[1]);
    for (long long ray = (long long)blockIdx.x * WARPS + warp; ray < total; ray += (long long)gridDim.x * WARPS) {
        const int n = (int)(ray / R);
        const float* pl = p.planes + (size_t)n * p.plane_bs;
        float* gpl = p.g_planes + (size_t)n * p.gplane_bs;
        Ray r;
        r.ox = p.origins[ray * 3]; r.oy = p.origins[ray * 3 + 1]; r.oz = p.origins[ray * 3 + 2];
        r.dx = p.dirs[ray * 3]; r.dy = p.dirs[ray * 3 + 1]; r.dz = p.dirs[ray * 3 + 2];
        gfe[lane] = p.g_feat[ray * NF + lane] * 2.f;                      // through rgb*2-1
        for (int i = lane; i < D; i += 32) dall[i] = p.depths_all[ray * D + i];
        __syncwarp();
        // ---- B1
        for (int i = lane; i < D; i += 32) {
            const Decoder* dec = dec0;
            asm volatile("" : "+l"(dec));
            float d = dall[i];
            float f[NF], hid[NH];
            gather_features(pl, p.W, p.H, r.ox + d * r.dx, r.oy + d * r.dy, r.oz + d * r.dz, scale, f);
            fc1(dec, f, hid);
#pragma unroll
            for (int h = 0; h < NH; h++) hid[h] = softplus_f(hid[h]);
            sig[i] = fc2_row(dec, hid, 0);
            float pd = 0.f;
#pragma unroll
            for (int o = 0; o < NF; o++) pd = fmaf(gfe[o], rgb_act(fc2_row(dec, hid, o + 1)), pd);
            pdot[i] = pd;
        }
        __syncwarp();
        // ---- B2
        for (int i = lane; i < D - 1; i += 32) {
            float delta = dall[i + 1] - dall[i];
            float sm = softplus_f((sig[i] + sig[i + 1]) * 0.5f - 1.f);
            alpha[i] = 1.f - expf(-(sm * delta));
        }
        __syncwarp();
        if (lane == 0) {
            float T = 1.f;
            for (int i = 0; i < D - 1; i++) { float a = alpha[i]; Tarr[i] = T; w[i] = a * T; T *= (1.f - a + 1e-10f); }
            w[D - 1] = 0.f;
        }
        __syncwarp();
        float ws = 0.f, wd = 0.f;
        for (int i = lane; i < D - 1; i += 32) { ws += w[i]; wd += w[i] * (0.5f * (dall[i] + dall[i + 1])); }
        ws = warp_sum(ws); wd = warp_sum(wd);
        const float depth = wd / ws;
        float gd = p.g_depth ? p.g_depth[ray] : 0.f;
        const bool depth_live = (ws > 0.f) && (depth == depth) && (depth >= lo) && (depth <= hi);
        // gw_i = dL/dw_i
        for (int i = lane; i < D - 1; i += 32) {
            float gw = 0.5f * (pdot[i] + pdot[i + 1]);
            if (depth_live) gw += gd * (0.5f * (dall[i] + dall[i + 1]) - depth) / ws;
            gmid[i] = gw;
        }
        __syncwarp();
        if (lane == 0) {                 // suffix scan: dalpha_i = gw_i T_i - (sum_{j>i} gw_j w_j) / (1 - alpha_i + 1e-10)
            float S = 0.f;
            for (int i = D - 2; i >= 0; i--) {
                float gw = gmid[i];
                gmid[i] = gw * Tarr[i] - S / (1.f - alpha[i] + 1e-10f);
                S += gw * w[i];
            }
        }
        __syncwarp();
        for (int i = lane; i < D - 1; i += 32) {
            float delta = dall[i + 1] - dall[i];
            float smid = (sig[i] + sig[i + 1]) * 0.5f - 1.f;
            // dalpha/dsp = delta * exp(-sp*delta) = delta * (1 - alpha);  dsp/dsmid = sigmoid(smid)
            gmid[i] = gmid[i] * delta * (1.f - alpha[i]) * sigmoid_f(smid);
        }
        __syncwarp();
        // ---- B3 (warp-uniform trip count: the scatter is cooperative)
        for (int i0 = 0; i0 < D; i0 += 32) {
            const int i = min(i0 + lane, D - 1);
            const bool valid = (i0 + lane) < D;
            const Decoder* dec = dec0;
            asm volatile("" : "+l"(dec));      // opaque per iteration: keeps ptxas from hoisting 4K weight loads out of the loop
            float d = dall[i];
            const float x = r.ox + d * r.dx, y = r.oy + d * r.dy, z = r.oz + d * r.dz;
            float gs = 0.5f * ((i > 0 ? gmid[i - 1] : 0.f) + (i < D - 1 ? gmid[i] : 0.f));      // dL/dsigma_i
            float a = 0.5f * ((i > 0 ? w[i - 1] : 0.f) + w[i]);                                   // colour coefficient
            float f[NF], hid[NH], dh[NH];
            gather_features(pl, p.W, p.H, x, y, z, scale, f);
            fc1(dec, f, hid);                                              // pre-activations
            const long long srow = ray * D + i;
            if (p.sc_f && valid) {
#pragma unroll
                for (int v = 0; v < NF / 4; v++) ((float4*)(p.sc_f + srow * NF))[v] = make_float4(f[4 * v], f[4 * v + 1], f[4 * v + 2], f[4 * v + 3]);
            }
#pragma unroll
            for (int h = 0; h < NH; h++) hid[h] = softplus_f(hid[h]);
            if (p.sc_hid && valid) {
#pragma unroll
                for (int v = 0; v < NH / 4; v++) ((float4*)(p.sc_hid + srow * NH))[v] = make_float4(hid[4 * v], hid[4 * v + 1], hid[4 * v + 2], hid[4 * v + 3]);
            }
            // output row 0 (sigma)
            {
                const float4* wr = (const float4*)(dec->w2);
#pragma unroll
                for (int v = 0; v < NH / 4; v++) { float4 ww = wr[v]; dh[4 * v] = gs * ww.x; dh[4 * v + 1] = gs * ww.y; dh[4 * v + 2] = gs * ww.z; dh[4 * v + 3] = gs * ww.w; }
                if (p.sc_dout && valid) p.sc_dout[srow * 36] = gs;
            }
#pragma unroll 4
            for (int o = 1; o < NO; o++) {
                float xo = fc2_row(dec, hid, o);
                float so = sigmoid_f(xo);
                float go = gfe[o - 1] * a * (1.f + 2.f * 0.001f) * so * (1.f - so);
                const float4* wr = (const float4*)(dec->w2 + o * NH);
#pragma unroll
                for (int v = 0; v < NH / 4; v++) {
                    float4 ww = wr[v];
                    dh[4 * v] = fmaf(go, ww.x, dh[4 * v]); dh[4 * v + 1] = fmaf(go, ww.y, dh[4 * v + 1]);
                    dh[4 * v + 2] = fmaf(go, ww.z, dh[4 * v + 2]); dh[4 * v + 3] = fmaf(go, ww.w, dh[4 * v + 3]);
                }
                if (p.sc_dout && valid) p.sc_dout[srow * 36 + o] = go;
            }
            if (p.sc_dout && valid) { p.sc_dout[srow * 36 + 33] = 0.f; p.sc_dout[srow * 36 + 34] = 0.f; p.sc_dout[srow * 36 + 35] = 0.f; }
#pragma unroll
            for (int h = 0; h < NH; h++) dh[h] *= (1.f - expf(-hid[h]));   // dpre: softplus'(x) = sigmoid(x) = 1 - exp(-softplus(x))
            if (p.sc_dpre && valid) {
#pragma unroll
                for (int v = 0; v < NH / 4; v++) ((float4*)(p.sc_dpre + srow * NH))[v] = make_float4(dh[4 * v], dh[4 * v + 1], dh[4 * v + 2], dh[4 * v + 3]);
            }
            // df = W1^T dpre
#pragma unroll
            for (int k = 0; k < NF; k++) f[k] = 0.f;
#pragma unroll
            for (int h = 0; h < NH; h++) {
                const float4* wr = (const float4*)(dec->w1 + h * NF);
                const float g = dh[h];
#pragma unroll
                for (int v = 0; v < NF / 4; v++) {
                    float4 ww = wr[v];
                    f[4 * v] = fmaf(g, ww.x, f[4 * v]); f[4 * v + 1] = fmaf(g, ww.y, f[4 * v + 1]);
                    f[4 * v + 2] = fmaf(g, ww.z, f[4 * v + 2]); f[4 * v + 3] = fmaf(g, ww.w, f[4 * v + 3]);
                }
            }
            if (p.g_planes) warp_scatter(gpl, scat, p.W, p.H, x, y, z, scale, f, valid, lane);
        }
        __syncwarp();
    }
}

// ================================================================= run_model (arbitrary points): renderer.py:142-149
struct PointParams {
    const float* planes; const float* coords;     // [N,H,W,96], [N,M,3]
    const float* w1; const float* b1; const float* w2; const float* b2; float w1_gain, w2_gain, b_gain;
    float* rgb; float* sigma;                     // [N,M,32], [N,M]
    const float* g_rgb; const float* g_sigma; float* g_planes;
    float* sc_f; float* sc_hid; float* sc_dpre; float* sc_dout;
    int n, m, H, W; float box_warp;
};

__device__ __forceinline__ void load_decoder_pp(Decoder* s, const PointParams& p) {
    for (int i = threadIdx.x; i < NH * NF; i += blockDim.x) s->w1[i] = p.w1[i] * p.w1_gain;
    for (int i = threadIdx.x; i < NH; i += blockDim.x) s->b1[i] = p.b1[i] * p.b_gain;
    for (int i = threadIdx.x; i < NO * NH; i += blockDim.x) s->w2[i] = p.w2[i] * p.w2_gain;
    for (int i = threadIdx.x; i < NO + 3; i += blockDim.x) s->b2[i] = i < NO ? p.b2[i] * p.b_gain : 0.f;
}

__global__ void __launch_bounds__(128) points_fwd_kernel(PointParams p) {
    __shared__ __align__(16) Decoder dec;
    load_decoder_pp(&dec, p);
    __syncthreads();
    const long long total = (long long)p.n * p.m;
    const float scale = 2.f / p.box_warp;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(i / p.m);
        const float* pl = p.planes + (size_t)n * p.H * p.W * 96;
        float f[NF], hid[NH];
        gather_features(pl, p.W, p.H, p.coords[3 * i], p.coords[3 * i + 1], p.coords[3 * i + 2], scale, f);
        fc1(&dec, f, hid);
#pragma unroll
        for (int h = 0; h < NH; h++) hid[h] = softplus_f(hid[h]);
        p.sigma[i] = fc2_row(&dec, hid, 0);
#pragma unroll 4
        for (int o = 1; o < NO; o++) p.rgb[i * NF + o - 1] = rgb_act(fc2_row(&dec, hid, o));
    }
}

__global__ void __launch_bounds__(128) points_bwd_kernel(PointParams p) {
    __shared__ __align__(16) Decoder dec_s;
    __shared__ __align__(16) float scat_all[4 * SCATTER_FLOATS];
    load_decoder_pp(&dec_s, p);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    float* scat = scat_all + (threadIdx.x >> 5) * SCATTER_FLOATS;
    const long long total = (long long)p.n * p.m;
    const float scale = 2.f / p.box_warp;
    // a warp handles 32 consecutive points of ONE image (m is padded per image by the trip count below)
    const long long per_img = ((long long)p.m + 31) / 32 * 32;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < per_img * p.n; j += (long long)gridDim.x * blockDim.x) {
        const Decoder* decp = &dec_s;
        asm volatile("" : "+l"(decp));      // see render_bwd_kernel: blocks hoisting of the weight loads
        const int n = (int)(j / per_img);
        const long long mloc = j % per_img;
        const bool valid = mloc < p.m;
        const long long i = (long long)n * p.m + (valid ? mloc : p.m - 1);
        const float* pl = p.planes + (size_t)n * p.H * p.W * 96;
        const float x = p.coords[3 * i], y = p.coords[3 * i + 1], z = p.coords[3 * i + 2];
        float f[NF], hid[NH], dh[NH];
        gather_features(pl, p.W, p.H, x, y, z, scale, f);
        fc1(decp, f, hid);
        if (p.sc_f && valid) {
#pragma unroll
            for (int k = 0; k < NF; k++) p.sc_f[i * NF + k] = f[k];
        }
#pragma unroll
        for (int h = 0; h < NH; h++) hid[h] = softplus_f(hid[h]);
        if (p.sc_hid && valid) {
#pragma unroll
            for (int h = 0; h < NH; h++) p.sc_hid[i * NH + h] = hid[h];
        }
        const float gs = p.g_sigma ? p.g_sigma[i] : 0.f;
#pragma unroll
        for (int h = 0; h < NH; h++) dh[h] = gs * decp->w2[h];
        if (p.sc_dout && valid) p.sc_dout[i * 36] = gs;
#pragma unroll 4
        for (int o = 1; o < NO; o++) {
            float so = sigmoid_f(fc2_row(decp, hid, o));
            float go = (p.g_rgb ? p.g_rgb[i * NF + o - 1] : 0.f) * (1.f + 2.f * 0.001f) * so * (1.f - so);
#pragma unroll
            for (int h = 0; h < NH; h++) dh[h] = fmaf(go, decp->w2[o * NH + h], dh[h]);
            if (p.sc_dout && valid) p.sc_dout[i * 36 + o] = go;
        }
        if (p.sc_dout && valid) { p.sc_dout[i * 36 + 33] = 0.f; p.sc_dout[i * 36 + 34] = 0.f; p.sc_dout[i * 36 + 35] = 0.f; }
#pragma unroll
        for (int h = 0; h < NH; h++) dh[h] *= (1.f - expf(-hid[h]));
        if (p.sc_dpre && valid) {
#pragma unroll
            for (int h = 0; h < NH; h++) p.sc_dpre[i * NH + h] = dh[h];
        }
#pragma unroll
        for (int k = 0; k < NF; k++) f[k] = 0.f;
#pragma unroll
        for (int h = 0; h < NH; h++)
#pragma unroll
            for (int k = 0; k < NF; k++) f[k] = fmaf(dh[h], decp->w1[h * NF + k], f[k]);
        if (p.g_planes) warp_scatter(p.g_planes + (size_t)n * p.H * p.W * 96, scat, p.W, p.H, x, y, z, scale, f, valid, lane);
    }
}

// ================================================================= stand-alone pieces (index-exact parity tests)
__global__ void ray_sampler_kernel(const float* cam, int n, int res, float* origins, float* dirs) {
    const long long total = (long long)n * res * res;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int nn = (int)(i / (res * res)), m = (int)(i % (res * res));
        Ray r = make_ray(cam + nn * 25, res, m);
        origins[3 * i] = r.ox; origins[3 * i + 1] = r.oy; origins[3 * i + 2] = r.oz;
        dirs[3 * i] = r.dx; dirs[3 * i + 1] = r.dy; dirs[3 * i + 2] = r.dz;
    }
}

// ray_marcher alone: colors [R,D,C], sigma [R,D], depths [R,D] -> rgb [R,C] (scaled to [-1,1]), depth [R] (unclamped), weights [R,D-1]
__global__ void __launch_bounds__(128) composite_kernel(const float* colors, const float* sigma, const float* depths, int rays, int D,
                                                        int C, float* rgb, float* depth, float* weights, int* minmax) {
    extern __shared__ float sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* d = sm + warp * 3 * D; float* sg = d + D; float* w = sg + D;
    int lmin = 0x7f800000, lmax = 0;
    for (int ray = blockIdx.x * 4 + warp; ray < rays; ray += gridDim.x * 4) {
        for (int i = lane; i < D; i += 32) {
            d[i] = depths[(size_t)ray * D + i]; sg[i] = sigma[(size_t)ray * D + i];
            lmin = min(lmin, __float_as_int(d[i])); lmax = max(lmax, __float_as_int(d[i]));
        }
        __syncwarp();
        warp_weights(d, sg, D, w, lane);
        for (int i = lane; i < D - 1; i += 32) weights[(size_t)ray * (D - 1) + i] = w[i];
        float ws = 0.f, wd = 0.f;
        for (int i = lane; i < D - 1; i += 32) { ws += w[i]; wd += w[i] * (0.5f * (d[i] + d[i + 1])); }
        ws = warp_sum(ws); wd = warp_sum(wd);
        if (lane == 0) depth[ray] = wd / ws;
        for (int c = lane; c < C; c += 32) {
            float acc = 0.f;
            for (int i = 0; i < D - 1; i++)
                acc += w[i] * (0.5f * (colors[((size_t)ray * D + i) * C + c] + colors[((size_t)ray * D + i + 1) * C + c]));
            rgb[(size_t)ray * C + c] = acc * 2.f - 1.f;
        }
        __syncwarp();
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { lmin = min(lmin, __shfl_xor_sync(0xffffffffu, lmin, o)); lmax = max(lmax, __shfl_xor_sync(0xffffffffu, lmax, o)); }
    if (lane == 0 && lmax != 0) { atomicMin(minmax, lmin); atomicMax(minmax + 1, lmax); }
}

// sample_importance alone: depths [R,Dc], weights [R,Dc-1], u [R,Df] -> fine [R,Df], inds [R,Df], cdf [R,Dc-2]
__global__ void __launch_bounds__(128) importance_kernel(const float* depths, const float* weights, const float* u, int rays, int Dc,
                                                         int Df, float* fine, int* inds, float* cdf_out) {
    extern __shared__ float sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* d = sm + warp * (3 * Dc + Df); float* w = d + Dc; float* cdf = w + Dc; float* fo = cdf + Dc;
    for (int ray = blockIdx.x * 4 + warp; ray < rays; ray += gridDim.x * 4) {
        for (int i = lane; i < Dc; i += 32) d[i] = depths[(size_t)ray * Dc + i];
        for (int i = lane; i < Dc - 1; i += 32) w[i] = weights[(size_t)ray * (Dc - 1) + i];
        __syncwarp();
        warp_importance(d, w, Dc, u + (size_t)ray * Df, Df, cdf, fo, inds + (size_t)ray * Df, lane);
        for (int t = lane; t < Df; t += 32) fine[(size_t)ray * Df + t] = fo[t];
        if (cdf_out) for (int k = lane; k < Dc - 2; k += 32) cdf_out[(size_t)ray * (Dc - 2) + k] = cdf[k];
        __syncwarp();
    }
}

// inverse-CDF alone from a given cdf (bit-exact `searchsorted(right=True)` check): bins [R,Nb], cdf [R,Nb-1+... ]
__global__ void inverse_cdf_kernel(const float* bins, const float* cdf, const float* u, int rays, int ncdf, int nbins, int Df,
                                   float* fine, int* inds) {
    const long long total = (long long)rays * Df;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int ray = (int)(i / Df);
        const float* c = cdf + (size_t)ray * ncdf; const float* b = bins + (size_t)ray * nbins;
        float uu = u[i];
        int ns = ncdf - 1;
        int lo = 0, hi = ncdf;
        while (lo < hi) { int mid = (lo + hi) >> 1; if (c[mid] > uu) hi = mid; else lo = mid + 1; }
        int below = max(lo - 1, 0), above = min(lo, ns);
        float denom = c[above] - c[below];
        if (denom < 1e-5f) denom = 1.f;
        fine[i] = b[below] + (uu - c[below]) / denom * (b[above] - b[below]);
        inds[i] = lo;
    }
}

// unify_samples alone: depths of cat([coarse, fine]) -> permutation (int32) such that out[k] = in[perm[k]]
__global__ void __launch_bounds__(128) unify_kernel(const float* dcs, const float* dfs, int rays, int Dc, int Df, int* perm, float* sorted) {
    extern __shared__ float sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* c = sm + warp * 2 * (Dc + Df); float* f = c + Dc; int* pc = (int*)(f + Df); int* pf = pc + Dc;
    for (int ray = blockIdx.x * 4 + warp; ray < rays; ray += gridDim.x * 4) {
        for (int i = lane; i < Dc; i += 32) c[i] = dcs[(size_t)ray * Dc + i];
        for (int i = lane; i < Df; i += 32) f[i] = dfs[(size_t)ray * Df + i];
        __syncwarp();
        warp_merge_ranks(c, Dc, f, Df, pc, pf, lane);
        const size_t o = (size_t)ray * (Dc + Df);
        for (int i = lane; i < Dc; i += 32) { perm[o + pc[i]] = i; sorted[o + pc[i]] = c[i]; }
        for (int i = lane; i < Df; i += 32) { perm[o + pf[i]] = Dc + i; sorted[o + pf[i]] = f[i]; }
        __syncwarp();
    }
}

// ================================================================= tensor-core (mma.sync 3xTF32) renderer, v2
// Same algorithm as render_fwd_kernel / render_bwd_kernel; differences:
//   * the decoder runs on tensor cores per 32-sample tile (raymarch_mma.cuh);
//   * gathered features are cached in a per-warp shared tile, so every sample is gathered ONCE per kernel
//     (forward: pass 2b re-reads pass 1 / 2a features; backward: B3 re-reads B1 features);
//   * pass 2b walks samples in storage order (coarse, then fine) and looks its colour coefficient up by merged rank --
//     compositing is a sum, so the order is free and the A-fragment loads stay conflict-free.
constexpr int GATHER_FLOATS = 2 * 32 * 12;
__device__ __forceinline__ int rup32(int v) { return (v + 31) & ~31; }

__device__ __forceinline__ size_t fwd2_warp_floats(int dc, int df) {
    const int D = dc + df;
    return (size_t)dc /*dcs*/ + rup32(dc) /*sigc*/ + rup32(df) /*sigf*/ + df /*fine*/ + dc /*cdf*/ + 3 * (size_t)D /*dall,sigm,w*/ +
           dc + df /*pos*/ + (size_t)(rup32(dc) + rup32(df)) * mma::FS + GATHER_FLOATS + 32 + 40;
}

__device__ __forceinline__ void gather_to_tile(const float* __restrict__ pl, int W, int H, const Ray& r, float d, float scale, float* row) {
    float f[NF];
    gather_features(pl, W, H, r.ox + d * r.dx, r.oy + d * r.dy, r.oz + d * r.dz, scale, f);
#pragma unroll
    for (int v = 0; v < NF / 4; v++) *(float4*)(row + 4 * v) = make_float4(f[4 * v], f[4 * v + 1], f[4 * v + 2], f[4 * v + 3]);
}

// Warp-cooperative tri-plane gather of a 32-sample tile.  Every lane first publishes its own sample's 12 (texel offset,
// weight) pairs; then 8 lanes serve one sample at a time, each lane owning 4 of the 32 channels, so every texel read is
// ONE coalesced 128-byte line (4 lines per warp instruction instead of 32 scattered ones): 8x fewer L1 wavefronts than
// the lane-per-sample gather.  Result rows go to the shared feature tile (row stride mma::FS).
__device__ __forceinline__ void warp_gather_tile(const float* __restrict__ pl, int W, int H, float x, float y, float z, float scale,
                                                 bool valid, float* Ftile, int* s_off, float* s_w, int lane) {
    float gc[3][2];
    plane_coords(x, y, z, scale, gc);
#pragma unroll
    for (int pp = 0; pp < 3; pp++) {
        Corner c;
        corners(gc[pp][0], gc[pp][1], W, H, c);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            s_off[lane * 12 + pp * 4 + q] = (valid && c.off[q] >= 0) ? c.off[q] + pp * NF : -1;
            s_w[lane * 12 + pp * 4 + q] = c.w[q] * (1.f / 3.f);          // mean over the 3 planes folded into the weights
        }
    }
    __syncwarp();
    const int sub = lane & 7, grp = lane >> 3;
#pragma unroll 2
    for (int k = 0; k < 8; k++) {
        const int smp = grp + 4 * k;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int c = 0; c < 12; c++) {
            const int off = s_off[smp * 12 + c];
            if (off >= 0) {
                const float ww = s_w[smp * 12 + c];
                const float4 v = __ldg((const float4*)(pl + off) + sub);
                acc.x = fmaf(v.x, ww, acc.x); acc.y = fmaf(v.y, ww, acc.y); acc.z = fmaf(v.z, ww, acc.z); acc.w = fmaf(v.w, ww, acc.w);
            }
        }
        *(float4*)(Ftile + smp * mma::FS + sub * 4) = acc;
    }
    __syncwarp();
}

__global__ void __launch_bounds__(WARPS * 32) render_fwd_mma_kernel(RenderParams p) {
    extern __shared__ __align__(16) float smem[];
    mma::DecM* dec = (mma::DecM*)smem;
    float* wbase = smem + (sizeof(mma::DecM) + 15) / 16 * 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int dc = p.dc, df = p.df, D = dc + df, dc32 = rup32(dc), df32 = rup32(df);
    float* b = wbase + warp * ((fwd2_warp_floats(dc, df) + 3) & ~(size_t)3);
    float* F = b; b += (size_t)(dc32 + df32) * mma::FS;           // 16-byte aligned rows first
    float* dcs = b; b += dc;
    float* sigc = b; b += dc32;
    float* sigf = b; b += df32;
    float* fine = b; b += df;
    float* cdf = b; b += dc;
    float* dall = b; b += D;
    float* sigm = b; b += D;
    float* w = b; b += D;
    int* pos_c = (int*)b; b += dc;
    int* pos_f = (int*)b; b += df;
    int* g_off = (int*)b; b += 32 * 12;
    float* g_w = b; b += 32 * 12;
    float* a_tile = b; b += 32;
    float* acc40 = b;
    mma::load_dec(dec, p.w1, p.b1, p.w2, p.b2, p.w1_gain, p.w2_gain, p.b_gain);
    __syncthreads();
    const int R = p.R;
    const long long total = (long long)p.n * R;
    const float scale = 2.f / p.box_warp;
    const int g = lane >> 2, t = lane & 3;
    int lmin = 0x7f800000, lmax = 0;
    for (long long ray = (long long)blockIdx.x * WARPS + warp; ray < total; ray += (long long)gridDim.x * WARPS) {
        const int n = (int)(ray / R);
        const float* pl = p.planes + (size_t)n * p.plane_bs;
        Ray r;
        r.ox = p.origins[ray * 3]; r.oy = p.origins[ray * 3 + 1]; r.oz = p.origins[ray * 3 + 2];
        r.dx = p.dirs[ray * 3]; r.dy = p.dirs[ray * 3 + 1]; r.dz = p.dirs[ray * 3 + 2];
        // ---- pass 1: coarse samples (gather once, sigma on tensor cores)
        for (int rb = 0; rb < dc32; rb += 32) {
            const int i = rb + lane;
            float d = 0.f;
            if (i < dc) { d = coarse_depth(p, i, p.jitter[ray * dc + i]); dcs[i] = d; }
            warp_gather_tile(pl, p.W, p.H, r.ox + d * r.dx, r.oy + d * r.dy, r.oz + d * r.dz, scale, i < dc, F + (size_t)rb * mma::FS, g_off, g_w, lane);
            mma::tile_sigma(dec, F + (size_t)rb * mma::FS, sigc + rb, lane);
        }
        __syncwarp();
        float depth, wsum;
        if (df > 0) {
            warp_weights(dcs, sigc, dc, w, lane);
            warp_importance(dcs, w, dc, p.u + ray * df, df, cdf, fine, nullptr, lane);
            warp_merge_ranks(dcs, dc, fine, df, pos_c, pos_f, lane);
            // ---- pass 2a: fine samples
            for (int rb = 0; rb < df32; rb += 32) {
                const int j = rb + lane;
                const float d = (j < df) ? fine[j] : 0.f;
                warp_gather_tile(pl, p.W, p.H, r.ox + d * r.dx, r.oy + d * r.dy, r.oz + d * r.dz, scale, j < df, F + (size_t)(dc32 + rb) * mma::FS, g_off, g_w, lane);
                mma::tile_sigma(dec, F + (size_t)(dc32 + rb) * mma::FS, sigf + rb, lane);
            }
            __syncwarp();
            for (int i = lane; i < dc; i += 32) { sigm[pos_c[i]] = sigc[i]; dall[pos_c[i]] = dcs[i]; }
            for (int j = lane; j < df; j += 32) { sigm[pos_f[j]] = sigf[j]; dall[pos_f[j]] = fine[j]; }
        } else {
            for (int i = lane; i < dc; i += 32) { sigm[i] = sigc[i]; dall[i] = dcs[i]; pos_c[i] = i; }
        }
        __syncwarp();
        warp_weights(dall, sigm, D, w, lane);
        warp_finalize(dall, w, D, depth, wsum, lane);          // w[] now holds the colour coefficients a_q
        // ---- pass 2b: colours, storage order, features from the shared tile
        for (int o = lane; o < 40; o += 32) acc40[o] = 0.f;
        for (int rb = 0; rb < dc32 + df32; rb += 32) {
            const int sr = rb + lane;
            const bool coarse = sr < dc32;
            const bool valid = coarse ? (sr < dc) : (sr - dc32 < df);
            a_tile[lane] = valid ? w[coarse ? pos_c[sr] : pos_f[sr - dc32]] : 0.f;
            __syncwarp();
            mma::tile_color(dec, F + (size_t)rb * mma::FS, a_tile, acc40, lane);
        }
        p.feat[ray * NF + lane] = acc40[lane + 1] * 2.f - 1.f;          // rgb*2-1 (ray_marcher.py:55)
        for (int i = lane; i < D; i += 32) {
            const float d = dall[i];
            lmin = min(lmin, float_as_ordered(d)); lmax = max(lmax, float_as_ordered(d));
            if (p.depths_all) p.depths_all[ray * D + i] = d;
            if (p.sigma_all) p.sigma_all[ray * D + i] = sigm[i];
        }
        if (lane == 0) { p.depth[ray] = depth; p.wsum[ray] = wsum; }
        __syncwarp();
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { lmin = min(lmin, __shfl_xor_sync(0xffffffffu, lmin, o)); lmax = max(lmax, __shfl_xor_sync(0xffffffffu, lmax, o)); }
    if (lane == 0 && lmax != 0) { atomicMin(p.minmax, lmin); atomicMax(p.minmax + 1, lmax); }
}

#include "raymarch_tc.cuh"
#include "raymarch_tc_bwd.cuh"

// scatter of a 32-row feature-gradient tile (row stride mma::FS); lane = row; `valid` rows only
__device__ __forceinline__ void warp_scatter_tile(float* __restrict__ gp, const float* dft, int* s_off, float* s_w, int W, int H, float x,
                                                  float y, float z, float scale, bool valid, int lane) {
    float gc[3][2];
    plane_coords(x, y, z, scale, gc);
#pragma unroll
    for (int pp = 0; pp < 3; pp++) {
        Corner c;
        corners(gc[pp][0], gc[pp][1], W, H, c);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            s_off[lane * 12 + pp * 4 + q] = (valid && c.off[q] >= 0) ? c.off[q] + pp * NF : -1;
            s_w[lane * 12 + pp * 4 + q] = c.w[q] * (1.f / 3.f);
        }
    }
    __syncwarp();
    const int sub = lane & 7;
    for (int tt = lane >> 3; tt < 32 * 12; tt += 4) {
        const int off = s_off[tt];
        if (off < 0) continue;
        const float ww = s_w[tt];
        const float4 v = *(const float4*)(dft + (tt / 12) * mma::FS + sub * 4);
        red_add_v4(gp + off + sub * 4, make_float4(v.x * ww, v.y * ww, v.z * ww, v.w * ww));
    }
    __syncwarp();
}

__device__ __forceinline__ size_t bwd2_warp_floats(int D) { return (size_t)rup32(D) * mma::FS + 32 * mma::FS + 7 * (size_t)rup32(D) + 32 + 2 * 32 * 12; }

__global__ void __launch_bounds__(WARPS * 32) render_bwd_mma_kernel(RenderParams p) {
    extern __shared__ __align__(16) float smem[];
    mma::DecM* dec = (mma::DecM*)smem;
    mma::DecMBwd* decb = (mma::DecMBwd*)(smem + (sizeof(mma::DecM) + 15) / 16 * 4);
    float* wbase = (float*)decb + (sizeof(mma::DecMBwd) + 15) / 16 * 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int D = p.dc + p.df, D32 = rup32(D);
    float* b = wbase + warp * ((bwd2_warp_floats(D) + 3) & ~(size_t)3);
    float* F = b; b += (size_t)D32 * mma::FS;
    float* DFt = b; b += 32 * mma::FS;
    float* dall = b; b += D32; float* sig = b; b += D32; float* w = b; b += D32; float* pdot = b; b += D32;
    float* alpha = b; b += D32; float* Tarr = b; b += D32; float* gmid = b; b += D32; float* gfe = b; b += 32;
    int* s_off = (int*)b; b += 32 * 12; float* s_w = b;
    mma::load_dec(dec, p.w1, p.b1, p.w2, p.b2, p.w1_gain, p.w2_gain, p.b_gain);
    mma::load_dec_bwd(decb, p.w2, p.w2_gain);
    __syncthreads();
    const int R = p.R;
    const long long total = (long long)p.n * R;
    const float scale = 2.f / p.box_warp;
    const float lo = __int_as_float(p.minmax[0]), hi = __int_as_float(p.minmax[1]);
    const int g = lane >> 2, t = lane & 3;
    for (long long ray = (long long)blockIdx.x * WARPS + warp; ray < total; ray += (long long)gridDim.x * WARPS) {
        const int n = (int)(ray / R);
        const float* pl = p.planes + (size_t)n * p.plane_bs;
        float* gpl = p.g_planes ? p.g_planes + (size_t)n * p.gplane_bs : nullptr;
        Ray r;
        r.ox = p.origins[ray * 3]; r.oy = p.origins[ray * 3 + 1]; r.oz = p.origins[ray * 3 + 2];
        r.dx = p.dirs[ray * 3]; r.dy = p.dirs[ray * 3 + 1]; r.dz = p.dirs[ray * 3 + 2];
        gfe[lane] = p.g_feat[ray * NF + lane] * 2.f;
        for (int i = lane; i < D; i += 32) dall[i] = p.depths_all[ray * D + i];
        __syncwarp();
        // ---- B1: gather once; sigma_i and p_i = <g_rgb, rgb_i>
        for (int rb = 0; rb < D32; rb += 32) {
            const int i = rb + lane;
            const float dd = (i < D) ? dall[i] : 0.f;
            warp_gather_tile(pl, p.W, p.H, r.ox + dd * r.dx, r.oy + dd * r.dy, r.oz + dd * r.dz, scale, i < D, F + (size_t)rb * mma::FS, s_off, s_w, lane);
            mma::tile_sigma_pdot(dec, F + (size_t)rb * mma::FS, gfe, sig + rb, pdot + rb, lane);     // arrays are padded to D32
        }
        __syncwarp();
        // ---- B2: compositing adjoint (same arithmetic as render_bwd_kernel)
        for (int i = lane; i < D - 1; i += 32) {
            float delta = dall[i + 1] - dall[i];
            float sm = softplus_f((sig[i] + sig[i + 1]) * 0.5f - 1.f);
            alpha[i] = 1.f - expf(-(sm * delta));
        }
        __syncwarp();
        if (lane == 0) {
            float T = 1.f;
            for (int i = 0; i < D - 1; i++) { float a = alpha[i]; Tarr[i] = T; w[i] = a * T; T *= (1.f - a + 1e-10f); }
            w[D - 1] = 0.f;
        }
        __syncwarp();
        float ws = 0.f, wd = 0.f;
        for (int i = lane; i < D - 1; i += 32) { ws += w[i]; wd += w[i] * (0.5f * (dall[i] + dall[i + 1])); }
        ws = warp_sum(ws); wd = warp_sum(wd);
        const float depth = wd / ws;
        const float gd = p.g_depth ? p.g_depth[ray] : 0.f;
        const bool depth_live = (ws > 0.f) && (depth == depth) && (depth >= lo) && (depth <= hi);
        for (int i = lane; i < D - 1; i += 32) {
            float gw = 0.5f * (pdot[i] + pdot[i + 1]);
            if (depth_live) gw += gd * (0.5f * (dall[i] + dall[i + 1]) - depth) / ws;
            gmid[i] = gw;
        }
        __syncwarp();
        if (lane == 0) {
            float S = 0.f;
            for (int i = D - 2; i >= 0; i--) {
                float gw = gmid[i];
                gmid[i] = gw * Tarr[i] - S / (1.f - alpha[i] + 1e-10f);
                S += gw * w[i];
            }
        }
        __syncwarp();
        for (int i = lane; i < D - 1; i += 32) {
            float delta = dall[i + 1] - dall[i];
            float smid = (sig[i] + sig[i + 1]) * 0.5f - 1.f;
            gmid[i] = gmid[i] * delta * (1.f - alpha[i]) * sigmoid_f(smid);
        }
        __syncwarp();
        // ---- B3: decoder backward per tile, features from the shared tile
        for (int rb = 0; rb < D32; rb += 32) {
            float hid[2][8][4];
            mma::fc1(dec, F + (size_t)rb * mma::FS, hid, lane);
            mma::softplus_inplace(hid);
            float dout[2][5][4];
            mma::fc2<5>(dec, hid, dout, lane);
#pragma unroll
            for (int mt = 0; mt < 2; mt++)
#pragma unroll
                for (int hh = 0; hh < 2; hh++) {
                    const int row = rb + 16 * mt + g + 8 * hh;
                    const bool valid = row < D;
                    float gs = 0.f, a = 0.f;
                    if (valid) {
                        gs = 0.5f * ((row > 0 ? gmid[row - 1] : 0.f) + (row < D - 1 ? gmid[row] : 0.f));
                        a = 0.5f * ((row > 0 ? w[row - 1] : 0.f) + w[row]);
                    }
#pragma unroll
                    for (int nn = 0; nn < 5; nn++)
#pragma unroll
                        for (int jj = 0; jj < 2; jj++) {
                            const int o = 8 * nn + 2 * t + jj;
                            float v = 0.f;
                            if (valid) {
                                if (o == 0) v = gs;
                                else if (o <= 32) { float so = mma::sigmoid_fast(dout[mt][nn][2 * hh + jj]); v = gfe[o - 1] * a * (1.f + 2.f * 0.001f) * so * (1.f - so); }
                            }
                            dout[mt][nn][2 * hh + jj] = v;
                        }
                }
            const long long srow0 = ray * D + rb;
            if (p.sc_dout) {
#pragma unroll
                for (int mt = 0; mt < 2; mt++)
#pragma unroll
                    for (int hh = 0; hh < 2; hh++) {
                        const int rr = 16 * mt + g + 8 * hh;
                        if (rb + rr < D) {
#pragma unroll
                            for (int nn = 0; nn < 5; nn++) {
                                const int o = 8 * nn + 2 * t;
                                if (o < 36) *(float2*)(p.sc_dout + (srow0 + rr) * 36 + o) = make_float2(dout[mt][nn][2 * hh], dout[mt][nn][2 * hh + 1]);
                            }
                        }
                    }
            }
            if (p.sc_hid) {
#pragma unroll
                for (int mt = 0; mt < 2; mt++)
#pragma unroll
                    for (int hh = 0; hh < 2; hh++) {
                        const int rr = 16 * mt + g + 8 * hh;
                        if (rb + rr < D) {
#pragma unroll
                            for (int nt = 0; nt < 8; nt++)
                                *(float2*)(p.sc_hid + (srow0 + rr) * NH + 8 * nt + 2 * t) = make_float2(hid[mt][nt][2 * hh], hid[mt][nt][2 * hh + 1]);
                        }
                    }
            }
            float dh[2][8][4];
            mma::bwd_fc2(decb, dout, dh, lane);
#pragma unroll
            for (int mt = 0; mt < 2; mt++)
#pragma unroll
                for (int nt = 0; nt < 8; nt++)
#pragma unroll
                    for (int j = 0; j < 4; j++) dh[mt][nt][j] *= (1.f - mma::ex2_ftz(-1.4426950408889634f * hid[mt][nt][j]));      // softplus' = 1 - exp(-softplus)
            if (p.sc_dpre) {
#pragma unroll
                for (int mt = 0; mt < 2; mt++)
#pragma unroll
                    for (int hh = 0; hh < 2; hh++) {
                        const int rr = 16 * mt + g + 8 * hh;
                        if (rb + rr < D) {
#pragma unroll
                            for (int nt = 0; nt < 8; nt++)
                                *(float2*)(p.sc_dpre + (srow0 + rr) * NH + 8 * nt + 2 * t) = make_float2(dh[mt][nt][2 * hh], dh[mt][nt][2 * hh + 1]);
                        }
                    }
            }
            if (p.sc_f) {
                for (int e = lane; e < 32 * 8; e += 32) {        // 32 rows x 8 float4
                    const int rr = e >> 3, v = e & 7;
                    if (rb + rr < D) *(float4*)(p.sc_f + (srow0 + rr) * NF + 4 * v) = *(const float4*)(F + (size_t)(rb + rr) * mma::FS + 4 * v);
                }
            }
            if (gpl) {
                float dfr[2][4][4];
                mma::bwd_fc1(dec, dh, dfr, lane);
#pragma unroll
                for (int mt = 0; mt < 2; mt++)
#pragma unroll
                    for (int hh = 0; hh < 2; hh++)
#pragma unroll
                        for (int nt = 0; nt < 4; nt++)
                            *(float2*)(DFt + (16 * mt + g + 8 * hh) * mma::FS + 8 * nt + 2 * t) = make_float2(dfr[mt][nt][2 * hh], dfr[mt][nt][2 * hh + 1]);
                __syncwarp();
                const int i = min(rb + lane, D - 1);
                const float d = dall[i];
                warp_scatter_tile(gpl, DFt, s_off, s_w, p.W, p.H, r.ox + d * r.dx, r.oy + d * r.dy, r.oz + d * r.dz, scale, rb + lane < D, lane);
            }
        }
        __syncwarp();
    }
}

size_t fwd2_smem_bytes(int dc, int df) {
    int D = dc + df;
    size_t per = (size_t)dc + ((dc + 31) & ~31) + ((df + 31) & ~31) + df + dc + 3 * (size_t)D + dc + df + (size_t)(((dc + 31) & ~31) + ((df + 31) & ~31)) * mma::FS + GATHER_FLOATS + 32 + 40;
    per = (per + 3) & ~(size_t)3;
    return ((sizeof(mma::DecM) + 15) / 16) * 16 + WARPS * per * sizeof(float);
}
size_t bwd2_smem_bytes(int D) {
    size_t D32 = (D + 31) & ~31;
    size_t per = D32 * mma::FS + 32 * mma::FS + 7 * D32 + 32 + 2 * 32 * 12;
    per = (per + 3) & ~(size_t)3;
    return ((sizeof(mma::DecM) + 15) / 16) * 16 + ((sizeof(mma::DecMBwd) + 15) / 16) * 16 + WARPS * per * sizeof(float);
}

size_t fwd_smem_bytes(int dc, int df) {
    size_t per_warp = (size_t)dc + (dc + df) + (dc + df) + dc + df + (dc + df) + dc + df + 32 * 33;
    return ((sizeof(Decoder) + 15) / 16) * 16 + WARPS * per_warp * sizeof(float);
}
size_t bwd_smem_bytes(int D) { return ((sizeof(Decoder) + 15) / 16) * 16 + WARPS * (7 * (size_t)D + 32 + SCATTER_FLOATS) * sizeof(float); }

int check_render(const RenderParams& p) {
    SPI_CHECK_ARG(p.planes && p.origins && p.dirs, "render: null pointer");
    SPI_CHECK_ARG(p.dc >= 4 && p.dc <= MAXDC, "render: depth_resolution must be in [4, %d]", MAXDC);
    SPI_CHECK_ARG(p.df >= 0 && p.df <= MAXDF, "render: depth_resolution_importance must be in [0, %d]", MAXDF);
    SPI_CHECK_ARG(p.n >= 0 && p.R >= 1, "render: bad shape");
    SPI_CHECK_ARG(((uintptr_t)p.planes & 15) == 0, "render: planes must be 16-byte aligned");
    return SPI_OK;
}

}  // namespace

#define FILL_DECODER(p)                                                                             \
    p.w1 = w1; p.b1 = b1; p.w2 = w2; p.b2 = b2; p.w1_gain = lr_mul * 0.17677669529663687f; p.w2_gain = lr_mul * 0.125f; \
    p.b_gain = lr_mul;

static int render_forward_impl(const float* planes, const float* origins, const float* dirs, const float* jitter, const float* u,
                               const float* w1, const float* b1, const float* w2, const float* b2, float lr_mul, float* feat,
                               float* depth, float* wsum, float* depths_all, float* sigma_all, int* minmax, int n,
                               int rays_per_image, long long plane_batch_stride, int plane_h, int plane_w, int dc, int df, float ray_start, float ray_end,
                               float box_warp, int disparity, float* sv_h, float* sv_o, float* sv_f, unsigned char* sv_src, cudaStream_t stream) {
    RenderParams p;
    memset(&p, 0, sizeof(p));
    p.sv_h = sv_h; p.sv_o = sv_o; p.sv_f = sv_f; p.sv_src = sv_src;
    p.planes = planes; p.origins = origins; p.dirs = dirs; p.jitter = jitter; p.u = u; FILL_DECODER(p);
    p.feat = feat; p.depth = depth; p.wsum = wsum; p.depths_all = depths_all; p.sigma_all = sigma_all; p.minmax = minmax;
    p.n = n; p.R = rays_per_image; p.H = plane_h; p.W = plane_w; p.dc = dc; p.df = df; p.plane_bs = plane_batch_stride;
    p.ray_start = ray_start; p.ray_end = ray_end; p.box_warp = box_warp; p.disparity = disparity;
    int rc = check_render(p);
    if (rc) return rc;
    SPI_CHECK_ARG(jitter && (df == 0 || u) && feat && depth && wsum && minmax, "render_forward: null pointer");
    if (n == 0) return SPI_OK;
    const bool simt = getenv("SPI_RENDER_SIMT") != nullptr;      // v1 SIMT kernels kept for A/B comparison
    long long rays = (long long)n * rays_per_image;
    // tcgen05 kernel: one 128-sample decoder tile per 32-sample round, D2 of every round resident in tensor memory (<= 6 rounds)
    const int tc_rounds = tcr::rounds_of(dc) + tcr::rounds_of(df);
    if (!simt && getenv("SPI_RENDER_MMA") == nullptr && tc_rounds <= tcr::MAX_ROUNDS) {
        const size_t smem_tc = tcr::smem_bytes(dc, df);
        cudaFuncSetAttribute(tcr::render_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tc);
        // up to 2 rounds: 256 TMEM columns per CTA, two CTAs per SM; more: all 512 columns, one CTA per SM
        long long groups = (rays + 3) / 4, cap = (tc_rounds <= 2 ? 2LL : 1LL) * spi_num_sms();
        int grid = (int)(groups < cap ? groups : cap);
        minmax_init_kernel<<<1, 1, 0, stream>>>(minmax);
        tcr::render_fwd_tc_kernel<<<grid, tcr::TC_THREADS, smem_tc, stream>>>(p, spi_tc_err_flag());
        depth_clamp_kernel<<<cdiv(rays, 256) > 1184 ? 1184 : cdiv(rays, 256), 256, 0, stream>>>(depth, minmax, rays);
        SPI_COUNT_LAUNCH(3);
        SPI_LAUNCH_CHECK("render_forward");
        return SPI_OK;
    }
    size_t smem = simt ? fwd_smem_bytes(dc, df) : fwd2_smem_bytes(dc, df);
    SPI_CHECK_ARG(smem <= 227 * 1024, "render_forward: %d+%d samples per ray need %zu B of shared memory (> 227 KB)", dc, df, smem);
    auto kern = simt ? render_fwd_kernel : render_fwd_mma_kernel;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, WARPS * 32, smem);
    if (occ < 1) occ = 1;
    long long want = (rays + WARPS - 1) / WARPS;
    long long cap = (long long)spi_num_sms() * occ;
    int grid = (int)(want < cap ? want : cap);
    minmax_init_kernel<<<1, 1, 0, stream>>>(minmax);
    kern<<<grid, WARPS * 32, smem, stream>>>(p);
    depth_clamp_kernel<<<cdiv(rays, 256) > 1184 ? 1184 : cdiv(rays, 256), 256, 0, stream>>>(depth, minmax, rays);
    SPI_COUNT_LAUNCH(3);
    SPI_LAUNCH_CHECK("render_forward");
    return SPI_OK;
}

// 1 when spi_render_forward_keep / spi_render_backward_kept run on the tcgen05 kernels for these depth resolutions
extern "C" int spi_render_keeps_activations(int dc, int df) {
    return (getenv("SPI_RENDER_SIMT") == nullptr && getenv("SPI_RENDER_MMA") == nullptr && getenv("SPI_RENDER_RECOMPUTE") == nullptr &&
            tcr::rounds_of(dc) + tcr::rounds_of(df) <= tcr::MAX_ROUNDS && dc + df <= tcb::MAXD) ? 1 : 0;
}

extern "C" int spi_render_forward(const float* planes, const float* origins, const float* dirs, const float* jitter, const float* u,
                                  const float* w1, const float* b1, const float* w2, const float* b2, float lr_mul, float* feat,
                                  float* depth, float* wsum, float* depths_all, float* sigma_all, int* minmax, int n,
                                  int rays_per_image, long long plane_batch_stride, int plane_h, int plane_w, int dc, int df, float ray_start, float ray_end,
                                  float box_warp, int disparity, cudaStream_t stream) {
    return render_forward_impl(planes, origins, dirs, jitter, u, w1, b1, w2, b2, lr_mul, feat, depth, wsum, depths_all, sigma_all, minmax, n,
                               rays_per_image, plane_batch_stride, plane_h, plane_w, dc, df, ray_start, ray_end, box_warp, disparity, nullptr, nullptr,
                               nullptr, nullptr, stream);
}

extern "C" int spi_render_forward_keep(const float* planes, const float* origins, const float* dirs, const float* jitter, const float* u,
                                       const float* w1, const float* b1, const float* w2, const float* b2, float lr_mul, float* feat,
                                       float* depth, float* wsum, float* depths_all, int* minmax, int n, int rays_per_image,
                                       long long plane_batch_stride, int plane_h, int plane_w, int dc, int df, float ray_start, float ray_end,
                                       float box_warp, int disparity, float* sv_h, float* sv_o, float* sv_f, unsigned char* sv_src,
                                       cudaStream_t stream) {
    SPI_CHECK_ARG(sv_h && sv_o && sv_src, "render_forward_keep: null activation buffer");
    SPI_CHECK_ARG(spi_render_keeps_activations(dc, df), "render_forward_keep: %d+%d samples per ray are outside the tcgen05 kernels' range", dc, df);
    return render_forward_impl(planes, origins, dirs, jitter, u, w1, b1, w2, b2, lr_mul, feat, depth, wsum, depths_all, nullptr, minmax, n,
                               rays_per_image, plane_batch_stride, plane_h, plane_w, dc, df, ray_start, ray_end, box_warp, disparity, sv_h, sv_o, sv_f,
                               sv_src, stream);
}

static int render_backward_impl(const float* planes, const float* origins, const float* dirs, const float* depths_all,
                                const int* minmax, const float* w1, const float* b1, const float* w2, const float* b2,
                                float lr_mul, const float* g_feat, const float* g_depth, float* g_planes, float* sc_f,
                                float* sc_hid, float* sc_dpre, float* sc_dout, int n, int rays_per_image,
                                long long plane_batch_stride, long long grad_batch_stride, int plane_h, int plane_w, int dc, int df, float box_warp,
                                const float* sv_h, const float* sv_o, const unsigned char* sv_src, cudaStream_t stream) {
    RenderParams p;
    memset(&p, 0, sizeof(p));
    p.sv_h = (float*)sv_h; p.sv_o = (float*)sv_o; p.sv_src = (unsigned char*)sv_src;
    p.gplane_bs = grad_batch_stride;
    p.planes = planes; p.origins = origins; p.dirs = dirs; FILL_DECODER(p);
    p.depths_all = (float*)depths_all; p.minmax = (int*)minmax; p.g_feat = g_feat; p.g_depth = g_depth; p.g_planes = g_planes;
    p.sc_f = sc_f; p.sc_hid = sc_hid; p.sc_dpre = sc_dpre; p.sc_dout = sc_dout;
    p.n = n; p.R = rays_per_image; p.H = plane_h; p.W = plane_w; p.dc = dc; p.df = df; p.box_warp = box_warp;
    p.plane_bs = plane_batch_stride;
    int rc = check_render(p);
    if (rc) return rc;
    SPI_CHECK_ARG(depths_all && minmax && g_feat, "render_backward: null pointer");
    if (n == 0) return SPI_OK;
    const bool simt = getenv("SPI_RENDER_SIMT") != nullptr;
    long long rays = (long long)n * rays_per_image;
    // tcgen05 kernel: hidden layer and outputs of the last two 32-sample rounds resident in tensor memory (<= 128 merged samples per ray)
    if (!simt && getenv("SPI_RENDER_MMA") == nullptr && dc + df <= tcb::MAXD) {
        cudaFuncSetAttribute(tcb::render_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tcb::SMEM_BYTES);
        long long groups = (rays + 3) / 4, cap = spi_num_sms();
        int grid = (int)(groups < cap ? groups : cap);
        tcb::render_bwd_tc_kernel<<<grid, tcb::THREADS, tcb::SMEM_BYTES, stream>>>(p, spi_tc_err_flag());
        SPI_COUNT_LAUNCH(1);
        SPI_LAUNCH_CHECK("render_backward");
        return SPI_OK;
    }
    size_t smem = simt ? bwd_smem_bytes(dc + df) : bwd2_smem_bytes(dc + df);
    SPI_CHECK_ARG(smem <= 227 * 1024, "render_backward: %d samples per ray need %zu B of shared memory (> 227 KB)", dc + df, smem);
    auto kern = simt ? render_bwd_kernel : render_bwd_mma_kernel;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, WARPS * 32, smem);
    if (occ < 1) occ = 1;
    long long want = (rays + WARPS - 1) / WARPS;
    long long cap = (long long)spi_num_sms() * occ;
    int grid = (int)(want < cap ? want : cap);
    kern<<<grid, WARPS * 32, smem, stream>>>(p);
    SPI_COUNT_LAUNCH(1);
    SPI_LAUNCH_CHECK("render_backward");
    return SPI_OK;
}

extern "C" int spi_render_backward(const float* planes, const float* origins, const float* dirs, const float* depths_all,
                                   const int* minmax, const float* w1, const float* b1, const float* w2, const float* b2,
                                   float lr_mul, const float* g_feat, const float* g_depth, float* g_planes, float* sc_f,
                                   float* sc_hid, float* sc_dpre, float* sc_dout, int n, int rays_per_image,
                                   long long plane_batch_stride, long long grad_batch_stride, int plane_h, int plane_w, int dc, int df, float box_warp,
                                   cudaStream_t stream) {
    return render_backward_impl(planes, origins, dirs, depths_all, minmax, w1, b1, w2, b2, lr_mul, g_feat, g_depth, g_planes, sc_f, sc_hid, sc_dpre,
                                sc_dout, n, rays_per_image, plane_batch_stride, grad_batch_stride, plane_h, plane_w, dc, df, box_warp, nullptr, nullptr,
                                nullptr, stream);
}

extern "C" int spi_render_backward_kept(const float* planes, const float* origins, const float* dirs, const float* depths_all,
                                        const int* minmax, const float* w1, const float* b1, const float* w2, const float* b2,
                                        float lr_mul, const float* g_feat, const float* g_depth, float* g_planes, float* sc_dpre,
                                        float* sc_dout, int n, int rays_per_image, long long plane_batch_stride, long long grad_batch_stride,
                                        int plane_h, int plane_w, int dc, int df, float box_warp, const float* sv_h, const float* sv_o,
                                        const unsigned char* sv_src, cudaStream_t stream) {
    SPI_CHECK_ARG(sv_h && sv_o && sv_src, "render_backward_kept: null activation buffer");
    SPI_CHECK_ARG(spi_render_keeps_activations(dc, df), "render_backward_kept: %d+%d samples per ray are outside the tcgen05 kernels' range", dc, df);
    return render_backward_impl(planes, origins, dirs, depths_all, minmax, w1, b1, w2, b2, lr_mul, g_feat, g_depth, g_planes, nullptr, nullptr, sc_dpre,
                                sc_dout, n, rays_per_image, plane_batch_stride, grad_batch_stride, plane_h, plane_w, dc, df, box_warp, sv_h, sv_o, sv_src,
                                stream);
}

extern "C" int spi_points_forward(const float* planes, const float* coords, const float* w1, const float* b1, const float* w2,
                                  const float* b2, float lr_mul, float* rgb, float* sigma, int n, int m, int plane_h, int plane_w,
                                  float box_warp, cudaStream_t stream) {
    SPI_CHECK_ARG(planes && coords && rgb && sigma, "points_forward: null pointer");
    PointParams p;
    memset(&p, 0, sizeof(p));
    p.planes = planes; p.coords = coords; FILL_DECODER(p);
    p.rgb = rgb; p.sigma = sigma; p.n = n; p.m = m; p.H = plane_h; p.W = plane_w; p.box_warp = box_warp;
    long long total = (long long)n * m;
    if (total == 0) return SPI_OK;
    long long cap = (long long)spi_num_sms() * 4, want = (total + 127) / 128;
    points_fwd_kernel<<<(int)(want < cap ? want : cap), 128, 0, stream>>>(p);
    SPI_COUNT_LAUNCH(1);
    SPI_LAUNCH_CHECK("points_forward");
    return SPI_OK;
}

extern "C" int spi_points_backward(const float* planes, const float* coords, const float* w1, const float* b1, const float* w2,
                                   const float* b2, float lr_mul, const float* g_rgb, const float* g_sigma, float* g_planes, float* sc_f,
                                   float* sc_hid, float* sc_dpre, float* sc_dout, int n, int m, int plane_h, int plane_w,
                                   float box_warp, cudaStream_t stream) {
    SPI_CHECK_ARG(planes && coords, "points_backward: null pointer");
    PointParams p;
    memset(&p, 0, sizeof(p));
    p.planes = planes; p.coords = coords; FILL_DECODER(p);
    p.g_rgb = g_rgb; p.g_sigma = g_sigma; p.g_planes = g_planes;
    p.sc_f = sc_f; p.sc_hid = sc_hid; p.sc_dpre = sc_dpre; p.sc_dout = sc_dout;
    p.n = n; p.m = m; p.H = plane_h; p.W = plane_w; p.box_warp = box_warp;
    long long total = (long long)n * m;
    if (total == 0) return SPI_OK;
    long long cap = (long long)spi_num_sms() * 4, want = (total + 127) / 128;
    points_bwd_kernel<<<(int)(want < cap ? want : cap), 128, 0, stream>>>(p);
    SPI_COUNT_LAUNCH(1);
    SPI_LAUNCH_CHECK("points_backward");
    return SPI_OK;
}

extern "C" int spi_ray_sampler(const float* cam, int n, int res, float* origins, float* dirs, cudaStream_t stream) {
    SPI_CHECK_ARG(cam && origins && dirs && res >= 1, "ray_sampler: bad argument");
    long long total = (long long)n * res * res;
    if (total == 0) return SPI_OK;
    ray_sampler_kernel<<<cdiv(total, 256), 256, 0, stream>>>(cam, n, res, origins, dirs);
    SPI_COUNT_LAUNCH(1);
    SPI_LAUNCH_CHECK("ray_sampler");
    return SPI_OK;
}

extern "C" int spi_ray_march(const float* colors, const float* sigma, const float* depths, int rays, int d, int c, float* rgb,
                             float* depth, float* weights, int* minmax, cudaStream_t stream) {
    SPI_CHECK_ARG(colors && sigma && depths && rgb && depth && weights && minmax, "ray_march: null pointer");
    SPI_CHECK_ARG(d >= 2 && d <= MAXDC + MAXDF, "ray_march: samples per ray must be in [2, %d]", MAXDC + MAXDF);
    if (rays == 0) return SPI_OK;
    minmax_init_kernel<<<1, 1, 0, stream>>>(minmax);
    composite_kernel<<<cdiv(rays, 4) > 2368 ? 2368 : cdiv(rays, 4), 128, 4 * 3 * d * sizeof(float), stream>>>(colors, sigma, depths, rays, d, c,
                                                                                                     rgb, depth, weights, minmax);
    depth_clamp_kernel<<<cdiv(rays, 256), 256, 0, stream>>>(depth, minmax, rays);
    SPI_COUNT_LAUNCH(3);
    SPI_LAUNCH_CHECK("ray_march");
    return SPI_OK;
}

extern "C" int spi_sample_importance(const float* depths, const float* weights, const float* u, int rays, int dc, int df, float* fine,
                                     int* inds, float* cdf, cudaStream_t stream) {
    SPI_CHECK_ARG(depths && weights && u && fine && inds, "sample_importance: null pointer");
    SPI_CHECK_ARG(dc >= 4 && dc <= MAXDC && df >= 1 && df <= MAXDF, "sample_importance: bad sample counts");
    if (rays == 0) return SPI_OK;
    importance_kernel<<<cdiv(rays, 4) > 2368 ? 2368 : cdiv(rays, 4), 128, 4 * (3 * dc + df) * sizeof(float), stream>>>(depths, weights, u, rays,
                                                                                                                dc, df, fine, inds, cdf);
    SPI_COUNT_LAUNCH(1);
    SPI_LAUNCH_CHECK("sample_importance");
    return SPI_OK;
}

extern "C" int spi_inverse_cdf(const float* bins, const float* cdf, const float* u, int rays, int ncdf, int nbins, int df, float* fine,
                               int* inds, cudaStream_t stream) {
    SPI_CHECK_ARG(bins && cdf && u && fine && inds && ncdf >= 2 && nbins >= ncdf, "inverse_cdf: bad argument");
    long long total = (long long)rays * df;
    if (total == 0) return SPI_OK;
    inverse_cdf_kernel<<<cdiv(total, 256), 256, 0, stream>>>(bins, cdf, u, rays, ncdf, nbins, df, fine, inds);
    SPI_COUNT_LAUNCH(1);
    SPI_LAUNCH_CHECK("inverse_cdf");
    return SPI_OK;
}

extern "C" int spi_unify_samples(const float* depths_coarse, const float* depths_fine, int rays, int dc, int df, int* perm,
                                 float* sorted, cudaStream_t stream) {
    SPI_CHECK_ARG(depths_coarse && depths_fine && perm && sorted, "unify_samples: null pointer");
    SPI_CHECK_ARG(dc >= 1 && dc <= MAXDC && df >= 1 && df <= MAXDF, "unify_samples: bad sample counts");
    if (rays == 0) return SPI_OK;
    unify_kernel<<<cdiv(rays, 4) > 2368 ? 2368 : cdiv(rays, 4), 128, 4 * 2 * (dc + df) * sizeof(float), stream>>>(depths_coarse, depths_fine, rays,
                                                                                                           dc, df, perm, sorted);
    SPI_COUNT_LAUNCH(1);
    SPI_LAUNCH_CHECK("unify_samples");
    return SPI_OK;
}
