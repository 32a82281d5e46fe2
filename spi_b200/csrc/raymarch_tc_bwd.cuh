// Fused renderer backward on tcgen05 / TMEM.  Included by raymarch.cu after raymarch_tc.cuh, inside its anonymous namespace.
//
// Same arithmetic as render_bwd_mma_kernel (adjoint of renderer.py:88-253 / ray_marcher.py:25-57 / triplane.py:123-135 given the
// merged sorted depths saved by the forward pass).  A CTA of 16 warps owns 4 rays at a time: warps q, q+4, q+8, q+12 share ray slot q
// (all four may touch TMEM lanes 32q..32q+31) and split the work of its samples four ways (gather / scatter: 8 samples each; TMEM
// read-outs: a quarter of the columns each), which keeps 16 warps per SM in flight although TMEM allows one CTA.  One "round" = the
// 32 merged samples [32 r, 32 r + 32) of each of the 4 rays = ONE M = 128 tile.  Per ray group:
//
//   B1, per round   gather -> F (smem, hi/lo) -> MMA1 -> +b1, softplus -> H (TMEM, kept) -> MMA2 -> D2 (TMEM, kept)
//                   sigma_i and p_i = <g_rgb, rgb_i> from D2
//   B2              compositing adjoint per ray (one warp of the quad, shuffle scans): d sigma_i, colour coefficients a_i
//   B3, per round   dOut = [a_i g_rgb rgb'(out), d sigma_i] -> TMEM (hi over D2 in place, lo beside) -> MMA3: dH = dOut W2
//                   dPre = dH (1 - exp(-H)) -> TMEM over H in place -> MMA4: dF = dPre W1 -> smem -> RED.v4 scatter into the
//                   12 texel lines of every sample (8 lanes per line)
//
// The hidden layer and the outputs of the last two rounds stay in tensor memory between B1 and B3 (512 columns, one CTA per SM):
// with up to 64 merged samples per ray nothing is recomputed; longer rays (up to 128 samples) re-run the forward part of their
// older rounds in B3.  All four GEMMs are 3xTF32 (hi*hi + lo*hi + hi*lo) with fp32 accumulation, like the forward pass.
// Decoder weight gradients: when requested the per-sample rows f | hid | dpre | dout are written to the scratch tensors and
// reduced by two TF32 GEMMs on the host side (unchanged contract of spi_render_backward).
// When the forward pass kept its activations (RenderParams::sv_*: hidden layer, pre-activation outputs, merged -> storage index)
// B1 shrinks to one 144-byte row read per sample and B3 reads H from global memory: no gather, no layers 1-2, half the MMA
// round-trips; dpre / dout rows are then written in storage order so that they pair with the kept f / hid rows.
#pragma once

namespace tcb {

using namespace tc05;

constexpr int THREADS = 512, NPART = 4;
constexpr int MAXD = 128;                          // merged samples per ray (4 rounds of 32)
constexpr int OFF_A_HI = 0, OFF_A_LO = 16384;
constexpr int OFF_W1_HI = 32768, OFF_W1_LO = OFF_W1_HI + 8192;
constexpr int W2_SLAB = 48 * 128;
constexpr int OFF_W2_HI = OFF_W1_LO + 8192, OFF_W2_LO = OFF_W2_HI + 2 * W2_SLAB;
constexpr int W2T_SLAB = 64 * 128;                 // B of MMA3: rows = hidden unit, K = (permuted) output index, 2 slabs of 32
constexpr int OFF_W2T_HI = OFF_W2_LO + 2 * W2_SLAB, OFF_W2T_LO = OFF_W2T_HI + 2 * W2T_SLAB;
constexpr int W1T_SLAB = 32 * 128;                 // B of MMA4: rows = feature, K = hidden unit, 2 slabs of 32
constexpr int OFF_W1T_HI = OFF_W2T_LO + 2 * W2T_SLAB, OFF_W1T_LO = OFF_W1T_HI + 2 * W1T_SLAB;
constexpr int OFF_BIAS = OFF_W1T_LO + 2 * W1T_SLAB;             // b1[64], b2 permuted [48]
constexpr int OFF_BARS = OFF_BIAS + (64 + 48) * 4;
constexpr int OFF_SLOT = OFF_BARS + 64;
constexpr int DFS = 36;                                          // row stride of the dF tile (floats)
constexpr int SLOT_FLOATS = 10 * MAXD + 32 + 32 * DFS;           // dall sig w pd[4] alpha Tarr gmid | gfe | dF tile
constexpr int OFF_OW = OFF_SLOT + 4 * SLOT_FLOATS * 4;           // per-slot (offset, weight) tables: 4 x 384 int2
constexpr int SMEM_BYTES = OFF_OW + 4 * 384 * 8 + 1024;
// tensor memory columns
constexpr uint32_t TM_COLS = 512;
__device__ __forceinline__ uint32_t tm_h_hi(int r) { return (uint32_t)(r * 128); }
__device__ __forceinline__ uint32_t tm_h_lo(int r) { return (uint32_t)(r * 128 + 64); }
__device__ __forceinline__ uint32_t tm_d2(int r) { return (uint32_t)(256 + r * 64); }
constexpr uint32_t TM_DOUT_LO = 384, TM_D3 = 448;

__device__ __forceinline__ void quad_sync(int slot) { slot_bar_sync<32 * NPART>(slot); }

// publish (texel offset, weight) of this lane's sample into the slot's table: warp `part` (< 3) of the quad writes plane `part`;
// out-of-range corners point at texel 0 with weight 0.  Callers quad_sync() before reading the table.
__device__ __forceinline__ void publish_corners(int W, int H, float x, float y, float z, float scale, bool valid, int2* s_ow, int lane, int part) {
    if (part < 3) {
        float gc[3][2];
        plane_coords(x, y, z, scale, gc);
        const float gx = part == 0 ? gc[0][0] : (part == 1 ? gc[1][0] : gc[2][0]);
        const float gy = part == 0 ? gc[0][1] : (part == 1 ? gc[1][1] : gc[2][1]);
        Corner c;
        corners(gx, gy, W, H, c);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const bool in = valid && c.off[q] >= 0;
            s_ow[lane * 12 + part * 4 + q] = make_int2(in ? c.off[q] + part * NF : 0, __float_as_int(in ? c.w[q] * (1.f / 3.f) : 0.f));
        }
    }
}

__global__ void __launch_bounds__(THREADS, 1) render_bwd_tc_kernel(RenderParams p, int* err) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    uint8_t* sm = raw + (base - smem_u32(raw));
    uint8_t* a_hi = sm + OFF_A_HI; uint8_t* a_lo = sm + OFF_A_LO;
    float* b1s = reinterpret_cast<float*>(sm + OFF_BIAS);
    float* b2s = b1s + 64;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + OFF_BARS);
    uint32_t* slot_addr = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int q = warp & 3, part = warp >> 2;
    float* sb = reinterpret_cast<float*>(sm + OFF_SLOT) + q * SLOT_FLOATS;
    float* dall = sb; float* sig = sb + MAXD; float* w = sb + 2 * MAXD; float* pd = sb + 3 * MAXD;      // pd[4][MAXD]: per-warp partial <g_rgb, rgb>
    float* alpha = sb + 7 * MAXD; float* Tarr = sb + 8 * MAXD; float* gmid = sb + 9 * MAXD; float* gfe = sb + 10 * MAXD; float* DFt = gfe + 32;
    int2* s_ow = reinterpret_cast<int2*>(sm + OFF_OW) + q * 384;

    if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    if (warp == 0) { __syncwarp(); tmem_alloc(slot_addr, TM_COLS); }
    // ---- decoder operands: gains folded, hi / lo split, K-major SWIZZLE_128B rows
    for (int i = tid; i < 64 * 32; i += THREADS) {
        const int h = i >> 5, f = i & 31;
        uint32_t hi, lo;
        split(p.w1[i] * p.w1_gain, hi, lo);
        uint32_t o = swz(h, f >> 2) + (f & 3) * 4;                                   // W1: rows h, K = f
        *reinterpret_cast<uint32_t*>(sm + OFF_W1_HI + o) = hi;
        *reinterpret_cast<uint32_t*>(sm + OFF_W1_LO + o) = lo;
        const int hk = h & 31;
        o = (h >> 5) * W1T_SLAB + swz(f, hk >> 2) + (hk & 3) * 4;                     // W1^T: rows f, K = h
        *reinterpret_cast<uint32_t*>(sm + OFF_W1T_HI + o) = hi;
        *reinterpret_cast<uint32_t*>(sm + OFF_W1T_LO + o) = lo;
    }
    for (int i = tid; i < 48 * 64; i += THREADS) {
        const int c = i >> 6, h = i & 63;                                             // c: permuted output index (colours 0..31, sigma 32)
        const int o_src = c < 32 ? c + 1 : (c == 32 ? 0 : -1);
        uint32_t hi, lo;
        split(o_src >= 0 ? p.w2[o_src * 64 + h] * p.w2_gain : 0.f, hi, lo);
        const int hk = h & 31, ck = c & 31;
        uint32_t o = (h >> 5) * W2_SLAB + swz(c, hk >> 2) + (hk & 3) * 4;             // W2: rows c, K = h
        *reinterpret_cast<uint32_t*>(sm + OFF_W2_HI + o) = hi;
        *reinterpret_cast<uint32_t*>(sm + OFF_W2_LO + o) = lo;
        o = (c >> 5) * W2T_SLAB + swz(h, ck >> 2) + (ck & 3) * 4;                     // W2^T: rows h, K = c
        *reinterpret_cast<uint32_t*>(sm + OFF_W2T_HI + o) = hi;
        *reinterpret_cast<uint32_t*>(sm + OFF_W2T_LO + o) = lo;
    }
    for (int i = tid; i < 16 * 64; i += THREADS) {                                    // K columns 48..63 of W2^T slab 1 are never read, but keep them finite
        const int c = 48 + (i >> 6), h = i & 63, ck = c & 31;
        const uint32_t o = W2T_SLAB + swz(h, ck >> 2) + (ck & 3) * 4;
        *reinterpret_cast<uint32_t*>(sm + OFF_W2T_HI + o) = 0u;
        *reinterpret_cast<uint32_t*>(sm + OFF_W2T_LO + o) = 0u;
    }
    if (tid < 64) b1s[tid] = p.b1[tid] * p.b_gain;
    if (tid < 48) b2s[tid] = tid < 32 ? p.b2[tid + 1] * p.b_gain : (tid == 32 ? p.b2[0] * p.b_gain : 0.f);
    fence_async_smem();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tm = *slot_addr;
    const uint32_t tlane = tm + ((uint32_t)(q * 32) << 16);
    const uint32_t id64 = idesc_tf32(128, 64), id48 = idesc_tf32(128, 48), id32 = idesc_tf32(128, 32);

    const int D = p.dc + p.df, rounds = (D + 31) >> 5;
    const int R = p.R;
    const long long total = (long long)p.n * R;
    const long long groups = (total + 3) >> 2;
    const float scale = 2.f / p.box_warp;
    const float dlo = __int_as_float(p.minmax[0]), dhi = __int_as_float(p.minmax[1]);
    uint32_t ph = 0;
    const int sub = lane & 7, grp = lane >> 3;
    const bool saved = p.sv_h != nullptr;          // hidden layer / outputs kept by the forward pass: no gather, no layers 1-2 here

    for (long long group = blockIdx.x; group < groups; group += gridDim.x) {
        const long long ray = group * 4 + q;
        const bool live = ray < total;
        const long long rr = live ? ray : total - 1;
        const int n = (int)(rr / R);
        const float* pl = p.planes + (size_t)n * p.plane_bs;
        float* gpl = p.g_planes ? p.g_planes + (size_t)n * p.gplane_bs : nullptr;
        Ray r;
        r.ox = p.origins[rr * 3]; r.oy = p.origins[rr * 3 + 1]; r.oz = p.origins[rr * 3 + 2];
        r.dx = p.dirs[rr * 3]; r.dy = p.dirs[rr * 3 + 1]; r.dz = p.dirs[rr * 3 + 2];
        if (part == 0) gfe[lane] = p.g_feat[rr * NF + lane] * 2.f;
        for (int i = part * 32 + lane; i < MAXD; i += 32 * NPART) dall[i] = i < D ? p.depths_all[rr * D + i] : 0.f;
        quad_sync(q);

        // forward part of round rd into TMEM slot sl: gather -> MMA1 -> softplus -> H[sl] -> MMA2 -> D2[sl]
        auto fwd_round = [&](const int rd, const int sl, const bool first) {
            const int i = rd * 32 + lane;
            const bool valid = i < D;
            const float d = dall[min(i, D - 1)];
            publish_corners(p.W, p.H, r.ox + d * r.dx, r.oy + d * r.dy, r.oz + d * r.dz, scale, valid, s_ow, lane, part);
            quad_sync(q);
            const float4* pls = reinterpret_cast<const float4*>(pl) + sub;
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const int smp = grp + 4 * (k + 2 * part);
                float4 v[12];
                float ww[12];
#pragma unroll
                for (int c = 0; c < 12; c++) {
                    const int2 ow = s_ow[smp * 12 + c];
                    ww[c] = __int_as_float(ow.y);
                    v[c] = __ldg(pls + (ow.x >> 2));
                }
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int c = 0; c < 12; c++) {
                    acc.x = fmaf(v[c].x, ww[c], acc.x); acc.y = fmaf(v[c].y, ww[c], acc.y); acc.z = fmaf(v[c].z, ww[c], acc.z); acc.w = fmaf(v[c].w, ww[c], acc.w);
                }
                uint4 hi, lo;
                split(acc.x, hi.x, lo.x); split(acc.y, hi.y, lo.y); split(acc.z, hi.z, lo.z); split(acc.w, hi.w, lo.w);
                const uint32_t o = swz(q * 32 + smp, sub);
                *reinterpret_cast<uint4*>(a_hi + o) = hi;
                *reinterpret_cast<uint4*>(a_lo + o) = lo;
                if (first && p.sc_f && live && rd * 32 + smp < D) *reinterpret_cast<float4*>(p.sc_f + (ray * D + rd * 32 + smp) * NF + 4 * sub) = acc;
            }
            fence_async_smem();
            fence_before();
            __syncthreads();
            if (tid == 0) {
                fence_after();
                uint32_t acc = 0;
#pragma unroll
                for (int pass = 0; pass < 3; pass++) {
                    const uint32_t a = smem_u32(pass == 1 ? a_lo : a_hi), wq = smem_u32(sm + (pass == 2 ? OFF_W1_LO : OFF_W1_HI));
#pragma unroll
                    for (int ks = 0; ks < 4; ks++) { mma_ss(tm + tm_h_hi(sl), desc_sw128(a + ks * 32), desc_sw128(wq + ks * 32), id64, acc); acc = 1; }
                }
                commit(bar);
            }
            if (!mbar_wait_bounded(bar, ph)) atomicExch(err, 11);
            ph ^= 1;
            fence_after();
            {   // +b1, softplus; this warp's 16 of the 64 hidden columns
                float v[16];
                uint32_t hh[16], hl[16];
                tmem_ld16(tlane + tm_h_hi(sl) + part * 16, v);
                tmem_wait_ld();
#pragma unroll
                for (int c = 0; c < 16; c++) { v[c] = mma::softplus_fast(v[c] + b1s[part * 16 + c]); split(v[c], hh[c], hl[c]); }
                tmem_st16(tlane + tm_h_hi(sl) + part * 16, hh);
                tmem_st16(tlane + tm_h_lo(sl) + part * 16, hl);
                if (first && p.sc_hid && live && valid) {
                    float4* dst = reinterpret_cast<float4*>(p.sc_hid + (ray * D + i) * NH + part * 16);
#pragma unroll
                    for (int c = 0; c < 4; c++) dst[c] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
                }
            }
            tmem_wait_st();
            fence_before();
            __syncthreads();
            if (tid == 0) {
                fence_after();
                uint32_t acc = 0;
#pragma unroll
                for (int pass = 0; pass < 3; pass++) {
                    const uint32_t a = tm + (pass == 1 ? tm_h_lo(sl) : tm_h_hi(sl)), wq = smem_u32(sm + (pass == 2 ? OFF_W2_LO : OFF_W2_HI));
#pragma unroll
                    for (int ks = 0; ks < 8; ks++) { mma_ts(tm + tm_d2(sl), a + ks * 8, desc_sw128(wq + (ks >> 2) * W2_SLAB + (ks & 3) * 32), id48, acc); acc = 1; }
                }
                commit(bar);
            }
            if (!mbar_wait_bounded(bar, ph)) atomicExch(err, 12);
            ph ^= 1;
            fence_after();
        };

        // ================================================================ B1: sigma_i and p_i of every round; the last two rounds stay resident
        for (int rd = 0; rd < rounds; rd++) {
            const int sl = rd & 1;
            const int i = rd * 32 + lane;
            float v[8];
            float sg = 0.f;
            if (saved) {      // pre-activation outputs (bias included) straight from the forward pass, storage-order rows
                const long long row = rr * D + (i < D ? p.sv_src[rr * D + i] : 0);
                const float4 a = __ldg(reinterpret_cast<const float4*>(p.sv_o + row * 36 + part * 8));
                const float4 b = __ldg(reinterpret_cast<const float4*>(p.sv_o + row * 36 + part * 8) + 1);
                v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
                if (part == 0) sg = __ldg(p.sv_o + row * 36 + 32);
            } else {
                fwd_round(rd, sl, true);
                tmem_ld8(tlane + tm_d2(sl) + part * 8, v);
                if (part == 0) tmem_ld1(tlane + tm_d2(sl) + 32, sg);
                tmem_wait_ld();
#pragma unroll
                for (int c = 0; c < 8; c++) v[c] += b2s[part * 8 + c];
                sg += b2s[32];
            }
            {   // p_i = <g_rgb, rgb_i>: 8 colour columns per warp of the quad; sigma by the first warp
                float pdv = 0.f;
#pragma unroll
                for (int c = 0; c < 8; c++) pdv = fmaf(gfe[part * 8 + c], mma::rgb_act_fast(v[c]), pdv);
                pd[part * MAXD + i] = pdv;
                if (part == 0) sig[i] = sg;
            }
        }
        quad_sync(q);
        // ================================================================ B2: compositing adjoint (first warp of the pair)
        if (part == 0) {
            float* pd0 = pd;
            for (int i = lane; i < D; i += 32) pd0[i] = (pd[i] + pd[MAXD + i]) + (pd[2 * MAXD + i] + pd[3 * MAXD + i]);
            __syncwarp();
            for (int i = lane; i < D - 1; i += 32) {
                float delta = dall[i + 1] - dall[i];
                float sm_ = softplus_f((sig[i] + sig[i + 1]) * 0.5f - 1.f);
                alpha[i] = 1.f - expf(-(sm_ * delta));
            }
            __syncwarp();
            {   // T_i = prod_{k<i} (1 - alpha_k + 1e-10), w_i = alpha_i T_i: exclusive product scan, 32-wide chunks with a carry
                float carry = 1.f;
                for (int b0 = 0; b0 < D - 1; b0 += 32) {
                    const int i = b0 + lane;
                    const float a = (i < D - 1) ? alpha[i] : 0.f;
                    float incl = (i < D - 1) ? (1.f - a + 1e-10f) : 1.f;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const float v = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o) incl *= v;
                    }
                    float excl = __shfl_up_sync(0xffffffffu, incl, 1);
                    if (lane == 0) excl = 1.f;
                    if (i < D - 1) { Tarr[i] = carry * excl; w[i] = a * (carry * excl); }
                    carry *= __shfl_sync(0xffffffffu, incl, 31);
                }
                if (lane == 0) w[D - 1] = 0.f;
            }
            __syncwarp();
            float ws = 0.f, wd = 0.f;
            for (int i = lane; i < D - 1; i += 32) { ws += w[i]; wd += w[i] * (0.5f * (dall[i] + dall[i + 1])); }
            ws = warp_sum(ws); wd = warp_sum(wd);
            const float depth = wd / ws;
            const float gd = p.g_depth ? p.g_depth[rr] : 0.f;
            const bool depth_live = (ws > 0.f) && (depth == depth) && (depth >= dlo) && (depth <= dhi);
            for (int i = lane; i < D - 1; i += 32) {
                float gw = 0.5f * (pd0[i] + pd0[i + 1]);
                if (depth_live) gw += gd * (0.5f * (dall[i] + dall[i + 1]) - depth) / ws;
                gmid[i] = gw;
            }
            __syncwarp();
            {   // S_i = sum_{k>i} gw_k w_k (exclusive suffix sum, chunks from the far end); d alpha_i = gw_i T_i - S_i / (1 - alpha_i + 1e-10)
                float carry = 0.f;
                for (int b0 = ((D - 2) >> 5) << 5; b0 >= 0; b0 -= 32) {
                    const int i = b0 + lane;
                    const bool in = i < D - 1;
                    const float gw = in ? gmid[i] : 0.f;
                    float incl = in ? gw * w[i] : 0.f;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const float v = __shfl_down_sync(0xffffffffu, incl, o);
                        if (lane + o < 32) incl += v;
                    }
                    float excl = __shfl_down_sync(0xffffffffu, incl, 1);
                    if (lane == 31) excl = 0.f;
                    const float S = carry + excl;
                    const float delta = in ? dall[i + 1] - dall[i] : 0.f;
                    const float smid = in ? (sig[i] + sig[i + 1]) * 0.5f - 1.f : 0.f;
                    if (in) gmid[i] = (gw * Tarr[i] - S / (1.f - alpha[i] + 1e-10f)) * delta * (1.f - alpha[i]) * sigmoid_f(smid);
                    carry += __shfl_sync(0xffffffffu, incl, 0);
                }
            }
            __syncwarp();
        }
        quad_sync(q);
        // ================================================================ B3: the two resident rounds first, older rounds are recomputed
        for (int rd = rounds - 1; rd >= 0; rd--) {
            const int sl = rd & 1;
            if (!saved && rd < rounds - 2) fwd_round(rd, sl, false);
            const int i = rd * 32 + lane;
            const bool valid = i < D;
            // row of this sample in the per-sample tensors: storage order when the forward pass kept its activations, merged order otherwise
            const long long row = rr * D + (saved ? (valid ? p.sv_src[rr * D + i] : 0) : min(i, D - 1));
            float gs = 0.f, a = 0.f;
            if (valid) {
                gs = 0.5f * ((i > 0 ? gmid[i - 1] : 0.f) + (i < D - 1 ? gmid[i] : 0.f));
                a = 0.5f * ((i > 0 ? w[i - 1] : 0.f) + w[i]);
            }
            {   // dOut: colours [8 part, 8 part + 8) (+ sigma and the zero padding by the first warp), hi over D2 in place
                float v[8];
                uint32_t hh[16], hl[16];
                if (saved) {
                    const float4 a4 = __ldg(reinterpret_cast<const float4*>(p.sv_o + row * 36 + part * 8));
                    const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.sv_o + row * 36 + part * 8) + 1);
                    v[0] = a4.x; v[1] = a4.y; v[2] = a4.z; v[3] = a4.w; v[4] = b4.x; v[5] = b4.y; v[6] = b4.z; v[7] = b4.w;
                } else {
                    tmem_ld8(tlane + tm_d2(sl) + part * 8, v);
                    tmem_wait_ld();
#pragma unroll
                    for (int c = 0; c < 8; c++) v[c] += b2s[part * 8 + c];
                }
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    const float so = mma::sigmoid_fast(v[c]);
                    v[c] = valid ? gfe[part * 8 + c] * a * (1.f + 2.f * 0.001f) * so * (1.f - so) : 0.f;
                    split(v[c], hh[c], hl[c]);
                }
                tmem_st8(tlane + tm_d2(sl) + part * 8, hh);
                tmem_st8(tlane + TM_DOUT_LO + part * 8, hl);
                if (p.sc_dout && live && valid) {
                    float* dst = p.sc_dout + row * 36;
#pragma unroll
                    for (int c = 0; c < 8; c++) dst[1 + part * 8 + c] = v[c];
                    if (part == 0) dst[0] = gs;
                    if (part == 3) { dst[33] = 0.f; dst[34] = 0.f; dst[35] = 0.f; }
                }
                if (part == 0) {
#pragma unroll
                    for (int c = 0; c < 16; c++) { hh[c] = 0u; hl[c] = 0u; }
                    split(gs, hh[0], hl[0]);
                    tmem_st16(tlane + tm_d2(sl) + 32, hh);
                    tmem_st16(tlane + TM_DOUT_LO + 32, hl);
                }
            }
            tmem_wait_st();
            fence_before();
            __syncthreads();
            if (tid == 0) {   // MMA3: dH[128x64] = dOut[128x48] W2 (K = 48 permuted outputs)
                fence_after();
                uint32_t acc = 0;
#pragma unroll
                for (int pass = 0; pass < 3; pass++) {
                    const uint32_t at = tm + (pass == 1 ? TM_DOUT_LO : tm_d2(sl)), wq = smem_u32(sm + (pass == 2 ? OFF_W2T_LO : OFF_W2T_HI));
#pragma unroll
                    for (int ks = 0; ks < 6; ks++) { mma_ts(tm + TM_D3, at + ks * 8, desc_sw128(wq + (ks >> 2) * W2T_SLAB + (ks & 3) * 32), id64, acc); acc = 1; }
                }
                commit(bar);
            }
            if (!mbar_wait_bounded(bar, ph)) atomicExch(err, 13);
            ph ^= 1;
            fence_after();
            {   // dPre = dH * softplus'(pre) = dH * (1 - exp(-hid)); this warp's 16 columns, over H in place
                const int c0 = part * 16;
                float dh[16], h1[16], h2[16];
                uint32_t hh[16], hl[16];
                tmem_ld16(tlane + TM_D3 + c0, dh);
                if (saved) {
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        const float4 h4 = __ldg(reinterpret_cast<const float4*>(p.sv_h + row * NH + c0) + c);
                        h1[4 * c] = h4.x; h1[4 * c + 1] = h4.y; h1[4 * c + 2] = h4.z; h1[4 * c + 3] = h4.w;
                        h2[4 * c] = 0.f; h2[4 * c + 1] = 0.f; h2[4 * c + 2] = 0.f; h2[4 * c + 3] = 0.f;
                    }
                } else {
                    tmem_ld16(tlane + tm_h_hi(sl) + c0, h1);
                    tmem_ld16(tlane + tm_h_lo(sl) + c0, h2);
                }
                tmem_wait_ld();
#pragma unroll
                for (int c = 0; c < 16; c++) {
                    dh[c] *= (1.f - mma::ex2_ftz(-1.4426950408889634f * (h1[c] + h2[c])));
                    split(dh[c], hh[c], hl[c]);
                }
                tmem_st16(tlane + tm_h_hi(sl) + c0, hh);
                tmem_st16(tlane + tm_h_lo(sl) + c0, hl);
                if (p.sc_dpre && live && valid) {
                    float4* dst = reinterpret_cast<float4*>(p.sc_dpre + row * NH + c0);
#pragma unroll
                    for (int c = 0; c < 4; c++) dst[c] = make_float4(dh[4 * c], dh[4 * c + 1], dh[4 * c + 2], dh[4 * c + 3]);
                }
            }
            tmem_wait_st();
            fence_before();
            __syncthreads();
            if (gpl) {
                if (tid == 0) {   // MMA4: dF[128x32] = dPre[128x64] W1
                    fence_after();
                    uint32_t acc = 0;
#pragma unroll
                    for (int pass = 0; pass < 3; pass++) {
                        const uint32_t at = tm + (pass == 1 ? tm_h_lo(sl) : tm_h_hi(sl)), wq = smem_u32(sm + (pass == 2 ? OFF_W1T_LO : OFF_W1T_HI));
#pragma unroll
                        for (int ks = 0; ks < 8; ks++) { mma_ts(tm + TM_D3, at + ks * 8, desc_sw128(wq + (ks >> 2) * W1T_SLAB + (ks & 3) * 32), id32, acc); acc = 1; }
                    }
                    commit(bar);
                }
                if (!mbar_wait_bounded(bar, ph)) atomicExch(err, 14);
                ph ^= 1;
                fence_after();
                {
                    float v[8];
                    tmem_ld8(tlane + TM_D3 + part * 8, v);
                    tmem_wait_ld();
                    *reinterpret_cast<float4*>(DFt + lane * DFS + part * 8) = make_float4(v[0], v[1], v[2], v[3]);
                    *reinterpret_cast<float4*>(DFt + lane * DFS + part * 8 + 4) = make_float4(v[4], v[5], v[6], v[7]);
                }
                const float d = dall[min(i, D - 1)];
                publish_corners(p.W, p.H, r.ox + d * r.dx, r.oy + d * r.dy, r.oz + d * r.dz, scale, valid && live, s_ow, lane, part);
                quad_sync(q);
                // scatter: this warp's 8 samples, 8 lanes per texel line (one 128-byte RED per line)
#pragma unroll 4
                for (int tt = part * 96 + grp; tt < part * 96 + 96; tt += 4) {
                    const int2 ow = s_ow[tt];
                    const float ww = __int_as_float(ow.y);
                    if (ww != 0.f) {
                        const float4 v = *reinterpret_cast<const float4*>(DFt + (tt / 12) * DFS + sub * 4);
                        red_add_v4(gpl + ow.x + sub * 4, make_float4(v.x * ww, v.y * ww, v.z * ww, v.w * ww));
                    }
                }
                quad_sync(q);      // the dF tile and the corner table are rewritten next round
            }
        }
        // the next group's MMAs overwrite H / D2
        fence_before();
        __syncthreads();
        fence_after();
    }
    fence_before();
    __syncthreads();
    if (warp == 0) { __syncwarp(); tmem_dealloc(tm, TM_COLS); }
}

}  // namespace tcb
