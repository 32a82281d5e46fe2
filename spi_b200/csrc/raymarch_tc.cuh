// Fused renderer forward on the 5th-generation tensor cores (tcgen05 + TMEM).  Included by raymarch.cu inside its anonymous
// namespace (uses RenderParams, Ray, corners / plane_coords, the per-ray warp stages and the MUFU activations).
//
// Same arithmetic as render_fwd_mma_kernel (renderer.py:88-253, ray_marcher.py:25-57, triplane.py:123-135), different machine
// mapping.  A CTA of 8 warps owns 4 rays at a time -- warps q and q+4 share ray slot q (both may touch TMEM lanes 32q..32q+31) and
// split its gather, its activation columns, the merge ranks and the two composite rounds -- one sample per lane, so the 4 x 32 samples of a "round"
// (coarse samples, then importance samples) form ONE M = 128 decoder tile:
//
//   gather     8 lanes per texel line (128 B), features -> hi / lo TF32 halves -> shared memory, SWIZZLE_128B K-major rows
//   layer 1    D1[128x64] = F W1^T   12 tcgen05.mma (3xTF32: hi*hi + lo*hi + hi*lo), accumulator in tensor memory
//   softplus   tcgen05.ld -> +b1, softplus (2 MUFU) -> hi / lo -> tcgen05.st back to TMEM (the hidden layer never touches smem)
//   layer 2    D2[128x48] = H W2^T   24 tcgen05.mma with the A operand read from TMEM; ALL 33 outputs at once
//   D2 of both rounds stays in TMEM; sigma (one column) feeds the per-ray stages (weights, inverse-CDF sampling, merge by rank),
//   and once the final weights are known the colours are read back, activated and composited with a 31-shuffle transpose-reduce.
//
// Compared with the mma.sync kernel every sample goes through the decoder exactly once (no sigma-only pre-pass + recompute), the
// SM's issue slots are left to the gather and the activations, and nothing of size [R*D, 32] is stored anywhere.
// TMEM columns (256 per CTA, two CTAs per SM): [0,48) D2 coarse | [64,112) D2 fine | [128,192) D1, then H_hi in place | [192,256) H_lo.
// Rows of W2 are permuted so that the 32 colours are columns 0..31 and sigma is column 32.
#pragma once

namespace tcr {

using namespace tc05;

constexpr int TC_THREADS = 256;                   // 8 warps: warps q and q + 4 share ray slot q (TMEM lanes 32q .. 32q+31)
constexpr int A_TILE = 128 * 128;                 // bytes of one 128 x 32 fp32 operand tile
constexpr int W1_TILE = 64 * 128;
constexpr int W2_SLAB = 48 * 128;                 // one K slab (32 hidden units) of W2
constexpr int OFF_A_HI = 0, OFF_A_LO = A_TILE, OFF_W1_HI = 2 * A_TILE, OFF_W1_LO = OFF_W1_HI + W1_TILE;
constexpr int OFF_W2_HI = OFF_W1_LO + W1_TILE, OFF_W2_LO = OFF_W2_HI + 2 * W2_SLAB;
constexpr int OFF_BIAS = OFF_W2_LO + 2 * W2_SLAB;  // b1[64], b2 permuted [48]
constexpr int OFF_BARS = OFF_BIAS + (64 + 48) * 4;
constexpr int OFF_OW = OFF_BARS + 64;              // per-slot (texel offset, weight) tables: 4 x 384 int2
constexpr int OFF_SLOT = OFF_OW + 4 * 384 * 8;     // per-slot float scratch starts here
constexpr uint32_t TM_D2 = 0, TM_D2_STRIDE = 64;   // D2 of round r at columns [64 r, 64 r + 48); H_hi / H_lo behind the last round
constexpr int MAX_ROUNDS = 6;                      // 64 * rounds + 128 <= 512 tensor-memory columns
__host__ __device__ inline int rounds_of(int d) { return (d + 31) >> 5; }

__host__ __device__ inline size_t slot_floats(int dc, int df) {
    const int D = dc + df;
    size_t f = (size_t)dc /*dcs*/ + 32 * rounds_of(dc) /*sigc*/ + 32 * rounds_of(df) /*sigf*/ + df /*fine*/ + dc /*cdf*/ + 3 * (size_t)D /*dall, sigm, w*/ +
               dc + df /*pos*/ + 64 /*feat halves*/;
    return (f + 3) & ~(size_t)3;
}
__host__ __device__ inline size_t smem_bytes(int dc, int df) { return OFF_SLOT + 4 * slot_floats(dc, df) * sizeof(float) + 1024; }

__device__ __forceinline__ void pair_sync(int slot) { slot_bar_sync<64>(slot); }

// Every lane publishes its own sample's (texel offset, weight) pairs into the slot's table -- the first warp of the pair writes
// planes 0 and 1, the second plane 2 -- then (after pair_sync) 8 lanes serve one sample, each owning 4 of the 32 channels, so a
// texel read is one coalesced 128-byte line.  Out-of-range corners read texel 0 with weight 0 (grid_sample's zeros padding): the
// inner loop has no branches and all 12 loads of a sample are in flight together.
__device__ __forceinline__ void publish_corners(int W, int H, float x, float y, float z, float scale, bool valid, int2* s_ow, int lane, int half) {
    float gc[3][2];
    plane_coords(x, y, z, scale, gc);
#pragma unroll
    for (int pp = 0; pp < 3; pp++) {
        if ((pp < 2) == (half == 0)) {
            Corner c;
            corners(gc[pp][0], gc[pp][1], W, H, c);
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const bool in = valid && c.off[q] >= 0;
                s_ow[lane * 12 + pp * 4 + q] = make_int2(in ? c.off[q] + pp * NF : 0, __float_as_int(in ? c.w[q] * (1.f / 3.f) : 0.f));
            }
        }
    }
}

// this warp's 16 samples (rows row0 + 16 half ...) of the A tiles, hi / lo halves, swizzled
__device__ __forceinline__ void gather_rows(const float* __restrict__ pl, const int2* s_ow, uint8_t* a_hi, uint8_t* a_lo, int row0, int lane, int half,
                                            float* frows, int nvalid) {
    const int sub = lane & 7, grp = lane >> 3;
    const float4* pls = reinterpret_cast<const float4*>(pl) + sub;
#pragma unroll 2
    for (int k = 0; k < 4; k++) {
        const int smp = grp + 4 * (k + 4 * half);
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
        uint4 h, l;
        split(acc.x, h.x, l.x); split(acc.y, h.y, l.y); split(acc.z, h.z, l.z); split(acc.w, h.w, l.w);
        const uint32_t o = swz(row0 + smp, sub);
        *reinterpret_cast<uint4*>(a_hi + o) = h;
        *reinterpret_cast<uint4*>(a_lo + o) = l;
        if (frows && smp < nvalid) *reinterpret_cast<float4*>(frows + smp * NF + 4 * sub) = acc;      // kept for the decoder weight gradients
    }
}

// v[c] (c = 0..31) per lane -> lane c returns sum over the 32 lanes of v[c]  (31 shuffles)
__device__ __forceinline__ float transpose_reduce(float (&v)[32], int lane) {
#pragma unroll
    for (int off = 16, n = 16; off >= 1; off >>= 1, n >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < n; i++) {
            const float send = up ? v[i] : v[i + n];
            const float keep = up ? v[i + n] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return v[0];
}

__global__ void __launch_bounds__(TC_THREADS, 2) render_fwd_tc_kernel(RenderParams p, int* err) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    uint8_t* sm = raw + (base - smem_u32(raw));
    uint8_t* a_hi = sm + OFF_A_HI; uint8_t* a_lo = sm + OFF_A_LO;
    uint8_t* w1_hi = sm + OFF_W1_HI; uint8_t* w1_lo = sm + OFF_W1_LO;
    uint8_t* w2_hi = sm + OFF_W2_HI; uint8_t* w2_lo = sm + OFF_W2_LO;
    float* b1s = reinterpret_cast<float*>(sm + OFF_BIAS);
    float* b2s = b1s + 64;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + OFF_BARS);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 2);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int q = warp & 3, half = warp >> 2;
    const int dc = p.dc, df = p.df, D = dc + df;
    int2* g_ow = reinterpret_cast<int2*>(sm + OFF_OW) + q * 384;
    float* b = reinterpret_cast<float*>(sm + OFF_SLOT) + q * slot_floats(dc, df);
    float* dcs = b; b += dc;
    const int RC = rounds_of(dc), RF = rounds_of(df), RT = RC + RF;
    float* sigc = b; b += 32 * RC;
    float* sigf = b; b += 32 * RF;
    float* fine = b; b += df;
    float* cdf = b; b += dc;
    float* dall = b; b += D;
    float* sigm = b; b += D;
    float* w = b; b += D;
    int* pos_c = (int*)b; b += dc;
    int* pos_f = (int*)b; b += df;
    float* featp = b;                                   // [2][32] partial composites of the two warps

    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_mbar_init(); }
    const uint32_t TM_H_HI = (uint32_t)(RT < 2 ? 2 : RT) * TM_D2_STRIDE, TM_H_LO = TM_H_HI + 64;
    const uint32_t TM_COLS = TM_H_HI + 128 <= 256 ? 256u : 512u;
    if (warp == 0) { __syncwarp(); tmem_alloc(slot, TM_COLS); }
    // decoder -> shared memory: gains folded, hi / lo split, canonical K-major SWIZZLE_128B rows
    for (int i = tid; i < 64 * 32; i += TC_THREADS) {
        const int n = i >> 5, k = i & 31;
        uint32_t h, l;
        split(p.w1[i] * p.w1_gain, h, l);
        const uint32_t o = swz(n, k >> 2) + (k & 3) * 4;
        *reinterpret_cast<uint32_t*>(w1_hi + o) = h;
        *reinterpret_cast<uint32_t*>(w1_lo + o) = l;
    }
    for (int i = tid; i < 48 * 64; i += TC_THREADS) {
        const int n = i >> 6, k = i & 63, kk = k & 31;
        const int o_src = n < 32 ? n + 1 : (n == 32 ? 0 : -1);        // colours first, sigma in column 32
        uint32_t h, l;
        split(o_src >= 0 ? p.w2[o_src * 64 + k] * p.w2_gain : 0.f, h, l);
        const uint32_t o = (k >> 5) * W2_SLAB + swz(n, kk >> 2) + (kk & 3) * 4;
        *reinterpret_cast<uint32_t*>(w2_hi + o) = h;
        *reinterpret_cast<uint32_t*>(w2_lo + o) = l;
    }
    if (tid < 64) b1s[tid] = p.b1[tid] * p.b_gain;
    if (tid < 48) b2s[tid] = tid < 32 ? p.b2[tid + 1] * p.b_gain : (tid == 32 ? p.b2[0] * p.b_gain : 0.f);
    fence_async_smem();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tm = *slot;
    const uint32_t tlane = tm + ((uint32_t)(q * 32) << 16);
    const uint32_t id1 = idesc_tf32(128, 64), id2 = idesc_tf32(128, 48);

    const int R = p.R;
    const long long total = (long long)p.n * R;
    const long long groups = (total + 3) >> 2;
    const float scale = 2.f / p.box_warp;
    uint32_t ph0 = 0, ph1 = 0;
    int lmin = 0x7f800000, lmax = 0;

    for (long long grp = blockIdx.x; grp < groups; grp += gridDim.x) {
        const long long ray = grp * 4 + q;
        const bool live = ray < total;
        const long long rr = live ? ray : total - 1;                    // dead slots shadow the last ray (no stores)
        const float* pl = p.planes + (size_t)(rr / R) * p.plane_bs;
        Ray r;
        r.ox = p.origins[rr * 3]; r.oy = p.origins[rr * 3 + 1]; r.oz = p.origins[rr * 3 + 2];
        r.dx = p.dirs[rr * 3]; r.dy = p.dirs[rr * 3 + 1]; r.dz = p.dirs[rr * 3 + 2];
        for (int round = 0; round < RT; round++) {
            // ---- sample positions + gather (coarse rounds first, then the importance rounds)
            const bool coarse = round < RC;
            const int idx = (coarse ? round : round - RC) * 32 + lane;
            float d = 0.f;
            bool valid;
            if (coarse) {
                valid = idx < dc;
                if (valid) { d = coarse_depth(p, idx, p.jitter[rr * dc + idx]); if (half == 0) dcs[idx] = d; }
            } else {
                valid = idx < df;
                if (valid) d = fine[idx];
            }
            // storage row of this lane's sample in the tensors kept for the backward pass
            const int st0 = coarse ? round * 32 : dc + (round - RC) * 32;
            const long long srow = ray * D + st0 + lane;
            const int nvalid = live ? min(32, (coarse ? dc : dc + df) - st0) : 0;
            publish_corners(p.W, p.H, r.ox + d * r.dx, r.oy + d * r.dy, r.oz + d * r.dz, scale, valid, g_ow, lane, half);
            pair_sync(q);
            gather_rows(pl, g_ow, a_hi, a_lo, q * 32, lane, half, p.sv_f ? p.sv_f + (ray * D + st0) * NF : nullptr, nvalid);
            fence_async_smem();
            fence_before();
            __syncthreads();
            // ---- layer 1 on the tensor core
            if (tid == 0) {
                fence_after();
                uint32_t acc = 0;
#pragma unroll
                for (int pass = 0; pass < 3; pass++) {
                    const uint32_t a = smem_u32(pass == 1 ? a_lo : a_hi), wq = smem_u32(pass == 2 ? w1_lo : w1_hi);
#pragma unroll
                    for (int ks = 0; ks < 4; ks++) { mma_ss(tm + TM_H_HI, desc_sw128(a + ks * 32), desc_sw128(wq + ks * 32), id1, acc); acc = 1; }
                }
                commit(&bars[0]);
            }
            if (!mbar_wait_bounded(&bars[0], ph0)) atomicExch(err, 1);
            ph0 ^= 1;
            fence_after();
            // ---- bias + softplus of this warp's 32 hidden columns, hi / lo halves back into TMEM (H_hi overwrites D1 in place)
#pragma unroll
            for (int ch = 0; ch < 2; ch++) {
                const int c0 = half * 32 + ch * 16;
                float v[16];
                uint32_t hh[16], hl[16];
                tmem_ld16(tlane + TM_H_HI + c0, v);
                tmem_wait_ld();
#pragma unroll
                for (int c = 0; c < 16; c++) { v[c] = mma::softplus_fast(v[c] + b1s[c0 + c]); split(v[c], hh[c], hl[c]); }
                tmem_st16(tlane + TM_H_HI + c0, hh);
                tmem_st16(tlane + TM_H_LO + c0, hl);
                if (p.sv_h && lane < nvalid) {
                    float4* dst = reinterpret_cast<float4*>(p.sv_h + srow * NH + c0);
#pragma unroll
                    for (int c = 0; c < 4; c++) dst[c] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
                }
            }
            tmem_wait_st();
            fence_before();
            __syncthreads();
            // ---- layer 2, A from tensor memory
            if (tid == 0) {
                fence_after();
                uint32_t acc = 0;
                const uint32_t dcol = tm + TM_D2 + round * TM_D2_STRIDE;
#pragma unroll
                for (int pass = 0; pass < 3; pass++) {
                    const uint32_t a = tm + (pass == 1 ? TM_H_LO : TM_H_HI), wq = smem_u32(pass == 2 ? w2_lo : w2_hi);
#pragma unroll
                    for (int ks = 0; ks < 8; ks++) {
                        mma_ts(dcol, a + ks * 8, desc_sw128(wq + (ks >> 2) * W2_SLAB + (ks & 3) * 32), id2, acc);
                        acc = 1;
                    }
                }
                commit(&bars[1]);
            }
            if (!mbar_wait_bounded(&bars[1], ph1)) atomicExch(err, 2);
            ph1 ^= 1;
            fence_after();
            if (p.sv_o) {          // pre-activation colours (+bias) of this warp's 16 columns, kept for the backward pass
                float v[16];
                tmem_ld16(tlane + TM_D2 + round * TM_D2_STRIDE + half * 16, v);
                tmem_wait_ld();
                if (lane < nvalid) {
                    float4* dst = reinterpret_cast<float4*>(p.sv_o + srow * 36 + half * 16);
#pragma unroll
                    for (int c = 0; c < 4; c++)
                        dst[c] = make_float4(v[4 * c] + b2s[half * 16 + 4 * c], v[4 * c + 1] + b2s[half * 16 + 4 * c + 1], v[4 * c + 2] + b2s[half * 16 + 4 * c + 2],
                                             v[4 * c + 3] + b2s[half * 16 + 4 * c + 3]);
                }
            }
            if (half == 0) {
                float sg;
                tmem_ld1(tlane + TM_D2 + round * TM_D2_STRIDE + 32, sg);
                tmem_wait_ld();
                sg += b2s[32];
                if (p.sv_o && lane < nvalid) p.sv_o[srow * 36 + 32] = sg;
                (coarse ? sigc : sigf)[idx] = sg;
                __syncwarp();
                if (round == RC - 1 && df > 0) {            // all coarse sigmas known: importance sampling
                    warp_weights(dcs, sigc, dc, w, lane);
                    warp_importance(dcs, w, dc, p.u + rr * df, df, cdf, fine, nullptr, lane);
                }
            }
            pair_sync(q);                               // fine[] / sigma arrays visible to both warps
        }
        // ---- merged order (ranks split between the two warps), final weights, colour coefficients
        if (df > 0) {
            if (half == 0) {
                for (int i = lane; i < dc; i += 32) {
                    const float v = dcs[i];
                    int rk = 0;
                    for (int k = 0; k < dc; k++) rk += (dcs[k] < v) || (dcs[k] == v && k < i);
                    for (int k = 0; k < df; k++) rk += (fine[k] < v);
                    pos_c[i] = rk; sigm[rk] = sigc[i]; dall[rk] = v;
                    if (p.sv_src && live) p.sv_src[ray * D + rk] = (unsigned char)i;
                }
            } else {
                for (int j = lane; j < df; j += 32) {
                    const float v = fine[j];
                    int rk = 0;
                    for (int k = 0; k < dc; k++) rk += (dcs[k] <= v);
                    for (int k = 0; k < df; k++) rk += (fine[k] < v) || (fine[k] == v && k < j);
                    pos_f[j] = rk; sigm[rk] = sigf[j]; dall[rk] = v;
                    if (p.sv_src && live) p.sv_src[ray * D + rk] = (unsigned char)(dc + j);
                }
            }
        } else if (half == 0) {
            for (int i = lane; i < dc; i += 32) { sigm[i] = sigc[i]; dall[i] = dcs[i]; pos_c[i] = i; if (p.sv_src && live) p.sv_src[ray * D + i] = (unsigned char)i; }
        }
        pair_sync(q);
        float depth = 0.f, wsum = 0.f;
        if (half == 0) {
            warp_weights(dall, sigm, D, w, lane);
            warp_finalize(dall, w, D, depth, wsum, lane);          // w[] now holds the colour coefficients a_q
        }
        pair_sync(q);
        // ---- composite: colours straight from TMEM, the rounds alternate between the two warps of the pair
        {
            float feat = 0.f;
            for (int round = half; round < RT; round += 2) {
                float v[32];
                tmem_ld32(tlane + TM_D2 + round * TM_D2_STRIDE, v);
                tmem_wait_ld();
                const bool coarse = round < RC;
                const int idx = (coarse ? round : round - RC) * 32 + lane;
                const bool valid = coarse ? idx < dc : idx < df;
                const float a = valid ? w[coarse ? pos_c[idx] : pos_f[idx]] : 0.f;
#pragma unroll
                for (int c = 0; c < 32; c++) v[c] = a * mma::rgb_act_fast(v[c] + b2s[c]);
                feat += transpose_reduce(v, lane);
            }
            featp[half * 32 + lane] = feat;
        }
        pair_sync(q);
        if (half == 0 && live) {
            p.feat[ray * NF + lane] = (featp[lane] + featp[32 + lane]) * 2.f - 1.f;        // rgb*2-1 (ray_marcher.py:55)
            for (int i = lane; i < D; i += 32) {
                const float dd = dall[i];
                lmin = min(lmin, float_as_ordered(dd)); lmax = max(lmax, float_as_ordered(dd));
                if (p.depths_all) p.depths_all[ray * D + i] = dd;
                if (p.sigma_all) p.sigma_all[ray * D + i] = sigm[i];
            }
            if (lane == 0) { p.depth[ray] = depth; p.wsum[ray] = wsum; }
        }
        // the next group's MMAs overwrite D2 and the slot scratch: every warp must be done with them
        fence_before();
        __syncthreads();
        fence_after();
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { lmin = min(lmin, __shfl_xor_sync(0xffffffffu, lmin, o)); lmax = max(lmax, __shfl_xor_sync(0xffffffffu, lmax, o)); }
    if (lane == 0 && lmax != 0) { atomicMin(p.minmax, lmin); atomicMax(p.minmax + 1, lmax); }
    fence_before();
    __syncthreads();
    if (warp == 0) { __syncwarp(); tmem_dealloc(tm, TM_COLS); }
}

}  // namespace tcr
