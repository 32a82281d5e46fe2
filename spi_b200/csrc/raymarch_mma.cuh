// Tensor-core evaluation of the OSG decoder (triplane.py:112-135) inside the fused renderer.
//
// The decoder is a genuine dense contraction ([R*D, 32] x [32, 64] -> softplus -> x [64, 33]); a warp owns a tile of
// 32 samples and evaluates it with mma.sync.m16n8k8 TF32.  To stay inside the 1e-3 render budget the products are
// formed with the 3xTF32 split (a = a_hi + a_lo, b = b_hi + b_lo; a_lo*b_hi + a_hi*b_lo + a_hi*b_hi, fp32
// accumulate), which restores ~fp32 accuracy at 3 MMAs per product -- still ~5x fewer issue slots than the SIMT
// formulation, and no per-FMA shared-memory weight load.
//
// Layout tricks
//   * sample features live in a per-warp shared tile F[rows][36] (stride 36 floats -> conflict-free A-fragment loads);
//   * the C fragment of layer 1 is reused directly as the A fragment of layer 2 by permuting layer 2's K order
//     (virtual k = t <-> hidden 8ks+2t, k = t+4 <-> hidden 8ks+2t+1), so the hidden activations never leave registers;
//   * weights sit once, in fp32, in shared memory; strides (36 / 72 / 40) make every B-fragment load conflict-free
//     (32-bit loads per warp, 64-bit loads per half-warp); the hi/lo split happens in registers.
#pragma once
#include "common.cuh"

namespace mma {

constexpr int NF = 32, NH = 64, NO = 33, NOP = 40;    // NOP: outputs padded to 5 n-tiles
constexpr int FS = 36;                                  // feature tile row stride (floats)
constexpr int W1S = 36, W2S = 72, W2TS = 40;

// fp32 weights, one copy: the hi/lo TF32 split of a B fragment is done in registers right after its load (6 ALU ops per 6
// MMAs) so that the shared-memory footprint stays small enough for 2-3 CTAs per SM.
struct DecM {
    float w1[NH * W1S];        // [h][k]
    float w2[NOP * W2S];       // [o][h], rows >= 33 are zero
    float b1[NH], b2[NOP];
};
struct DecMBwd {               // extra for the backward pass
    float w2t[NH * W2TS];      // [h][o]
};

// hi/lo split for 3xTF32.  `cvt.rna.tf32.f32` is emulated on sm_100 (FSETP + IADD3 + SEL + LOP3 per conversion, 9
// instructions per split -- 2.5x the HMMA count of a tile); a truncating split is 2 instructions: hi = x with the 13 low
// mantissa bits cleared, lo = x - hi (exact).  The tensor core reads only the top 19 bits of a tf32 operand, so lo goes in
// unmasked; the dropped terms (lo*lo and lo's own truncation) are <= 2^-20 relative.
__device__ __forceinline__ void split(float x, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}
// not `volatile`: the instruction is a pure function of its operands, and ptxas must be free to interleave the three
// dependent MMAs of one accumulator with those of its neighbours
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma3(float (&d)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], uint32_t bh0, uint32_t bh1,
                                     uint32_t bl0, uint32_t bl1) {
    mma_tf32(d, al, bh0, bh1);
    mma_tf32(d, ah, bl0, bl1);
    mma_tf32(d, ah, bh0, bh1);
}

// one-time staging of the decoder into shared memory (gains folded, hi/lo split)
__device__ __forceinline__ void load_dec(DecM* s, const float* w1, const float* b1, const float* w2, const float* b2, float g1, float g2,
                                         float gb) {
    for (int i = threadIdx.x; i < NH * W1S; i += blockDim.x) {
        int h = i / W1S, k = i % W1S;
        s->w1[i] = (k < NF) ? w1[h * NF + k] * g1 : 0.f;
    }
    for (int i = threadIdx.x; i < NOP * W2S; i += blockDim.x) {
        int o = i / W2S, h = i % W2S;
        s->w2[i] = (o < NO && h < NH) ? w2[o * NH + h] * g2 : 0.f;
    }
    for (int i = threadIdx.x; i < NH; i += blockDim.x) s->b1[i] = b1[i] * gb;
    for (int i = threadIdx.x; i < NOP; i += blockDim.x) s->b2[i] = i < NO ? b2[i] * gb : 0.f;
}
__device__ __forceinline__ void load_dec_bwd(DecMBwd* s, const float* w2, float g2) {
    for (int i = threadIdx.x; i < NH * W2TS; i += blockDim.x) {
        int h = i / W2TS, o = i % W2TS;
        s->w2t[i] = (o < NO) ? w2[o * NH + h] * g2 : 0.f;
    }
}

// ---- layer 1: pre[32 x 64] = F[32 x 32] W1^T + b1.  acc[mt][nt][j]: row 16mt + g + 8(j>>1), hidden 8nt + 2t + (j&1)
__device__ __forceinline__ void fc1(const DecM* d, const float* F, float (&acc)[2][8][4], int lane) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int nt = 0; nt < 8; nt++) {
        const float bb0 = d->b1[8 * nt + 2 * t], bb1 = d->b1[8 * nt + 2 * t + 1];
#pragma unroll
        for (int mt = 0; mt < 2; mt++) { acc[mt][nt][0] = bb0; acc[mt][nt][1] = bb1; acc[mt][nt][2] = bb0; acc[mt][nt][3] = bb1; }
    }
#pragma unroll 1
    for (int ks = 0; ks < 4; ks++) {
        uint32_t ah[2][4], al[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; mt++) {
            const float* r0 = F + (16 * mt + g) * FS + 8 * ks + t;
            const float* r1 = r0 + 8 * FS;
            split(r0[0], ah[mt][0], al[mt][0]); split(r1[0], ah[mt][1], al[mt][1]);
            split(r0[4], ah[mt][2], al[mt][2]); split(r1[4], ah[mt][3], al[mt][3]);
        }
#pragma unroll
        for (int nt = 0; nt < 8; nt++) {
            const int o = (8 * nt + g) * W1S + 8 * ks + t;
            uint32_t bh0, bh1, bl0, bl1;
            split(d->w1[o], bh0, bl0); split(d->w1[o + 4], bh1, bl1);
#pragma unroll
            for (int mt = 0; mt < 2; mt++) mma3(acc[mt][nt], ah[mt], al[mt], bh0, bh1, bl0, bl1);
        }
    }
}

// ---- layer 2: out[32 x 8*NT2] = hid[32 x 64] W2^T + b2 (first NT2 n-tiles).  out[mt][n][j]: row as above, output 8n + 2t + (j&1)
template <int NT2>
__device__ __forceinline__ void fc2(const DecM* d, const float (&hid)[2][8][4], float (&out)[2][NT2][4], int lane) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int n = 0; n < NT2; n++) {
        const float bb0 = d->b2[8 * n + 2 * t], bb1 = d->b2[8 * n + 2 * t + 1];
#pragma unroll
        for (int mt = 0; mt < 2; mt++) { out[mt][n][0] = bb0; out[mt][n][1] = bb1; out[mt][n][2] = bb0; out[mt][n][3] = bb1; }
    }
#pragma unroll
    for (int ks = 0; ks < 8; ks++) {
        uint32_t ah[2][4], al[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; mt++) {     // C fragment of layer 1 -> A fragment (a0=c0, a1=c2, a2=c1, a3=c3)
            split(hid[mt][ks][0], ah[mt][0], al[mt][0]); split(hid[mt][ks][2], ah[mt][1], al[mt][1]);
            split(hid[mt][ks][1], ah[mt][2], al[mt][2]); split(hid[mt][ks][3], ah[mt][3], al[mt][3]);
        }
#pragma unroll
        for (int n = 0; n < NT2; n++) {
            const int o = (8 * n + g) * W2S + 8 * ks + 2 * t;
            const float2 bw = *(const float2*)(d->w2 + o);
            uint32_t bh0, bh1, bl0, bl1;
            split(bw.x, bh0, bl0); split(bw.y, bh1, bl1);
#pragma unroll
            for (int mt = 0; mt < 2; mt++) mma3(out[mt][n], ah[mt], al[mt], bh0, bh1, bl0, bl1);
        }
    }
}

// MUFU-based activations, one MUFU per transcendental (the .ftz forms skip the denormal pre/post-scaling that `__expf` /
// `__logf` / `__fdividef` wrap around them): |abs err| <~ 5e-7 on softplus / sigmoid outputs.
__device__ __forceinline__ float ex2_ftz(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float lg2_ftz(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float rcp_ftz(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float softplus_fast(float x) {
    return fmaf(lg2_ftz(1.f + ex2_ftz(-1.4426950408889634f * fabsf(x))), 0.6931471805599453f, fmaxf(x, 0.f));
}
__device__ __forceinline__ float sigmoid_fast(float x) { return rcp_ftz(1.f + ex2_ftz(-1.4426950408889634f * x)); }
__device__ __forceinline__ float rgb_act_fast(float x) { return sigmoid_fast(x) * (1.f + 2.f * 0.001f) - 0.001f; }

__device__ __forceinline__ void softplus_inplace(float (&a)[2][8][4]) {
#pragma unroll
    for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int nt = 0; nt < 8; nt++)
#pragma unroll
            for (int j = 0; j < 4; j++) a[mt][nt][j] = softplus_fast(a[mt][nt][j]);
}

// sigma of the 32 rows of a tile -> sig_tile[row] (written by the lanes with t == 0)
__device__ __noinline__ void tile_sigma(const DecM* d, const float* F, float* sig_tile, int lane) {
    float hid[2][8][4];
    fc1(d, F, hid, lane);
    softplus_inplace(hid);
    float out[2][1][4];
    fc2<1>(d, hid, out, lane);
    if ((lane & 3) == 0) {
        const int g = lane >> 2;
#pragma unroll
        for (int mt = 0; mt < 2; mt++) { sig_tile[16 * mt + g] = out[mt][0][0]; sig_tile[16 * mt + g + 8] = out[mt][0][2]; }
    }
}

// full decoder on a 32-row tile; colour accumulation sum_rows a[row] * rgb(row, c) added into acc40[o] (o = 1..32 used).
// a_tile[32]: colour coefficient per row (0 for padding rows); also returns sigma per row when sig_tile != nullptr.
__device__ __noinline__ void tile_color(const DecM* d, const float* F, const float* a_tile, float* acc40, int lane) {
    const int g = lane >> 2, t = lane & 3;
    float hid[2][8][4];
    fc1(d, F, hid, lane);
    softplus_inplace(hid);
    float out[2][5][4];
    fc2<5>(d, hid, out, lane);
    float racc[5][2];
#pragma unroll
    for (int nn = 0; nn < 5; nn++) { racc[nn][0] = 0.f; racc[nn][1] = 0.f; }
#pragma unroll
    for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int hh = 0; hh < 2; hh++) {
            const float a = a_tile[16 * mt + g + 8 * hh];
#pragma unroll
            for (int nn = 0; nn < 5; nn++)
#pragma unroll
                for (int jj = 0; jj < 2; jj++) {
                    // padding rows hold garbage features: select, never multiply a NaN by zero
                    const float c = rgb_act_fast(out[mt][nn][2 * hh + jj]);
                    racc[nn][jj] += (a != 0.f) ? a * c : 0.f;
                }
        }
#pragma unroll
    for (int nn = 0; nn < 5; nn++)
#pragma unroll
        for (int jj = 0; jj < 2; jj++) {
            float v = racc[nn][jj];
            v += __shfl_xor_sync(0xffffffffu, v, 4); v += __shfl_xor_sync(0xffffffffu, v, 8); v += __shfl_xor_sync(0xffffffffu, v, 16);
            if (g == 0) acc40[8 * nn + 2 * t + jj] += v;
        }
    __syncwarp();
}

// backward, first sweep: sigma per row and p = <g_rgb, rgb(row)> per row of a 32-row tile -> sig_tile[32], pdot_tile[32]
__device__ __noinline__ void tile_sigma_pdot(const DecM* d, const float* F, const float* gfe, float* sig_tile, float* pdot_tile, int lane) {
    const int g = lane >> 2, t = lane & 3;
    float hid[2][8][4];
    fc1(d, F, hid, lane);
    softplus_inplace(hid);
    float out[2][5][4];
    fc2<5>(d, hid, out, lane);
#pragma unroll
    for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int hh = 0; hh < 2; hh++) {
            float pd = 0.f;
#pragma unroll
            for (int nn = 0; nn < 5; nn++)
#pragma unroll
                for (int jj = 0; jj < 2; jj++) {
                    const int o = 8 * nn + 2 * t + jj;
                    if (o >= 1 && o <= 32) pd = fmaf(gfe[o - 1], rgb_act_fast(out[mt][nn][2 * hh + jj]), pd);
                }
            pd += __shfl_xor_sync(0xffffffffu, pd, 1); pd += __shfl_xor_sync(0xffffffffu, pd, 2);
            if (t == 0) { pdot_tile[16 * mt + g + 8 * hh] = pd; sig_tile[16 * mt + g + 8 * hh] = out[mt][0][2 * hh]; }
        }
    __syncwarp();
}

// ---- backward GEMMs -------------------------------------------------------------------------------------------------
// dhid[32 x 64] = dout[32 x 40] W2  (K = outputs; A = dout via the C->A trick, B from the transposed copy W2T[h][o])
__device__ __forceinline__ void bwd_fc2(const DecMBwd* d, const float (&dout)[2][5][4], float (&dh)[2][8][4], int lane) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int nt = 0; nt < 8; nt++)
#pragma unroll
            for (int j = 0; j < 4; j++) dh[mt][nt][j] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 5; ks++) {
        uint32_t ah[2][4], al[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; mt++) {
            split(dout[mt][ks][0], ah[mt][0], al[mt][0]); split(dout[mt][ks][2], ah[mt][1], al[mt][1]);
            split(dout[mt][ks][1], ah[mt][2], al[mt][2]); split(dout[mt][ks][3], ah[mt][3], al[mt][3]);
        }
#pragma unroll
        for (int nt = 0; nt < 8; nt++) {
            const int o = (8 * nt + g) * W2TS + 8 * ks + 2 * t;
            const float2 bw = *(const float2*)(d->w2t + o);
            uint32_t bh0, bh1, bl0, bl1;
            split(bw.x, bh0, bl0); split(bw.y, bh1, bl1);
#pragma unroll
            for (int mt = 0; mt < 2; mt++) mma3(dh[mt][nt], ah[mt], al[mt], bh0, bh1, bl0, bl1);
        }
    }
}

// df[32 x 32] = dpre[32 x 64] W1  (K = hidden via the C->A trick; B[k][n] = W1[h = 8ks+2t(+1)][feature 8nt+g])
__device__ __forceinline__ void bwd_fc1(const DecM* d, const float (&dp)[2][8][4], float (&df)[2][4][4], int lane) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int nt = 0; nt < 4; nt++)
#pragma unroll
            for (int j = 0; j < 4; j++) df[mt][nt][j] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 8; ks++) {
        uint32_t ah[2][4], al[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; mt++) {
            split(dp[mt][ks][0], ah[mt][0], al[mt][0]); split(dp[mt][ks][2], ah[mt][1], al[mt][1]);
            split(dp[mt][ks][1], ah[mt][2], al[mt][2]); split(dp[mt][ks][3], ah[mt][3], al[mt][3]);
        }
#pragma unroll
        for (int nt = 0; nt < 4; nt++) {
            const int o0 = (8 * ks + 2 * t) * W1S + 8 * nt + g;
            uint32_t bh0, bh1, bl0, bl1;
            split(d->w1[o0], bh0, bl0); split(d->w1[o0 + W1S], bh1, bl1);
#pragma unroll
            for (int mt = 0; mt < 2; mt++) mma3(df[mt][nt], ah[mt], al[mt], bh0, bh1, bl0, bl1);
        }
    }
}

}  // namespace mma
