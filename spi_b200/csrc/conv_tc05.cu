// Implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a only.
//
// Replaces, for the dense stride-1 layers of the path, the cuDNN kernels behind `_conv2d_wrapper`
// (eg3d/torch_utils/ops/conv2d_resample.py:30-43) as called by `modulated_conv2d` (eg3d/training/networks_stylegan2.py:34-91),
// the super-resolution blocks (eg3d/training/superresolution.py:279-290) and the VGG feature extractors of the losses
// (spi/criteria/lpips/networks.py:53-63, spi/criteria/bbox_cx_loss.py:60-93).
//
//   y[n, oy, ox, o] = epilogue( sum_{ky,kx,i} x[n, oy+ky-pad, ox+kx-pad, i] * w[g(n), o, ky, kx, i] )        (correlation)
//
// Layout: x, y channels-last fp32 ([N,H,W,C] in memory), w [G][O][KH][KW][I] with G = N (per-sample modulated weights) or 1.
// GEMM view: M = output pixels (tile = 8 x 16 patch of one image = 128 rows), N = output channels (tile BN), K = KH*KW*I walked
// as (tap, 32-channel chunk).  One k-block of the A operand is ONE TMA box load of the shifted patch -- (32 ch, 16, 8, 1) at
// (c, x0+kx-pad, y0+ky-pad, n) -- whose out-of-bounds texels the TMA unit zero-fills, so the padding costs nothing and no
// im2col matrix ever exists.  The box lands in shared memory as 128 rows of 128 bytes in the SWIZZLE_128B pattern, which is
// exactly the K-major canonical layout a tcgen05.mma shared-memory descriptor reads (kind::tf32: fp32 bits, low 13 mantissa
// bits ignored).  Accumulators live in tensor memory, double-buffered, so the epilogue of tile t overlaps the MMAs of t+1.
//
// Warp roles (192 threads, persistent CTAs, static round-robin over tiles):
//   warp 0   TMA producer        (one lane) : waits empty[s], arms full[s] with the stage's byte count, issues the A and B loads
//   warp 1   MMA issuer          (one lane) : waits full[s], 4 x tcgen05.mma (K = 8 each), tcgen05.commit -> empty[s]; after the last
//                                             k-block commit -> tfull[buf]
//   warps 2-5 epilogue (128 threads, thread = one pixel row of the tile): tcgen05.ld 32 columns, fused epilogue
//             (+noise*strength, +bias, relu/lrelu, gain, clamp = networks_stylegan2.py:320-329 / bias_act.py:54), swizzled staging
//             tile in shared memory, TMA store (clipped at the image border by the unit), arrive tempty[buf].
#include <cuda.h>

#include "common.cuh"
#include "tc05.cuh"

namespace {

using namespace tc05;

constexpr int TW = 16, TH = 8;            // pixel patch of one M tile
constexpr int A_BYTES = 128 * 128;        // 128 pixels x 32 channels x 4 B
constexpr int NTHREADS = 192;

struct ConvParams {
    int n, h, w, ci, co, kh, kw, pad;
    int tiles_x, tiles_y, tiles_o, total;
    int per_sample;                        // 1: weight group = sample index
    const float* bias;                     // [co] or NULL
    const float* noise;                    // [h*w] or NULL
    const float* noise_strength;           // device scalar (NULL = 1)
    int act;                               // 0 linear, 1 relu, 2 lrelu(slope)
    float slope, gain, clamp;              // clamp < 0: off
    int* err;                              // device flag: set to non-zero if a barrier wait timed out
};

template <int BN, int STAGES>
struct Smem {
    static constexpr int B_BYTES = BN * 128;
    static constexpr int STAGE = A_BYTES + B_BYTES;
    static constexpr int C_OFF = STAGES * STAGE;
    static constexpr int BAR_OFF = C_OFF + 2 * A_BYTES;
    static constexpr int TOTAL = BAR_OFF + (2 * STAGES + 4) * 8 + 16 + 1024;   // +1024: manual alignment slack
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(NTHREADS, 1)
conv_fprop_tc05_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                       const __grid_constant__ CUtensorMap tm_y, const ConvParams p) {
    using S = Smem<BN, STAGES>;
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    uint8_t* sm = raw + (base - smem_u32(raw));
    uint8_t* sC = sm + S::C_OFF;
    uint64_t* full = reinterpret_cast<uint64_t*>(sm + S::BAR_OFF);
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = empty + STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* slot = reinterpret_cast<uint32_t*>(tempty + 2);
    constexpr uint32_t TCOLS = 2 * BN < 32 ? 32 : 2 * BN;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; b++) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], 1); }
        fence_mbar_init();
        tma_prefetch_desc(&tm_x);
        tma_prefetch_desc(&tm_w);
        tma_prefetch_desc(&tm_y);
    }
    if (warp == 0) { __syncwarp(); tmem_alloc(slot, TCOLS); }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tm = *slot;

    const int cchunks = p.ci >> 5;
    const int kblocks = p.kh * p.kw * cchunks;

    if (warp == 0) {
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            bool ok = true;
            for (int t = blockIdx.x; t < p.total && ok; t += gridDim.x) {
                const int ot = t % p.tiles_o; int r = t / p.tiles_o;
                const int tx = r % p.tiles_x; r /= p.tiles_x;
                const int ty = r % p.tiles_y; const int n = r / p.tiles_y;
                const int x0 = tx * TW - p.pad, y0 = ty * TH - p.pad;
                const int wrow = (p.per_sample ? n * p.co : 0) + ot * BN;
                for (int kb = 0; kb < kblocks; kb++) {
                    const int tap = kb / cchunks, cc = kb - tap * cchunks;
                    const int ky = tap / p.kw, kx = tap - ky * p.kw;
                    if (!mbar_wait_bounded(&empty[s], ph ^ 1)) { atomicExch(p.err, 1); ok = false; break; }
                    mbar_expect_tx(&full[s], S::STAGE);
                    tma_load_4d(sm + s * S::STAGE, &tm_x, cc * 32, x0 + kx, y0 + ky, n, &full[s]);
                    tma_load_3d(sm + s * S::STAGE + A_BYTES, &tm_w, cc * 32, tap, wrow, &full[s]);
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = idesc_tf32(128, BN);
            int s = 0; uint32_t ph = 0;
            int local = 0;
            bool ok = true;
            for (int t = blockIdx.x; t < p.total && ok; t += gridDim.x, local++) {
                const int buf = local & 1;
                const uint32_t tph = (local >> 1) & 1;
                if (!mbar_wait_bounded(&tempty[buf], tph ^ 1)) { atomicExch(p.err, 2); break; }
                fence_after();
                const uint32_t dcol = tm + buf * BN;
                for (int kb = 0; kb < kblocks; kb++) {
                    if (!mbar_wait_bounded(&full[s], ph)) { atomicExch(p.err, 3); ok = false; break; }
                    fence_after();
                    const uint32_t a = smem_u32(sm + s * S::STAGE), b = a + A_BYTES;
#pragma unroll
                    for (int ks = 0; ks < 4; ks++) mma_ss(dcol, desc_sw128(a + ks * 32), desc_sw128(b + ks * 32), idesc, (kb | ks) ? 1u : 0u);
                    commit(&empty[s]);
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
                if (ok) commit(&tfull[buf]);
            }
        }
    } else {
        const int q = warp & 3;                       // TMEM lane quarter this warp may touch
        const int row = q * 32 + lane;                // pixel row of the tile
        const int py_in = row / TW, px_in = row % TW;
        const bool leader = (warp == 2 && lane == 0);
        const float strength = p.noise ? (p.noise_strength ? *p.noise_strength : 1.f) : 0.f;
        int local = 0, cidx = 0;
        for (int t = blockIdx.x; t < p.total; t += gridDim.x, local++) {
            const int ot = t % p.tiles_o; int r = t / p.tiles_o;
            const int tx = r % p.tiles_x; r /= p.tiles_x;
            const int ty = r % p.tiles_y; const int n = r / p.tiles_y;
            const int buf = local & 1;
            const uint32_t tph = (local >> 1) & 1;
            if (!mbar_wait_bounded(&tfull[buf], tph)) { atomicExch(p.err, 4); break; }
            fence_after();
            const int py = ty * TH + py_in, px = tx * TW + px_in;
            float nz = 0.f;
            if (p.noise && py < p.h && px < p.w) nz = p.noise[py * p.w + px] * strength;
            const uint32_t tbase = tm + ((uint32_t)(q * 32) << 16) + buf * BN;
#pragma unroll 1
            for (int c = 0; c < BN / 32; c++, cidx++) {
                const int o0 = ot * BN + c * 32;
                if (o0 >= p.co) break;
                float v[32];
                tmem_ld32(tbase + c * 32, v);
                tmem_wait_ld();
                if (leader) tma_wait_group_read<1>();          // the staging buffer written two chunks ago has been drained
                named_bar_sync(1, 128);
                uint8_t* stg = sC + (cidx & 1) * A_BYTES;
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    float e[4];
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        float a = v[4 * j + k] + nz;
                        if (p.bias) { const int o = o0 + 4 * j + k; a += (o < p.co) ? __ldg(p.bias + o) : 0.f; }
                        if (p.act == 1) a = fmaxf(a, 0.f);
                        else if (p.act == 2) a = a < 0.f ? a * p.slope : a;
                        a *= p.gain;
                        if (p.clamp >= 0.f) a = fminf(fmaxf(a, -p.clamp), p.clamp);
                        e[k] = a;
                    }
                    *reinterpret_cast<float4*>(stg + swz(row, j)) = make_float4(e[0], e[1], e[2], e[3]);
                }
                fence_async_smem();
                named_bar_sync(1, 128);
                if (leader) {
                    tma_store_4d(&tm_y, stg, o0, tx * TW, ty * TH, n);
                    tma_commit_group();
                }
            }
            // every thread's TMEM reads of this buffer have completed (wait::ld above, then the barrier): hand it back
            fence_before();
            named_bar_sync(1, 128);
            if (leader) mbar_arrive(&tempty[buf]);
        }
        if (leader) tma_wait_all();
    }
    fence_before();
    __syncthreads();
    if (warp == 0) { __syncwarp(); tmem_dealloc(tm, TCOLS); }
}

// w [G][O][T][I] -> wt [G][I][T][O] with the taps reversed (T-1-t): the weights of the data-gradient convolution
// (conv2d backward w.r.t. the input of a stride-1 'same' correlation = correlation of dy with the flipped, transposed kernel)
__global__ void weight_flip_transpose_kernel(const float* __restrict__ w, float* __restrict__ wt, int o, int taps, int i) {
    __shared__ float tile[32][33];
    const int g = blockIdx.z / taps, t = blockIdx.z % taps;
    const int o0 = blockIdx.y * 32, i0 = blockIdx.x * 32;
    const float* src = w + ((size_t)g * o * taps + t) * i;                    // + oo * taps * i + ii
    float* dst = wt + ((size_t)g * i * taps + (taps - 1 - t)) * o;            // + ii * taps * o + oo
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int oo = o0 + r, ii = i0 + threadIdx.x;
        tile[r][threadIdx.x] = (oo < o && ii < i) ? src[(size_t)oo * taps * i + ii] : 0.f;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int ii = i0 + r, oo = o0 + threadIdx.x;
        if (ii < i && oo < o) dst[(size_t)ii * taps * o + oo] = tile[threadIdx.x][r];
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// rank-`rank` fp32 tensor map, innermost dimension first; strides in bytes for dimensions 1..rank-1
bool make_map(CUtensorMap* m, const void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box, int tf32_round) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint32_t ones[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(m, tf32_round ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(ptr), dims,
                    strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

template <int BN, int STAGES>
int launch(const CUtensorMap& mx, const CUtensorMap& mw, const CUtensorMap& my, const ConvParams& p, int max_ctas, cudaStream_t stream) {
    using S = Smem<BN, STAGES>;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(conv_fprop_tc05_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL) != cudaSuccess) {
            spi_set_error("spi_conv2d_tc: cannot reserve %d bytes of shared memory", S::TOTAL);
            return SPI_ERR_CUDA;
        }
        configured = true;
    }
    int grid = p.total < max_ctas ? p.total : max_ctas;
    conv_fprop_tc05_kernel<BN, STAGES><<<grid, NTHREADS, S::TOTAL, stream>>>(mx, mw, my, p);
    SPI_COUNT_LAUNCH(1);
    SPI_LAUNCH_CHECK("spi_conv2d_tc");
    return SPI_OK;
}

}  // namespace

extern "C" int spi_conv2d_tc_supported(int h, int w, int ci, int co, int kh, int kw) {
    return (ci % 32 == 0 && ci >= 32 && co % 32 == 0 && co >= 32 && kh == kw && (kh == 1 || kh == 3) && w >= 16 && h >= 8) ? 1 : 0;
}

extern "C" int spi_conv2d_tc(const float* x, const float* w, float* y, int n, int h, int wd, int ci, int co, int kh, int kw, int per_sample,
                             const float* bias, const float* noise, const float* noise_strength, int act, float slope, float gain, float clamp,
                             int flags, cudaStream_t stream) {
    SPI_CHECK_ARG(x && w && y, "spi_conv2d_tc: null tensor");
    SPI_CHECK_ARG(spi_conv2d_tc_supported(h, wd, ci, co, kh, kw), "spi_conv2d_tc: unsupported shape h=%d w=%d ci=%d co=%d k=%dx%d", h, wd, ci, co, kh, kw);
    SPI_CHECK_ARG(act >= 0 && act <= 2, "spi_conv2d_tc: act must be 0 (linear), 1 (relu) or 2 (lrelu)");
    SPI_CHECK_ARG((((uintptr_t)x | (uintptr_t)w | (uintptr_t)y) & 15) == 0, "spi_conv2d_tc: tensors must be 16-byte aligned");
    const int tf32_round = (flags & 1) ? 0 : 1;
    const int g = per_sample ? n : 1;
    CUtensorMap mx, mw, my;
    {
        cuuint64_t dims[4] = {(cuuint64_t)ci, (cuuint64_t)wd, (cuuint64_t)h, (cuuint64_t)n};
        cuuint64_t str[3] = {(cuuint64_t)ci * 4, (cuuint64_t)wd * ci * 4, (cuuint64_t)h * wd * ci * 4};
        cuuint32_t box[4] = {32, TW, TH, 1};
        if (!make_map(&mx, x, 4, dims, str, box, tf32_round)) { spi_set_error("spi_conv2d_tc: cuTensorMapEncodeTiled(x) failed"); return SPI_ERR_CUDA; }
    }
    const int bn = (co % 256 == 0) ? 256 : (co % 128 == 0 ? 128 : (co >= 192 ? 256 : (co > 64 ? 128 : 64)));
    {
        cuuint64_t dims[3] = {(cuuint64_t)ci, (cuuint64_t)(kh * kw), (cuuint64_t)g * co};
        cuuint64_t str[2] = {(cuuint64_t)ci * 4, (cuuint64_t)kh * kw * ci * 4};
        cuuint32_t box[3] = {32, 1, (cuuint32_t)bn};
        if (!make_map(&mw, w, 3, dims, str, box, tf32_round)) { spi_set_error("spi_conv2d_tc: cuTensorMapEncodeTiled(w) failed"); return SPI_ERR_CUDA; }
    }
    {
        cuuint64_t dims[4] = {(cuuint64_t)co, (cuuint64_t)wd, (cuuint64_t)h, (cuuint64_t)n};
        cuuint64_t str[3] = {(cuuint64_t)co * 4, (cuuint64_t)wd * co * 4, (cuuint64_t)h * wd * co * 4};
        cuuint32_t box[4] = {32, TW, TH, 1};
        if (!make_map(&my, y, 4, dims, str, box, 0)) { spi_set_error("spi_conv2d_tc: cuTensorMapEncodeTiled(y) failed"); return SPI_ERR_CUDA; }
    }
    ConvParams p;
    p.n = n; p.h = h; p.w = wd; p.ci = ci; p.co = co; p.kh = kh; p.kw = kw; p.pad = (kh - 1) / 2;
    p.tiles_x = cdiv(wd, TW); p.tiles_y = cdiv(h, TH); p.tiles_o = cdiv(co, bn);
    p.total = p.tiles_x * p.tiles_y * p.tiles_o * n;
    p.per_sample = per_sample ? 1 : 0;
    p.bias = bias; p.noise = noise; p.noise_strength = noise_strength;
    p.act = act; p.slope = slope; p.gain = gain; p.clamp = clamp;
    p.err = spi_tc_err_flag();
    const int sms = spi_num_sms();
    if (bn == 256) return launch<256, 4>(mx, mw, my, p, sms, stream);
    if (bn == 128) return launch<128, 6>(mx, mw, my, p, sms, stream);
    return launch<64, 8>(mx, mw, my, p, sms, stream);
}

/* kept for callers of the convolution alone: same flag as spi_tc_error() */
extern "C" int spi_conv2d_tc_error(void) { return spi_tc_error(); }

extern "C" int spi_conv_weight_flip_transpose(const float* w, float* wt, int g, int o, int taps, int i, cudaStream_t stream) {
    SPI_CHECK_ARG(w && wt && g > 0 && o > 0 && taps > 0 && i > 0, "spi_conv_weight_flip_transpose: bad arguments");
    dim3 grid(cdiv(i, 32), cdiv(o, 32), g * taps), block(32, 8);
    SPI_CHECK_ARG(grid.z <= 65535 && grid.y <= 65535, "spi_conv_weight_flip_transpose: too many groups");
    weight_flip_transpose_kernel<<<grid, block, 0, stream>>>(w, wt, o, taps, i);
    SPI_COUNT_LAUNCH(1);
    SPI_LAUNCH_CHECK("spi_conv_weight_flip_transpose");
    return SPI_OK;
}
