// Weight gradient of the convolutions of conv_tc2.cu on tcgen05 + TMEM + TMA (sm_100a): replaces cuDNN's wgrad kernels behind
// `conv2d_gradfix` / autograd of `_conv2d_wrapper` (eg3d/torch_utils/ops/conv2d_resample.py:30-43, conv2d_gradfix.py) for the modulated
// convolutions of the generator and the super-resolution blocks (eg3d/training/networks_stylegan2.py:34-91).
//
//   dW[g, m, tap, c] = sum over images of group g and pixels p of  U[n, p, m] * V[n, p + shift(tap), c]
//
// U = the operand that is NOT shifted by the tap (stride-1 conv: dy, m = output channel; stride-2 transposed conv: x, m = input
// channel), V = the shifted one (stride-1: x through its halo patch; transposed: dy through its four parity views).  GEMM view: M = 128
// channels of U, N = 32 channels of V per tap, K = pixels.  Both operands are channel-contiguous in memory, i.e. MN-major for the
// tensor core: a TMA box [8 px][16 rows] x 32 channels lands as 128 rows of 128 bytes (SWIZZLE_128B_ATOM_32B), and 8 consecutive pixel
// rows are two MN-major swizzle atoms (4 K-rows x 32 MN-elements each) = the K = 8 of one tf32 instruction.  One tcgen05.mma (K = 8 for tf32) therefore consumes one image row of
// the 16 x 8 pixel tile:  A = 4 atoms (128 channels of U) at a leading-dimension stride of 16 KB (one box per 32-channel chunk),
// B = the patch row shifted by the tap.  Taps that differ only in kx are ONE instruction: their B atoms are the same patch rows
// displaced by one pixel = 128 bytes, so a leading-dimension byte offset of 128 enumerates them (N = 96 for a 3x3 kernel: three
// overlapping atoms).  A 3x3 wgrad is 3 MMAs (ky) of 128 x 96 x 8 per image row, accumulating into 3 x 96 = 288 TMEM columns that stay
// resident for the CTA's whole pixel range; the partial sums of the CTAs that share an output block meet through TMA reduce-adds.
//
// Work item (one CTA) = (group g, 128-channel block of U, 32-channel chunk of V, contiguous range of pixel tiles).
// Warp roles: warp 0 TMA producer of U tiles (2 stages), warp 1 TMA producer of V patches (2-4 stages), warp 2 MMA issuer,
// warp 3 TMEM allocation, warps 4-7 epilogue (TMEM -> swizzled staging -> TMA reduce-add into dW).
#include <cuda.h>

#include "common.cuh"
#include "tc05.cuh"

namespace {

using namespace tc05;

constexpr int NTW = 256;
constexpr int U_BYTES = 4 * 16384;            // 128 pixels x 128 channels of U (four 32-channel boxes)
constexpr int P_BYTES = 18 * 16 * 128;        // largest V patch: 18 rows x 16 pixels x 32 channels
constexpr int UST = 2;
constexpr int SMW_U = 0;
constexpr int SMW_P = UST * U_BYTES;          // patch ring: (total - U) / patch bytes buffers
constexpr int SMW_PREGION = 96 * 1024;        // 96 KB of patches (2 x 36 KB halo patches, or 5 x 16 KB for 1x1)
constexpr int SMW_BAR = SMW_P + SMW_PREGION;
constexpr int SMW_TOTAL = SMW_BAR + 1024 + 1024;
constexpr int MAXWV = 4, MAXWM = 3, MAXWO = 9, MAXPS = 5;

struct WMma { int dy, dx, n, dcol; };                     // patch row / pixel shift, N (32 x taps in the instruction), accumulator column
struct WView { int vmap, oy, ox, nmma; WMma mma[MAXWM]; int coff; };       // coff: channel offset of this view inside V (folded 32-channel chunks)
struct WOut { int dcol, tap, coff; };
struct WgradArgs {
    CUtensorMap umap;                 // U: {channels, W, H, N}, box {32, 8, 16, 1}
    CUtensorMap vmap[MAXWV];          // V views: box {32, px, rows, 1}
    CUtensorMap omap;                 // dW: {c (V channels), taps, m (U channels), G}, box {32, 1, 128, 1}
    WView views[MAXWV];
    WOut outs[MAXWO];
    int nviews, nouts;
    int n, groups, imgs_per_group, tiles_x, tiles_y, tiles_per_group, tiles_per_item, splits;
    int mblocks, cchunks, uchunks_total;      // 128-blocks of U channels, 32-chunks of V channels, 32-chunks of U channels
    int px, patch_bytes, nps, patch_stride;
    float* usum;                      // optional: column sums of U over all pixels ([groups][U channels], accumulated with atomics)
    int cu;
    int* err;
};

__device__ __forceinline__ void mma_mn(uint32_t tmem_d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}\n" ::"r"(tmem_d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tma_reduce_add_4d_w(const void* tmap, const void* src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(tmap),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
// instruction descriptor, kind::tf32, both operands MN-major (bits 15, 16)
__device__ __forceinline__ uint32_t idesc_tf32_mn(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__global__ void __launch_bounds__(NTW, 1) conv_wgrad_tc2_kernel(const __grid_constant__ WgradArgs a) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    uint8_t* sm = raw + (base - smem_u32(raw));
    uint64_t* full_u = reinterpret_cast<uint64_t*>(sm + SMW_BAR);
    uint64_t* empty_u = full_u + UST;
    uint64_t* full_p = empty_u + UST;
    uint64_t* empty_p = full_p + MAXPS;
    uint64_t* done = empty_p + MAXPS;
    uint32_t* slot = reinterpret_cast<uint32_t*>(done + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // work item
    int it = blockIdx.x;
    const int split = it % a.splits; it /= a.splits;
    const int cc = it % a.cchunks; it /= a.cchunks;
    const int mb = it % a.mblocks; const int g = it / a.mblocks;
    const int t0 = split * a.tiles_per_item;
    const int t1 = min(a.tiles_per_group, t0 + a.tiles_per_item);
    const int uchunks = min(4, a.uchunks_total - mb * 4);          // valid 32-channel chunks of this 128-block

    // column sums of U (the bias gradient that goes with a tall-skinny weight gradient): the epilogue warps, idle during the main loop, read
    // every U tile out of shared memory; one CTA per (group, M block) does it (cc == 0), and its U stages wait for that reader as well
    const bool do_sum = a.usum != nullptr && cc == 0;
    if (tid == 0) {
        for (int s = 0; s < UST; s++) { mbar_init(&full_u[s], 1); mbar_init(&empty_u[s], do_sum ? 2 : 1); }
        for (int s = 0; s < MAXPS; s++) { mbar_init(&full_p[s], 1); mbar_init(&empty_p[s], 1); }
        mbar_init(done, 1);
        fence_mbar_init();
        tma_prefetch_desc(&a.umap);
        tma_prefetch_desc(&a.omap);
        for (int v = 0; v < a.nviews; v++) tma_prefetch_desc(&a.vmap[v]);
    }
    if (warp == 3) { __syncwarp(); tmem_alloc(slot, 512); }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tm = *slot;
    const int tiles_img = a.tiles_x * a.tiles_y;

    if (warp == 0) {
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int t = t0; t < t1; t++) {
                const int img = g * a.imgs_per_group + t / tiles_img;
                const int r = t % tiles_img;
                const int x0 = (r % a.tiles_x) * 8, y0 = (r / a.tiles_x) * 16;
                if (!mbar_wait_bounded(&empty_u[s], ph ^ 1)) { atomicExch(a.err, 21); return; }
                mbar_expect_tx(&full_u[s], (uint32_t)uchunks * 16384u);
                for (int j = 0; j < uchunks; j++)
                    tma_load_4d(sm + SMW_U + s * U_BYTES + j * 16384, &a.umap, (mb * 4 + j) * 32, x0, y0, img, &full_u[s]);
                if (++s == UST) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int t = t0; t < t1; t++) {
                const int img = g * a.imgs_per_group + t / tiles_img;
                const int r = t % tiles_img;
                const int x0 = (r % a.tiles_x) * 8, y0 = (r / a.tiles_x) * 16;
                for (int v = 0; v < a.nviews; v++) {
                    if (!mbar_wait_bounded(&empty_p[s], ph ^ 1)) { atomicExch(a.err, 22); return; }
                    mbar_expect_tx(&full_p[s], (uint32_t)a.patch_bytes);
                    tma_load_4d(sm + SMW_P + s * a.patch_stride, &a.vmap[a.views[v].vmap], cc * 32 + a.views[v].coff, x0 + a.views[v].ox, y0 + a.views[v].oy, img, &full_p[s]);
                    if (++s == a.nps) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 2) {
        if (lane == 0) {
            // MN-major tf32 descriptors (layout type 1 = SWIZZLE_128B_BASE32B, the only one the unit accepts for transposed 32-bit
            // operands: atoms of 4 K-rows x 128 B, 32-byte chunks XOR-ed with the row; tools/mn_probe.cu): hi = SBO 512 (second half of
            // the K = 8 step) | version 1 | layout 1; lo = address >> 4 | LBO << 16.
            // A: LBO = 16 KB (next 32-channel box of U).  B: LBO = 128 B (next tap in kx = the same rows one pixel further).
            const uint32_t hi = (512u >> 4) | (1u << 14) | (1u << 29);
            const uint32_t a_lo0 = (smem_u32(sm + SMW_U) >> 4) | ((16384u >> 4) << 16);
            const uint32_t b_lo0 = (smem_u32(sm + SMW_P) >> 4) | ((128u >> 4) << 16);
            const uint32_t rowp = (uint32_t)a.px * 128u >> 4;         // patch row pitch in 16-byte units
            int us = 0, ps = 0; uint32_t uph = 0, pph = 0;
            uint32_t first = 1;
            for (int t = t0; t < t1; t++) {
                if (!mbar_wait_bounded(&full_u[us], uph)) { atomicExch(a.err, 23); return; }
                const uint32_t alo = a_lo0 + (uint32_t)(us * (U_BYTES >> 4));
                for (int v = 0; v < a.nviews; v++) {
                    if (!mbar_wait_bounded(&full_p[ps], pph)) { atomicExch(a.err, 24); return; }
                    fence_after();
                    const uint32_t blo = b_lo0 + (uint32_t)(ps * (a.patch_stride >> 4));
                    const int nm = a.views[v].nmma;
                    for (int m = 0; m < nm; m++) {
                        const WMma w = a.views[v].mma[m];
                        const uint32_t idesc = idesc_tf32_mn(128, w.n);
                        const uint32_t b0 = blo + (uint32_t)w.dy * rowp + (uint32_t)w.dx * 8u;
                        const uint32_t d = tm + (uint32_t)w.dcol;
#pragma unroll 4
                        for (int r = 0; r < 16; r++)
                            mma_mn(d, alo + (uint32_t)r * 64u, hi, b0 + (uint32_t)r * rowp, hi, idesc, (first && r == 0) ? 0u : 1u);
                    }
                    commit(&empty_p[ps]);
                    if (++ps == a.nps) { ps = 0; pph ^= 1; }
                }
                first = 0;
                commit(&empty_u[us]);
                if (++us == UST) { us = 0; uph ^= 1; }
            }
            commit(done);
        }
    } else if (warp >= 4) {
        const int q = warp & 3;
        const int row = q * 32 + lane;                 // channel of U inside the 128-block = TMEM lane
        const bool leader = (warp == 4 && lane == 0);
        if (do_sum) {
            // thread `row` owns channel `row` of the 128-block: its values sit at one 4-byte column of the 128 rows of box `row / 32`
            // (SWIZZLE_128B_ATOM_32B: 32-byte chunk index XOR-ed with the row) -- a warp reads 128 contiguous bytes per row, conflict-free
            const int c = row & 31;
            const bool have = (row >> 5) < uchunks;
            float acc = 0.f;
            int us = 0; uint32_t uph = 0;
            for (int t = t0; t < t1; t++) {
                if (!mbar_wait_bounded(&full_u[us], uph)) { atomicExch(a.err, 26); break; }
                if (have) {
                    const uint8_t* reg = sm + SMW_U + us * U_BYTES + (row >> 5) * 16384 + (c & 7) * 4;
#pragma unroll 8
                    for (int r = 0; r < 128; r++) acc += *reinterpret_cast<const float*>(reg + r * 128 + ((((c >> 3) ^ r) & 3) << 5));
                }
                named_bar_sync(2, 128);
                if (leader) mbar_arrive(&empty_u[us]);
                if (++us == UST) { us = 0; uph ^= 1; }
            }
            if (have && mb * 128 + row < a.cu) atomicAdd(a.usum + (size_t)g * a.cu + mb * 128 + row, acc);
        }
        if (t1 > t0) {
            if (!mbar_wait_bounded(done, 0)) { atomicExch(a.err, 25); }
            fence_after();
            // all operand stages are idle now: the U region doubles as the staging area (two 16 KB tiles)
            for (int k = 0; k < a.nouts; k++) {
                float v[32];
                tmem_ld32(tm + ((uint32_t)(q * 32) << 16) + (uint32_t)a.outs[k].dcol, v);
                tmem_wait_ld();
                if (leader) tma_wait_group_read<1>();
                named_bar_sync(1, 128);
                uint8_t* stg = sm + SMW_U + (k & 1) * 16384;
#pragma unroll
                for (int j = 0; j < 8; j++)
                    *reinterpret_cast<float4*>(stg + swz(row, j)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                fence_async_smem();
                named_bar_sync(1, 128);
                if (leader) {
                    tma_reduce_add_4d_w(&a.omap, stg, cc * 32 + a.outs[k].coff, a.outs[k].tap, mb * 128, g);
                    tma_commit_group();
                }
            }
            if (leader) tma_wait_group_read<0>();      // the staging tiles have been read; the stores themselves complete with the grid (as CUTLASS' tma_store_wait<0>)
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 3) { __syncwarp(); tmem_dealloc(tm, 512); }
}

typedef CUresult (*EncodeTiledFnW)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFnW encode_fn_w() {
    static EncodeTiledFnW fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFnW>(ptr);
    }
    return fn;
}

// `operand`: tf32-rounded load in the SWIZZLE_128B_ATOM_32B pattern (what an MN-major tf32 matrix descriptor reads); else a plain
// fp32 SWIZZLE_128B map (the dW reduce-store)
bool map4(CUtensorMap* m, const float* ptr, const cuuint64_t* dims, const cuuint64_t* str, const cuuint32_t* box, int operand) {
    EncodeTiledFnW fn = encode_fn_w();
    if (!fn) return false;
    cuuint32_t ones[4] = {1, 1, 1, 1};
    return fn(m, operand ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(ptr), dims, str, box, ones,
              CU_TENSOR_MAP_INTERLEAVE_NONE, operand ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool map_img(CUtensorMap* m, const float* ptr, int c, long long wv, long long hv, int n, long long spx, long long srow, long long simg, int box_px,
             int box_rows) {
    cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)wv, (cuuint64_t)hv, (cuuint64_t)n};
    cuuint64_t str[3] = {(cuuint64_t)spx, (cuuint64_t)srow, (cuuint64_t)simg};
    cuuint32_t box[4] = {32, (cuuint32_t)box_px, (cuuint32_t)box_rows, 1};
    return map4(m, ptr, dims, str, box, 1);
}

}  // namespace

// Weight gradient.  mode 0: y = stride-1 'same' correlation (k = 1 or 3) of x [N][H][W][Ci] with w [G][Co][k*k][Ci]; dy [N][H][W][Co];
//                           dw [G][Co][k*k][Ci]  (U = dy, V = x).
//                   mode 1: y = stride-2 transposed 3x3 convolution of x [N][H][W][Ci] (spi_conv_transpose2d_s2_tc2); dy [N][2H+1][2W+1][Co];
//                           dw is written TRANSPOSED as [G][Ci][9][Co]  (U = x, V = the four parity views of dy).
// G = N when per_sample (one gradient per image), else 1 (summed over the batch).  dw is overwritten; mode + 4: dw is already zero on entry
// (a slice of the caller's zero arena -- the partial sums meet in dw through reduce-adds), the fill is skipped.
static int wgrad_impl(const float* x, const float* dy, float* dw, int n, int h, int wd, int ci, int co, int k, int per_sample, int mode,
                      float* usum, bool fold_v, cudaStream_t stream) {
    const bool prezeroed = (mode & 4) != 0;
    mode &= 3;
    SPI_CHECK_ARG(x && dy && dw, "spi_conv_wgrad_tc2: null tensor");
    // (k = 1, mode 0: ci may be any multiple of 4 >= 32 -- the last 32-channel chunk of x is zero-filled and its dw columns clipped by TMA;
    //  that is the tall-skinny reduction dw[co][ci] = sum_rows dy[row][co] * x[row][ci] of the renderer's decoder gradients)
    const bool ragged_ok = (k == 1 && mode == 0);
    SPI_CHECK_ARG(ci >= 32 && co >= 32 && (ci % 32 == 0 || (ci % 4 == 0 && ragged_ok)) && (co % 32 == 0 || (co % 4 == 0 && ragged_ok)),
                  "spi_conv_wgrad_tc2: channel counts must be multiples of 32 (ci=%d co=%d)", ci, co);
    SPI_CHECK_ARG(!usum || mode == 0, "spi_conv_wgrad_tc2: column sums are taken of dy (mode 0 only)");
    SPI_CHECK_ARG((mode == 0 && (k == 1 || k == 3)) || (mode == 1 && k == 3), "spi_conv_wgrad_tc2: unsupported mode %d / k %d", mode, k);
    SPI_CHECK_ARG((((uintptr_t)x | (uintptr_t)dy | (uintptr_t)dw) & 15) == 0, "spi_conv_wgrad_tc2: tensors must be 16-byte aligned");
    WgradArgs a;
    memset(&a, 0, sizeof(a));
    const int groups = per_sample ? n : 1;
    const int taps = k * k;
    const int cu = mode == 0 ? co : ci;          // channels of U (M side), of V (N side)
    const int cv = mode == 0 ? ci : co;
    const float* U = mode == 0 ? dy : x;
    const float* V = mode == 0 ? x : dy;
    a.n = n; a.groups = groups; a.imgs_per_group = n / groups;
    a.tiles_x = cdiv(wd, 8); a.tiles_y = cdiv(h, 16);
    a.tiles_per_group = a.tiles_x * a.tiles_y * a.imgs_per_group;
    a.mblocks = cdiv(cu, 128); a.cchunks = cdiv(cv, 32); a.uchunks_total = cdiv(cu, 32);
    a.usum = usum; a.cu = cu;
    bool ok = map_img(&a.umap, U, cu, wd, h, n, (long long)cu * 4, (long long)wd * cu * 4, (long long)h * wd * cu * 4, 8, 16);
    {
        cuuint64_t dims[4] = {(cuuint64_t)cv, (cuuint64_t)taps, (cuuint64_t)cu, (cuuint64_t)groups};
        cuuint64_t str[3] = {(cuuint64_t)cv * 4, (cuuint64_t)taps * cv * 4, (cuuint64_t)cu * taps * cv * 4};
        cuuint32_t box[4] = {32, 1, 128, 1};
        ok = ok && map4(&a.omap, dw, dims, str, box, 0);
    }
    if (mode == 0 && fold_v) {
        // tall-skinny form (k = 1): ONE CTA column takes all 32-channel chunks of V (a view per chunk, 32 accumulator columns each), so U is
        // read once however wide V is
        SPI_CHECK_ARG(k == 1 && cv <= 32 * MAXWV, "spi_conv_wgrad_tc2: folded form needs k = 1 and at most %d V channels", 32 * MAXWV);
        a.px = 8;
        a.patch_bytes = 16 * a.px * 128;
        ok = ok && map_img(&a.vmap[0], V, cv, wd, h, n, (long long)cv * 4, (long long)wd * cv * 4, (long long)h * wd * cv * 4, a.px, 16);
        a.nviews = a.cchunks;
        for (int q = 0; q < a.nviews; q++) {
            WView& v = a.views[q];
            v.vmap = 0; v.oy = 0; v.ox = 0; v.nmma = 1; v.coff = 32 * q;
            v.mma[0] = WMma{0, 0, 32, 32 * q};
            a.outs[a.nouts++] = WOut{32 * q, 0, 32 * q};
        }
        a.cchunks = 1;
    } else if (mode == 0) {
        const int halo = k / 2;
        a.px = 8 + 2 * halo;
        const int rows = 16 + 2 * halo;
        a.patch_bytes = rows * a.px * 128;
        ok = ok && map_img(&a.vmap[0], V, cv, wd, h, n, (long long)cv * 4, (long long)wd * cv * 4, (long long)h * wd * cv * 4, a.px, rows);
        a.nviews = 1;
        WView& v = a.views[0];
        v.vmap = 0; v.oy = -halo; v.ox = -halo; v.nmma = 0;
        for (int ky = 0; ky < k; ky++) {
            v.mma[v.nmma++] = WMma{ky, 0, 32 * k, ky * 32 * k};          // the k taps of this row in one instruction (N = 32 k)
            for (int kx = 0; kx < k; kx++) a.outs[a.nouts++] = WOut{ky * 32 * k + kx * 32, ky * k + kx, 0};
        }
    } else {
        const int hi = 2 * h + 1, wi = 2 * wd + 1;
        a.px = 9;
        const int rows = 17;
        a.patch_bytes = rows * a.px * 128;
        int dcol = 0;
        for (int q = 0; q < 4 && ok; q++) {
            const int py = q >> 1, px = q & 1;
            ok = map_img(&a.vmap[q], V + ((size_t)py * wi + px) * cv, cv, (wi - px + 1) / 2, (hi - py + 1) / 2, n, (long long)2 * cv * 4,
                         (long long)2 * wi * cv * 4, (long long)hi * wi * cv * 4, a.px, rows);
            WView& v = a.views[q];
            v.vmap = q; v.oy = 0; v.ox = 0; v.nmma = 0;
            const int nkx = px ? 1 : 2;                   // taps of this parity in x: kx = px, px + 2
            for (int ky = py; ky < 3; ky += 2) {
                v.mma[v.nmma++] = WMma{ky >> 1, 0, 32 * nkx, dcol};
                for (int j = 0; j < nkx; j++) a.outs[a.nouts++] = WOut{dcol + 32 * j, ky * 3 + px + 2 * j, 0};
                dcol += 32 * nkx;
            }
        }
        a.nviews = 4;
    }
    if (!ok) { spi_set_error("spi_conv_wgrad_tc2: cuTensorMapEncodeTiled failed"); return SPI_ERR_CUDA; }
    a.patch_stride = (a.patch_bytes + 1023) & ~1023;
    a.nps = SMW_PREGION / a.patch_stride;
    if (a.nps > MAXPS) a.nps = MAXPS;
    // split the pixel tiles of a group so that the grid fills the chip about twice
    const int sms = spi_num_sms();
    const int columns = groups * a.mblocks * a.cchunks;
    int splits = (2 * sms) / columns;          // floor: one item more than two full waves would cost a third wave
    if (splits > a.tiles_per_group) splits = a.tiles_per_group;
    if (splits < 1) splits = 1;
    a.tiles_per_item = cdiv(a.tiles_per_group, splits);
    a.splits = cdiv(a.tiles_per_group, a.tiles_per_item);
    a.err = spi_tc_err_flag();
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(conv_wgrad_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMW_TOTAL) != cudaSuccess) {
            spi_set_error("spi_conv_wgrad_tc2: cannot reserve %d bytes of shared memory", SMW_TOTAL);
            return SPI_ERR_CUDA;
        }
        configured = true;
    }
    if (!prezeroed) cudaMemsetAsync(dw, 0, (size_t)groups * cu * taps * cv * 4, stream);
    if (usum) cudaMemsetAsync(usum, 0, (size_t)groups * cu * 4, stream);
    conv_wgrad_tc2_kernel<<<columns * a.splits, NTW, SMW_TOTAL, stream>>>(a);
    SPI_COUNT_LAUNCH(1);
    SPI_LAUNCH_CHECK("spi_conv_wgrad_tc2");
    return SPI_OK;
}

extern "C" int spi_conv_wgrad_tc2(const float* x, const float* dy, float* dw, int n, int h, int wd, int ci, int co, int k, int per_sample, int mode,
                                  cudaStream_t stream) {
    return wgrad_impl(x, dy, dw, n, h, wd, ci, co, k, per_sample, mode, nullptr, false, stream);
}

// Tall-skinny reduction over `rows` rows (a multiple of 8): out[m][c] = sum_r u[r][m] * v[r][c] and, when usum is given, usum[m] = sum_r u[r][m].
// u [rows][cu], v [rows][cv] row-major, cu / cv multiples of 4 and >= 32; TF32 operands, fp32 accumulation; out [cu][cv] and usum [cu] are
// overwritten.  This is the 1x1 weight gradient above with the rows as pixels; for cv <= 128 one CTA column takes all of v, so u and v are read once.
// Used for the decoder gradients of the renderer (dW = d_pre^T f, db = sum d_pre; OSGDecoder, eg3d/training/triplane.py:112-135).
extern "C" int spi_rows_outer_sum(const float* u, const float* v, long long rows, int cu, int cv, float* out, float* usum, cudaStream_t stream) {
    SPI_CHECK_ARG(rows >= 8 && rows % 8 == 0 && rows / 8 <= 2147483647LL, "spi_rows_outer_sum: rows must be a positive multiple of 8");
    return wgrad_impl(v, u, out, 1, (int)(rows / 8), 8, cv, cu, 1, 0, 0, usum, cv <= 32 * MAXWV, stream);
}
