// Implicit-GEMM convolution engine, second generation (tcgen05 + TMEM + TMA, sm_100a): the dense contractions of the path
//   * stride-1 'same' 3x3 / 1x1 correlation (forward of every non-resampling SynthesisLayer, ToRGB, the VGG16 / VGG19 layers,
//     and -- with flipped, transposed weights -- their data gradients)          eg3d/torch_utils/ops/conv2d_resample.py:30-43,
//     eg3d/training/networks_stylegan2.py:34-91, eg3d/training/superresolution.py:279-290, spi/criteria/lpips/networks.py:53-63
//   * stride-2 transposed 3x3 convolution (forward of the up-sampling layers, conv2d_resample.py:114-119), executed as its four
//     output-parity phases, each a small stride-1 correlation written through a strided tensor map
//   * stride-2 3x3 correlation (data gradient of the former), reading the four input-parity views through strided tensor maps.
//
// What changed against conv_tc05.cu (which moved one 16 KB A box + one B tile per (tap, 32-channel chunk) and was bound by L2 -> SM
// bandwidth at 96-128 B/clk/SM): the A operand of ALL taps of a 32-channel chunk is ONE halo patch.  A CTA owns MT (1 or 2) M tiles
// of 16 rows x 8 pixels side by side; the patch [R rows][PX pixels][32 ch] (R <= 18, PX = 16 or 24) is fetched by one TMA box load
// (out-of-bounds texels zero-filled: padding is free) and lands as R*PX rows of 128 bytes in the SWIZZLE_128B pattern.  Because an
// M tile is 8 pixels wide, the 8-row groups of the tile for tap (dy, dx) start at patch + ((g + dy) * PX + 8 s + dx) * 128: a uniform
// stride of PX * 128 bytes -- exactly what a K-major shared-memory matrix descriptor expresses (SBO = PX * 128).  The unit applies the
// swizzle to absolute shared-memory address bits (measured: profiles/r2_conv2_probe.txt), so neither the tap shift nor a pitch that
// is not a multiple of the 1024-byte atom disturbs the pattern TMA wrote: PX is just the tile width plus the halo (18 for 3x3).  The nine taps are nine descriptors into the same
// patch; A traffic drops from 9 x 16 KB to 27 KB per M tile and chunk, and each B (weight) tile is used by MT M tiles.
// 3x3, Cout tile 128, MT = 2: (55 + 147) KB per 4608 MMA cycles = 44 B/clk/SM.
// Tile forms: Cout % 256 == 0 on a chip-filling map: one M tile x N = 256 channels; Cout = 128 tiles on a chip-filling map: operand roles
// swapped (A = weight tile, B = 256 pixels of a [34 rows][10 px] patch: M128 x N256 again, accumulator = channels x pixels, per-warp
// epilogue stores); otherwise MT = 2 M tiles x N <= 128; maps too small for 148 tiles split Cin over CTAs (reduce-add stores).
//
// Warp roles (256 threads, persistent CTAs, static round-robin over tiles):
//   warp 0  patch producer (one lane): empty_a -> TMA box load of the next (view, chunk) patch -> full_a          (2 buffers)
//   warp 1  weight producer (one lane): empty_b -> TMA load of the [BN x 32] weight tile of (tap, chunk) -> full_b (4-8 stages)
//   warp 2  MMA issuer (one lane): per weight stage MT x 4 tcgen05.mma.kind::tf32 (K = 8), commit -> empty_b; after the last tap of
//           a patch commit -> empty_a; after the last k-block commit -> tfull[buf]
//   warp 3  allocates / frees tensor memory (512 columns: MT accumulators of BN columns, double-buffered)
//   warps 4-7  epilogue: tcgen05.ld, fused +noise*strength +bias, relu / lrelu, gain, clamp, swizzled staging tile, TMA store
//           (clipped at the image border by the unit), arrive tempty[buf].
#include <cuda.h>

#include "common.cuh"
#include "tc05.cuh"

namespace {

using namespace tc05;

constexpr int NT2 = 256;
constexpr int BST = 16;                       // most weight stages
constexpr int STG_BYTES = 128 * 128;          // epilogue staging tile (128 pixels x 32 channels)
constexpr int MAXV = 4, MAXP = 4, MAXT = 9;

struct Tap { int dy, dx, wtap; };             // patch row / pixel offset of the tap, index into the weight tensor's tap axis
struct View { int amap, ntaps, oy, ox; Tap taps[MAXT]; };     // patch origin = tile origin + (oy, ox)
struct Program {                              // one output grid (a phase of the transposed convolution, or the whole image)
    int nviews, omap, tiles_x, tiles_y, tile_begin, fuse_epilogue;
    View views[MAXV];
};
struct Conv2Args {
    CUtensorMap amap[MAXV];
    CUtensorMap omap[MAXP];
    CUtensorMap wmap;
    Program prog[MAXP];
    int nprog, total;
    int n, ci, co, bn, tiles_o, mt, px, patch_bytes, per_sample, dbg, splitk, cps, npb, patch_stride, nbst, b_stride, sm_b, sm_stg, pair;
    int swap, trows;                  // swap: output channels on M, 256 pixels (32 rows x 8) on N; trows = image rows per tile (16 or 32)
    const float* bias; const float* noise; const float* noise_strength;
    int noise_w, out_h, out_w;
    int act; float slope, gain, clamp;
    int* err;
};

// Shared memory: [barriers + step tables: 4 KB][patch ring: npb x patch_stride][weight stages: nbst x b_stride][staging: 2 x 16 KB];
// the three data regions are sized per launch (plan_smem) from the patch and weight-tile sizes.
constexpr int SM_BAR = 0;
constexpr int SM_PATCH = 4096;
constexpr int MAXPB = 8;                       // patch ring depth (smaller patches -> more buffers)
constexpr int SM_TOTAL = 227 * 1024;           // everything an SM can give one CTA
constexpr int SM_DATA = SM_TOTAL - 1024 - SM_PATCH;

// D[tmem] (+)= A[smem] * B[smem], descriptors handed over as 32-bit halves (no 64-bit arithmetic on the issuing thread)
__device__ __forceinline__ void mma_lohi(uint32_t tmem_d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}\n" ::"r"(tmem_d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(accumulate)
        : "memory");
}

// TMA reduce-store: global[box] += shared tile (fp32 add performed by the L2), used by the split-K partial sums
__device__ __forceinline__ void tma_reduce_add_4d(const void* tmap, const void* src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(tmap),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}

// ---- CTA-pair (cta_group::2) helpers: two CTAs of a cluster on neighbouring SMs execute ONE M = 256 MMA; each holds its own 128 rows
// of A and half of the B tile, the issuing (rank 0) CTA's barriers collect the TMA bytes of both, tcgen05.commit multicasts the
// "stage free" / "accumulator ready" arrivals to both.  Shared-memory addresses of rank 1 differ from rank 0's by bit 24.
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(void* dst, const void* tmap, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
                     smem_u32(dst)),
                 "l"(tmap), "r"(smem_u32(bar) & PEER_MASK), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(void* dst, const void* tmap, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                     smem_u32(dst)),
                 "l"(tmap), "r"(smem_u32(bar) & PEER_MASK), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void mma_lohi_2sm(uint32_t tmem_d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], da, db, %5, p;\n\t}\n" ::"r"(tmem_d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive (once the MMAs issued so far have retired) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {          // arrive on rank 0's copy of the barrier from either CTA
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & PEER_MASK) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

struct TileCoord { int q, n, ty, tx, ot, ks; };
__device__ __forceinline__ TileCoord decode_tile(const Conv2Args& a, int t) {
    TileCoord c;
    c.q = 0;
#pragma unroll
    for (int i = 1; i < MAXP; i++) if (i < a.nprog && t >= a.prog[i].tile_begin) c.q = i;
    const Program& P = a.prog[c.q];
    int r = t - P.tile_begin;
    c.ot = r % a.tiles_o; r /= a.tiles_o;
    c.ks = r % a.splitk; r /= a.splitk;
    c.tx = r % P.tiles_x; r /= P.tiles_x;
    c.ty = r % P.tiles_y; c.n = r / P.tiles_y;
    return c;
}

template <bool PAIR>
__global__ void __launch_bounds__(NT2, 1) conv_tc2_kernel(const __grid_constant__ Conv2Args a) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    uint8_t* sm = raw + (base - smem_u32(raw));
    uint64_t* full_a = reinterpret_cast<uint64_t*>(sm + SM_BAR);
    uint64_t* empty_a = full_a + MAXPB;
    uint64_t* full_b = empty_a + MAXPB;
    uint64_t* empty_b = full_b + BST;
    uint64_t* tfull = empty_b + BST;
    uint64_t* tempty = tfull + 2;
    uint32_t* slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // PAIR: rank of this CTA in its pair, tile loop over pairs; each CTA owns ONE 8-pixel-wide M tile (a.mt == 1) at x + 8 * rank
    const int rank = PAIR ? (int)cluster_ctarank() : 0;
    const int tile0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, tstride = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int xmul = PAIR ? 16 : a.mt * 8;
    // per-program step list {A offset inside the patch (16-byte units), weight tap, flags, patch x/y origin}: the hot loops read it
    // with one LDS instead of chasing the kernel parameters through constant-memory loads
    int4* steps = reinterpret_cast<int4*>(sm + SM_BAR + 512);           // [MAXP][MAXV * MAXT]
    int* nsteps = reinterpret_cast<int*>(sm + SM_BAR + 512 + MAXP * MAXV * MAXT * 16);
    if (tid < MAXP) {
        int cnt = 0;
        if (tid < a.nprog) {
            const Program& P = a.prog[tid];
            for (int v = 0; v < P.nviews; v++)
                for (int k = 0; k < P.views[v].ntaps; k++) {
                    const Tap tp = P.views[v].taps[k];
                    const int flags = (k == 0 ? 1 : 0) | (k == P.views[v].ntaps - 1 ? 2 : 0) | (P.views[v].amap << 2);
                    steps[tid * MAXV * MAXT + cnt++] = make_int4((tp.dy * a.px + tp.dx) * 8, tp.wtap, flags,
                                                                 ((P.views[v].oy & 0xffff) << 16) | (P.views[v].ox & 0xffff));
                }
        }
        nsteps[tid] = cnt;
    }
    if (tid == 0) {
        for (int s = 0; s < MAXPB; s++) { mbar_init(&full_a[s], 1); mbar_init(&empty_a[s], 1); }
        for (int s = 0; s < 2; s++) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], PAIR ? 2 : 1); }
        for (int s = 0; s < BST; s++) { mbar_init(&full_b[s], 1); mbar_init(&empty_b[s], 1); }
        fence_mbar_init();
        for (int i = 0; i < MAXV; i++) tma_prefetch_desc(&a.amap[i]);
        for (int i = 0; i < a.nprog; i++) tma_prefetch_desc(&a.omap[i]);
        tma_prefetch_desc(&a.wmap);
    }
    if (warp == 3) { __syncwarp(); if (PAIR) tmem_alloc_2sm(slot, 512); else tmem_alloc(slot, 512); }
    fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();          // both CTAs' barriers are initialised before any remote arrival / TMA signal
    fence_after();
    const uint32_t tm = *slot;
    const int cchunks = a.ci >> 5;
    const uint32_t b_bytes = (PAIR ? 64u : (uint32_t)a.bn) * 128u;

    if (warp == 0) {
        if (lane == 0) {
            int pb = 0; uint32_t ph = 0;
            for (int t = tile0; t < a.total; t += tstride) {
                const TileCoord c = decode_tile(a, t);
                const int4* st = steps + c.q * MAXV * MAXT;
                const int ns = nsteps[c.q];
                const int x0 = c.tx * xmul + rank * 8, y0 = c.ty * a.trows;
                const int c1 = min(cchunks, (c.ks + 1) * a.cps);
                for (int cc = c.ks * a.cps; cc < c1; cc++)
                    for (int j = 0; j < ns; j++) {
                        const int4 sp = st[j];
                        if (!(sp.z & 1)) continue;
                        const int ox = (int)(short)(sp.w & 0xffff), oy = (int)(short)((unsigned)sp.w >> 16);
                        if (!mbar_wait_bounded(&empty_a[pb], ph ^ 1)) { atomicExch(a.err, 11); return; }
                        if (PAIR) {
                            if (rank == 0) mbar_expect_tx(&full_a[pb], 2u * (uint32_t)a.patch_bytes);      // both CTAs' patches land on rank 0's barrier
                            tma_load_4d_2sm(sm + SM_PATCH + pb * a.patch_stride, &a.amap[(sp.z >> 2) & 3], cc * 32, x0 + ox, y0 + oy, c.n, &full_a[pb]);
                        } else {
                            mbar_expect_tx(&full_a[pb], (uint32_t)a.patch_bytes);
                            tma_load_4d(sm + SM_PATCH + pb * a.patch_stride, &a.amap[(sp.z >> 2) & 3], cc * 32, x0 + ox, y0 + oy, c.n, &full_a[pb]);
                        }
                        if (++pb == a.npb) { pb = 0; ph ^= 1; }
                    }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int t = tile0; t < a.total; t += tstride) {
                const TileCoord c = decode_tile(a, t);
                const int4* st = steps + c.q * MAXV * MAXT;
                const int ns = nsteps[c.q];
                const int wrow = (a.per_sample ? c.n * a.co : 0) + c.ot * a.bn + (PAIR ? rank * 64 : 0);
                const int c1 = min(cchunks, (c.ks + 1) * a.cps);
                for (int cc = c.ks * a.cps; cc < c1; cc++)
                    for (int j = 0; j < ns; j++) {
                        const int wtap = st[j].y;
                        if (!mbar_wait_bounded(&empty_b[s], ph ^ 1)) { atomicExch(a.err, 12); return; }
                        if (PAIR) {
                            if (rank == 0) mbar_expect_tx(&full_b[s], 2u * b_bytes);
                            tma_load_3d_2sm(sm + a.sm_b + s * a.b_stride, &a.wmap, cc * 32, wtap, wrow, &full_b[s]);
                        } else {
                            mbar_expect_tx(&full_b[s], b_bytes);
                            tma_load_3d(sm + a.sm_b + s * a.b_stride, &a.wmap, cc * 32, wtap, wrow, &full_b[s]);
                        }
                        if (++s == a.nbst) { s = 0; ph ^= 1; }
                    }
            }
        }
    } else if (warp == 2) {
        if (lane == 0 && rank == 0) {
            const uint32_t idesc = idesc_tf32(PAIR ? 256 : 128, a.swap ? 256 : a.bn);
            const bool swp = !PAIR && a.swap;
            // descriptor words: hi = SBO | version 1 | SWIZZLE_128B, lo = (address >> 4) | LBO 1; advancing an operand only adds to lo
            const uint32_t a_hi = (((uint32_t)a.px * 128u) >> 4) | (1u << 14) | (2u << 29);
            const uint32_t b_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
            const uint32_t pa_lo0 = (smem_u32(sm + SM_PATCH) >> 4) | (1u << 16);
            const uint32_t bb_lo0 = (smem_u32(sm + a.sm_b) >> 4) | (1u << 16);
            const int mt = a.mt;
            int s = 0, pb = 0, local = 0; uint32_t ph = 0, pph = 0;
            for (int t = tile0; t < a.total; t += tstride, local++) {
                const TileCoord c = decode_tile(a, t);
                const int4* st = steps + c.q * MAXV * MAXT;
                const int ns = nsteps[c.q];
                const int buf = local & 1;
                if (!mbar_wait_bounded(&tempty[buf], ((local >> 1) & 1) ^ 1)) { atomicExch(a.err, 13); return; }
                fence_after();
                const uint32_t dcol = tm + buf * 256;
                uint32_t acc = 0;
                const int c1 = min(cchunks, (c.ks + 1) * a.cps);
                for (int cc = c.ks * a.cps; cc < c1; cc++) {
                    int4 sp = st[0];
                    for (int j = 0; j < ns; j++) {
                        const int4 cur = sp;
                        if (j + 1 < ns) sp = st[j + 1];
                        if (cur.z & 1) {
                            if (!mbar_wait_bounded(&full_a[pb], pph)) { atomicExch(a.err, 14); return; }
                        }
                        if (!mbar_wait_bounded(&full_b[s], ph)) { atomicExch(a.err, 15); return; }
                        fence_after();
                        const uint32_t alo = pa_lo0 + (uint32_t)(pb * (a.patch_stride >> 4)) + (uint32_t)cur.x;
                        const uint32_t blo = bb_lo0 + (uint32_t)(s * (a.b_stride >> 4));
                        if (PAIR) {
                            mma_lohi_2sm(dcol, alo, a_hi, blo, b_hi, idesc, acc);
                            mma_lohi_2sm(dcol, alo + 2, a_hi, blo + 2, b_hi, idesc, 1u);
                            mma_lohi_2sm(dcol, alo + 4, a_hi, blo + 4, b_hi, idesc, 1u);
                            mma_lohi_2sm(dcol, alo + 6, a_hi, blo + 6, b_hi, idesc, 1u);
                        } else if (swp) {
                            // swapped roles: A = the [128 co x 32 ci] weight tile, B = 256 pixels of the patch (32 rows x 8 px, 8-pixel groups at
                            // a stride of PX * 128 bytes): D[co lane][pixel column], one N = 256 instruction per K step
                            mma_lohi(dcol, blo, b_hi, alo, a_hi, idesc, acc);
                            mma_lohi(dcol, blo + 2, b_hi, alo + 2, a_hi, idesc, 1u);
                            mma_lohi(dcol, blo + 4, b_hi, alo + 4, a_hi, idesc, 1u);
                            mma_lohi(dcol, blo + 6, b_hi, alo + 6, a_hi, idesc, 1u);
                        } else {
                        mma_lohi(dcol, alo, a_hi, blo, b_hi, idesc, acc);
                        mma_lohi(dcol, alo + 2, a_hi, blo + 2, b_hi, idesc, 1u);
                        mma_lohi(dcol, alo + 4, a_hi, blo + 4, b_hi, idesc, 1u);
                        mma_lohi(dcol, alo + 6, a_hi, blo + 6, b_hi, idesc, 1u);
                        }
                        if (!PAIR && mt == 2) {
                            mma_lohi(dcol + 128, alo + 64, a_hi, blo, b_hi, idesc, acc);
                            mma_lohi(dcol + 128, alo + 66, a_hi, blo + 2, b_hi, idesc, 1u);
                            mma_lohi(dcol + 128, alo + 68, a_hi, blo + 4, b_hi, idesc, 1u);
                            mma_lohi(dcol + 128, alo + 70, a_hi, blo + 6, b_hi, idesc, 1u);
                        }
                        acc = 1;
                        if (PAIR) commit_pair(&empty_b[s]); else commit(&empty_b[s]);
                        if (++s == a.nbst) { s = 0; ph ^= 1; }
                        if (cur.z & 2) {
                            if (PAIR) commit_pair(&empty_a[pb]); else commit(&empty_a[pb]);
                            if (++pb == a.npb) { pb = 0; pph ^= 1; }
                        }
                    }
                }
                if (PAIR) commit_pair(&tfull[buf]); else commit(&tfull[buf]);
            }
        }
    } else if (warp >= 4) {
        const int q = warp & 3;
        const int row = q * 32 + lane;                 // pixel row of the M tile = TMEM lane
        const int py_in = row >> 3, px_in = row & 7;
        const bool leader = (warp == 4 && lane == 0);
        const float strength = a.noise ? (a.noise_strength ? *a.noise_strength : 1.f) : 0.f;
        // epilogue constants in registers: act(r) = max(r, r * neg) covers linear (neg 1), relu (0) and lrelu (slope); clamp off = +inf
        const float neg = a.act == 1 ? 0.f : (a.act == 2 ? a.slope : 1.f);
        const float gain = a.gain, cl = a.clamp >= 0.f ? a.clamp : __int_as_float(0x7f800000);
        float* sbias = reinterpret_cast<float*>(sm + SM_BAR + 3072);          // bias of the tile's BN output channels (<= 256 floats)
        uint8_t* sC = sm + a.sm_stg;
        int local = 0, cidx = 0;
        for (int t = tile0; t < a.total; t += tstride, local++) {
            const TileCoord c = decode_tile(a, t);
            const Program& P = a.prog[c.q];
            const int buf = local & 1;
            if (!mbar_wait_bounded(&tfull[buf], (local >> 1) & 1)) { atomicExch(a.err, 16); break; }
            fence_after();
            const bool fuse = P.fuse_epilogue != 0;
            if (!PAIR && a.swap) {
                // D[co lane][pixel column]: this thread owns output channel og of 32 pixels per chunk (4 image rows x 8 px).  Each warp
                // stages its own [32 px][32 ch] tile (a lane writes one float per pixel row: 32 lanes = one 128-byte row, conflict-free)
                // and stores it with its own TMA box {32 ch, 8 px, 4 rows}: no barrier between the four warps inside a tile.
                const int og = c.ot * 128 + row;
                const float bv = (fuse && a.bias && og < a.co) ? __ldg(a.bias + og) : 0.f;
                const int xo = c.tx * 8, yo = c.ty * 32;
                const uint32_t tbase = tm + ((uint32_t)(q * 32) << 16) + buf * 256;
                const uint32_t so = (uint32_t)((lane & 3) * 4);
                const int kc = lane >> 2;
#pragma unroll 1
                for (int cb = 0; cb < 8; cb++, cidx++) {
                    if (yo + 4 * cb >= a.out_h) break;
                    float v[32];
                    tmem_ld32(tbase + cb * 32, v);
                    tmem_wait_ld();
                    if (lane == 0) tma_wait_group_read<1>();       // this warp's staging buffer of two chunks ago has been read
                    __syncwarp();
                    uint8_t* stg = sC + q * 8192 + (cidx & 1) * 4096;
                    if (fuse) {
#pragma unroll
                        for (int j = 0; j < 32; j++) {
                            float nzv = 0.f;
                            if (a.noise) {
                                const int py = yo + 4 * cb + (j >> 3), px = xo + (j & 7);
                                if (py < a.out_h && px < a.out_w) nzv = __ldg(a.noise + py * a.noise_w + px) * strength;      // same address for the whole warp
                            }
                            const float e = v[j] + (nzv + bv);
                            *reinterpret_cast<float*>(stg + swz(j, kc) + so) = fminf(fmaxf(fmaxf(e, e * neg) * gain, -cl), cl);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; j++) *reinterpret_cast<float*>(stg + swz(j, kc) + so) = v[j];
                    }
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0 && !(a.dbg & 16)) {
                        tma_store_4d(&a.omap[1], stg, c.ot * 128 + q * 32, xo, yo + 4 * cb, c.n);
                        tma_commit_group();
                    }
                }
                fence_before();
                named_bar_sync(1, 128);
                if (leader) mbar_arrive(&tempty[buf]);
                continue;
            }
            if (fuse) {          // read by everybody after the first barrier of the chunk loop; the previous tile's readers are past their last barrier
                for (int o = row; o < a.bn; o += 128) {
                    const int og = c.ot * a.bn + o;
                    sbias[o] = (a.bias && og < a.co) ? __ldg(a.bias + og) : 0.f;
                }
            }
            for (int m = 0; m < a.mt; m++) {
                const int xo = PAIR ? c.tx * 16 + rank * 8 : (c.tx * a.mt + m) * 8, yo = c.ty * 16;
                float nz = 0.f;
                if (fuse && a.noise) {
                    const int py = yo + py_in, px = xo + px_in;
                    if (py < a.out_h && px < a.out_w) nz = a.noise[py * a.noise_w + px] * strength;
                }
                const uint32_t tbase = tm + ((uint32_t)(q * 32) << 16) + buf * 256 + m * 128;
#pragma unroll 1
                for (int cb = 0; cb < a.bn / 32; cb++, cidx++) {
                    const int o0 = c.ot * a.bn + cb * 32;
                    if (o0 >= a.co) break;
                    float v[32];
                    tmem_ld32(tbase + cb * 32, v);
                    tmem_wait_ld();
                    if (leader) tma_wait_group_read<1>();          // the staging buffer written two chunks ago has been drained
                    named_bar_sync(1, 128);
                    uint8_t* stg = sC + (cidx & 1) * STG_BYTES;
                    if (fuse) {
                        const float4* b4 = reinterpret_cast<const float4*>(sbias + cb * 32);
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            const float4 bb = b4[j];                    // same address for the whole warp: one broadcast LDS.128
                            float e[4] = {v[4 * j] + (nz + bb.x), v[4 * j + 1] + (nz + bb.y), v[4 * j + 2] + (nz + bb.z), v[4 * j + 3] + (nz + bb.w)};
#pragma unroll
                            for (int k = 0; k < 4; k++) e[k] = fminf(fmaxf(fmaxf(e[k], e[k] * neg) * gain, -cl), cl);
                            *reinterpret_cast<float4*>(stg + swz(row, j)) = make_float4(e[0], e[1], e[2], e[3]);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; j++)
                            *reinterpret_cast<float4*>(stg + swz(row, j)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    }
                    fence_async_smem();
                    named_bar_sync(1, 128);
                    if (leader && !(a.dbg & 16)) {
                        if (a.splitk > 1) tma_reduce_add_4d(&a.omap[P.omap], stg, o0, xo, yo, c.n);
                        else tma_store_4d(&a.omap[P.omap], stg, o0, xo, yo, c.n);
                        tma_commit_group();
                    }
                }
            }
            fence_before();
            named_bar_sync(1, 128);
            if (leader) { if (PAIR) mbar_arrive_leader(&tempty[buf]); else mbar_arrive(&tempty[buf]); }
        }
        if (leader || (a.swap && lane == 0)) tma_wait_group_read<0>();      // the staging tiles have been read; the stores themselves complete with the grid (as CUTLASS' tma_store_wait<0>)
    }
    fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();          // no CTA of the pair leaves (or frees tensor memory) while the other may still signal it
    if (warp == 3) { __syncwarp(); if (PAIR) tmem_dealloc_2sm(tm, 512); else tmem_dealloc(tm, 512); }
}

// w [G][O][T][I] -> wt [G][I][T'][O], T' = T-1-t when `reverse` (data gradient of a stride-1 'same' correlation) else t
__global__ void weight_transpose_kernel(const float* __restrict__ w, float* __restrict__ wt, int o, int taps, int i, int reverse) {
    __shared__ float tile[32][33];
    const int g = blockIdx.z / taps, t = blockIdx.z % taps;
    const int o0 = blockIdx.y * 32, i0 = blockIdx.x * 32;
    const float* src = w + ((size_t)g * o * taps + t) * i;
    float* dst = wt + ((size_t)g * i * taps + (reverse ? taps - 1 - t : t)) * o;
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

EncodeTiledFn encode_fn2() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

bool make_map2(CUtensorMap* m, const void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box, int tf32_round) {
    EncodeTiledFn fn = encode_fn2();
    if (!fn) return false;
    cuuint32_t ones[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(m, tf32_round ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(ptr), dims,
                    strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// NHWC-like view: `c` channels innermost, pixel stride `spx` bytes, row stride `srow` bytes, image stride `simg` bytes
bool map_image(CUtensorMap* m, const float* ptr, int c, long long wv, long long hv, int n, long long spx, long long srow, long long simg,
               int box_px, int box_rows, int tf32_round) {
    cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)wv, (cuuint64_t)hv, (cuuint64_t)n};
    cuuint64_t str[3] = {(cuuint64_t)spx, (cuuint64_t)srow, (cuuint64_t)simg};
    cuuint32_t box[4] = {32, (cuuint32_t)box_px, (cuuint32_t)box_rows, 1};
    return make_map2(m, ptr, 4, dims, str, box, tf32_round);
}

int pick_bn(int co) { return co % 128 == 0 ? 128 : (co % 96 == 0 && co <= 96 ? 96 : (co % 64 == 0 && co < 128 ? 64 : (co > 128 ? 128 : (co + 31) / 32 * 32))); }

// Small feature maps cannot fill 148 SMs with 256-pixel x 128-channel tiles: narrow the Cout tile to 64 and split the Cin chunks of
// every tile over `splitk` CTAs.  Partial sums meet in the output through TMA reduce-adds, so the output is zeroed first and no fused
// epilogue is possible (`allow_split` is false when the caller asked for one).  `tiles_pn` = pixel tiles x images.
void plan_tiles(Conv2Args& a, int tiles_pn, bool allow_split) {
    const int cchunks = a.ci / 32, sms = spi_num_sms();
    if (!a.swap && a.bn == 128 && (long long)tiles_pn * a.tiles_o * (allow_split ? cchunks : 1) < sms) { a.bn = 64; a.tiles_o = cdiv(a.co, a.bn); }
    const int tiles = tiles_pn * a.tiles_o;
    a.splitk = 1; a.cps = cchunks;
    if (!allow_split || tiles >= sms * 3 / 4 || cchunks < 2) return;
    double best = (double)tiles / ((double)cdiv(tiles, sms) * sms);
    for (int cps = cchunks - 1; cps >= 1; cps--) {
        const int s = cdiv(cchunks, cps);
        if (s > 32) break;
        if (cdiv(cchunks, s) != cps) continue;
        const long long items = (long long)tiles * s;
        const double eff = (double)items / ((double)cdiv(items, sms) * sms);
        if (eff > best + 1e-9) { best = eff; a.splitk = s; a.cps = cps; }
        if (eff >= 0.8) break;
    }
}

// flags bit 9 (512): the caller hands in an output that is already zero (a slice of its per-iteration zero arena), no fill needed
void zero_split_output(const Conv2Args& a, float* y, size_t y_bytes, cudaStream_t stream) {
    if (a.splitk > 1 && !(a.dbg & 512)) cudaMemsetAsync(y, 0, y_bytes, stream);
}

int launch2(Conv2Args& a, cudaStream_t stream) {
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(conv_tc2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL) != cudaSuccess ||
            cudaFuncSetAttribute(conv_tc2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL) != cudaSuccess) {
            spi_set_error("spi_conv_tc2: cannot reserve %d bytes of shared memory", SM_TOTAL);
            return SPI_ERR_CUDA;
        }
        configured = true;
    }
    a.err = spi_tc_err_flag();
    {   // shared-memory plan: weight stages first (they turn over once per tap), then as many patch buffers as still fit
        const int P = (a.patch_bytes + 1023) & ~1023, B = ((a.pair ? 64 : a.bn) * 128 + 1023) & ~1023;
        int taps = 0;
        for (int v = 0; v < a.prog[0].nviews; v++) taps += a.prog[0].views[v].ntaps;
        a.npb = taps <= 2 ? 4 : 2;                       // 1x1 layers: a patch lasts one or two MMAs batches, the ring must be deeper
        a.nbst = (SM_DATA - 2 * STG_BYTES - a.npb * P) / B;
        if (a.nbst > BST) a.nbst = BST;
        if (a.nbst < 2) { spi_set_error("spi_conv_tc2: shared-memory plan failed (patch %d B, weight tile %d B)", P, B); return SPI_ERR_ARG; }
        while (a.npb < 4 && SM_DATA - 2 * STG_BYTES - a.nbst * B - (a.npb + 1) * P >= 0) a.npb++;
        a.patch_stride = P; a.b_stride = B;
        a.sm_b = SM_PATCH + a.npb * P;
        a.sm_stg = a.sm_b + a.nbst * B;
    }
    const int sms = spi_num_sms();
    if (a.pair) {          // clusters of two CTAs (neighbouring SMs), one M = 256 tile per pair
        const int pairs = a.total < sms / 2 ? a.total : sms / 2;
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(2 * pairs); cfg.blockDim = dim3(NT2); cfg.dynamicSmemBytes = SM_TOTAL; cfg.stream = stream;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
        cfg.attrs = &attr; cfg.numAttrs = 1;
        if (cudaLaunchKernelEx(&cfg, conv_tc2_kernel<true>, a) != cudaSuccess) {
            spi_set_error("spi_conv_tc2: cluster launch failed: %s", cudaGetErrorString(cudaGetLastError()));
            return SPI_ERR_CUDA;
        }
    } else {
        const int grid = a.total < sms ? a.total : sms;
        conv_tc2_kernel<false><<<grid, NT2, SM_TOTAL, stream>>>(a);
    }
    SPI_COUNT_LAUNCH(1);
    SPI_LAUNCH_CHECK("spi_conv_tc2");
    return SPI_OK;
}

// Tile shape.  Cout % 256 == 0 and a map large enough to fill the chip: one 128-pixel M tile x 256 output channels (an MMA of N = 256
// reads 96 B/clk of operands from shared memory against 128 B/clk at N = 128: the weight-tile re-read is what bounds N = 128);
// otherwise two M tiles x (at most) 128 channels, which halves the weight traffic per pixel instead.
void common_args(Conv2Args& a, int n, int ci, int co, int per_sample, int wv_out, int hv_out, int flags, bool allow_split = true) {
    memset(&a, 0, sizeof(a));
    a.n = n; a.ci = ci; a.co = co; a.per_sample = per_sample ? 1 : 0;
    const long long t256 = (long long)cdiv(wv_out, 8) * cdiv(hv_out, 16) * n * (co / 256);
    // (small maps reach the chip-filling tile count by splitting Cin, which the caller must allow: no fused epilogue then)
    const bool fill = t256 >= spi_num_sms() * 3 / 4 || (allow_split && !(flags & 32) && t256 * (ci / 32) >= spi_num_sms() / 2 && t256 >= 8);
    if (co % 256 == 0 && fill && !(flags & 128)) {
        a.bn = 256; a.mt = 1;
    } else {
        a.bn = pick_bn(co);
        a.mt = (wv_out > 8 && !(flags & 4)) ? 2 : 1;
    }
    a.tiles_o = cdiv(co, a.bn);
    a.trows = 16;
    a.dbg = flags;
    a.gain = 1.f; a.clamp = -1.f;
    // CTA pairs (cta_group::2, M = 256): the two M tiles of a 16-pixel-wide tile go to the two CTAs of a cluster, each loads half of the
    // weight tile -- an N = 128 layer then reads 96 B/clk of operands per SM instead of 128.  Large maps only (no split, full waves).
    const long long pair_tiles = (long long)cdiv(wv_out, 16) * cdiv(hv_out, 16) * n * a.tiles_o;
    if ((flags & 256) && a.bn == 128 && a.mt == 2 && pair_tiles >= spi_num_sms()) { a.pair = 1; a.mt = 1; }
}

bool map_weights(Conv2Args& a, const float* w, int taps, int g, int tf32_round) {
    cuuint64_t dims[3] = {(cuuint64_t)a.ci, (cuuint64_t)taps, (cuuint64_t)g * a.co};
    cuuint64_t str[2] = {(cuuint64_t)a.ci * 4, (cuuint64_t)taps * a.ci * 4};
    cuuint32_t box[3] = {32, 1, (cuuint32_t)(a.pair ? 64 : a.bn)};
    return make_map2(&a.wmap, w, 3, dims, str, box, tf32_round);
}

}  // namespace

extern "C" int spi_conv_tc2_supported(int ci, int co) { return (ci % 32 == 0 && ci >= 32 && co % 32 == 0 && co >= 32) ? 1 : 0; }

// Stride-1 'same' correlation, k = 1 or 3.  x [N][H][W][Ci], w [G][Co][k*k][Ci] (G = N if per_sample else 1), y [N][H][W][Co].
// flags: bit 0 = feed raw fp32 bits to the tensor core (truncation) instead of round-to-nearest TF32 on load.
// (Descriptor base offset stays 0 although tap-shifted operand rows start inside a 1024-byte swizzle atom: measured on B200, the
// unit swizzles on absolute shared-memory address bits, so the pattern TMA wrote is the pattern the MMA reads; setting the field
// from the shift gives wrong results -- profiles/r2_conv2_probe.txt.)
static int conv_s1(bool plan_only, const float* x, const float* w, float* y, int n, int h, int wd, int ci, int co, int k, int per_sample,
                   const float* bias, const float* noise, const float* noise_strength, int act, float slope, float gain, float clamp,
                   int flags, cudaStream_t stream) {
    SPI_CHECK_ARG(plan_only || (x && w && y), "spi_conv2d_tc2: null tensor");
    SPI_CHECK_ARG(spi_conv_tc2_supported(ci, co) && (k == 1 || k == 3), "spi_conv2d_tc2: unsupported shape ci=%d co=%d k=%d", ci, co, k);
    SPI_CHECK_ARG(act >= 0 && act <= 2, "spi_conv2d_tc2: act must be 0 (linear), 1 (relu) or 2 (lrelu)");
    SPI_CHECK_ARG((((uintptr_t)x | (uintptr_t)w | (uintptr_t)y) & 15) == 0, "spi_conv2d_tc2: tensors must be 16-byte aligned");
    const int rnd = (flags & 1) ? 0 : 1;
    Conv2Args a;
    const bool epi = bias || noise || act != 0 || gain != 1.f || clamp >= 0.f;
    common_args(a, n, ci, co, per_sample, wd, h, flags, !epi);
    const int halo = k / 2;
    // Cout = 128 tiles on a map that fills the chip: swap the operand roles -- the [128 co x 32 ci] weight tile becomes A and 256 pixels
    // (32 rows x 8 px of the halo patch) become B, so that every instruction is M128 x N256 (96 B/clk of shared-memory operand reads
    // instead of the 128 B/clk that bound the N = 128 form at 57 % tensor pipe).  flags bit 10 (1024) keeps the unswapped form.
    if (!(flags & 1024) && a.bn == 128 && !a.pair && co % 128 == 0 &&
        (long long)cdiv(wd, 8) * cdiv(h, 32) * n * (co / 128) >= spi_num_sms() * 3 / 4) {
        a.swap = 1; a.mt = 1; a.trows = 32;
    }
    a.px = a.mt * 8 + (halo ? ((flags & 64) ? 8 : 2 * halo) : 0);
    const int rows = a.trows + 2 * halo;
    a.patch_bytes = rows * a.px * 128;
    const int tile_w = a.pair ? 16 : a.mt * 8;
    plan_tiles(a, cdiv(wd, tile_w) * cdiv(h, a.trows) * n, !epi && !(flags & 32) && !a.swap);
    if (plan_only) return a.splitk;
    zero_split_output(a, y, (size_t)n * h * wd * co * 4, stream);
    if (!map_image(&a.amap[0], x, ci, wd, h, n, (long long)ci * 4, (long long)wd * ci * 4, (long long)h * wd * ci * 4, a.px, rows, rnd) ||
        !map_image(&a.omap[0], y, co, wd, h, n, (long long)co * 4, (long long)wd * co * 4, (long long)h * wd * co * 4, 8, 16, 0) ||
        (a.swap && !map_image(&a.omap[1], y, co, wd, h, n, (long long)co * 4, (long long)wd * co * 4, (long long)h * wd * co * 4, 8, 4, 0)) ||
        !map_weights(a, w, k * k, per_sample ? n : 1, rnd)) {
        spi_set_error("spi_conv2d_tc2: cuTensorMapEncodeTiled failed");
        return SPI_ERR_CUDA;
    }
    for (int i = 1; i < MAXV; i++) a.amap[i] = a.amap[0];
    Program& P = a.prog[0];
    P.nviews = 1; P.omap = 0; P.tile_begin = 0; P.fuse_epilogue = 1;
    P.tiles_x = cdiv(wd, tile_w); P.tiles_y = cdiv(h, a.trows);
    View& V = P.views[0];
    V.amap = 0; V.oy = -halo; V.ox = -halo; V.ntaps = k * k;
    for (int ky = 0; ky < k; ky++)
        for (int kx = 0; kx < k; kx++) V.taps[ky * k + kx] = Tap{ky, (flags & 8) ? 0 : kx, ky * k + kx};
    a.nprog = 1;
    a.total = P.tiles_x * P.tiles_y * a.tiles_o * n * a.splitk;
    a.bias = bias; a.noise = noise; a.noise_strength = noise_strength; a.noise_w = wd; a.out_h = h; a.out_w = wd;
    a.act = act; a.slope = slope; a.gain = gain; a.clamp = clamp;
    return launch2(a, stream);
}

// Stride-2 transposed 3x3 convolution, no padding: y[n, 2 iy + ky, 2 ix + kx, o] += x[n, iy, ix, i] * w[g, o, ky*3+kx, i].
// x [N][H][W][Ci], w [G][Co][9][Ci], y [N][2H+1][2W+1][Co].  Four output-parity phases in one launch (heaviest first).
static int conv_t2(bool plan_only, const float* x, const float* w, float* y, int n, int h, int wd, int ci, int co, int per_sample,
                   int flags, cudaStream_t stream) {
    SPI_CHECK_ARG(plan_only || (x && w && y), "spi_conv_transpose2d_s2_tc2: null tensor");
    SPI_CHECK_ARG(spi_conv_tc2_supported(ci, co), "spi_conv_transpose2d_s2_tc2: unsupported shape ci=%d co=%d", ci, co);
    SPI_CHECK_ARG((((uintptr_t)x | (uintptr_t)w | (uintptr_t)y) & 15) == 0, "spi_conv_transpose2d_s2_tc2: tensors must be 16-byte aligned");
    const int rnd = (flags & 1) ? 0 : 1;
    const int ho = 2 * h + 1, wo = 2 * wd + 1;
    Conv2Args a;
    common_args(a, n, ci, co, per_sample, wd + 1, h + 1, flags & ~256);
    a.px = a.mt * 8 + ((flags & 64) ? 8 : 1);
    const int rows = 17;
    a.patch_bytes = rows * a.px * 128;
    {
        int tp = 0;
        for (int q = 0; q < 4; q++) tp += cdiv(wd + 1 - (q & 1), a.mt * 8) * cdiv(h + 1 - (q >> 1), 16) * n;
        plan_tiles(a, tp, !(flags & 32));
    }
    if (plan_only) return a.splitk;
    zero_split_output(a, y, (size_t)n * ho * wo * co * 4, stream);
    bool ok = map_image(&a.amap[0], x, ci, wd, h, n, (long long)ci * 4, (long long)wd * ci * 4, (long long)h * wd * ci * 4, a.px, rows, rnd) &&
              map_weights(a, w, 9, per_sample ? n : 1, rnd);
    for (int i = 1; i < MAXV; i++) a.amap[i] = a.amap[0];
    int begin = 0;
    for (int q = 0; q < 4 && ok; q++) {
        const int py = q >> 1, px = q & 1;                // phase order (0,0) 4 taps, (0,1) 2, (1,0) 2, (1,1) 1
        const int hr = h + 1 - py, wr = wd + 1 - px;      // rows / columns of the phase
        ok = map_image(&a.omap[q], y + ((size_t)py * wo + px) * co, co, wr, hr, n, (long long)2 * co * 4, (long long)2 * wo * co * 4,
                       (long long)ho * wo * co * 4, 8, 16, 0);
        Program& P = a.prog[q];
        P.nviews = 1; P.omap = q; P.fuse_epilogue = 0;
        P.tiles_x = cdiv(wr, a.mt * 8); P.tiles_y = cdiv(hr, 16);
        P.tile_begin = begin;
        begin += P.tiles_x * P.tiles_y * a.tiles_o * n * a.splitk;
        View& V = P.views[0];
        V.amap = 0; V.oy = -1; V.ox = -1; V.ntaps = 0;
        for (int ky = py; ky < 3; ky += 2)
            for (int kx = px; kx < 3; kx += 2) V.taps[V.ntaps++] = Tap{1 - (ky >> 1), 1 - (kx >> 1), ky * 3 + kx};
    }
    if (!ok) { spi_set_error("spi_conv_transpose2d_s2_tc2: cuTensorMapEncodeTiled failed"); return SPI_ERR_CUDA; }
    a.nprog = 4; a.total = begin;
    return launch2(a, stream);
}

// Stride-2 3x3 correlation, no padding: y[n, j, i, o] = sum x[n, 2j + ky, 2i + kx, c] * w[g, o, ky*3+kx, c].
// x [N][2H+1][2W+1][Ci], w [G][Co][9][Ci], y [N][H][W][Co].  The four input-parity views are strided tensor maps.
static int conv_s2(bool plan_only, const float* x, const float* w, float* y, int n, int h, int wd, int ci, int co, int per_sample, int flags,
                   cudaStream_t stream) {
    SPI_CHECK_ARG(plan_only || (x && w && y), "spi_conv2d_s2_tc2: null tensor");
    SPI_CHECK_ARG(spi_conv_tc2_supported(ci, co), "spi_conv2d_s2_tc2: unsupported shape ci=%d co=%d", ci, co);
    SPI_CHECK_ARG((((uintptr_t)x | (uintptr_t)w | (uintptr_t)y) & 15) == 0, "spi_conv2d_s2_tc2: tensors must be 16-byte aligned");
    const int rnd = (flags & 1) ? 0 : 1;
    const int hi = 2 * h + 1, wi = 2 * wd + 1;
    Conv2Args a;
    common_args(a, n, ci, co, per_sample, wd, h, flags & ~256);
    a.px = a.mt * 8 + ((flags & 64) ? 8 : 1);
    const int rows = 17;
    a.patch_bytes = rows * a.px * 128;
    plan_tiles(a, cdiv(wd, a.mt * 8) * cdiv(h, 16) * n, !(flags & 32));
    if (plan_only) return a.splitk;
    zero_split_output(a, y, (size_t)n * h * wd * co * 4, stream);
    bool ok = map_weights(a, w, 9, per_sample ? n : 1, rnd) &&
              map_image(&a.omap[0], y, co, wd, h, n, (long long)co * 4, (long long)wd * co * 4, (long long)h * wd * co * 4, 8, 16, 0);
    Program& P = a.prog[0];
    P.nviews = 4; P.omap = 0; P.fuse_epilogue = 0; P.tile_begin = 0;
    P.tiles_x = cdiv(wd, a.mt * 8); P.tiles_y = cdiv(h, 16);
    for (int q = 0; q < 4 && ok; q++) {
        const int py = q >> 1, px = q & 1;
        ok = map_image(&a.amap[q], x + ((size_t)py * wi + px) * ci, ci, (wi - px + 1) / 2, (hi - py + 1) / 2, n, (long long)2 * ci * 4,
                       (long long)2 * wi * ci * 4, (long long)hi * wi * ci * 4, a.px, rows, rnd);
        View& V = P.views[q];
        V.amap = q; V.oy = 0; V.ox = 0; V.ntaps = 0;
        for (int ky = py; ky < 3; ky += 2)
            for (int kx = px; kx < 3; kx += 2) V.taps[V.ntaps++] = Tap{ky >> 1, kx >> 1, ky * 3 + kx};
    }
    if (!ok) { spi_set_error("spi_conv2d_s2_tc2: cuTensorMapEncodeTiled failed"); return SPI_ERR_CUDA; }
    a.nprog = 1;
    a.total = P.tiles_x * P.tiles_y * a.tiles_o * n * a.splitk;
    return launch2(a, stream);
}

extern "C" int spi_conv2d_tc2(const float* x, const float* w, float* y, int n, int h, int wd, int ci, int co, int k, int per_sample,
                              const float* bias, const float* noise, const float* noise_strength, int act, float slope, float gain, float clamp,
                              int flags, cudaStream_t stream) {
    return conv_s1(false, x, w, y, n, h, wd, ci, co, k, per_sample, bias, noise, noise_strength, act, slope, gain, clamp, flags, stream);
}
extern "C" int spi_conv_transpose2d_s2_tc2(const float* x, const float* w, float* y, int n, int h, int wd, int ci, int co, int per_sample,
                                           int flags, cudaStream_t stream) {
    return conv_t2(false, x, w, y, n, h, wd, ci, co, per_sample, flags, stream);
}
extern "C" int spi_conv2d_s2_tc2(const float* x, const float* w, float* y, int n, int h, int wd, int ci, int co, int per_sample, int flags,
                                 cudaStream_t stream) {
    return conv_s2(false, x, w, y, n, h, wd, ci, co, per_sample, flags, stream);
}

// How many Cin splits the call of the given form (0 stride-1 'same', 1 stride-2 transposed, 2 stride-2; same n/h/wd/ci/co/k/flags, epilogue
// != 0 when a fused epilogue is requested) will use.  > 1 means the output is accumulated with reduce-adds and must be zero on entry: the
// entry points fill it themselves unless flags bit 9 (512) says the caller already did.  Negative: invalid arguments.  No device work.
extern "C" int spi_conv_tc2_splits(int form, int n, int h, int wd, int ci, int co, int k, int per_sample, int epilogue, int flags) {
    int rc;
    if (form == 0) rc = conv_s1(true, nullptr, nullptr, nullptr, n, h, wd, ci, co, k, per_sample, nullptr, nullptr, nullptr, epilogue ? 2 : 0, 0.2f, 1.f, -1.f, flags, nullptr);
    else if (form == 1) rc = conv_t2(true, nullptr, nullptr, nullptr, n, h, wd, ci, co, per_sample, flags, nullptr);
    else if (form == 2) rc = conv_s2(true, nullptr, nullptr, nullptr, n, h, wd, ci, co, per_sample, flags, nullptr);
    else { spi_set_error("spi_conv_tc2_splits: form must be 0, 1 or 2"); return SPI_ERR_ARG; }
    return rc;
}

extern "C" int spi_conv_weight_transpose(const float* w, float* wt, int g, int o, int taps, int i, int reverse, cudaStream_t stream) {
    SPI_CHECK_ARG(w && wt && g > 0 && o > 0 && taps > 0 && i > 0, "spi_conv_weight_transpose: bad arguments");
    dim3 grid(cdiv(i, 32), cdiv(o, 32), g * taps), block(32, 8);
    SPI_CHECK_ARG(grid.z <= 65535 && grid.y <= 65535, "spi_conv_weight_transpose: too many groups");
    weight_transpose_kernel<<<grid, block, 0, stream>>>(w, wt, o, taps, i, reverse ? 1 : 0);
    SPI_COUNT_LAUNCH(1);
    SPI_LAUNCH_CHECK("spi_conv_weight_transpose");
    return SPI_OK;
}
