// Stand-alone probe of MN-major tcgen05.mma.kind::tf32 operands (not part of the library): one MMA, M = 128, K = 8, N in {32, 96},
// operands laid out in shared memory exactly as a TMA box of [pixels][32 channels] lands (128-byte rows, SWIZZLE_128B), every
// combination of {K-major, MN-major} x {LBO / SBO assignment}, checked against a host evaluation.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o spi_b200/build/mn_probe tools/mn_probe.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../spi_b200/csrc/tc05.cuh"

using namespace tc05;

__device__ __forceinline__ void mma64(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(da),
                 "l"(db), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ uint64_t mkdesc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout = 2) {
    return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) | ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46) |
           ((uint64_t)layout << 61);
}
// MN-major tf32 operands must use SWIZZLE_128B_BASE32B (layout type 1; TMA: CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): 32-byte chunks of a
// 128-byte row XOR-ed with (row & 3); atoms are 4 K-rows deep
__device__ __forceinline__ uint32_t sw32(int r, int c) { return (uint32_t)(r * 128 + ((((c >> 3) ^ r) & 3) << 5) + (c & 7) * 4); }
// byte offset of element (row r, float column c) in a region of 128-byte rows with the SWIZZLE_128B pattern (1024-byte atoms)
__device__ __forceinline__ uint32_t sw(int r, int c) { return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((c >> 2) ^ r) & 7) << 4) + (c & 3) * 4); }

// cfg: bit0 A MN-major, bit1 B MN-major, bit2 swap LBO/SBO roles for MN-major operands, bits 4.. : B pixel shift, bit 8: N = 96 (LBO_B = 128)
__global__ void __launch_bounds__(128) probe(const float* Amk, const float* P, float* Dout, int cfg) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    uint8_t* sm = raw + (base - smem_u32(raw));
    uint8_t* sA = sm;                     // MN-major: 4 atoms (32 m each) of [8 k-rows][128 B] at 16 KB stride; K-major: [128 m-rows][32 B used of 128]
    uint8_t* sB = sm + 65536;             // patch: 32 pixel rows x 128 B
    uint64_t* bar = (uint64_t*)(sm + 65536 + 16384);
    uint32_t* slot = (uint32_t*)(bar + 2);
    const int tid = threadIdx.x, warp = tid >> 5;
    const bool amn = cfg & 1, bmn = cfg & 2, swap = cfg & 4;
    const int shift = (cfg >> 4) & 15;
    const int N = (cfg & 256) ? 96 : 32;
    if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(slot, 128);
    for (int i = tid; i < 65536 / 4; i += 128) ((float*)sA)[i] = 0.f;
    for (int i = tid; i < 16384 / 4; i += 128) ((float*)sB)[i] = 0.f;
    __syncthreads();
    // A[m][k], m < 128, k < 8
    for (int i = tid; i < 128 * 8; i += 128) {
        const int m = i >> 3, k = i & 7;
        const float v = Amk[i];
        if (amn) *(float*)(sA + (m >> 5) * 16384 + sw32(k, m & 31)) = v;      // row = k (pixel), column = channel inside the 32-chunk
        else *(float*)(sA + sw(m, k)) = v;                                   // K-major: row = m, first 8 floats of the 128-byte row
    }
    // P[r][c]: 32 pixel rows x 32 channels.  B[k][n] = P[k + shift + n / 32][n % 32] (MN-major view) ; K-major view: row = n, col = k
    for (int i = tid; i < 32 * 32; i += 128) {
        const int r = i >> 5, c = i & 31;
        if (bmn) *(float*)(sB + sw32(r, c)) = P[i];
    }
    if (!bmn) {
        for (int i = tid; i < N * 8; i += 128) {
            const int n = i >> 3, k = i & 7;
            *(float*)(sB + sw(n, k)) = P[(k + shift + n / 32) * 32 + (n & 31)];
        }
    }
    fence_async_smem();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tm = *slot;
    if (tid == 0) {
        uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        if (amn) idesc |= 1u << 15;
        if (bmn) idesc |= 1u << 16;
        uint64_t da, db;
        if (amn) da = swap ? mkdesc(smem_u32(sA), 512, 16384, 1) : mkdesc(smem_u32(sA), 16384, 512, 1);
        else da = mkdesc(smem_u32(sA), 16, 1024);
        if (bmn) db = swap ? mkdesc(smem_u32(sB) + shift * 128, 512, 128, 1) : mkdesc(smem_u32(sB) + shift * 128, 128, 512, 1);
        else db = mkdesc(smem_u32(sB), 16, 1024);
        mma64(tm, da, db, idesc, 0);
        commit(bar);
    }
    mbar_wait(bar, 0);
    fence_after();
    float v[32];
    for (int c = 0; c < N / 32; c++) {
        tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + c * 32, v);
        tmem_wait_ld();
        for (int j = 0; j < 32; j++) Dout[tid * 96 + c * 32 + j] = v[j];
    }
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 128);
}

int main() {
    std::vector<float> A(128 * 8), P(32 * 32), D(128 * 96);
    srand(3);
    auto rnd = []() { return (float)((rand() % 2001) - 1000) / 1024.f; };       // exactly representable in TF32
    for (auto& v : A) v = rnd();
    for (auto& v : P) v = rnd();
    float *dA, *dP, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dP, P.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dP, P.data(), P.size() * 4, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    const int cfgs[] = {0, 1, 2, 3, 5, 6, 7, 3 | (1 << 4), 3 | (3 << 4), 2 | (1 << 4), 3 | 256, 3 | 256 | (1 << 4), 2 | 256, 7 | 256, 0 | 256, 1 | 256};
    for (int cfg : cfgs) {
        cudaMemset(dD, 0xff, D.size() * 4);
        probe<<<1, 128, 96 * 1024>>>(dA, dP, dD, cfg);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        const int shift = (cfg >> 4) & 15, N = (cfg & 256) ? 96 : 32;
        double maxerr = 0, maxref = 0, maxabs = 0;
        for (int m = 0; m < 128; m++)
            for (int n = 0; n < N; n++) {
                double ref = 0;
                for (int k = 0; k < 8; k++) ref += (double)A[m * 8 + k] * P[(k + shift + n / 32) * 32 + (n & 31)];
                maxerr = fmax(maxerr, fabs(ref - D[m * 96 + n]));
                maxref = fmax(maxref, fabs(ref));
                maxabs = fmax(maxabs, fabs((double)D[m * 96 + n]));
            }
        printf("cfg A=%s B=%s swap=%d shift=%d N=%d: %s  max|D|=%.3f max|ref|=%.3f max err=%.3e  %s\n", (cfg & 1) ? "MN" : "K ", (cfg & 2) ? "MN" : "K ", (cfg >> 2) & 1,
               shift, N, cudaGetErrorString(e), maxabs, maxref, maxerr, maxerr < 1e-3 ? "OK" : "WRONG");
    }
    return 0;
}
