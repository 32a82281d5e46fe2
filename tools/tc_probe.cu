// Stand-alone probe of the tcgen05 building blocks in spi_b200/csrc/tc05.cuh (not part of the library):
//   D1[128x64] = A[128x32] W1^T            A, W1 from shared memory (SWIZZLE_128B K-major), 3xTF32
//   D2[128x48] = softplus(D1 + b1) W2^T    A from tensor memory (written with tcgen05.st), W2 in two 32-wide K slabs, 3xTF32
// against an fp64 host evaluation.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o spi_b200/build/tc_probe tools/tc_probe.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../spi_b200/csrc/tc05.cuh"

using namespace tc05;

__global__ void __launch_bounds__(128) probe(const float* A, const float* W1, const float* b1, const float* W2, float* D1out, float* D2out,
                                             int* err, int passes) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    uint8_t* sm = raw + (base - smem_u32(raw));
    uint8_t* a_hi = sm;                 // 128 x 128 B
    uint8_t* a_lo = sm + 16384;
    uint8_t* w1_hi = sm + 32768;        // 64 x 128 B
    uint8_t* w1_lo = sm + 40960;
    uint8_t* w2_hi = sm + 49152;        // 2 slabs x 48 x 128 B
    uint8_t* w2_lo = sm + 61440;
    uint64_t* bars = (uint64_t*)(sm + 73728);
    uint32_t* slot = (uint32_t*)(sm + 73728 + 64);
    const int tid = threadIdx.x, warp = tid >> 5;

    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(slot, 256);
    {   // A: one row per thread
        for (int kc = 0; kc < 8; kc++) {
            uint32_t h[4], l[4];
            for (int e = 0; e < 4; e++) split(A[tid * 32 + 4 * kc + e], h[e], l[e]);
            *(uint4*)(a_hi + swz(tid, kc)) = make_uint4(h[0], h[1], h[2], h[3]);
            *(uint4*)(a_lo + swz(tid, kc)) = make_uint4(l[0], l[1], l[2], l[3]);
        }
        for (int i = tid; i < 64 * 32; i += 128) {
            int n = i / 32, k = i % 32;
            uint32_t h, l;
            split(W1[i], h, l);
            *(uint32_t*)(w1_hi + swz(n, k >> 2) + (k & 3) * 4) = h;
            *(uint32_t*)(w1_lo + swz(n, k >> 2) + (k & 3) * 4) = l;
        }
        for (int i = tid; i < 48 * 64; i += 128) {
            int n = i / 64, k = i % 64, kk = k & 31;
            uint32_t h, l;
            split(n < 33 ? W2[n * 64 + k] : 0.f, h, l);
            *(uint32_t*)(w2_hi + (k >> 5) * 6144 + swz(n, kk >> 2) + (kk & 3) * 4) = h;
            *(uint32_t*)(w2_lo + (k >> 5) * 6144 + swz(n, kk >> 2) + (kk & 3) * 4) = l;
        }
    }
    fence_async_smem();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tm = *slot;
    if (tid == 0) {
        const uint32_t id1 = idesc_tf32(128, 64);
        uint32_t acc = 0;
        for (int pass = 0; pass < passes; pass++) {
            const uint8_t* a = pass == 1 ? a_lo : a_hi;
            const uint8_t* w = pass == 2 ? w1_lo : w1_hi;
            for (int ks = 0; ks < 4; ks++) {
                mma_ss(tm, desc_sw128(smem_u32(a) + ks * 32), desc_sw128(smem_u32(w) + ks * 32), id1, acc);
                acc = 1;
            }
        }
        commit(&bars[0]);
    }
    if (!mbar_wait(&bars[0], 0)) { if (tid == 0) atomicOr(err, 1); }
    fence_after();
    const uint32_t lane_base = tm + ((uint32_t)(warp * 32) << 16);
    float v[64];
    tmem_ld32(lane_base, v);
    tmem_ld32(lane_base + 32, v + 32);
    tmem_wait_ld();
    uint32_t hh[64], hl[64];
    for (int c = 0; c < 64; c++) {
        D1out[tid * 64 + c] = v[c];
        float x = v[c] + b1[c];
        float sp = fmaxf(x, 0.f) + log1pf(expf(-fabsf(x)));
        split(sp, hh[c], hl[c]);
    }
    tmem_st32(lane_base + 64, hh);
    tmem_st32(lane_base + 96, hh + 32);
    tmem_st32(lane_base + 128, hl);
    tmem_st32(lane_base + 160, hl + 32);
    tmem_wait_st();
    fence_before();
    __syncthreads();
    if (tid == 0) {
        fence_after();
        const uint32_t id2 = idesc_tf32(128, 48);
        uint32_t acc = 0;
        for (int pass = 0; pass < passes; pass++) {
            const uint32_t a = tm + (pass == 1 ? 128 : 64);
            const uint8_t* w = pass == 2 ? w2_lo : w2_hi;
            for (int ks = 0; ks < 8; ks++) {
                mma_ts(tm + 192, a + ks * 8, desc_sw128(smem_u32(w) + (ks >> 2) * 6144 + (ks & 3) * 32), id2, acc);
                acc = 1;
            }
        }
        commit(&bars[1]);
    }
    if (!mbar_wait(&bars[1], 0)) { if (tid == 0) atomicOr(err, 2); }
    fence_after();
    float o[48];
    tmem_ld32(lane_base + 192, o);
    tmem_ld16(lane_base + 224, o + 32);
    tmem_wait_ld();
    for (int c = 0; c < 48; c++) D2out[tid * 48 + c] = o[c];
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 256);
}

static float frand() { return (float)rand() / RAND_MAX * 2.f - 1.f; }

int main() {
    std::vector<float> A(128 * 32), W1(64 * 32), b1(64), W2(33 * 64);
    srand(1);
    for (auto& x : A) x = frand() * 2.f;
    for (auto& x : W1) x = frand();
    for (auto& x : b1) x = frand();
    for (auto& x : W2) x = frand();
    float *dA, *dW1, *db1, *dW2, *dD1, *dD2;
    int* derr;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dW1, W1.size() * 4); cudaMalloc(&db1, 256); cudaMalloc(&dW2, W2.size() * 4);
    cudaMalloc(&dD1, 128 * 64 * 4); cudaMalloc(&dD2, 128 * 48 * 4); cudaMalloc(&derr, 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dW1, W1.data(), W1.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(db1, b1.data(), 256, cudaMemcpyHostToDevice);
    cudaMemcpy(dW2, W2.data(), W2.size() * 4, cudaMemcpyHostToDevice);
    const int smem = 73728 + 128 + 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    std::vector<double> R1(128 * 64), R2(128 * 48, 0.0);
    for (int m = 0; m < 128; m++)
        for (int n = 0; n < 64; n++) {
            double s = 0;
            for (int k = 0; k < 32; k++) s += (double)A[m * 32 + k] * W1[n * 32 + k];
            R1[m * 64 + n] = s;
        }
    for (int m = 0; m < 128; m++)
        for (int n = 0; n < 33; n++) {
            double s = 0;
            for (int k = 0; k < 64; k++) {
                double x = R1[m * 64 + k] + b1[k];
                double sp = fmax(x, 0.0) + log1p(exp(-fabs(x)));
                s += sp * W2[n * 64 + k];
            }
            R2[m * 48 + n] = s;
        }
    for (int passes = 1; passes <= 3; passes += 2) {
        cudaMemset(derr, 0, 4); cudaMemset(dD1, 0, 128 * 64 * 4); cudaMemset(dD2, 0, 128 * 48 * 4);
        probe<<<1, 128, smem>>>(dA, dW1, db1, dW2, dD1, dD2, derr, passes);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<float> D1(128 * 64), D2(128 * 48);
        int err = 0;
        cudaMemcpy(D1.data(), dD1, D1.size() * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(D2.data(), dD2, D2.size() * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(&err, derr, 4, cudaMemcpyDeviceToHost);
        double e1 = 0, e2 = 0, m1 = 0, m2 = 0;
        for (size_t i = 0; i < R1.size(); i++) { e1 = fmax(e1, fabs(D1[i] - R1[i])); m1 = fmax(m1, fabs(R1[i])); }
        for (size_t i = 0; i < R2.size(); i++) { e2 = fmax(e2, fabs(D2[i] - R2[i])); m2 = fmax(m2, fabs(R2[i])); }
        printf("passes=%d cuda=%s flag=%d  D1 max|err| %.3e (max|ref| %.2f)   D2 max|err| %.3e (max|ref| %.2f)\n", passes,
               cudaGetErrorString(e), err, e1, m1, e2, m2);
        if (e1 > 1e-2 * m1) {
            printf("D1 row0: "); for (int c = 0; c < 8; c++) printf("%.4f/%.4f ", D1[c], R1[c]); printf("\n");
            printf("D1 row9: "); for (int c = 0; c < 8; c++) printf("%.4f/%.4f ", D1[9 * 64 + c], R1[9 * 64 + c]); printf("\n");
            printf("D1 row77: "); for (int c = 0; c < 8; c++) printf("%.4f/%.4f ", D1[77 * 64 + c], R1[77 * 64 + c]); printf("\n");
        }
        if (e2 > 1e-2 * m2) {
            printf("D2 row0: "); for (int c = 0; c < 8; c++) printf("%.4f/%.4f ", D2[c], R2[c]); printf("\n");
            printf("D2 row77: "); for (int c = 28; c < 36; c++) printf("%.4f/%.4f ", D2[77 * 48 + c], R2[77 * 48 + c]); printf("\n");
        }
    }
    return 0;
}
