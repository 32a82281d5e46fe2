// spi_b200 common device/host helpers (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define SPI_OK 0
#define SPI_ERR_ARG (-1)       // invalid argument (python shim raises RuntimeError, like TORCH_CHECK)
#define SPI_ERR_UNSUPPORTED (-2)
#define SPI_ERR_CUDA (-3)

#define SPI_DT_F32 0
#define SPI_DT_F16 1
#define SPI_DT_F64 2

void spi_set_error(const char* fmt, ...);
int* spi_tc_err_flag();          // device int, non-zero after a tcgen05 pipeline barrier timed out (lib.cu)
extern "C" int spi_tc_error();

#define SPI_CHECK_ARG(cond, ...)                      \
    do {                                              \
        if (!(cond)) {                                \
            spi_set_error(__VA_ARGS__);               \
            return SPI_ERR_ARG;                       \
        }                                             \
    } while (0)

#define SPI_LAUNCH_CHECK(name)                                                      \
    do {                                                                            \
        cudaError_t e__ = cudaGetLastError();                                       \
        if (e__ != cudaSuccess) {                                                   \
            spi_set_error("%s: CUDA launch failed: %s", name, cudaGetErrorString(e__)); \
            return SPI_ERR_CUDA;                                                    \
        }                                                                           \
    } while (0)

// Launch counter: every kernel launch of this library bumps it (bench.py reports it as gpu_launches).
extern unsigned long long g_spi_launches;
#define SPI_COUNT_LAUNCH(n) (g_spi_launches += (n))

static inline int spi_num_sms() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------- device helpers
#ifdef __CUDACC__

template <class T> struct Acc { typedef float t; };
template <> struct Acc<double> { typedef double t; };

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// streaming (read-once) 128-bit load / store that do not pollute L1
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream(float4* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w));
}
// vector reduction into global memory (sm_90+): one 16-byte RED instead of four
__device__ __forceinline__ void red_add_v4(float* p, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ float softplus_f(float x) {  // torch.nn.functional.softplus, beta=1, threshold=20
    return x > 20.f ? x : log1pf(expf(x));
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

#endif  // __CUDACC__
