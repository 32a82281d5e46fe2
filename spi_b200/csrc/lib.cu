// libspi_b200: library-wide state (error string, launch counter, version).
#include "common.cuh"
#include <stdarg.h>

static thread_local char g_err[512] = "";
unsigned long long g_spi_launches = 0;

void spi_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* spi_last_error() { return g_err; }
extern "C" unsigned long long spi_launch_count() { return g_spi_launches; }
extern "C" void spi_reset_launch_count() { g_spi_launches = 0; }
extern "C" int spi_abi_version() { return 3; }       // 3: spi_lpips_tap_* take per-sample weights; grouped (style / modulation bank) entry points

// Device flag shared by the tcgen05 kernels (conv_tc05.cu, raymarch_tc.cuh): set when a bounded mbarrier wait timed out.
int* spi_tc_err_flag() {
    static int* flag = nullptr;
    if (!flag) {
        cudaMalloc(&flag, sizeof(int));
        cudaMemset(flag, 0, sizeof(int));
    }
    return flag;
}
extern "C" int spi_tc_error() {
    int v = 0;
    cudaDeviceSynchronize();
    cudaMemcpy(&v, spi_tc_err_flag(), sizeof(int), cudaMemcpyDeviceToHost);
    if (v) cudaMemset(spi_tc_err_flag(), 0, sizeof(int));
    return v;
}
