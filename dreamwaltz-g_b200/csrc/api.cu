// libdwg_sm100.so: error reporting and library-level entry points.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace dwg {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace dwg

extern "C" const char* dwg_last_error(void) { return dwg::g_err; }
extern "C" int dwg_version(void) { return 100; }
extern "C" int dwg_device_cc(void) {
    int dev = 0;
    cudaDeviceProp p;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&p, dev) != cudaSuccess) {
        dwg::set_error("dwg_device_cc: no CUDA device");
        return DWG_ERR_CUDA;
    }
    return p.major * 10 + p.minor;
}
