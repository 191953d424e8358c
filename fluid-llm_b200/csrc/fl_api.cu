// fl_api.cu -- library-level entry points of libfluidgrid.so (error text, version, device probe).
#include "fl_common.cuh"

static thread_local char g_err[512] = "";

void fl_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* fl_last_error(void) { return g_err; }
extern "C" int fl_abi_version(void) { return FL_ABI_VERSION; }
extern "C" int fl_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
