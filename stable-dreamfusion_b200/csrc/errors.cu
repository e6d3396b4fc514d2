// errors.cu — last-error string of the C-ABI (include/sdf_b200.h: sdf_last_error).
#include "common.cuh"
#include <stdarg.h>

static thread_local char g_err[512] = "";

void sdf_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

SDF_API const char* sdf_last_error(void) { return g_err; }

SDF_API int sdf_abi_version(void) { return 2; }

int sdf_num_sms() {
    static const int n = [] {
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        return v;
    }();
    return n;
}

#include <cstdlib>
bool sdf_pdl_enabled() {
    static const bool on = [] { const char* e = getenv("SDF_PDL"); return !(e && e[0] == '0'); }();
    return on;
}
