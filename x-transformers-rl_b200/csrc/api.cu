// api.cu -- error slot, version and device probes of the C ABI (include/xtrl.h).
#include "common.cuh"

namespace xtrl {
static thread_local char g_err[512] = "";
char *error_slot() { return g_err; }
int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
}  // namespace xtrl

extern "C" int xtrl_abi_version(void) { return XTRL_ABI_VERSION; }
extern "C" const char *xtrl_last_error(void) { return xtrl::error_slot(); }
extern "C" int xtrl_device_sm(void) {
    int dev = 0, major = 0, minor = 0;
    XTRL_CUDA(cudaGetDevice(&dev));
    XTRL_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    XTRL_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    return major * 10 + minor;
}
