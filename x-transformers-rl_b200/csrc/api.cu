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
int num_sms() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
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

extern "C" int xtrl_device_num_sms(void) { return xtrl::num_sms(); }

extern "C" int64_t xtrl_struct_size(int32_t which) {
    switch (which) {
        case 0: return sizeof(xtrl_model_desc);
        case 1: return sizeof(xtrl_weight_layout);
        case 2: return sizeof(xtrl_step_io);
        case 3: return sizeof(xtrl_loss_cfg);
        case 4: return sizeof(xtrl_loss_io);
        case 5: return sizeof(xtrl_pool_desc);
        case 6: return sizeof(xtrl_opt_cfg);
        case 7: return sizeof(xtrl_ff_io);
        case 8: return sizeof(xtrl_attn_io);
        case 9: return sizeof(xtrl_attn_core_io);
        case 10: return sizeof(xtrl_mlp_io);
        case 11: return sizeof(xtrl_prep_io);
        case 12: return sizeof(xtrl_prep_bwd_io);
        case 13: return sizeof(xtrl_peer_ctx);
        case 14: return sizeof(xtrl_seg2d);
        default: return -1;
    }
}
