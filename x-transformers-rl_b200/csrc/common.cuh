// common.cuh -- shared plumbing of libxtrl.so (error slot, launch checks, small device helpers).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/xtrl.h"

namespace xtrl {

char *error_slot();                                   // thread-local buffer (api.cu)
int   fail(int code, const char *fmt, ...);           // formats into the slot, returns code

#define XTRL_REQUIRE(cond, code, ...) do { if (!(cond)) return ::xtrl::fail((code), __VA_ARGS__); } while (0)
#define XTRL_CUDA(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) \
    return ::xtrl::fail(XTRL_E_CUDA, "%s failed: %s", #call, cudaGetErrorString(_e)); } while (0)
#define XTRL_LAUNCH_CHECK(name) do { cudaError_t _e = cudaGetLastError(); if (_e != cudaSuccess) \
    return ::xtrl::fail(XTRL_E_CUDA, "launch of %s failed: %s", (name), cudaGetErrorString(_e)); } while (0)

int   num_sms();                                     // SM count of the current device, queried once (api.cu); 148 on B200

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
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// streaming (read-once) 128-bit load that does not pollute L1
__device__ __forceinline__ float4 ld_stream_f4(const float4 *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint4 ld_stream_u4(const uint4 *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

}  // namespace xtrl
