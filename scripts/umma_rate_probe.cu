// umma_rate_probe.cu -- cycles per tcgen05.mma for the small shapes of the attention core (one CTA, one issuing thread):
// M=128, K=16, N in {16, 64, 128}; operands K-major or MN-major; accumulating into one TMEM tile or rotating over 4.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/umma_rate_probe scripts/umma_rate_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../x-transformers-rl_b200/csrc/umma.cuh"
using namespace xtrl::umma;

__global__ void __launch_bounds__(128) probe(long long *out, int N, int mn_major, int rotate, int reps) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *bar = (uint64_t *)(smem + 96 * 1024);
    uint32_t *tslot = (uint32_t *)(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 96 * 1024 / 4; i += 128) ((uint32_t *)smem)[i] = 0x3c003c00u;
    if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(tslot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tslot;
    if (tid == 0) {
        uint32_t idesc = idesc_bf16(128, N);
        if (mn_major) idesc |= (1u << 15) | (1u << 16);
        const unsigned char *sA = smem, *sB = smem + 48 * 1024;
        const uint64_t ad0 = mn_major ? smem_desc(sA, 128, 2048) : smem_desc(sA, 2048, 128);
        const uint64_t bd0 = mn_major ? smem_desc(sB, 128, 2048) : smem_desc(sB, 2048, 128);
        const uint32_t step = mn_major ? 256 / 16 : 4096 / 16;          // descriptor start address advances in 16-byte units
        long long t0 = clock64();
#pragma unroll 8
        for (int r = 0; r < reps; ++r) {
            const uint32_t s = r & 7;
            mma_bf16_ss(tbase + (rotate ? (r & 3) * 128 : 0), ad0 + s * step, bd0 + s * step, idesc, 1);
        }
        long long t1 = clock64();
        mma_commit(bar);
        mbar_wait(bar, 0);
        long long t2 = clock64();
        out[0] = t1 - t0; out[1] = t2 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

int main() {
    long long *d; cudaMalloc(&d, 16);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024 + 64);
    const int reps = 256;
    for (int mn = 0; mn < 2; ++mn) for (int N : {16, 64, 128}) for (int rot = 0; rot < 2; ++rot) {
        long long h[2];
        for (int it = 0; it < 2; ++it) {
            probe<<<1, 128, 96 * 1024 + 64>>>(d, N, mn, rot, reps);
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("CUDA error\n"); return 1; }
            cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        }
        printf("M=128 N=%3d K=16 %s %s: issue %.1f cyc/mma, complete %.1f cyc/mma\n", N, mn ? "MN-major" : "K-major ", rot ? "4 accumulators" : "1 accumulator ",
               (double)h[0] / reps, (double)h[1] / reps);
    }
    return 0;
}
