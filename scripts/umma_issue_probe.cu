// umma_issue_probe.cu -- is the ~48-cycle cost of a small tcgen05.mma a property of the issuing thread or of the tensor core?
// W warps (1, 2, 4) each issue `reps` MMAs (M=128, N=16|64|128, K=16; SS or TS) into their own accumulators at the same time.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/umma_issue_probe scripts/umma_issue_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../x-transformers-rl_b200/csrc/umma.cuh"
using namespace xtrl::umma;

template <int N, bool TS>
__global__ void __launch_bounds__(256) probe(long long *out, int W, int reps) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *bar = (uint64_t *)(smem + 96 * 1024);
    uint32_t *tslot = (uint32_t *)(bar + 8);
    __shared__ long long t_start[8], t_end[8];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 96 * 1024 / 4; i += 256) ((uint32_t *)smem)[i] = 0x3c003c00u;
    if (tid == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(tslot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tslot;
    if (warp < W && lane == 0) {
        const uint32_t idesc = idesc_bf16(128, N);
        const unsigned char *sA = smem + warp * 8192, *sB = smem + 48 * 1024 + warp * 8192;
        const uint64_t ad = smem_desc(sA, 2048, 128), bd = smem_desc(sB, 2048, 128);
        const uint32_t td = tbase + warp * 128, ta = tbase + 256 + warp * 32;
        const long long t0 = clock64();
#pragma unroll 8
        for (int r = 0; r < reps; ++r) {
            if (TS) mma_bf16_ts(td, ta + (r & 3) * 8, bd, idesc, 1);
            else mma_bf16_ss(td, ad, bd, idesc, 1);
        }
        mma_commit(&bar[warp]);
        mbar_wait(&bar[warp], 0);
        t_start[warp] = t0; t_end[warp] = clock64();
    }
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
        long long a = t_start[0], b = t_end[0];
        for (int w = 1; w < W; ++w) { a = t_start[w] < a ? t_start[w] : a; b = t_end[w] > b ? t_end[w] : b; }
        out[0] = b - a;
    }
    if (warp == 0) tmem_dealloc(tbase, 512);
}

template <int N, bool TS>
void run(long long *d) {
    cudaFuncSetAttribute(probe<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024 + 128);
    const int reps = 512;
    for (int W : {1, 2, 4}) {
        long long h = 0;
        for (int it = 0; it < 2; ++it) {
            probe<N, TS><<<1, 256, 96 * 1024 + 128>>>(d, W, reps);
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(cudaGetLastError())); return; }
            cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        }
        printf("%s M=128 N=%3d K=16, %d issuing warp(s): %.1f cycles per MMA per warp, %.1f cycles per MMA overall\n", TS ? "TS" : "SS", N, W,
               (double)h / reps, (double)h / reps / W);
    }
}

int main() {
    long long *d; cudaMalloc(&d, 16);
    run<16, false>(d); run<64, false>(d); run<128, false>(d);
    run<16, true>(d); run<64, true>(d); run<128, true>(d);
    return 0;
}
