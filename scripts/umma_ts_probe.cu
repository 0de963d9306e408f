// umma_ts_probe.cu -- tcgen05.mma with the A operand in tensor memory (written by tcgen05.st), B from a token-major shared-memory
// image read MN-major (the layout the attention core keeps q / k / v / dO in): numerics against a host reference and cycles per MMA.
//   D[128 x 16] = A[128 x K] * Bt,  A[m][k] in TMEM (lane m, column k/2 holds bf16 pair (k, k+1)),  B image: token k, dim n.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/umma_ts_probe scripts/umma_ts_probe.cu
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
#include "../x-transformers-rl_b200/csrc/umma.cuh"
using namespace xtrl::umma;

constexpr int KTOT = 128, R = 512, kChunk = R * 16;       // B image: 2 chunks of 8 dims x 512 rows x 16 B
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t *>(&h);
}
__host__ __device__ inline float aval(int m, int k) { return (float)(((m * 7 + k * 3) % 13) - 6) * 0.25f; }
__host__ __device__ inline float bval(int k, int n) { return (float)(((k * 5 + n * 11) % 9) - 4) * 0.5f; }

__global__ void __launch_bounds__(160) probe(float *out, long long *cyc, int reps) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *sB = smem;                                     // token-major image of B (rows = K tokens)
    uint64_t *bar = (uint64_t *)(smem + 2 * kChunk);
    uint32_t *tslot = (uint32_t *)(bar + 2);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < R * 2; i += 160) {
        const int r = i >> 1, c = i & 1;
        uint32_t w[4];
        for (int j = 0; j < 4; ++j) w[j] = r < KTOT ? pack2(bval(r, c * 8 + 2 * j), bval(r, c * 8 + 2 * j + 1)) : 0u;
        *reinterpret_cast<uint4 *>(sB + c * kChunk + (r >> 3) * 128 + (r & 7) * 16) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(tslot, 256);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tslot;
    constexpr uint32_t cA = 64, cD = 0;                           // A: 64 columns (K = 128), D: 16 columns
    if (warp < 4) {                                               // thread <-> TMEM lane (row m of A)
        const int m = warp * 32 + lane;
        const uint32_t tl = tbase + ((uint32_t)(warp * 32) << 16);
#pragma unroll
        for (int c0 = 0; c0 < KTOT / 2; c0 += 16) {
            uint32_t w[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) w[j] = pack2(aval(m, 2 * (c0 + j)), aval(m, 2 * (c0 + j) + 1));
            tmem_st16(tl + cA + c0, w);
        }
        tmem_st_wait();
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 4 && lane == 0) {
        tc_fence_after();
        const uint64_t bcols = smem_desc(sB, 128, kChunk);         // MN-major B: [N = 16 x K = 16 tokens], +16 per K step
        const uint32_t id = idesc_bf16(128, 16) | (1u << 16);      // B MN-major
        for (int s = 0; s < KTOT / 16; ++s) mma_bf16_ts(tbase + cD, tbase + cA + s * 8, bcols + s * 16, id, s > 0);
        mma_commit(&bar[0]);
        mbar_wait(&bar[0], 0);
        // rate: issue `reps` TS MMAs into a scratch accumulator
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r) mma_bf16_ts(tbase + 32, tbase + cA + (r & 7) * 8, bcols + (r & 7) * 16, id, 1);
        long long t1 = clock64();
        mma_commit(&bar[1]);
        mbar_wait(&bar[1], 0);
        long long t2 = clock64();
        cyc[0] = t1 - t0; cyc[1] = t2 - t0;
        // the SS form of the same product for comparison (A = the B image itself read K-major: timing only)
        const uint64_t arows = smem_desc(sB, kChunk, 128);
        long long t3 = clock64();
        for (int r = 0; r < reps; ++r) mma_bf16_ss(tbase + 48, arows, bcols + (r & 7) * 16, id, 1);
        long long t4 = clock64();
        mma_commit(&bar[0]);
        mbar_wait(&bar[0], 1);
        long long t5 = clock64();
        cyc[2] = t4 - t3; cyc[3] = t5 - t3;
        tc_fence_before();
    }
    __syncthreads();
    tc_fence_after();
    if (warp < 4) {
        float v[16];
        tmem_ld16(tbase + ((uint32_t)(warp * 32) << 16) + cD, v);
        for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * 16 + j] = v[j];
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 256);
}

int main() {
    float *d; long long *c;
    cudaMalloc(&d, 128 * 16 * 4); cudaMalloc(&c, 32);
    const int smem = 2 * kChunk + 64, reps = 256;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe<<<1, 160, smem>>>(d, c, reps);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    static float h[128 * 16]; long long hc[4];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost); cudaMemcpy(hc, c, 32, cudaMemcpyDeviceToHost);
    double worst = 0.;
    for (int m = 0; m < 128; ++m) for (int n = 0; n < 16; ++n) {
        double ref = 0.;
        for (int k = 0; k < KTOT; ++k) ref += (double)aval(m, k) * (double)bval(k, n);
        worst = fmax(worst, fabs(ref - (double)h[m * 16 + n]));
    }
    printf("TS mma (A in TMEM via tcgen05.st, B MN-major image): max abs err vs host = %g (%s)\n", worst, worst < 1e-3 ? "exact" : "MISMATCH");
    printf("TS M=128 N=16 K=16: issue %.1f cyc/mma, complete %.1f cyc/mma\n", (double)hc[0] / reps, (double)hc[1] / reps);
    printf("SS M=128 N=16 K=16: issue %.1f cyc/mma, complete %.1f cyc/mma\n", (double)hc[2] / reps, (double)hc[3] / reps);
    return worst < 1e-3 ? 0 : 2;
}
