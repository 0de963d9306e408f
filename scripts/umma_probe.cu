// umma_probe.cu -- standalone check of the tcgen05 helpers in csrc/umma.cuh on one CTA:
//   D[128 x 16] = A[128 x K] * B[16 x K]^T  (bf16 operands, fp32 accumulate in TMEM), K in {48,128,192},
// A arrives by a 1-D TMA bulk copy of a host-packed canonical image, B is written by the threads.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe scripts/umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "../x-transformers-rl_b200/csrc/umma.cuh"
using namespace xtrl::umma;

__global__ void __launch_bounds__(128) probe(const __nv_bfloat16 *a_img, const __nv_bfloat16 *b_rowmajor, float *d_out, int K) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *sA = smem;                                  // 128 rows x K bf16, canonical image
    unsigned char *sB = smem + 128 * K * 2;                    // 16 rows x K
    uint64_t *bars = (uint64_t *)(sB + 16 * K * 2);            // [0] weights landed, [1] mma done
    uint32_t *tslot = (uint32_t *)(bars + 2);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(tslot, 32);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tslot;
    if (tid == 0) { mbar_expect_tx(&bars[0], 128 * K * 2); bulk_g2s(sA, a_img, 128 * K * 2, &bars[0]); }
    for (int i = tid; i < 16 * K; i += 128) {
        const int e = i / K, k = i % K;
        *(__nv_bfloat16 *)(sB + img_off_bf16(16, e, k)) = b_rowmajor[i];
    }
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
        mbar_wait(&bars[0], 0);
        tc_fence_after();
        const uint32_t idesc = idesc_bf16(128, 16);
        for (int s = 0; s < K / 16; ++s) {
            const uint64_t ad = smem_desc(sA + s * 2 * (128 * 16), 128 * 16, 128);
            const uint64_t bd = smem_desc(sB + s * 2 * (16 * 16), 16 * 16, 128);
            mma_bf16_ss(tbase, ad, bd, idesc, s > 0);
        }
        mma_commit(&bars[1]);
    }
    mbar_wait(&bars[1], 0);
    tc_fence_after();
    float v[16];
    tmem_ld16(tbase + ((uint32_t)(warp * 32) << 16), v);
    for (int n = 0; n < 16; ++n) d_out[(warp * 32 + lane) * 16 + n] = v[n];
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 32);
}

int main() {
    int rc = 0;
    for (int K : {48, 128, 192}) {
        std::vector<float> A(128 * K), B(16 * K);
        for (auto &x : A) x = (rand() % 2001 - 1000) / 1000.f;
        for (auto &x : B) x = (rand() % 2001 - 1000) / 1000.f;
        std::vector<__nv_bfloat16> aimg(128 * K), brow(16 * K);
        std::vector<float> Ar(128 * K), Br(16 * K);
        for (int r = 0; r < 128; ++r) for (int k = 0; k < K; ++k) {
            __nv_bfloat16 h = __float2bfloat16(A[r * K + k]);
            Ar[r * K + k] = __bfloat162float(h);
            aimg[img_off_bf16(128, r, k) / 2] = h;
        }
        for (int i = 0; i < 16 * K; ++i) { brow[i] = __float2bfloat16(B[i]); Br[i] = __bfloat162float(brow[i]); }
        __nv_bfloat16 *da, *db; float *dd;
        cudaMalloc(&da, aimg.size() * 2); cudaMalloc(&db, brow.size() * 2); cudaMalloc(&dd, 128 * 16 * 4);
        cudaMemcpy(da, aimg.data(), aimg.size() * 2, cudaMemcpyHostToDevice);
        cudaMemcpy(db, brow.data(), brow.size() * 2, cudaMemcpyHostToDevice);
        cudaMemset(dd, 0, 128 * 16 * 4);
        const size_t smem = 128 * K * 2 + 16 * K * 2 + 64;
        cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        probe<<<1, 128, smem>>>(da, db, dd, K);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("K=%d CUDA error: %s\n", K, cudaGetErrorString(e)); return 2; }
        std::vector<float> D(128 * 16);
        cudaMemcpy(D.data(), dd, D.size() * 4, cudaMemcpyDeviceToHost);
        double worst = 0;
        for (int m = 0; m < 128; ++m) for (int n = 0; n < 16; ++n) {
            double ref = 0;
            for (int k = 0; k < K; ++k) ref += (double)Ar[m * K + k] * Br[n * K + k];
            worst = fmax(worst, fabs(ref - D[m * 16 + n]));
        }
        printf("K=%d max abs err %.3e  (D[0][0]=%f D[5][3]=%f D[127][15]=%f)\n", K, worst, D[0], D[5 * 16 + 3], D[127 * 16 + 15]);
        if (worst > 1e-3) rc = 1;
        cudaFree(da); cudaFree(db); cudaFree(dd);
    }
    printf(rc ? "UMMA PROBE FAILED\n" : "UMMA PROBE OK\n");
    return rc;
}
