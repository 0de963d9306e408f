// umma_probe_mn.cu -- check of MN-major ("transposed") tcgen05 operands over the SAME token-major canonical
// images the forward GEMMs use:   D[128 features x 64] = P[:, m0:m0+128]^T * Q   with K = 128 tokens,
// P image = [128 tokens x 192 features] (K-major for a token GEMM), Q image = [128 tokens x 64].
// variant 0: LBO = 128 (8-token blocks), SBO = rows*16 (8-feature blocks); variant 1: swapped.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe_mn scripts/umma_probe_mn.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "../x-transformers-rl_b200/csrc/umma.cuh"
using namespace xtrl::umma;

constexpr int TOK = 128, FP = 192, FQ = 64;

__global__ void __launch_bounds__(128) probe(const __nv_bfloat16 *p_img, const __nv_bfloat16 *q_img, float *d_out, int m0, int variant) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *sP = smem;
    unsigned char *sQ = smem + TOK * FP * 2;
    uint64_t *bars = (uint64_t *)(sQ + TOK * FQ * 2);
    uint32_t *tslot = (uint32_t *)(bars + 2);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(tslot, 64);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tslot;
    if (tid == 0) {
        mbar_expect_tx(&bars[0], TOK * FP * 2 + TOK * FQ * 2);
        bulk_g2s(sP, p_img, TOK * FP * 2, &bars[0]);
        bulk_g2s(sQ, q_img, TOK * FQ * 2, &bars[0]);
    }
    __syncthreads();
    if (tid == 0) {
        mbar_wait(&bars[0], 0);
        tc_fence_after();
        const uint32_t idesc = idesc_bf16(128, FQ) | (1u << 15) | (1u << 16);      // both operands MN-major
        const uint32_t tokblk = 128, featblk = TOK * 16;
        for (int s = 0; s < TOK / 16; ++s) {
            const unsigned char *pa = sP + (m0 >> 3) * featblk + s * 2 * tokblk;
            const unsigned char *pb = sQ + s * 2 * tokblk;
            const uint64_t ad = variant == 0 ? smem_desc(pa, tokblk, featblk) : smem_desc(pa, featblk, tokblk);
            const uint64_t bd = variant == 0 ? smem_desc(pb, tokblk, featblk) : smem_desc(pb, featblk, tokblk);
            mma_bf16_ss(tbase, ad, bd, idesc, s > 0);
        }
        mma_commit(&bars[1]);
    }
    mbar_wait(&bars[1], 0);
    tc_fence_after();
    for (int c = 0; c < FQ; c += 16) {
        float v[16];
        tmem_ld16(tbase + ((uint32_t)(warp * 32) << 16) + c, v);
        for (int n = 0; n < 16; ++n) d_out[(warp * 32 + lane) * FQ + c + n] = v[n];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 64);
}

int main() {
    std::vector<float> P(TOK * FP), Q(TOK * FQ);
    for (auto &x : P) x = (rand() % 2001 - 1000) / 1000.f;
    for (auto &x : Q) x = (rand() % 2001 - 1000) / 1000.f;
    std::vector<__nv_bfloat16> pimg(TOK * FP), qimg(TOK * FQ);
    for (int t = 0; t < TOK; ++t) {
        for (int f = 0; f < FP; ++f) { __nv_bfloat16 h = __float2bfloat16(P[t * FP + f]); P[t * FP + f] = __bfloat162float(h); pimg[img_off_bf16(TOK, t, f) / 2] = h; }
        for (int f = 0; f < FQ; ++f) { __nv_bfloat16 h = __float2bfloat16(Q[t * FQ + f]); Q[t * FQ + f] = __bfloat162float(h); qimg[img_off_bf16(TOK, t, f) / 2] = h; }
    }
    __nv_bfloat16 *dp, *dq; float *dd;
    cudaMalloc(&dp, pimg.size() * 2); cudaMalloc(&dq, qimg.size() * 2); cudaMalloc(&dd, 128 * FQ * 4);
    cudaMemcpy(dp, pimg.data(), pimg.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dq, qimg.data(), qimg.size() * 2, cudaMemcpyHostToDevice);
    const size_t smem = TOK * FP * 2 + TOK * FQ * 2 + 64;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int ok_variant = -1;
    for (int variant : {0, 1}) for (int m0 : {0, 64}) {
        cudaMemset(dd, 0, 128 * FQ * 4);
        probe<<<1, 128, smem>>>(dp, dq, dd, m0, variant);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("variant %d m0 %d CUDA error: %s\n", variant, m0, cudaGetErrorString(e)); return 2; }
        std::vector<float> D(128 * FQ);
        cudaMemcpy(D.data(), dd, D.size() * 4, cudaMemcpyDeviceToHost);
        double worst = 0;
        for (int m = 0; m < 128; ++m) for (int n = 0; n < FQ; ++n) {
            double ref = 0;
            for (int t = 0; t < TOK; ++t) ref += (double)P[t * FP + m0 + m] * Q[t * FQ + n];
            worst = fmax(worst, fabs(ref - D[m * FQ + n]));
        }
        printf("variant %d m0 %d: max abs err %.3e (D[0][0]=%f D[5][3]=%f)\n", variant, m0, worst, D[0], D[5 * FQ + 3]);
        if (worst < 1e-3 && m0 == 64) ok_variant = variant;
    }
    printf(ok_variant >= 0 ? "UMMA MN PROBE OK variant %d\n" : "UMMA MN PROBE FAILED %d\n", ok_variant);
    return ok_variant >= 0 ? 0 : 1;
}
