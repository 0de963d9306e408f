// prmt_probe.cu -- what PTX prmt.b32 (default mode) does with selector bit 3 on sm_100a
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(unsigned *out, unsigned a) {
    unsigned sels[6] = {0x9988, 0xbbaa, 0x8888, 0x9999, 0x3210, 0x1100};
    for (int i = 0; i < 6; ++i) { unsigned d; asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(0u), "r"(sels[i])); out[i] = d; }
}
int main() {
    unsigned *d, h[6]; cudaMalloc(&d, 24);
    for (unsigned a : {0x80ff7f01u, 0x017f8000u}) {
        k<<<1, 1>>>(d, a); cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
        printf("a=%08x: 9988->%08x bbaa->%08x 8888->%08x 9999->%08x 3210->%08x 1100->%08x\n", a, h[0], h[1], h[2], h[3], h[4], h[5]);
    }
    return 0;
}
