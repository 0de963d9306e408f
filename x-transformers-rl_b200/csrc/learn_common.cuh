// learn_common.cuh -- pieces shared by the learn-side fused sub-layer kernels (learn_ff.cu, learn_attn.cu):
// 128-token tiles, thread (row, half) ownership, bf16 operand images, LayerNorm over D = 48 in registers.
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace xtrl {
namespace learn {

constexpr int D = 48, TOK = 128, NT = 256;
constexpr int kCore = TOK * 16;        // bytes between 8-feature chunks of a 128-token image
constexpr uint32_t kIdA = 1u << 15, kIdB = 1u << 16;                 // MN-major operand flags of the idesc

__device__ __forceinline__ uint32_t pack2(float a, float b) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t *>(&h);
}
__device__ __forceinline__ void unpack8(const uint4 u, float (&o)[8]) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { o[2 * i] = __uint_as_float(w[i] << 16); o[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
}
// row statistics of LayerNorm over 48 fp32 values held in registers
__device__ __forceinline__ void ln_stats(const float (&xr)[D], float &mean, float &rstd) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < D; ++i) s += xr[i];
    mean = s * (1.f / D);
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < D; ++i) { const float d = xr[i] - mean; v = fmaf(d, d, v); }
    rstd = rsqrtf(v * (1.f / D) + 1e-5f);
}
__device__ __forceinline__ void load_row(const float *p, bool ok, float (&xr)[D]) {
#pragma unroll
    for (int i = 0; i < D / 4; ++i) {
        const float4 v = ok ? __ldg(reinterpret_cast<const float4 *>(p) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        xr[4 * i] = v.x; xr[4 * i + 1] = v.y; xr[4 * i + 2] = v.z; xr[4 * i + 3] = v.w;
    }
}
// 8 consecutive features of this thread's row -> one 16-byte piece of a 128-token image
__device__ __forceinline__ void put8(unsigned char *img, int row, int chunk, const float (&v)[8]) {
    *reinterpret_cast<uint4 *>(img + chunk * kCore + (row >> 3) * 128 + (row & 7) * 16) =
        make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
}


// packed fp32 pairs (sm_100: FFMA2 / FMUL2 / FADD2 issue two fp32 operations per instruction)
__device__ __forceinline__ uint64_t pk2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// pull the 128-byte line of `p` into L2 (no register, no wait).  Every persistent CTA of the sub-layer kernels prefetches the
// rows of its NEXT tile while it works on the current one: all SMs load their tiles at the same moment, so without this each
// tile starts with a burst of DRAM round trips that nothing overlaps (ncu: 34 % of attn_pre_bwd's stall samples).
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// rows [t0, t0 + rows) of a row-major fp32 [N][48] tensor: 192 bytes per row, one prefetch per 128-byte line
__device__ __forceinline__ void prefetch_rows48(const float *base, int t0, int rows, int tid, int nthreads) {
    const char *p = reinterpret_cast<const char *>(base + (size_t)t0 * D);
    for (int o = tid * 128; o < rows * D * 4; o += nthreads * 128) prefetch_l2(p + o);
}

// 16 bytes global -> shared without a register round trip (LDGSTS); ok = false writes zeros.  A thread queues all the
// pieces of its tile row back to back and waits once: one memory round trip per tile instead of one per piece.
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc, bool ok) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    const int n = ok ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Row tiles by TMA.  The 128 rows of a tile of a row-major fp32 [N][48] tensor are 24 KB contiguous in global memory, so a
// tile moves by ONE bulk copy issued by one thread instead of 12 row-strided 16-byte loads per thread (every warp
// instruction of those touches 32 lines: ~3000 L1 requests per tile and operand).  The fetch of tile i+1 is issued as soon
// as every thread has read tile i out of the staging buffer, so it lands under the MMAs and the epilogue of tile i.
constexpr int kRowTile = TOK * D * 4;
__device__ __forceinline__ void rows_fetch(float *sX, const float *x, int tile, int N, uint64_t *bar) {        // one thread
    const int t0 = tile * TOK, rows = min(TOK, N - t0);
    const uint32_t bytes = (uint32_t)rows * D * 4;
    umma::mbar_expect_tx(bar, bytes);
    umma::bulk_g2s(sX, x + (size_t)t0 * D, bytes, bar);
}
__device__ __forceinline__ void rows_read(const float *sX, int row, bool ok, float (&xr)[D]) {                 // this thread's row
    const float4 *p = reinterpret_cast<const float4 *>(sX + row * D);
#pragma unroll
    for (int i = 0; i < D / 4; ++i) {
        const float4 v = ok ? p[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        xr[4 * i] = v.x; xr[4 * i + 1] = v.y; xr[4 * i + 2] = v.z; xr[4 * i + 3] = v.w;
    }
}
__device__ __forceinline__ void rows_store(float *dst, const float *sOut, int tile, int N) {                   // one thread
    const int t0 = tile * TOK, rows = min(TOK, N - t0);
    umma::bulk_s2g(dst + (size_t)t0 * D, sOut, (uint32_t)rows * D * 4);
}

// Weight-gradient flush.  The accumulators live in TMEM with lane = weight row, so a thread that adds its own values to a
// row-major [rows][ncols] gradient makes every warp instruction touch 32 different lines -- 32 L2 atomic transactions per
// 32 values, and with 148 CTAs flushing ~19k values each the flush cost as much as the tile loop (ncu: 44 % of
// attn_pre_bwd's samples).  Instead the rows are staged in shared memory (odd pitch: conflict-free transposed writes) and
// added by flat index: consecutive lanes, consecutive addresses, one transaction per 128 bytes.
// addr(r, c) -> destination of element (r, c) or nullptr.
template <class F>
__device__ __forceinline__ void flush_rows(const float *stage, int pitch, int nrows, int ncols, int tid, int nthreads, F addr) {
    for (int f = tid; f < nrows * ncols; f += nthreads) {
        const int r = f / ncols, c = f - r * ncols;
        float *p = addr(r, c);
        const float v = stage[r * pitch + c];
        if (p && v != 0.f) atomicAdd(p, v);
    }
}

__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

}  // namespace learn
}  // namespace xtrl
