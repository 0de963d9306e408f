// learn_ff.cu -- learn-side feed-forward block of the sequence model, forward and backward, on tcgen05.
//
//     y = x + W2 * drop(gelu(W1 * (LN(x) * gamma) + b1)) + b2           (tokens x 48 -> 192 -> 48)
//
// Replaces, inside Agent.learn's teacher-forced pass (x_transformers_rl.py:928-935; x-transformers
// Decoder feed-forward sub-layer): LayerNorm -> Linear(D,4D) -> GELU -> Dropout ->
// Linear(4D,D) -> +residual, and their backward.  SURVEY 8f row 1.
//
// A CTA walks 128-token tiles; TMEM lane = token row.  256 threads: thread (row, half) owns half of the
// row's columns in every epilogue.  Operands are bf16 canonical no-swizzle images in shared memory, fp32
// accumulators in TMEM:
//   fwd:  PRE[128x192] = Xn * W1^T          H = drop(gelu(PRE + b1))       OUT[128x48] = H * W2^T
//   bwd:  recompute PRE / H (nothing but x is saved by the forward), then
//         dH[128x192]  = dY * W2            dPRE = dH . gelu'(PRE) . mask
//         dXn[128x48]  = dPRE * W1          dx = dy + LN'(dXn * gamma)
//         dW1 | db1   += dPRE^T * [Xn | 1]  dW2^T | db2 += [H | 1]^T * dY  -- accumulated in TMEM across all
//         tiles of the CTA; the token-major images the token GEMMs already use are re-read as MN-major
//         ("transposed") operands, so no transposed copy is ever written; bias gradients fall out of a
//         constant-one column.  One atomicAdd pass per CTA at the end.
// HBM traffic per token: fwd 192 B read + 192 B written; bwd 384 B read + 192 B written (fp32 residual
// stream); weights stay in shared memory for the life of the CTA.
#include "common.cuh"
#include "learn_common.cuh"

namespace xtrl {
namespace ff {
using namespace umma;
using namespace learn;

constexpr int F = 192;
constexpr int XW = 64;                 // Xn / dY image width: 48 features + a ones column (48) + zero pad
constexpr int HW = 200;                // H image width: 192 features + a ones column (192) + zero pad
constexpr int kW1 = F * D * 2;         // W1 image: rows = 192 outputs, K = 48
constexpr int kW2 = D * F * 2;         // W2 image: rows = 48 outputs, K = 192

// exact-form GELU x * Phi(x) and its derivative Phi(x) + x * phi(x), with erf from Abramowitz-Stegun 7.1.26
// (|error| < 1.5e-7, far inside the bf16 class): the exp(-x^2 / 2) it needs is the phi(x) of the derivative, so the pair
// costs one ex2, one rcp and ~10 FMAs instead of erff + expf.
__device__ __forceinline__ void gelu_both(float x, float &g, float &dg) {
    const float z = fabsf(x) * 0.70710678118654752f;
    const float t = __frcp_rn(fmaf(0.3275911f, z, 1.f));
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-0.72134752044448170f * x * x));      // exp(-x^2 / 2)
    float poly = fmaf(1.061405429f, t, -1.453152027f);
    poly = fmaf(poly, t, 1.421413741f);
    poly = fmaf(poly, t, -0.284496736f);
    poly = fmaf(poly, t, 0.254829592f);
    const float half_erfc = 0.5f * poly * t * e;                                           // 0.5 * erfc(|x| / sqrt 2)
    const float cdf = x >= 0.f ? 1.f - half_erfc : half_erfc;
    g = x * cdf;
    dg = fmaf(x * 0.3989422804014327f, e, cdf);
}
__device__ __forceinline__ float gelu(float x) { float g, d; gelu_both(x, g, d); return g; }
// The same pair of functions on packed fp32 pairs: half the floating-point instructions (FFMA2 / FMUL2 / FADD2).
// g = x * Phi(x), dg = Phi(x) + x * phi(x);  Phi = 0.5 + copysign(0.5 - half_erfc(|x|), x).
__device__ __forceinline__ void gelu_both2(uint64_t x2, uint64_t &g2, uint64_t &d2) {
    float x0, x1;
    upk2(x2, x0, x1);
    const uint64_t z = mul2(pk2(fabsf(x0), fabsf(x1)), pk2(0.70710678118654752f, 0.70710678118654752f));
    float u0, u1;
    upk2(fma2(pk2(0.3275911f, 0.3275911f), z, pk2(1.f, 1.f)), u0, u1);
    const uint64_t t = pk2(__frcp_rn(u0), __frcp_rn(u1));
    float a0, a1;
    upk2(mul2(mul2(x2, x2), pk2(-0.72134752044448170f, -0.72134752044448170f)), a0, a1);
    float e0, e1;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(a0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(a1));
    const uint64_t e = pk2(e0, e1);                                                        // exp(-x^2 / 2)
    uint64_t poly = fma2(pk2(0.5f * 1.061405429f, 0.5f * 1.061405429f), t, pk2(0.5f * -1.453152027f, 0.5f * -1.453152027f));
    poly = fma2(poly, t, pk2(0.5f * 1.421413741f, 0.5f * 1.421413741f));
    poly = fma2(poly, t, pk2(0.5f * -0.284496736f, 0.5f * -0.284496736f));
    poly = fma2(poly, t, pk2(0.5f * 0.254829592f, 0.5f * 0.254829592f));
    float h0, h1;
    upk2(fma2(mul2(poly, t), pk2(-e0, -e1), pk2(0.5f, 0.5f)), h0, h1);                     // 0.5 - 0.5 erfc(|x| / sqrt 2) >= 0
    const uint64_t cdf = add2(pk2(copysignf(h0, x0), copysignf(h1, x1)), pk2(0.5f, 0.5f));
    g2 = mul2(x2, cdf);
    d2 = fma2(mul2(x2, pk2(0.3989422804014327f, 0.3989422804014327f)), e, cdf);
}
// Dropout decisions of the four hidden units 4 g .. 4 g + 3 of token t: one hash -> four 7-bit uniforms; the sign bit of byte j
// of the result is set <=> keep (p is quantised to 1/128, as in the attention core).  add4 = (128 - threshold) in every byte.
__device__ __forceinline__ uint32_t fold_key(const int64_t *rng_step, uint32_t key) {
    const uint64_t s = (uint64_t)(rng_step ? *rng_step : 0) * 0x9E3779B97F4A7C15ull;
    return (uint32_t)(s >> 32) ^ (uint32_t)s ^ (key * 0x632BE5ABu);
}
__device__ __forceinline__ uint32_t keep4(uint32_t t, int g, uint32_t key, uint32_t add4) {
    uint32_t h = (t * (uint32_t)(F / 4) + (uint32_t)g) * 0x9E3779B1u ^ key;
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 15;
    return (h & 0x7f7f7f7fu) + add4;
}
__device__ __forceinline__ uint32_t prmt_sign(uint32_t a, uint32_t sel) {                 // selector nibble 8 | j: replicate the sign of byte j
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(0u), "r"(sel));
    return d;
}
// 16 pre-activations of token t (columns col0 .. col0 + 15, accumulator values + bias) -> GELU values g and, if WITH_D, the
// derivative d, both already multiplied by the dropout keep factor (0 or drop_scale)
template <bool WITH_D>
__device__ __forceinline__ void act16(const float (&pre)[16], const float *bias, uint32_t t, int col0, uint32_t key, uint32_t add4, float scale,
                                      float (&g)[16], float (&d)[16]) {
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) {
        const float4 b = *reinterpret_cast<const float4 *>(bias + 4 * q4);
        const uint32_t kw = add4 ? keep4(t, (col0 >> 2) + q4, key, add4) : 0u;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int j = 4 * q4 + 2 * h;
            uint64_t g2, d2;
            gelu_both2(add2(pk2(pre[j], pre[j + 1]), h ? pk2(b.z, b.w) : pk2(b.x, b.y)), g2, d2);
            if (add4) {                                                                   // keep factor as a float: mask & scale bits
                const uint64_t k2 = pk2(__uint_as_float(prmt_sign(kw, 0x8888u + 0x1111u * (2 * h)) & __float_as_uint(scale)),
                                        __uint_as_float(prmt_sign(kw, 0x8888u + 0x1111u * (2 * h + 1)) & __float_as_uint(scale)));
                g2 = mul2(g2, k2);
                if (WITH_D) d2 = mul2(d2, k2);
            }
            upk2(g2, g[j], g[j + 1]);
            if (WITH_D) upk2(d2, d[j], d[j + 1]);
        }
    }

}

struct Args {
    const float *x, *dy, *gamma, *b1, *b2;
    const unsigned char *wimg;
    float *y, *dx, *dgamma, *dW1, *db1, *dW2, *db2;
    const int64_t *rng_step;
    uint32_t rng_key, drop_add4;              // (128 - threshold) replicated in 4 bytes; 0 = no dropout
    float drop_scale;
    int N;
};

// ------------------------------------------------------------------------------------------ forward
constexpr int kFwdSmem = kW1 + kW2 + TOK * D * 2 + TOK * F * 2 + (F + 2 * D) * 4 + 64;

__global__ void __launch_bounds__(NT, 2) ff_fwd_kernel(const Args a) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *sW1 = smem, *sW2 = sW1 + kW1, *sXn = sW2 + kW2, *sH = sXn + TOK * D * 2;
    float *sB1 = reinterpret_cast<float *>(sH + TOK * F * 2), *sB2 = sB1 + F, *sG = sB2 + D;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sG + D);
    uint32_t *tslot = reinterpret_cast<uint32_t *>(bars + 3);
    // fp32 row tiles move by bulk copy (rows_fetch / rows_store) through the two halves of the H image, which is idle from
    // the end of one tile's second GEMM to the first epilogue of the next: x of the next tile lands in the lower half, y of
    // this tile leaves from the upper half
    float *sXs = reinterpret_cast<float *>(sH), *sYs = sXs + TOK * D;
    static_assert(2 * kRowTile <= TOK * F * 2, "row-tile staging fits the H image");
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = warp & 3, half = warp >> 2, row = q * 32 + lane;

    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(tslot, 256);
    for (int i = tid; i < F; i += NT) sB1[i] = a.b1[i];
    if (tid < D) { sB2[tid] = a.b2[tid]; sG[tid] = a.gamma[tid]; }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tslot, tlane = tbase + ((uint32_t)(q * 32) << 16);
    const int ntiles = (a.N + TOK - 1) / TOK;
    if (tid == 0) {
        mbar_expect_tx(&bars[0], kW1 + kW2); bulk_g2s(sW1, a.wimg, kW1 + kW2, &bars[0]);
        if ((int)blockIdx.x < ntiles) rows_fetch(sXs, a.x, blockIdx.x, a.N, &bars[2]);
    }
    const uint32_t key = fold_key(a.rng_step, a.rng_key);
    uint32_t ph = 0, phx = 0;
    int pend = -1;                                                               // tile whose y still sits in sYs
    bool first = true;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int t = tile * TOK + row;
        const bool ok = t < a.N;
        float xr[D];
        mbar_wait(&bars[2], phx); phx ^= 1;
        rows_read(sXs, row, ok, xr);
        float mean, rstd, own[24];
        ln_stats(xr, mean, rstd);
#pragma unroll
        for (int i = 0; i < 24; ++i) own[i] = half ? xr[24 + i] : xr[i];           // this thread's half of the row
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = (own[c * 8 + j] - mean) * rstd * sG[half * 24 + c * 8 + j];
            put8(sXn, row, half * 3 + c, v);
        }
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            if (pend >= 0) rows_store(a.y, sYs, pend, a.N);
            if (first) mbar_wait(&bars[0], 0);
            tc_fence_after();
#pragma unroll
            for (int s = 0; s < D / 16; ++s)
                mma_bf16_ss(tbase, smem_desc(sXn + s * 2 * kCore, kCore, 128), smem_desc(sW1 + s * 2 * F * 16, F * 16, 128),
                            idesc_bf16(128, F), s > 0);
            if (pend >= 0) bulk_wait_read();                                     // the epilogue below overwrites sYs: commit only once it is read
            mma_commit(&bars[1]);
        }
        first = false;
        mbar_wait(&bars[1], ph); ph ^= 1;
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < 6; ++c) {
            const int col0 = half * 96 + c * 16;
            float v[16];
            tmem_ld16(tlane + col0, v);
            {
                float g[16], dummy[16];
                act16<false>(v, sB1 + col0, (uint32_t)t, col0, key, a.drop_add4, a.drop_scale, g, dummy);
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = g[j];
            }
            float lo[8], hi[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) { lo[j] = v[j]; hi[j] = v[8 + j]; }
            put8(sH, row, col0 / 8, lo);
            put8(sH, row, col0 / 8 + 1, hi);
        }
        tc_fence_before();
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
#pragma unroll
            for (int s = 0; s < F / 16; ++s)
                mma_bf16_ss(tbase + F, smem_desc(sH + s * 2 * kCore, kCore, 128), smem_desc(sW2 + s * 2 * D * 16, D * 16, 128),
                            idesc_bf16(128, D), s > 0);
            mma_commit(&bars[1]);
        }
        mbar_wait(&bars[1], ph); ph ^= 1;
        tc_fence_after();
        if (tid == 0 && tile + (int)gridDim.x < ntiles) rows_fetch(sXs, a.x, tile + gridDim.x, a.N, &bars[2]);   // the H image is idle now
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float v[8];
            tmem_ld8(tlane + F + half * 24 + c * 8, v);
            {
                float4 *dst = reinterpret_cast<float4 *>(sYs + row * D + half * 24 + c * 8);
                const int f = half * 24 + c * 8;
                const float *o = own + c * 8;
                dst[0] = make_float4(v[0] + sB2[f] + o[0], v[1] + sB2[f + 1] + o[1], v[2] + sB2[f + 2] + o[2], v[3] + sB2[f + 3] + o[3]);
                dst[1] = make_float4(v[4] + sB2[f + 4] + o[4], v[5] + sB2[f + 5] + o[5], v[6] + sB2[f + 6] + o[6], v[7] + sB2[f + 7] + o[7]);
            }
        }
        pend = tile;
        tc_fence_before();
    }
    if (pend >= 0) {                                                             // the last tile's y
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) { rows_store(a.y, sYs, pend, a.N); bulk_wait_all(); }
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 256);
}

// ----------------------------------------------------------------------------------------- backward
// shared-memory order matters: the dW MMAs read 128 "feature rows" starting at feature 128 of the H and
// dPRE images (64 real + 64 beyond the image); what follows each image is more shared memory of this
// CTA, and the accumulator rows those bytes feed (>= 64, resp. > 64) are never read.
constexpr int kBwdH = TOK * HW * 2, kBwdDP = TOK * F * 2, kBwdX = TOK * XW * 2;
constexpr int kBwdSmem = kBwdH + kBwdDP + 2 * kBwdX + kW1 + kW2 + 2 * kRowTile + (F + D) * 4 + 64;

// 512 threads: thread (row, cq) owns a QUARTER of the row's columns in every epilogue (48 of the 192 hidden units; 16 of the
// 48 features for cq < 3).  The epilogues -- GELU, its derivative and the dropout hash on 128 x 192 values per tile -- are the
// bulk of this kernel's time and are latency-bound, so sixteen warps per SM (one CTA per SM: shared memory) nearly halve it
// against eight.  Rows are re-read from L2 where a second use is far from the first instead of living in registers.
constexpr int NT4 = 512;

__global__ void __launch_bounds__(NT4, 1) ff_bwd_kernel(const Args a) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *sH = smem, *sDP = sH + kBwdH, *sXn = sDP + kBwdDP, *sDY = sXn + kBwdX;
    unsigned char *sW1 = sDY + kBwdX, *sW2 = sW1 + kW1;
    float *sX32 = reinterpret_cast<float *>(sW2 + kW2);                          // two fp32 row tiles of x (TMA staging, see rows_fetch):
    float *sB1 = sX32 + 2 * TOK * D, *sG = sB1 + F;                              // a tile's rows are read at its start and again at its end
    uint64_t *bars = reinterpret_cast<uint64_t *>(sG + D);
    uint32_t *tslot = reinterpret_cast<uint32_t *>(bars + 4);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = warp & 3, cq = warp >> 2, row = q * 32 + lane;
    constexpr uint32_t cPre = 0, cDxn = 192, cA = 256, cB = 320, cC = 384, cD = 448;

    // bars: 0 weights, 1 the GEMMs the next epilogue waits for, 2 x tile landed, 3 the weight-gradient GEMMs of a tile (the
    // last readers of its images: they run under the LayerNorm backward and the next tile's row statistics)
    if (tid == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(tslot, 512);
    for (int i = tid; i < F; i += NT4) sB1[i] = a.b1[i];
    if (tid < D) sG[tid] = a.gamma[tid];
    if (cq == 0) {                         // constant columns: [1, 0 x 7] then zeros
        const float one[8] = {1.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, zero[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        put8(sXn, row, 6, one); put8(sXn, row, 7, zero);
        put8(sDY, row, 6, one); put8(sDY, row, 7, zero);
        put8(sH, row, 24, one);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tslot, tlane = tbase + ((uint32_t)(q * 32) << 16);
    const int ntiles = (a.N + TOK - 1) / TOK;
    if (tid == 0) {
        mbar_expect_tx(&bars[0], kW1 + kW2); bulk_g2s(sW1, a.wimg, kW1 + kW2, &bars[0]);
        if ((int)blockIdx.x < ntiles) rows_fetch(sX32, a.x, blockIdx.x, a.N, &bars[2]);
    }
    const uint32_t key = fold_key(a.rng_step, a.rng_key);
    uint32_t ph = 0, phx = 0;
    int it = 0;
    bool first = true;
    float dg_acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) dg_acc[i] = 0.f;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int t = tile * TOK + row;
        const bool ok = t < a.N;
        const float *sx = sX32 + (it & 1) * TOK * D;
        if (tile + (int)gridDim.x < ntiles) {                                    // next tile's dy rows -> L2
            const int t0 = (tile + gridDim.x) * TOK, nr = min(TOK, a.N - t0);
            prefetch_rows48(a.dy, t0, nr, tid, NT4);
        }
        float mean, rstd, dyr[16];
        {
            float xr[D];
            mbar_wait(&bars[2], phx); phx ^= 1;
            rows_read(sx, row, ok, xr);
            ln_stats(xr, mean, rstd);
            if (cq < 3) {                                                        // my 16 features of the Xn and dY images
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 v = ok ? __ldg(reinterpret_cast<const float4 *>(a.dy + (size_t)t * D + cq * 16) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
                    dyr[4 * i] = v.x; dyr[4 * i + 1] = v.y; dyr[4 * i + 2] = v.z; dyr[4 * i + 3] = v.w;
                }
            }
            if (it > 0) { mbar_wait(&bars[3], (it - 1) & 1); tc_fence_after(); }   // the previous tile's images are free
            if (cq < 3) {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    float v[8], w[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int f = cq * 16 + c * 8 + j;
                        float xv = xr[0];
#pragma unroll
                        for (int m = 0; m < 3; ++m) if (cq == m) xv = xr[m * 16 + c * 8 + j];      // static register indices
                        v[j] = (xv - mean) * rstd * sG[f];
                        w[j] = dyr[c * 8 + j];
                    }
                    put8(sXn, row, cq * 2 + c, v);
                    put8(sDY, row, cq * 2 + c, w);
                }
            }
        }
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {                                                          // PRE = Xn * W1^T
            if (first) mbar_wait(&bars[0], 0);
            tc_fence_after();
#pragma unroll
            for (int s = 0; s < D / 16; ++s)
                mma_bf16_ss(tbase + cPre, smem_desc(sXn + s * 2 * kCore, kCore, 128), smem_desc(sW1 + s * 2 * F * 16, F * 16, 128),
                            idesc_bf16(128, F), s > 0);
            mma_commit(&bars[1]);
            // the other x buffer was last read at the end of the previous tile, which every thread has left
            if (tile + (int)gridDim.x < ntiles) rows_fetch(sX32 + ((it + 1) & 1) * TOK * D, a.x, tile + gridDim.x, a.N, &bars[2]);
        }
        mbar_wait(&bars[1], ph); ph ^= 1;
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < 3; ++c) {                                            // H and gelu'.mask, my 48 hidden units
            const int col0 = cq * 48 + c * 16;
            float v[16], d[16];
            {
                float pre[16];
                tmem_ld16(tlane + cPre + col0, pre);
                act16<true>(pre, sB1 + col0, (uint32_t)t, col0, key, a.drop_add4, a.drop_scale, v, d);
            }
            float lo[8], hi[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) { lo[j] = v[j]; hi[j] = v[8 + j]; }
            put8(sH, row, col0 / 8, lo);
            put8(sH, row, col0 / 8 + 1, hi);
#pragma unroll
            for (int j = 0; j < 8; ++j) { lo[j] = d[j]; hi[j] = d[8 + j]; }
            put8(sDP, row, col0 / 8, lo);                                        // parked in dPRE's slot
            put8(sDP, row, col0 / 8 + 1, hi);
        }
        tc_fence_before();
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
#pragma unroll
            for (int s = 0; s < D / 16; ++s)                                     // dH = dY * W2 (W2 image read MN-major)
                mma_bf16_ss(tbase + cPre, smem_desc(sDY + s * 2 * kCore, kCore, 128), smem_desc(sW2 + s * 256, 128, D * 16),
                            idesc_bf16(128, F) | kIdB, s > 0);
            mma_commit(&bars[1]);
        }
        mbar_wait(&bars[1], ph); ph ^= 1;
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < 3; ++c) {                                            // dPRE = dH . gelu'.mask
            const int col0 = cq * 48 + c * 16;
            float v[16], lo[8], hi[8];
            tmem_ld16(tlane + cPre + col0, v);
            unsigned char *p = sDP + (col0 / 8) * kCore + (row >> 3) * 128 + (row & 7) * 16;
            unpack8(*reinterpret_cast<const uint4 *>(p), lo);
            unpack8(*reinterpret_cast<const uint4 *>(p + kCore), hi);
#pragma unroll
            for (int j = 0; j < 8; ++j) { lo[j] *= v[j]; hi[j] *= v[8 + j]; }
            put8(sDP, row, col0 / 8, lo);
            put8(sDP, row, col0 / 8 + 1, hi);
        }
        tc_fence_before();
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
#pragma unroll
            for (int s = 0; s < F / 16; ++s)                                     // dXn = dPRE * W1 (W1 image read MN-major)
                mma_bf16_ss(tbase + cDxn, smem_desc(sDP + s * 2 * kCore, kCore, 128), smem_desc(sW1 + s * 256, 128, F * 16),
                            idesc_bf16(128, D) | kIdB, s > 0);
            mma_commit(&bars[1]);                                                // the LayerNorm backward can start
#pragma unroll
            for (int s = 0; s < TOK / 16; ++s) {                                 // dW2^T | db2 += [H|1]^T * dY
                const uint64_t bd = smem_desc(sDY + s * 256, 128, kCore);
                mma_bf16_ss(tbase + cC, smem_desc(sH + s * 256, 128, kCore), bd, idesc_bf16(128, XW) | kIdA | kIdB, !first || s > 0);
                mma_bf16_ss(tbase + cD, smem_desc(sH + 16 * kCore + s * 256, 128, kCore), bd, idesc_bf16(128, XW) | kIdA | kIdB, !first || s > 0);
            }
#pragma unroll
            for (int s = 0; s < TOK / 16; ++s) {                                 // dW1 | db1 += dPRE^T * [Xn|1]
                const uint64_t bd = smem_desc(sXn + s * 256, 128, kCore);
                mma_bf16_ss(tbase + cA, smem_desc(sDP + s * 256, 128, kCore), bd, idesc_bf16(128, XW) | kIdA | kIdB, !first || s > 0);
                mma_bf16_ss(tbase + cB, smem_desc(sDP + 16 * kCore + s * 256, 128, kCore), bd, idesc_bf16(128, XW) | kIdA | kIdB, !first || s > 0);
            }
            mma_commit(&bars[3]);
        }
        first = false;
        mbar_wait(&bars[1], ph); ph ^= 1;
        tc_fence_after();
        if (cq < 3) {                                                            // LayerNorm backward + residual, my 16 features
            float c1 = 0.f, c2 = 0.f, dxo[16], xho[16];
#pragma unroll
            for (int c = 0; c < 6; ++c) {                                        // the row sums need all 48 features: stream them
                float v[8];
                tmem_ld8(tlane + cDxn + c * 8, v);
                const float4 x0 = ok ? *(reinterpret_cast<const float4 *>(sx + row * D + c * 8)) : make_float4(0.f, 0.f, 0.f, 0.f);
                const float4 x1 = ok ? *(reinterpret_cast<const float4 *>(sx + row * D + c * 8) + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
                const float xs[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float xh = (xs[j] - mean) * rstd, dh = v[j] * sG[c * 8 + j];
                    c1 += dh; c2 = fmaf(dh, xh, c2);
                    if ((c >> 1) == cq) { dxo[(c & 1) * 8 + j] = v[j]; xho[(c & 1) * 8 + j] = xh; }
                }
            }
            c1 *= (1.f / D); c2 *= (1.f / D);
            float o[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                dg_acc[i] = fmaf(dxo[i], xho[i], dg_acc[i]);
                o[i] = dyr[i] + rstd * (dxo[i] * sG[cq * 16 + i] - c1 - xho[i] * c2);
            }
            if (ok) {
                float4 *dst = reinterpret_cast<float4 *>(a.dx + (size_t)t * D + cq * 16);
#pragma unroll
                for (int i = 0; i < 4; ++i) dst[i] = make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
            }
        }
        tc_fence_before();
    }
    // ---- flush the per-CTA weight-gradient accumulators ----
    if (cq < 3) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float s = warp_sum(dg_acc[i]);
            if (lane == 0 && s != 0.f) atomicAdd(a.dgamma + cq * 16 + i, s);
        }
    }
    if (!first) {                                                                // this CTA saw at least one tile
        mbar_wait(&bars[3], (it - 1) & 1);                                       // the last weight-gradient GEMMs
        tc_fence_after();
        float *stage = reinterpret_cast<float *>(sH);                            // dW1 rows [192][D + 1] fp32 (see flush_rows); the H image is done with
        static_assert(F * (D + 1) * 4 <= kBwdH + kBwdDP, "stage fits the tile images");
        __syncthreads();
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {                                            // quarter cq: columns 16 cq .. 16 cq + 15 (48 = the bias column)
            const int col0 = cq * 16 + c * 8;
            if (col0 > D) break;
            float va[8], vb[8], vc[8], vd[8];
            tmem_ld8(tlane + cA + col0, va);
            tmem_ld8(tlane + cB + col0, vb);
            tmem_ld8(tlane + cC + col0, vc);
            tmem_ld8(tlane + cD + col0, vd);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int i = col0 + j;
                if (i < D) {
                    stage[row * (D + 1) + i] = va[j];
                    atomicAdd(a.dW2 + i * F + row, vc[j]);                         // lanes = consecutive addresses already
                    if (row < 64) { stage[(128 + row) * (D + 1) + i] = vb[j]; atomicAdd(a.dW2 + i * F + 128 + row, vd[j]); }
                    if (row == 64) atomicAdd(a.db2 + i, vd[j]);
                } else if (i == D) {
                    atomicAdd(a.db1 + row, va[j]);
                    if (row < 64) atomicAdd(a.db1 + 128 + row, vb[j]);
                }
            }
        }
        __syncthreads();
        flush_rows(stage, D + 1, F, D, tid, NT4, [&](int r, int c) { return a.dW1 + r * D + c; });
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

__global__ void ff_pack_kernel(const float *__restrict__ w1, const float *__restrict__ w2, unsigned char *__restrict__ img) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < F * D) {
        const int j = i / D, k = i % D;                                          // w1 [192][48]
        *reinterpret_cast<__nv_bfloat16 *>(img + img_off_bf16(F, j, k)) = __float2bfloat16(w1[i]);
        const int o = i / F, h = i % F;                                          // w2 [48][192]
        *reinterpret_cast<__nv_bfloat16 *>(img + kW1 + img_off_bf16(D, o, h)) = __float2bfloat16(w2[i]);
    }
}

constexpr int kPackLayers = 8;
struct PackN { const float *w1[kPackLayers], *w2[kPackLayers]; unsigned char *img[kPackLayers]; };
__global__ void ff_pack_n_kernel(const PackN a) {                                // blockIdx.y = layer
    const int i = blockIdx.x * blockDim.x + threadIdx.x, l = blockIdx.y;
    if (i < F * D) {
        *reinterpret_cast<__nv_bfloat16 *>(a.img[l] + img_off_bf16(F, i / D, i % D)) = __float2bfloat16(a.w1[l][i]);
        *reinterpret_cast<__nv_bfloat16 *>(a.img[l] + kW1 + img_off_bf16(D, i / F, i % F)) = __float2bfloat16(a.w2[l][i]);
    }
}

static int check_dims(int dim, int hidden) {
    XTRL_REQUIRE(dim == D && hidden == F, XTRL_E_UNSUPPORTED, "xtrl_ff_*: built for dim=%d hidden=%d (got %d, %d)", D, F, dim, hidden);
    return XTRL_OK;
}
static int fill(const xtrl_ff_io *io, int64_t N, Args &a) {
    XTRL_REQUIRE(N <= (int64_t)1 << 24, XTRL_E_BADARG, "xtrl_ff_*: at most 2^24 tokens per call (N=%lld)", (long long)N);
    XTRL_REQUIRE(io->dropout >= 0.f && io->dropout < 1.f, XTRL_E_BADARG, "xtrl_ff_*: dropout %f not in [0,1)", io->dropout);
    a.x = io->x; a.dy = io->dy; a.gamma = io->ln_gamma; a.b1 = io->b1; a.b2 = io->b2;
    a.wimg = static_cast<const unsigned char *>(io->w_img);
    a.y = io->y; a.dx = io->dx; a.dgamma = io->d_ln_gamma; a.dW1 = io->d_w1; a.db1 = io->d_b1; a.dW2 = io->d_w2; a.db2 = io->d_b2;
    a.rng_step = io->rng_step; a.rng_key = io->rng_key;
    const uint32_t thr = (uint32_t)(io->dropout * 128.f + 0.5f);             // the drop probability is quantised to 1/128
    a.drop_add4 = thr ? (128u - thr) * 0x01010101u : 0u;
    a.drop_scale = 1.f / (1.f - (float)thr / 128.f);
    a.N = (int)N;
    return XTRL_OK;
}

}  // namespace ff
}  // namespace xtrl

extern "C" int64_t xtrl_ff_img_bytes(int32_t dim, int32_t hidden) {
    using namespace xtrl::ff;
    return (dim == D && hidden == F) ? (int64_t)(kW1 + kW2) : -1;
}

extern "C" int xtrl_ff_pack(const float *w1, const float *w2, void *img, int32_t dim, int32_t hidden, void *stream) {
    using namespace xtrl; using namespace xtrl::ff;
    if (int rc = check_dims(dim, hidden)) return rc;
    XTRL_REQUIRE(w1 && w2 && img, XTRL_E_BADARG, "xtrl_ff_pack: null pointer");
    ff_pack_kernel<<<(F * D + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w1, w2, static_cast<unsigned char *>(img));
    XTRL_LAUNCH_CHECK("ff_pack_kernel");
    return XTRL_OK;
}

extern "C" int xtrl_ff_pack_layers(const float *const *w1, const float *const *w2, void *const *img, int32_t nlayers, int32_t dim,
                                   int32_t hidden, void *stream) {
    using namespace xtrl; using namespace xtrl::ff;
    if (int rc = check_dims(dim, hidden)) return rc;
    XTRL_REQUIRE(w1 && w2 && img && nlayers >= 1 && nlayers <= kPackLayers, XTRL_E_BADARG,
                 "xtrl_ff_pack_layers: null pointer or nlayers not in 1..%d (got %d)", kPackLayers, nlayers);
    PackN a;
    for (int l = 0; l < nlayers; ++l) {
        XTRL_REQUIRE(w1[l] && w2[l] && img[l], XTRL_E_BADARG, "xtrl_ff_pack_layers: null pointer (layer %d)", l);
        a.w1[l] = w1[l]; a.w2[l] = w2[l]; a.img[l] = static_cast<unsigned char *>(img[l]);
    }
    ff_pack_n_kernel<<<dim3((F * D + 255) / 256, nlayers), 256, 0, (cudaStream_t)stream>>>(a);
    XTRL_LAUNCH_CHECK("ff_pack_n_kernel");
    return XTRL_OK;
}

extern "C" int xtrl_ff_fwd(const xtrl_ff_io *io, int64_t N, int32_t dim, int32_t hidden, void *stream) {
    using namespace xtrl; using namespace xtrl::ff;
    if (int rc = check_dims(dim, hidden)) return rc;
    XTRL_REQUIRE(io && N >= 0, XTRL_E_BADARG, "xtrl_ff_fwd: bad arguments");
    if (N == 0) return XTRL_OK;
    XTRL_REQUIRE(io->x && io->ln_gamma && io->b1 && io->b2 && io->w_img && io->y, XTRL_E_BADARG, "xtrl_ff_fwd: null pointer");
    XTRL_REQUIRE((((uintptr_t)io->x | (uintptr_t)io->y) & 15) == 0, XTRL_E_BADARG, "xtrl_ff_fwd: x and y must be 16-byte aligned (row tiles move by bulk copy)");
    Args a;
    if (int rc = fill(io, N, a)) return rc;
    static bool attr = false;
    if (!attr) { XTRL_CUDA(cudaFuncSetAttribute(ff_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmem)); attr = true; }
    const int ntiles = (int)((N + TOK - 1) / TOK);
    const int grid = ntiles < 2 * num_sms() ? ntiles : 2 * num_sms();
    ff_fwd_kernel<<<grid, NT, kFwdSmem, (cudaStream_t)stream>>>(a);
    XTRL_LAUNCH_CHECK("ff_fwd_kernel");
    return XTRL_OK;
}

extern "C" int xtrl_ff_bwd(const xtrl_ff_io *io, int64_t N, int32_t dim, int32_t hidden, void *stream) {
    using namespace xtrl; using namespace xtrl::ff;
    if (int rc = check_dims(dim, hidden)) return rc;
    XTRL_REQUIRE(io && N >= 0, XTRL_E_BADARG, "xtrl_ff_bwd: bad arguments");
    if (N == 0) return XTRL_OK;
    XTRL_REQUIRE(io->x && io->dy && io->ln_gamma && io->b1 && io->w_img && io->dx && io->d_ln_gamma && io->d_w1 && io->d_b1 &&
                 io->d_w2 && io->d_b2, XTRL_E_BADARG, "xtrl_ff_bwd: null pointer");
    XTRL_REQUIRE(((uintptr_t)io->x & 15) == 0, XTRL_E_BADARG, "xtrl_ff_bwd: x must be 16-byte aligned (row tiles move by bulk copy)");
    Args a;
    if (int rc = fill(io, N, a)) return rc;
    static bool attr = false;
    if (!attr) { XTRL_CUDA(cudaFuncSetAttribute(ff_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem)); attr = true; }
    const int ntiles = (int)((N + TOK - 1) / TOK);
    const int grid = ntiles < num_sms() ? ntiles : num_sms();
    ff_bwd_kernel<<<grid, NT4, kBwdSmem, (cudaStream_t)stream>>>(a);
    XTRL_LAUNCH_CHECK("ff_bwd_kernel");
    return XTRL_OK;
}
