// learn_attn.cu -- learn-side attention sub-layer of the sequence model around the attention core, forward and
// backward, on tcgen05 (the core itself -- softmax(q k^T) v over the full sequence -- is a separate call).
//
//   pre :  [q | k | v] = (LN(x) * gamma) * Wqkv^T,  partial rotary (first dh/2 dims of every head) on q and k,
//          written as bf16 [b][H][n][16] -- the layout the attention core reads.
//   post:  y = x + mask * (o * Wo^T),   o read as bf16 [b][H][n][16], mask = position < episode length.
//
// Replaces, inside Agent.learn's teacher-forced pass (x_transformers_rl.py:928-935; x-transformers Decoder
// attention sub-layer): LayerNorm, to_q / to_k / to_v, rotary, the head split / merge
// transposes, to_out, the padded-row zeroing and the residual add -- and their backward.  SURVEY 8f row 1.
// Built for the plain attention block (no value gate, no value residual); those variants keep the torch ops.
//
// Same machinery as learn_ff.cu: 128-token tiles, TMEM lane = token, thread (row, half) owns half of the
// columns, bf16 canonical images in shared memory, weight gradients accumulated in TMEM across the CTA's
// tiles from the token-major images re-read as MN-major operands, one atomicAdd pass per CTA.
#include "learn_common.cuh"

namespace xtrl {
namespace attn {
using namespace umma;
using namespace learn;

struct Args {
    const float *x, *gamma, *cos_tab, *sin_tab;           // tables [n][4]
    const unsigned char *wimg;                            // [Wqkv image | Wo image]
    __nv_bfloat16 *q, *k, *v;                             // pre fwd out
    const __nv_bfloat16 *dq, *dk, *dv;                    // pre bwd in
    float *dx, *dgamma, *dwq, *dwk, *dwv;                 // pre bwd out
    const __nv_bfloat16 *o;                               // post in
    const int32_t *lens;
    float *y;                                             // post fwd out
    const float *dy;                                      // post bwd in
    __nv_bfloat16 *d_o;                                   // post bwd out
    float *dwo;
    int N, n;
};

// rotate (fwd) or un-rotate (bwd: the transpose) the first 8 of 16 head dims, pairs (2i, 2i+1)
__device__ __forceinline__ void rotary8(float (&v)[16], const float4 cs, const float4 sn, bool inverse) {
    const float c[4] = {cs.x, cs.y, cs.z, cs.w};
    const float s[4] = {sn.x, sn.y, sn.z, sn.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float si = inverse ? -s[i] : s[i];
        const float a = v[2 * i], b = v[2 * i + 1];
        v[2 * i] = a * c[i] - b * si;
        v[2 * i + 1] = b * c[i] + a * si;
    }
}
__device__ __forceinline__ void store16(__nv_bfloat16 *p, const float (&v)[16]) {
    uint4 *d = reinterpret_cast<uint4 *>(p);
    d[0] = make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
    d[1] = make_uint4(pack2(v[8], v[9]), pack2(v[10], v[11]), pack2(v[12], v[13]), pack2(v[14], v[15]));
}
// 16 bf16 of one (token, head) -> two 16-byte pieces of a 128-token image (chunks 2*blk, 2*blk+1)
__device__ __forceinline__ void copy_block(unsigned char *img, int row, int blk, const __nv_bfloat16 *src, bool ok) {
    const uint4 z = make_uint4(0, 0, 0, 0);
    const uint4 a = ok ? __ldg(reinterpret_cast<const uint4 *>(src)) : z, b = ok ? __ldg(reinterpret_cast<const uint4 *>(src) + 1) : z;
    unsigned char *p = img + (2 * blk) * kCore + (row >> 3) * 128 + (row & 7) * 16;
    *reinterpret_cast<uint4 *>(p) = a;
    *reinterpret_cast<uint4 *>(p + kCore) = b;
}

template <int H> struct Dims {
    static constexpr int I = 16 * H, NQ = 3 * I;
    static constexpr int kWqkv = NQ * D * 2, kWo = D * I * 2;
    static constexpr int kMT = (NQ + 127) / 128;          // 128-row tiles of the Wqkv gradient
};

// -------------------------------------------------------------------------------------------- pre fwd
template <int H>
__global__ void __launch_bounds__(NT, 1) attn_pre_fwd_kernel(const Args a) {
    using P = Dims<H>;
    constexpr int NQ = P::NQ, TCOLS = NQ > 256 ? 512 : 256;
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *sW = smem, *sXn = sW + P::kWqkv;
    float *sG = reinterpret_cast<float *>(sXn + TOK * D * 2);
    uint64_t *bars = reinterpret_cast<uint64_t *>(sG + D);
    uint32_t *tslot = reinterpret_cast<uint32_t *>(bars + 2);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = warp & 3, half = warp >> 2, row = q * 32 + lane;
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(tslot, TCOLS);
    if (tid < D) sG[tid] = a.gamma[tid];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tslot, tlane = tbase + ((uint32_t)(q * 32) << 16);
    if (tid == 0) { mbar_expect_tx(&bars[0], P::kWqkv); bulk_g2s(sW, a.wimg, P::kWqkv, &bars[0]); }
    uint32_t ph = 0;
    bool first = true;
    const int ntiles = (a.N + TOK - 1) / TOK;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int t = tile * TOK + row;
        const bool ok = t < a.N;
        const int e = t / a.n, pos = t - e * a.n;
        float xr[D], own[24], mean, rstd;
        load_row(a.x + (size_t)t * D, ok, xr);
        ln_stats(xr, mean, rstd);
#pragma unroll
        for (int i = 0; i < 24; ++i) own[i] = half ? xr[24 + i] : xr[i];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = (own[c * 8 + j] - mean) * rstd * sG[half * 24 + c * 8 + j];
            put8(sXn, row, half * 3 + c, v);
        }
        const float4 cs = ok ? __ldg(reinterpret_cast<const float4 *>(a.cos_tab) + pos) : make_float4(1.f, 1.f, 1.f, 1.f);
        const float4 sn = ok ? __ldg(reinterpret_cast<const float4 *>(a.sin_tab) + pos) : make_float4(0.f, 0.f, 0.f, 0.f);
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            if (first) mbar_wait(&bars[0], 0);
            tc_fence_after();
#pragma unroll
            for (int c0 = 0; c0 < NQ; c0 += 256) {
                const int nc = NQ - c0 < 256 ? NQ - c0 : 256;
#pragma unroll
                for (int s = 0; s < D / 16; ++s)
                    mma_bf16_ss(tbase + c0, smem_desc(sXn + s * 2 * kCore, kCore, 128),
                                smem_desc(sW + (c0 / 8) * 128 + s * 2 * NQ * 16, NQ * 16, 128), idesc_bf16(128, nc), s > 0);
            }
            mma_commit(&bars[1]);
        }
        first = false;
        mbar_wait(&bars[1], ph); ph ^= 1;
        tc_fence_after();
#pragma unroll 1
        for (int bb = 0; bb < 3 * H / 2; ++bb) {
            const int bi = half * (3 * H / 2) + bb, which = bi / H, h = bi - which * H;
            float v[16];
            tmem_ld16(tlane + bi * 16, v);
            if (which < 2) rotary8(v, cs, sn, false);
            __nv_bfloat16 *dst = (which == 0 ? a.q : which == 1 ? a.k : a.v) + ((size_t)(e * H + h) * a.n + pos) * 16;
            if (ok) store16(dst, v);
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, TCOLS);
}

// -------------------------------------------------------------------------------------------- pre bwd
template <int H>
__global__ void __launch_bounds__(NT, 1) attn_pre_bwd_kernel(const Args a) {
    using P = Dims<H>;
    constexpr int NQ = P::NQ, I = P::I;
    constexpr uint32_t cDxn = 0, cAcc = 64;
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *sDQ = smem, *sXn = sDQ + TOK * NQ * 2, *sW = sXn + TOK * D * 2;     // sDQ first: its MN-major over-read lands in sXn / sW
    float *sG = reinterpret_cast<float *>(sW + P::kWqkv);
    uint64_t *bars = reinterpret_cast<uint64_t *>(sG + D);
    uint32_t *tslot = reinterpret_cast<uint32_t *>(bars + 2);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = warp & 3, half = warp >> 2, row = q * 32 + lane;
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(tslot, 256);
    if (tid < D) sG[tid] = a.gamma[tid];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tslot, tlane = tbase + ((uint32_t)(q * 32) << 16);
    if (tid == 0) { mbar_expect_tx(&bars[0], P::kWqkv); bulk_g2s(sW, a.wimg, P::kWqkv, &bars[0]); }
    uint32_t ph = 0;
    bool first = true;
    float dg_acc[24];
#pragma unroll
    for (int i = 0; i < 24; ++i) dg_acc[i] = 0.f;
    const int ntiles = (a.N + TOK - 1) / TOK;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int t = tile * TOK + row;
        const bool ok = t < a.N;
        const int e = t / a.n, pos = t - e * a.n;
        float xr[D], own[24], mean, rstd;
        load_row(a.x + (size_t)t * D, ok, xr);
        ln_stats(xr, mean, rstd);
#pragma unroll
        for (int i = 0; i < D; ++i) xr[i] = (xr[i] - mean) * rstd;
#pragma unroll
        for (int i = 0; i < 24; ++i) own[i] = half ? xr[24 + i] : xr[i];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = own[c * 8 + j] * sG[half * 24 + c * 8 + j];
            put8(sXn, row, half * 3 + c, v);
        }
        const float4 cs = ok ? __ldg(reinterpret_cast<const float4 *>(a.cos_tab) + pos) : make_float4(1.f, 1.f, 1.f, 1.f);
        const float4 sn = ok ? __ldg(reinterpret_cast<const float4 *>(a.sin_tab) + pos) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
        for (int bb = 0; bb < 3 * H / 2; ++bb) {
            const int bi = half * (3 * H / 2) + bb, which = bi / H, h = bi - which * H;
            const __nv_bfloat16 *src = (which == 0 ? a.dq : which == 1 ? a.dk : a.dv) + ((size_t)(e * H + h) * a.n + pos) * 16;
            if (which < 2) {
                float lo[8], hi[8], v[16];
                const uint4 z = make_uint4(0, 0, 0, 0);
                unpack8(ok ? __ldg(reinterpret_cast<const uint4 *>(src)) : z, lo);
                unpack8(ok ? __ldg(reinterpret_cast<const uint4 *>(src) + 1) : z, hi);
#pragma unroll
                for (int j = 0; j < 8; ++j) { v[j] = lo[j]; v[8 + j] = hi[j]; }
                rotary8(v, cs, sn, true);
#pragma unroll
                for (int j = 0; j < 8; ++j) lo[j] = v[j];
                put8(sDQ, row, 2 * bi, lo);
                put8(sDQ, row, 2 * bi + 1, hi);
            } else {
                copy_block(sDQ, row, bi, src, ok);
            }
        }
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            if (first) mbar_wait(&bars[0], 0);
            tc_fence_after();
#pragma unroll
            for (int s = 0; s < NQ / 16; ++s)                                    // dXn = dQKV * Wqkv (image read MN-major)
                mma_bf16_ss(tbase + cDxn, smem_desc(sDQ + s * 2 * kCore, kCore, 128), smem_desc(sW + s * 256, 128, NQ * 16),
                            idesc_bf16(128, D) | kIdB, s > 0);
#pragma unroll
            for (int s = 0; s < TOK / 16; ++s) {                                 // dWqkv += dQKV^T * Xn
                const uint64_t bd = smem_desc(sXn + s * 256, 128, kCore);
#pragma unroll
                for (int m = 0; m < P::kMT; ++m)
                    mma_bf16_ss(tbase + cAcc + m * D, smem_desc(sDQ + m * 16 * kCore + s * 256, 128, kCore), bd,
                                idesc_bf16(128, D) | kIdA | kIdB, !first || s > 0);
            }
            mma_commit(&bars[1]);
        }
        first = false;
        mbar_wait(&bars[1], ph); ph ^= 1;
        tc_fence_after();
        {
            float dxn[D];
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                float v[8];
                tmem_ld8(tlane + cDxn + c * 8, v);
#pragma unroll
                for (int j = 0; j < 8; ++j) dxn[c * 8 + j] = v[j];
            }
            float c1 = 0.f, c2 = 0.f;
#pragma unroll
            for (int i = 0; i < D; ++i) { const float dh = dxn[i] * sG[i]; c1 += dh; c2 = fmaf(dh, xr[i], c2); }
            c1 *= (1.f / D); c2 *= (1.f / D);
            float o[24];
#pragma unroll
            for (int i = 0; i < 24; ++i) {
                const float dxi = half ? dxn[24 + i] : dxn[i];
                dg_acc[i] = fmaf(dxi, own[i], dg_acc[i]);
                o[i] = rstd * (dxi * sG[half * 24 + i] - c1 - own[i] * c2);
            }
            if (ok) {
                float4 *dst = reinterpret_cast<float4 *>(a.dx + (size_t)t * D + half * 24);
#pragma unroll
                for (int i = 0; i < 6; ++i) dst[i] = make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
            }
        }
        tc_fence_before();
    }
#pragma unroll
    for (int i = 0; i < 24; ++i) {
        const float s = warp_sum(dg_acc[i]);
        if (lane == 0 && s != 0.f) atomicAdd(a.dgamma + half * 24 + i, s);
    }
    if (!first) {
        tc_fence_after();
#pragma unroll 1
        for (int m = 0; m < P::kMT; ++m) {
            const int j = m * 128 + row, which = j / I, jj = j - which * I;
            float *dst = (which == 0 ? a.dwq : which == 1 ? a.dwk : a.dwv) + jj * D + half * 24;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float v[8];
                tmem_ld8(tlane + cAcc + m * D + half * 24 + c * 8, v);
                if (j < NQ) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) atomicAdd(dst + c * 8 + i, v[i]);
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 256);
}

// ------------------------------------------------------------------------------------------- post fwd
template <int H>
__global__ void __launch_bounds__(NT, 2) attn_post_fwd_kernel(const Args a) {
    using P = Dims<H>;
    constexpr int I = P::I;
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *sO = smem, *sW = sO + TOK * I * 2;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sW + P::kWo);
    uint32_t *tslot = reinterpret_cast<uint32_t *>(bars + 2);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = warp & 3, half = warp >> 2, row = q * 32 + lane;
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(tslot, 64);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tslot, tlane = tbase + ((uint32_t)(q * 32) << 16);
    if (tid == 0) { mbar_expect_tx(&bars[0], P::kWo); bulk_g2s(sW, a.wimg + P::kWqkv, P::kWo, &bars[0]); }
    uint32_t ph = 0;
    bool first = true;
    const int ntiles = (a.N + TOK - 1) / TOK;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int t = tile * TOK + row;
        const bool ok = t < a.N;
        const int e = t / a.n, pos = t - e * a.n;
        float own[24];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const float4 v = ok ? __ldg(reinterpret_cast<const float4 *>(a.x + (size_t)t * D + half * 24) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            own[4 * i] = v.x; own[4 * i + 1] = v.y; own[4 * i + 2] = v.z; own[4 * i + 3] = v.w;
        }
        const bool live = ok && pos < __ldg(a.lens + (ok ? e : 0));
#pragma unroll
        for (int hh = 0; hh < H / 2; ++hh) {
            const int h = half * (H / 2) + hh;
            copy_block(sO, row, h, a.o + ((size_t)(e * H + h) * a.n + pos) * 16, ok);
        }
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            if (first) mbar_wait(&bars[0], 0);
            tc_fence_after();
#pragma unroll
            for (int s = 0; s < I / 16; ++s)
                mma_bf16_ss(tbase, smem_desc(sO + s * 2 * kCore, kCore, 128), smem_desc(sW + s * 2 * D * 16, D * 16, 128), idesc_bf16(128, D), s > 0);
            mma_commit(&bars[1]);
        }
        first = false;
        mbar_wait(&bars[1], ph); ph ^= 1;
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float v[8];
            tmem_ld8(tlane + half * 24 + c * 8, v);
            if (ok) {
                float4 *dst = reinterpret_cast<float4 *>(a.y + (size_t)t * D + half * 24 + c * 8);
                const float *o = own + c * 8;
                const float m = live ? 1.f : 0.f;
                dst[0] = make_float4(fmaf(m, v[0], o[0]), fmaf(m, v[1], o[1]), fmaf(m, v[2], o[2]), fmaf(m, v[3], o[3]));
                dst[1] = make_float4(fmaf(m, v[4], o[4]), fmaf(m, v[5], o[5]), fmaf(m, v[6], o[6]), fmaf(m, v[7], o[7]));
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 64);
}

// ------------------------------------------------------------------------------------------- post bwd
template <int H>
__global__ void __launch_bounds__(NT, 2) attn_post_bwd_kernel(const Args a) {
    using P = Dims<H>;
    constexpr int I = P::I;
    constexpr uint32_t cDo = 0, cAcc = 128;
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *sO = smem, *sDY = sO + TOK * I * 2, *sW = sDY + TOK * D * 2;     // sO first: its over-read (H = 4) lands in sDY / sW
    uint64_t *bars = reinterpret_cast<uint64_t *>(sW + P::kWo);
    uint32_t *tslot = reinterpret_cast<uint32_t *>(bars + 2);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = warp & 3, half = warp >> 2, row = q * 32 + lane;
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(tslot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tslot, tlane = tbase + ((uint32_t)(q * 32) << 16);
    if (tid == 0) { mbar_expect_tx(&bars[0], P::kWo); bulk_g2s(sW, a.wimg + P::kWqkv, P::kWo, &bars[0]); }
    uint32_t ph = 0;
    bool first = true;
    const int ntiles = (a.N + TOK - 1) / TOK;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int t = tile * TOK + row;
        const bool ok = t < a.N;
        const int e = t / a.n, pos = t - e * a.n;
        const bool live = ok && pos < __ldg(a.lens + (ok ? e : 0));
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float v[8];
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 *src = reinterpret_cast<const float4 *>(a.dy + (size_t)t * D + half * 24 + c * 8);
            const float4 lo = live ? __ldg(src) : z, hi = live ? __ldg(src + 1) : z;
            v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w; v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
            put8(sDY, row, half * 3 + c, v);
        }
#pragma unroll
        for (int hh = 0; hh < H / 2; ++hh) {
            const int h = half * (H / 2) + hh;
            copy_block(sO, row, h, a.o + ((size_t)(e * H + h) * a.n + pos) * 16, ok);
        }
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            if (first) mbar_wait(&bars[0], 0);
            tc_fence_after();
#pragma unroll
            for (int s = 0; s < D / 16; ++s)                                     // dO = dOut * Wo (image read MN-major)
                mma_bf16_ss(tbase + cDo, smem_desc(sDY + s * 2 * kCore, kCore, 128), smem_desc(sW + s * 256, 128, D * 16),
                            idesc_bf16(128, I) | kIdB, s > 0);
#pragma unroll
            for (int s = 0; s < TOK / 16; ++s)                                   // dWo^T += O^T * dOut
                mma_bf16_ss(tbase + cAcc, smem_desc(sO + s * 256, 128, kCore), smem_desc(sDY + s * 256, 128, kCore),
                            idesc_bf16(128, D) | kIdA | kIdB, !first || s > 0);
            mma_commit(&bars[1]);
        }
        first = false;
        mbar_wait(&bars[1], ph); ph ^= 1;
        tc_fence_after();
#pragma unroll
        for (int hh = 0; hh < H / 2; ++hh) {
            const int h = half * (H / 2) + hh;
            float v[16];
            tmem_ld16(tlane + cDo + h * 16, v);
            if (ok) store16(a.d_o + ((size_t)(e * H + h) * a.n + pos) * 16, v);
        }
        tc_fence_before();
    }
    if (!first) {
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float v[8];
            tmem_ld8(tlane + cAcc + half * 24 + c * 8, v);
            if (row < I) {
#pragma unroll
                for (int i = 0; i < 8; ++i) atomicAdd(a.dwo + (half * 24 + c * 8 + i) * I + row, v[i]);
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 256);
}

template <int H>
__global__ void attn_pack_kernel(const float *__restrict__ wq, const float *__restrict__ wk, const float *__restrict__ wv,
                                 const float *__restrict__ wo, unsigned char *__restrict__ img) {
    using P = Dims<H>;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P::NQ * D) {
        const int j = i / D, k = i % D, which = j / P::I, jj = j % P::I;         // Wqkv rows: q, k, v outputs; K = 48
        const float w = (which == 0 ? wq : which == 1 ? wk : wv)[jj * D + k];
        *reinterpret_cast<__nv_bfloat16 *>(img + img_off_bf16(P::NQ, j, k)) = __float2bfloat16(w);
    }
    if (i < D * P::I) {
        const int r = i / P::I, c = i % P::I;                                    // Wo [48][I]: rows = outputs, K = I
        *reinterpret_cast<__nv_bfloat16 *>(img + P::kWqkv + img_off_bf16(D, r, c)) = __float2bfloat16(wo[i]);
    }
}

template <int H> struct Smem {
    using P = Dims<H>;
    static constexpr int pre_fwd = P::kWqkv + TOK * D * 2 + D * 4 + 64;
    static constexpr int pre_bwd = TOK * P::NQ * 2 + TOK * D * 2 + P::kWqkv + D * 4 + 64;
    static constexpr int post_fwd = TOK * P::I * 2 + P::kWo + 64;
    static constexpr int post_bwd = TOK * P::I * 2 + TOK * D * 2 + P::kWo + 64;
};

static int fill(const xtrl_attn_io *io, int32_t b, int32_t n, Args &a) {
    XTRL_REQUIRE((int64_t)b * n <= (int64_t)1 << 24, XTRL_E_BADARG, "xtrl_attn_*: at most 2^24 tokens per call");
    a.x = io->x; a.gamma = io->ln_gamma; a.cos_tab = io->cos_tab; a.sin_tab = io->sin_tab;
    a.wimg = static_cast<const unsigned char *>(io->w_img);
    a.q = static_cast<__nv_bfloat16 *>(io->q); a.k = static_cast<__nv_bfloat16 *>(io->k); a.v = static_cast<__nv_bfloat16 *>(io->v);
    a.dq = static_cast<const __nv_bfloat16 *>(io->dq); a.dk = static_cast<const __nv_bfloat16 *>(io->dk); a.dv = static_cast<const __nv_bfloat16 *>(io->dv);
    a.dx = io->dx; a.dgamma = io->d_ln_gamma; a.dwq = io->d_wq; a.dwk = io->d_wk; a.dwv = io->d_wv;
    a.o = static_cast<const __nv_bfloat16 *>(io->o); a.lens = io->lens; a.y = io->y; a.dy = io->dy;
    a.d_o = static_cast<__nv_bfloat16 *>(io->d_o); a.dwo = io->d_wo;
    a.N = b * n; a.n = n;
    return XTRL_OK;
}

template <typename K>
static int launch(K kernel, const char *name, int smem, int ctas_per_sm, const Args &a, void *stream) {
    static const void *seen[16];                    // kernels whose shared-memory limit is already raised
    static int nseen = 0;
    bool known = false;
    for (int i = 0; i < nseen; ++i) known |= seen[i] == (const void *)kernel;
    if (!known) {
        XTRL_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        if (nseen < 16) seen[nseen++] = (const void *)kernel;
    }
    const int ntiles = (a.N + TOK - 1) / TOK, cap = ctas_per_sm * kNumSMs;
    kernel<<<ntiles < cap ? ntiles : cap, NT, smem, (cudaStream_t)stream>>>(a);
    XTRL_LAUNCH_CHECK(name);
    return XTRL_OK;
}

}  // namespace attn
}  // namespace xtrl

static int attn_check(int32_t dim, int32_t heads, int32_t dim_head) {
    using namespace xtrl;
    XTRL_REQUIRE(dim == learn::D && dim_head == 16 && (heads == 4 || heads == 8), XTRL_E_UNSUPPORTED,
                 "xtrl_attn_*: built for dim=48, dim_head=16, heads in {4, 8} (got %d, %d, %d)", dim, dim_head, heads);
    return XTRL_OK;
}

extern "C" int64_t xtrl_attn_img_bytes(int32_t dim, int32_t heads, int32_t dim_head) {
    using namespace xtrl::attn;
    if (dim != xtrl::learn::D || dim_head != 16 || (heads != 4 && heads != 8)) return -1;
    return heads == 4 ? Dims<4>::kWqkv + Dims<4>::kWo : Dims<8>::kWqkv + Dims<8>::kWo;
}

extern "C" int xtrl_attn_pack(const float *wq, const float *wk, const float *wv, const float *wo, void *img,
                              int32_t dim, int32_t heads, int32_t dim_head, void *stream) {
    using namespace xtrl; using namespace xtrl::attn;
    if (int rc = attn_check(dim, heads, dim_head)) return rc;
    XTRL_REQUIRE(wq && wk && wv && wo && img, XTRL_E_BADARG, "xtrl_attn_pack: null pointer");
    unsigned char *p = static_cast<unsigned char *>(img);
    if (heads == 4) attn_pack_kernel<4><<<(Dims<4>::NQ * learn::D + 255) / 256, 256, 0, (cudaStream_t)stream>>>(wq, wk, wv, wo, p);
    else attn_pack_kernel<8><<<(Dims<8>::NQ * learn::D + 255) / 256, 256, 0, (cudaStream_t)stream>>>(wq, wk, wv, wo, p);
    XTRL_LAUNCH_CHECK("attn_pack_kernel");
    return XTRL_OK;
}

extern "C" int xtrl_attn_pre_fwd(const xtrl_attn_io *io, int32_t b, int32_t n, int32_t heads, void *stream) {
    using namespace xtrl; using namespace xtrl::attn;
    if (int rc = attn_check(learn::D, heads, 16)) return rc;
    XTRL_REQUIRE(io && b >= 0 && n >= 0, XTRL_E_BADARG, "xtrl_attn_pre_fwd: bad arguments");
    if ((int64_t)b * n == 0) return XTRL_OK;
    XTRL_REQUIRE(io->x && io->ln_gamma && io->cos_tab && io->sin_tab && io->w_img && io->q && io->k && io->v, XTRL_E_BADARG,
                 "xtrl_attn_pre_fwd: null pointer");
    Args a;
    if (int rc = fill(io, b, n, a)) return rc;
    return heads == 4 ? launch(attn_pre_fwd_kernel<4>, "attn_pre_fwd_kernel", Smem<4>::pre_fwd, 1, a, stream)
                      : launch(attn_pre_fwd_kernel<8>, "attn_pre_fwd_kernel", Smem<8>::pre_fwd, 1, a, stream);
}

extern "C" int xtrl_attn_pre_bwd(const xtrl_attn_io *io, int32_t b, int32_t n, int32_t heads, void *stream) {
    using namespace xtrl; using namespace xtrl::attn;
    if (int rc = attn_check(learn::D, heads, 16)) return rc;
    XTRL_REQUIRE(io && b >= 0 && n >= 0, XTRL_E_BADARG, "xtrl_attn_pre_bwd: bad arguments");
    if ((int64_t)b * n == 0) return XTRL_OK;
    XTRL_REQUIRE(io->x && io->ln_gamma && io->cos_tab && io->sin_tab && io->w_img && io->dq && io->dk && io->dv && io->dx &&
                 io->d_ln_gamma && io->d_wq && io->d_wk && io->d_wv, XTRL_E_BADARG, "xtrl_attn_pre_bwd: null pointer");
    Args a;
    if (int rc = fill(io, b, n, a)) return rc;
    return heads == 4 ? launch(attn_pre_bwd_kernel<4>, "attn_pre_bwd_kernel", Smem<4>::pre_bwd, 1, a, stream)
                      : launch(attn_pre_bwd_kernel<8>, "attn_pre_bwd_kernel", Smem<8>::pre_bwd, 1, a, stream);
}

extern "C" int xtrl_attn_post_fwd(const xtrl_attn_io *io, int32_t b, int32_t n, int32_t heads, void *stream) {
    using namespace xtrl; using namespace xtrl::attn;
    if (int rc = attn_check(learn::D, heads, 16)) return rc;
    XTRL_REQUIRE(io && b >= 0 && n >= 0, XTRL_E_BADARG, "xtrl_attn_post_fwd: bad arguments");
    if ((int64_t)b * n == 0) return XTRL_OK;
    XTRL_REQUIRE(io->x && io->o && io->lens && io->w_img && io->y, XTRL_E_BADARG, "xtrl_attn_post_fwd: null pointer");
    Args a;
    if (int rc = fill(io, b, n, a)) return rc;
    return heads == 4 ? launch(attn_post_fwd_kernel<4>, "attn_post_fwd_kernel", Smem<4>::post_fwd, 2, a, stream)
                      : launch(attn_post_fwd_kernel<8>, "attn_post_fwd_kernel", Smem<8>::post_fwd, 2, a, stream);
}

extern "C" int xtrl_attn_post_bwd(const xtrl_attn_io *io, int32_t b, int32_t n, int32_t heads, void *stream) {
    using namespace xtrl; using namespace xtrl::attn;
    if (int rc = attn_check(learn::D, heads, 16)) return rc;
    XTRL_REQUIRE(io && b >= 0 && n >= 0, XTRL_E_BADARG, "xtrl_attn_post_bwd: bad arguments");
    if ((int64_t)b * n == 0) return XTRL_OK;
    XTRL_REQUIRE(io->dy && io->o && io->lens && io->w_img && io->d_o && io->d_wo, XTRL_E_BADARG, "xtrl_attn_post_bwd: null pointer");
    Args a;
    if (int rc = fill(io, b, n, a)) return rc;
    return heads == 4 ? launch(attn_post_bwd_kernel<4>, "attn_post_bwd_kernel", Smem<4>::post_bwd, 2, a, stream)
                      : launch(attn_post_bwd_kernel<8>, "attn_post_bwd_kernel", Smem<8>::post_bwd, 2, a, stream);
}
