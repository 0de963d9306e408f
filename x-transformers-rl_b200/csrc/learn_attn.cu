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
// Value gates and the learned value-residual mix (c2's block) ride along as extra output blocks of the same GEMM:
// pre also emits the gate pre-activation (bf16 [b][H][n][16], bias added) and the mix logits (bf16 [b*n][16], first H
// valid), pre_bwd takes their gradients as extra K blocks; the sigmoid / lerp / multiply on those small tensors stay
// element-wise ops of the caller.  Image order: [Wo | Wq Wk Wv (Wg) (Wm, zero-padded to 16 rows)].
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
    const float *bg, *bm;                                 // gate / mix biases (pre fwd)
    __nv_bfloat16 *g, *mix;                               // pre fwd out: gate pre-activation [b][H][n][16], mix logits [N][16]
    const __nv_bfloat16 *dg, *dmix;                       // pre bwd in
    float *dwg, *dbg, *dwm, *dbm;                         // pre bwd out
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

// NB = 3 (q, k, v) or 4 (+ gate) blocks of I outputs; MIX adds one 16-wide block of per-head mix logits
template <int H, int NB = 3, int MIX = 0> struct Dims {
    static constexpr int I = 16 * H, NQ = NB * I + 16 * MIX;
    static constexpr int NBLK = NB * H + MIX;             // 16-column output blocks
    static constexpr int kWqkv = NQ * D * 2, kWo = D * I * 2;
    static constexpr int kMT = (NQ + 127) / 128;          // 128-row tiles of the projection-weight gradient
    static constexpr int XW = 64;                         // Xn image width in the backward: 48 features + ones column + pad
};

template <int H, int NB, int MIX> struct Smem {
    using P = Dims<H, NB, MIX>;
    static constexpr int pre_fwd = P::kWqkv + TOK * D * 2 + kRowTile + (D + P::I + 16) * 4 + 64;
    static constexpr int pre_bwd_base = TOK * P::NQ * 2 + TOK * P::XW * 2 + P::kWqkv + kRowTile + D * 4 + 64;
    static constexpr bool pre_bwd_ostage = pre_bwd_base + 2 * kRowTile <= 227 * 1024;
    static constexpr int pre_bwd = pre_bwd_base + (pre_bwd_ostage ? 2 * kRowTile : 0);
    static constexpr int post_fwd = TOK * P::I * 2 + P::kWo + 64;
    static constexpr int post_bwd = TOK * P::I * 2 + TOK * D * 2 + P::kWo + 64;
};

// -------------------------------------------------------------------------------------------- pre fwd
template <int H, int NB, int MIX>
__global__ void __launch_bounds__(NT, 1) attn_pre_fwd_kernel(const Args a) {
    using P = Dims<H, NB, MIX>;
    constexpr int NQ = P::NQ, I = P::I, NBLK = P::NBLK, TCOLS = NQ > 256 ? 512 : 256;
    constexpr int BLK0 = (NBLK + 1) / 2;                                    // blocks of half 0; half 1 takes the rest
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *sW = smem, *sXn = sW + P::kWqkv;
    float *sX32 = reinterpret_cast<float *>(sXn + TOK * D * 2);                      // fp32 row tile (TMA staging)
    float *sG = sX32 + TOK * D, *sBias = sG + D;                                     // bias of the gate block / mix block
    uint64_t *bars = reinterpret_cast<uint64_t *>(sBias + I + 16);
    uint32_t *tslot = reinterpret_cast<uint32_t *>(bars + 3);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = warp & 3, half = warp >> 2, row = q * 32 + lane;
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(tslot, TCOLS);
    if (tid < D) sG[tid] = a.gamma[tid];
    for (int i = tid; i < I + 16; i += NT)
        sBias[i] = i < I ? (NB == 4 ? a.bg[i] : 0.f) : ((MIX && i - I < H) ? a.bm[i - I] : 0.f);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tslot, tlane = tbase + ((uint32_t)(q * 32) << 16);
    const int ntiles = (a.N + TOK - 1) / TOK;
    if (tid == 0) {
        mbar_expect_tx(&bars[0], P::kWqkv); bulk_g2s(sW, a.wimg + P::kWo, P::kWqkv, &bars[0]);
        if ((int)blockIdx.x < ntiles) rows_fetch(sX32, a.x, blockIdx.x, a.N, &bars[2]);
    }
    uint32_t ph = 0, phx = 0;
    bool first = true;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int t = tile * TOK + row;
        const bool ok = t < a.N;
        const int e = t / a.n, pos = t - e * a.n;
        float xr[D], own[24], mean, rstd;
        mbar_wait(&bars[2], phx); phx ^= 1;
        rows_read(sX32, row, ok, xr);
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
            if (tile + (int)gridDim.x < ntiles) rows_fetch(sX32, a.x, tile + gridDim.x, a.N, &bars[2]);   // every thread has read its row
        }
        first = false;
        mbar_wait(&bars[1], ph); ph ^= 1;
        tc_fence_after();
        const int b0 = half ? BLK0 : 0, b1 = half ? NBLK : BLK0;
#pragma unroll 1
        for (int bi = b0; bi < b1; ++bi) {
            float v[16];
            tmem_ld16(tlane + bi * 16, v);
            if (bi < NB * H) {
                const int which = bi / H, h = bi - which * H;
                if (which < 2) rotary8(v, cs, sn, false);
                if (NB == 4 && which == 3) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] += sBias[h * 16 + j];
                }
                __nv_bfloat16 *dst = (which == 0 ? a.q : which == 1 ? a.k : which == 2 ? a.v : a.g) + ((size_t)(e * H + h) * a.n + pos) * 16;
                if (ok) store16(dst, v);
            } else {                                                          // mix logits, one per head
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] += sBias[I + j];
                if (ok) store16(a.mix + (size_t)t * 16, v);
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, TCOLS);
}

// -------------------------------------------------------------------------------------------- pre bwd
template <int H, int NB, int MIX>
__global__ void __launch_bounds__(NT, 1) attn_pre_bwd_kernel(const Args a) {
    using P = Dims<H, NB, MIX>;
    constexpr int NQ = P::NQ, I = P::I, NBLK = P::NBLK, XW = P::XW;
    constexpr int BLK0 = (NBLK + 1) / 2;
    constexpr uint32_t cDxn = 0, cAcc = 64;
    constexpr int TCOLS = 64 + P::kMT * XW > 256 ? 512 : 256;
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *sDQ = smem, *sXn = sDQ + TOK * NQ * 2, *sW = sXn + TOK * XW * 2;    // sDQ first: its MN-major over-read lands in sXn / sW
    float *sX32 = reinterpret_cast<float *>(sW + P::kWqkv);                             // fp32 row tile of x (TMA staging)
    constexpr bool OST = Smem<H, NB, MIX>::pre_bwd_ostage;                              // dx leaves through shared memory too (two buffers)
    float *sOut = sX32 + TOK * D;
    float *sG = sOut + (OST ? 2 * TOK * D : 0);
    uint64_t *bars = reinterpret_cast<uint64_t *>(sG + D);
    uint32_t *tslot = reinterpret_cast<uint32_t *>(bars + 4);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = warp & 3, half = warp >> 2, row = q * 32 + lane;
    if (tid == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1); fence_mbar_init(); }   // 3: residual-gradient tile landed
    if (warp == 0) tmem_alloc(tslot, TCOLS);
    if (tid < D) sG[tid] = a.gamma[tid];
    if (half == 0) {                                                          // constant ones column (bias gradients) + zero pad
        const float one[8] = {1.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, zero[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        put8(sXn, row, 6, one); put8(sXn, row, 7, zero);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tslot, tlane = tbase + ((uint32_t)(q * 32) << 16);
    const int ntiles = (a.N + TOK - 1) / TOK;
    if (tid == 0) {
        mbar_expect_tx(&bars[0], P::kWqkv); bulk_g2s(sW, a.wimg + P::kWo, P::kWqkv, &bars[0]);
        if ((int)blockIdx.x < ntiles) rows_fetch(sX32, a.x, blockIdx.x, a.N, &bars[2]);
    }
    uint32_t ph = 0, phx = 0, phr = 0;
    bool first = true;
    int pend = -1, it = 0;                                                    // tile whose dx still sits in sOut[(it - 1) & 1]
    const bool res = a.dy != nullptr;                                         // dx = dy (the block's residual path) + the LayerNorm path
    float dg_acc[24];
#pragma unroll
    for (int i = 0; i < 24; ++i) dg_acc[i] = 0.f;
    const int b0 = half ? BLK0 : 0, b1 = half ? NBLK : BLK0;
    // this thread's gradient blocks of a tile, global -> the dY image, all in flight at once (sDQ must be free: the MMAs
    // of the previous tile have completed)
    auto fetch_dy = [&](int tile) {
        const int t = tile * TOK + row;
        const bool ok = t < a.N;
        const int tc = ok ? t : 0, e = tc / a.n, pos = tc - e * a.n;
#pragma unroll 4
        for (int bi = b0; bi < b1; ++bi) {
            const int which = bi / H, h = bi - which * H;
            const __nv_bfloat16 *src = bi >= NB * H ? a.dmix + (size_t)tc * 16
                : (which == 0 ? a.dq : which == 1 ? a.dk : which == 2 ? a.dv : a.dg) + ((size_t)(e * H + h) * a.n + pos) * 16;
            unsigned char *p = sDQ + (2 * bi) * kCore + (row >> 3) * 128 + (row & 7) * 16;
            cp_async16(p, src, ok);
            cp_async16(p + kCore, src + 8, ok);
        }
        cp_async_commit();
    };
    if ((int)blockIdx.x < ntiles) fetch_dy(blockIdx.x);
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int t = tile * TOK + row;
        const bool ok = t < a.N;
        const int e = t / a.n, pos = t - e * a.n;
        float xr[D], own[24], mean, rstd;
        mbar_wait(&bars[2], phx); phx ^= 1;
        rows_read(sX32, row, ok, xr);
        const float4 cs = ok ? __ldg(reinterpret_cast<const float4 *>(a.cos_tab) + pos) : make_float4(1.f, 1.f, 1.f, 1.f);
        const float4 sn = ok ? __ldg(reinterpret_cast<const float4 *>(a.sin_tab) + pos) : make_float4(0.f, 0.f, 0.f, 0.f);
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
        cp_async_wait_all();
        // un-rotate the first 8 dims of this thread's q / k blocks where they lie (the thread's own copies: no barrier)
#pragma unroll 2
        for (int bi = b0; bi < b1; ++bi) {
            if (bi >= 2 * H) break;
            uint4 *p = reinterpret_cast<uint4 *>(sDQ + (2 * bi) * kCore + (row >> 3) * 128 + (row & 7) * 16);
            float lo[8];
            unpack8(*p, lo);
            const float c[4] = {cs.x, cs.y, cs.z, cs.w}, sv[4] = {sn.x, sn.y, sn.z, sn.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float x0 = lo[2 * i], x1 = lo[2 * i + 1];
                lo[2 * i] = x0 * c[i] + x1 * sv[i];
                lo[2 * i + 1] = x1 * c[i] - x0 * sv[i];
            }
            *p = make_uint4(pack2(lo[0], lo[1]), pack2(lo[2], lo[3]), pack2(lo[4], lo[5]), pack2(lo[6], lo[7]));
        }
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            if (OST && pend >= 0) { bulk_wait_read(); rows_store(a.dx, sOut + ((it - 1) & 1) * TOK * D, pend, a.N); }
            if (OST && res) rows_fetch(sOut + (it & 1) * TOK * D, a.dy, tile, a.N, &bars[3]);   // lands where this tile's dx is formed
            if (first) mbar_wait(&bars[0], 0);
            tc_fence_after();
#pragma unroll
            for (int s = 0; s < NQ / 16; ++s)                                    // dXn = dY * W (image read MN-major)
                mma_bf16_ss(tbase + cDxn, smem_desc(sDQ + s * 2 * kCore, kCore, 128), smem_desc(sW + s * 256, 128, NQ * 16),
                            idesc_bf16(128, D) | kIdB, s > 0);
#pragma unroll
            for (int s = 0; s < TOK / 16; ++s) {                                 // dW | db += dY^T * [Xn | 1]
                const uint64_t bd = smem_desc(sXn + s * 256, 128, kCore);
#pragma unroll
                for (int m = 0; m < P::kMT; ++m)
                    mma_bf16_ss(tbase + cAcc + m * XW, smem_desc(sDQ + m * 16 * kCore + s * 256, 128, kCore), bd,
                                idesc_bf16(128, XW) | kIdA | kIdB, !first || s > 0);
            }
            mma_commit(&bars[1]);
            if (tile + (int)gridDim.x < ntiles) rows_fetch(sX32, a.x, tile + gridDim.x, a.N, &bars[2]);   // every thread has read its row
        }
        first = false;
        mbar_wait(&bars[1], ph); ph ^= 1;
        tc_fence_after();
        if (tile + (int)gridDim.x < ntiles) fetch_dy(tile + gridDim.x);          // flies under the epilogue and the next row load
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
            if (OST || ok) {
                float4 *dst = reinterpret_cast<float4 *>(OST ? sOut + (it & 1) * TOK * D + row * D + half * 24 : a.dx + (size_t)t * D + half * 24);
                if (res) {
                    if (OST) { mbar_wait(&bars[3], phr); phr ^= 1; }
                    const float4 *rsd = OST ? dst : reinterpret_cast<const float4 *>(a.dy + (size_t)t * D + half * 24);
#pragma unroll
                    for (int i = 0; i < 6; ++i) {
                        const float4 r4 = (OST && !ok) ? make_float4(0.f, 0.f, 0.f, 0.f) : rsd[i];
                        o[4 * i] += r4.x; o[4 * i + 1] += r4.y; o[4 * i + 2] += r4.z; o[4 * i + 3] += r4.w;
                    }
                }
#pragma unroll
                for (int i = 0; i < 6; ++i) dst[i] = make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
            }
        }
        pend = tile; ++it;
        tc_fence_before();
    }
    if (OST && pend >= 0) {                                                    // the last tile's dx
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) rows_store(a.dx, sOut + ((it - 1) & 1) * TOK * D, pend, a.N);
    }
#pragma unroll
    for (int i = 0; i < 24; ++i) {
        const float s = warp_sum(dg_acc[i]);
        if (lane == 0 && s != 0.f) atomicAdd(a.dgamma + half * 24 + i, s);
    }
    if (!first) {
        tc_fence_after();
        float *stage = reinterpret_cast<float *>(sDQ);                            // [128][D + 1] fp32, the dY image is done with
        constexpr int PITCH = D + 1;
        auto wrow = [&](int j, float *&bdst) -> float * {                         // output row j of [Wq; Wk; Wv; (Wg); (Wm)]
            const int which = j / I, jj = j - which * I;
            bdst = nullptr;
            if (j < NB * I) {
                if (which == 3) bdst = a.dbg + jj;
                return (which == 0 ? a.dwq : which == 1 ? a.dwk : which == 2 ? a.dwv : a.dwg) + jj * D;
            }
            if (MIX && j - NB * I < H) { bdst = a.dbm + (j - NB * I); return a.dwm + (j - NB * I) * D; }
            return nullptr;
        };
#pragma unroll 1
        for (int m = 0; m < P::kMT; ++m) {
            float *bdst;
            wrow(m * 128 + row, bdst);
            __syncthreads();
#pragma unroll
            for (int c = 0; c < 4; ++c) {                                         // half 0: columns 0..31, half 1: 32..63
                const int col0 = half * 32 + c * 8;
                if (col0 > D) break;
                float v[8];
                tmem_ld8(tlane + cAcc + m * XW + col0, v);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (col0 + i < D) stage[row * PITCH + col0 + i] = v[i];
                    else if (col0 + i == D && bdst) atomicAdd(bdst, v[i]);
                }
            }
            __syncthreads();
            flush_rows(stage, PITCH, 128, D, tid, NT, [&](int r, int c) { float *bd; float *w = wrow(m * 128 + r, bd); return w ? w + c : nullptr; });
        }
        tc_fence_before();
    }
    if (OST && tid == 0) bulk_wait_all();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, TCOLS);
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
    if (tid == 0) { mbar_expect_tx(&bars[0], P::kWo); bulk_g2s(sW, a.wimg, P::kWo, &bars[0]); }
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
    if (tid == 0) { mbar_expect_tx(&bars[0], P::kWo); bulk_g2s(sW, a.wimg, P::kWo, &bars[0]); }
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

template <int H, int NB, int MIX>
__device__ __forceinline__ void attn_pack_one(const float *__restrict__ wq, const float *__restrict__ wk, const float *__restrict__ wv,
                                              const float *__restrict__ wo, const float *__restrict__ wg, const float *__restrict__ wm,
                                              unsigned char *__restrict__ img, int i) {
    using P = Dims<H, NB, MIX>;
    if (i < D * P::I) {
        const int r = i / P::I, c = i % P::I;                                    // Wo [48][I]: rows = outputs, K = I
        *reinterpret_cast<__nv_bfloat16 *>(img + img_off_bf16(D, r, c)) = __float2bfloat16(wo[i]);
    }
    if (i < P::NQ * D) {
        const int j = i / D, k = i % D;                                          // projection rows: q, k, v, (gate), (mix padded to 16); K = 48
        float w = 0.f;
        if (j < NB * P::I) {
            const int which = j / P::I, jj = j % P::I;
            w = (which == 0 ? wq : which == 1 ? wk : which == 2 ? wv : wg)[jj * D + k];
        } else if (j - NB * P::I < H) {
            w = wm[(j - NB * P::I) * D + k];
        }
        *reinterpret_cast<__nv_bfloat16 *>(img + P::kWo + img_off_bf16(P::NQ, j, k)) = __float2bfloat16(w);
    }
}
template <int H, int NB, int MIX>
__global__ void attn_pack_kernel(const float *__restrict__ wq, const float *__restrict__ wk, const float *__restrict__ wv,
                                 const float *__restrict__ wo, const float *__restrict__ wg, const float *__restrict__ wm,
                                 unsigned char *__restrict__ img) {
    attn_pack_one<H, NB, MIX>(wq, wk, wv, wo, wg, wm, img, blockIdx.x * blockDim.x + threadIdx.x);
}
// the same for up to kPackLayers layers in one launch (blockIdx.y = layer): the images follow the weights every optimiser step
constexpr int kPackLayers = 8;
struct PackN { const float *wq[kPackLayers], *wk[kPackLayers], *wv[kPackLayers], *wo[kPackLayers], *wg[kPackLayers], *wm[kPackLayers]; unsigned char *img[kPackLayers]; };
template <int H, int NB, int MIX>
__global__ void attn_pack_n_kernel(const PackN a) {
    const int l = blockIdx.y;
    attn_pack_one<H, NB, MIX>(a.wq[l], a.wk[l], a.wv[l], a.wo[l], a.wg[l], a.wm[l], a.img[l], blockIdx.x * blockDim.x + threadIdx.x);
}


static int fill(const xtrl_attn_io *io, int32_t b, int32_t n, Args &a) {
    XTRL_REQUIRE((int64_t)b * n <= (int64_t)1 << 24, XTRL_E_BADARG, "xtrl_attn_*: at most 2^24 tokens per call");
    a.x = io->x; a.gamma = io->ln_gamma; a.cos_tab = io->cos_tab; a.sin_tab = io->sin_tab;
    a.wimg = static_cast<const unsigned char *>(io->w_img);
    a.q = static_cast<__nv_bfloat16 *>(io->q); a.k = static_cast<__nv_bfloat16 *>(io->k); a.v = static_cast<__nv_bfloat16 *>(io->v);
    a.dq = static_cast<const __nv_bfloat16 *>(io->dq); a.dk = static_cast<const __nv_bfloat16 *>(io->dk); a.dv = static_cast<const __nv_bfloat16 *>(io->dv);
    a.dx = io->dx; a.dgamma = io->d_ln_gamma; a.dwq = io->d_wq; a.dwk = io->d_wk; a.dwv = io->d_wv;
    a.o = static_cast<const __nv_bfloat16 *>(io->o); a.lens = io->lens; a.y = io->y; a.dy = io->dy;
    a.d_o = static_cast<__nv_bfloat16 *>(io->d_o); a.dwo = io->d_wo;
    a.bg = io->gate_bias; a.bm = io->mix_bias;
    a.g = static_cast<__nv_bfloat16 *>(io->gate); a.mix = static_cast<__nv_bfloat16 *>(io->mix);
    a.dg = static_cast<const __nv_bfloat16 *>(io->d_gate); a.dmix = static_cast<const __nv_bfloat16 *>(io->d_mix);
    a.dwg = io->d_wg; a.dbg = io->d_bg; a.dwm = io->d_wm; a.dbm = io->d_bm;
    a.N = b * n; a.n = n;
    return XTRL_OK;
}

template <typename K>
static int launch(K kernel, const char *name, int smem, int ctas_per_sm, const Args &a, void *stream) {
    static const void *seen[32];                    // kernels whose shared-memory limit is already raised
    static int nseen = 0;
    bool known = false;
    for (int i = 0; i < nseen; ++i) known |= seen[i] == (const void *)kernel;
    if (!known) {
        XTRL_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        if (nseen < 32) seen[nseen++] = (const void *)kernel;
    }
    const int ntiles = (a.N + TOK - 1) / TOK, cap = ctas_per_sm * num_sms();
    kernel<<<ntiles < cap ? ntiles : cap, NT, smem, (cudaStream_t)stream>>>(a);
    XTRL_LAUNCH_CHECK(name);
    return XTRL_OK;
}

// the built (heads, gate, mix) combinations; everything else is XTRL_E_UNSUPPORTED
static bool variant_ok(int heads, int gate, int mix) {
    return (heads == 4 || heads == 8) && (gate == 0 || gate == 1) && (mix == 0 || mix == 1) && !(heads == 8 && gate && mix);
}
#define XTRL_ATTN_VARIANT(heads, gate, mix, CALL)                                   \
    do {                                                                             \
        if ((heads) == 4 && !(gate) && !(mix)) { CALL(4, 3, 0); }                    \
        else if ((heads) == 4 && !(gate) && (mix)) { CALL(4, 3, 1); }                \
        else if ((heads) == 4 && (gate) && !(mix)) { CALL(4, 4, 0); }                \
        else if ((heads) == 4 && (gate) && (mix)) { CALL(4, 4, 1); }                 \
        else if ((heads) == 8 && !(gate) && !(mix)) { CALL(8, 3, 0); }               \
        else if ((heads) == 8 && !(gate) && (mix)) { CALL(8, 3, 1); }                \
        else { CALL(8, 4, 0); }                                                      \
    } while (0)

}  // namespace attn
}  // namespace xtrl

static int attn_check(int32_t dim, int32_t heads, int32_t dim_head, int32_t gate, int32_t mix) {
    using namespace xtrl;
    XTRL_REQUIRE(dim == learn::D && dim_head == 16 && attn::variant_ok(heads, gate, mix), XTRL_E_UNSUPPORTED,
                 "xtrl_attn_*: built for dim=48, dim_head=16, heads in {4, 8}, not (8 heads + gate + mix) (got dim=%d dim_head=%d heads=%d gate=%d mix=%d)",
                 dim, dim_head, heads, gate, mix);
    return XTRL_OK;
}

extern "C" int64_t xtrl_attn_img_bytes(int32_t dim, int32_t heads, int32_t dim_head, int32_t gate, int32_t mix) {
    using namespace xtrl::attn;
    if (dim != xtrl::learn::D || dim_head != 16 || !variant_ok(heads, gate, mix)) return -1;
    const int I = 16 * heads, NQ = (gate ? 4 : 3) * I + (mix ? 16 : 0);
    return (int64_t)xtrl::learn::D * I * 2 + (int64_t)NQ * xtrl::learn::D * 2;
}

extern "C" int xtrl_attn_pack(const float *wq, const float *wk, const float *wv, const float *wo, const float *wg, const float *wm,
                              void *img, int32_t dim, int32_t heads, int32_t dim_head, void *stream) {
    using namespace xtrl; using namespace xtrl::attn;
    const int gate = wg != nullptr, mix = wm != nullptr;
    if (int rc = attn_check(dim, heads, dim_head, gate, mix)) return rc;
    XTRL_REQUIRE(wq && wk && wv && wo && img, XTRL_E_BADARG, "xtrl_attn_pack: null pointer");
    unsigned char *p = static_cast<unsigned char *>(img);
#define CALL(H, NB, MIX) attn_pack_kernel<H, NB, MIX><<<(Dims<H, NB, MIX>::NQ * learn::D + 255) / 256, 256, 0, (cudaStream_t)stream>>>(wq, wk, wv, wo, wg, wm, p)
    XTRL_ATTN_VARIANT(heads, gate, mix, CALL);
#undef CALL
    XTRL_LAUNCH_CHECK("attn_pack_kernel");
    return XTRL_OK;
}

extern "C" int xtrl_attn_pack_layers(const float *const *wq, const float *const *wk, const float *const *wv, const float *const *wo,
                                     const float *const *wg, const float *const *wm, void *const *img, int32_t nlayers,
                                     int32_t dim, int32_t heads, int32_t dim_head, void *stream) {
    using namespace xtrl; using namespace xtrl::attn;
    const int gate = wg != nullptr, mix = wm != nullptr;
    if (int rc = attn_check(dim, heads, dim_head, gate, mix)) return rc;
    XTRL_REQUIRE(wq && wk && wv && wo && img && nlayers >= 1 && nlayers <= kPackLayers, XTRL_E_BADARG,
                 "xtrl_attn_pack_layers: null pointer or nlayers not in 1..%d (got %d)", kPackLayers, nlayers);
    PackN a;
    for (int l = 0; l < nlayers; ++l) {
        XTRL_REQUIRE(wq[l] && wk[l] && wv[l] && wo[l] && img[l] && (!gate || wg[l]) && (!mix || wm[l]), XTRL_E_BADARG,
                     "xtrl_attn_pack_layers: null pointer (layer %d)", l);
        a.wq[l] = wq[l]; a.wk[l] = wk[l]; a.wv[l] = wv[l]; a.wo[l] = wo[l]; a.wg[l] = gate ? wg[l] : nullptr; a.wm[l] = mix ? wm[l] : nullptr;
        a.img[l] = static_cast<unsigned char *>(img[l]);
    }
#define CALL(H, NB, MIX) attn_pack_n_kernel<H, NB, MIX><<<dim3((Dims<H, NB, MIX>::NQ * learn::D + 255) / 256, nlayers), 256, 0, (cudaStream_t)stream>>>(a)
    XTRL_ATTN_VARIANT(heads, gate, mix, CALL);
#undef CALL
    XTRL_LAUNCH_CHECK("attn_pack_n_kernel");
    return XTRL_OK;
}

extern "C" int xtrl_attn_pre_fwd(const xtrl_attn_io *io, int32_t b, int32_t n, int32_t heads, void *stream) {
    using namespace xtrl; using namespace xtrl::attn;
    XTRL_REQUIRE(io && b >= 0 && n >= 0, XTRL_E_BADARG, "xtrl_attn_pre_fwd: bad arguments");
    const int gate = io->gate != nullptr, mix = io->mix != nullptr;
    if (int rc = attn_check(learn::D, heads, 16, gate, mix)) return rc;
    if ((int64_t)b * n == 0) return XTRL_OK;
    XTRL_REQUIRE(io->x && io->ln_gamma && io->cos_tab && io->sin_tab && io->w_img && io->q && io->k && io->v && (!gate || io->gate_bias) &&
                 (!mix || io->mix_bias), XTRL_E_BADARG, "xtrl_attn_pre_fwd: null pointer");
    XTRL_REQUIRE(((uintptr_t)io->x & 15) == 0, XTRL_E_BADARG, "xtrl_attn_pre_fwd: x must be 16-byte aligned (row tiles move by bulk copy)");
    Args a;
    if (int rc = fill(io, b, n, a)) return rc;
#define CALL(H, NB, MIX) return launch(attn_pre_fwd_kernel<H, NB, MIX>, "attn_pre_fwd_kernel", Smem<H, NB, MIX>::pre_fwd, 1, a, stream)
    XTRL_ATTN_VARIANT(heads, gate, mix, CALL);
#undef CALL
}

extern "C" int xtrl_attn_pre_bwd(const xtrl_attn_io *io, int32_t b, int32_t n, int32_t heads, void *stream) {
    using namespace xtrl; using namespace xtrl::attn;
    XTRL_REQUIRE(io && b >= 0 && n >= 0, XTRL_E_BADARG, "xtrl_attn_pre_bwd: bad arguments");
    const int gate = io->d_gate != nullptr, mix = io->d_mix != nullptr;
    if (int rc = attn_check(learn::D, heads, 16, gate, mix)) return rc;
    if ((int64_t)b * n == 0) return XTRL_OK;
    XTRL_REQUIRE(io->x && io->ln_gamma && io->cos_tab && io->sin_tab && io->w_img && io->dq && io->dk && io->dv && io->dx &&
                 io->d_ln_gamma && io->d_wq && io->d_wk && io->d_wv && (!gate || (io->d_wg && io->d_bg)) && (!mix || (io->d_wm && io->d_bm)),
                 XTRL_E_BADARG, "xtrl_attn_pre_bwd: null pointer");
    XTRL_REQUIRE((((uintptr_t)io->x | (uintptr_t)io->dx | (uintptr_t)io->dy) & 15) == 0, XTRL_E_BADARG,
                 "xtrl_attn_pre_bwd: x, dx and dy must be 16-byte aligned (row tiles move by bulk copy)");
    Args a;
    if (int rc = fill(io, b, n, a)) return rc;
#define CALL(H, NB, MIX) return launch(attn_pre_bwd_kernel<H, NB, MIX>, "attn_pre_bwd_kernel", Smem<H, NB, MIX>::pre_bwd, 1, a, stream)
    XTRL_ATTN_VARIANT(heads, gate, mix, CALL);
#undef CALL
}

extern "C" int xtrl_attn_post_fwd(const xtrl_attn_io *io, int32_t b, int32_t n, int32_t heads, void *stream) {
    using namespace xtrl; using namespace xtrl::attn;
    if (int rc = attn_check(learn::D, heads, 16, 0, 0)) return rc;
    XTRL_REQUIRE(io && b >= 0 && n >= 0, XTRL_E_BADARG, "xtrl_attn_post_fwd: bad arguments");
    if ((int64_t)b * n == 0) return XTRL_OK;
    XTRL_REQUIRE(io->x && io->o && io->lens && io->w_img && io->y, XTRL_E_BADARG, "xtrl_attn_post_fwd: null pointer");
    Args a;
    if (int rc = fill(io, b, n, a)) return rc;
    return heads == 4 ? launch(attn_post_fwd_kernel<4>, "attn_post_fwd_kernel", Smem<4, 3, 0>::post_fwd, 2, a, stream)
                      : launch(attn_post_fwd_kernel<8>, "attn_post_fwd_kernel", Smem<8, 3, 0>::post_fwd, 2, a, stream);
}

extern "C" int xtrl_attn_post_bwd(const xtrl_attn_io *io, int32_t b, int32_t n, int32_t heads, void *stream) {
    using namespace xtrl; using namespace xtrl::attn;
    if (int rc = attn_check(learn::D, heads, 16, 0, 0)) return rc;
    XTRL_REQUIRE(io && b >= 0 && n >= 0, XTRL_E_BADARG, "xtrl_attn_post_bwd: bad arguments");
    if ((int64_t)b * n == 0) return XTRL_OK;
    XTRL_REQUIRE(io->dy && io->o && io->lens && io->w_img && io->d_o && io->d_wo, XTRL_E_BADARG, "xtrl_attn_post_bwd: null pointer");
    Args a;
    if (int rc = fill(io, b, n, a)) return rc;
    return heads == 4 ? launch(attn_post_bwd_kernel<4>, "attn_post_bwd_kernel", Smem<4, 3, 0>::post_bwd, 2, a, stream)
                      : launch(attn_post_bwd_kernel<8>, "attn_post_bwd_kernel", Smem<8, 3, 0>::post_bwd, 2, a, stream);
}
