// learn_mlp.cu -- the two-layer head MLPs of the learn-side model, forward and backward, on tcgen05:
//
//     y = W2 * act(W1 * [src0 | src1 | (src2)] + b1) + b2,    act = SiLU on the first n_silu hidden units, identity on the rest
//
// Replaces, inside Agent.learn (x_transformers_rl.py:348-369 actor / critic heads, :322-331 to_pred / to_pred_done,
// called from :516-549): torch.cat of the 48-wide inputs, Linear, SiLU, Linear and their backward (bias-gradient
// reductions included).  One kernel serves the critic head (K_in 96|144 -> 96 -> 100 bins), the actor head (-> A or 2A)
// and the world-model head with the done logit riding along as one linear hidden unit (-> 2(S+1) + 1): the caller hands
// over zero-padded dense weights (hid_p x in_dim, out_p x hid_p) as operand images.  SURVEY 8f row 1 (heads).
//
// Same machinery as learn_ff.cu: 128-token tiles, TMEM lane = token, thread (row, half) owns half the columns; the
// backward recomputes the hidden layer, and dW1|db1 = dPRE^T [X|1], dW2|db2 = dY^T [H|1] accumulate in TMEM across
// the CTA's tiles from the token-major images re-read MN-major; one atomicAdd pass per CTA into dense padded buffers.
#include "learn_common.cuh"

namespace xtrl {
namespace mlp {
using namespace umma;
using namespace learn;

struct Args {
    const float *src[3];
    int per_episode2;                                     // src[2] is [episodes][48], row = token / n
    const float *b1, *b2;
    const unsigned char *wimg;                            // [W1 image (rows hid_p, K in_dim) | W2 image (rows out_p, K hid_p)]
    __nv_bfloat16 *y;                                     // [N][out_cols]
    const __nv_bfloat16 *gy;
    float *dsrc[3];                                       // [N][48] each
    float dscale[3];                                      // d_src[s] = (dacc bit s ? d_src[s] : 0) + dscale[s] * dX_s
    int dacc;
    float *dW1, *db1, *dW2, *db2;                         // dense padded, +=
    int N, n, n_silu, out_cols;
};

__device__ __forceinline__ float sigmoidf(float x) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * x));
    return __frcp_rn(1.f + e);
}

template <int KIN, int HID, int NOUT> struct Dims {
    static constexpr int NSRC = KIN / D;
    static constexpr int XW = KIN + 16, HW = HID + 16;    // image widths incl. the ones column (bias gradients) + pad
    static constexpr int kW1 = HID * KIN * 2, kW2 = NOUT * HID * 2;
    static constexpr int kX = TOK * XW * 2, kH = TOK * HW * 2, kDP = TOK * HID * 2, kDY = TOK * NOUT * 2;
    static constexpr int fwd_smem = kW1 + kW2 + TOK * KIN * 2 + TOK * HID * 2 + (HID + NOUT) * 4 + 64;
    // backward order: dY, dPRE first -- their MN-major reads run past the image (M = 128 > NOUT / HID) into what follows
    static constexpr int bwd_base = kDY + kDP + kH + kX + kW1 + kW2 + HID * 4 + 64;
    // two sources: their dX tiles leave (and, when accumulating, first arrive) by bulk copy through two row-tile buffers
    static constexpr bool ost = NSRC == 2 && bwd_base + 2 * kRowTile <= 227 * 1024;
    static constexpr int bwd_smem = bwd_base + (ost ? 2 * kRowTile : 0);
    static_assert(kDP + kH + kX >= (128 - NOUT) * kCore / 8 && kH + kX >= (128 - HID) * kCore / 8, "over-read stays inside the CTA's shared memory");
};

// this thread's 24 columns of every 48-wide source -> bf16 image chunks src * 6 + half * 3 + c
template <int NSRC>
__device__ __forceinline__ void build_x(unsigned char *sX, const Args &a, int t, bool ok, int row, int half) {
#pragma unroll
    for (int s = 0; s < NSRC; ++s) {
        const size_t r = (s == 2 && a.per_episode2) ? (size_t)(t / a.n) : (size_t)t;
        const float4 *p = reinterpret_cast<const float4 *>(a.src[s] + r * D + half * 24);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 lo = ok ? __ldg(p + 2 * c) : z, hi = ok ? __ldg(p + 2 * c + 1) : z;
            const float v[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
            put8(sX, row, s * 6 + half * 3 + c, v);
        }
    }
}

// ------------------------------------------------------------------------------------------ forward
template <int KIN, int HID, int NOUT>
__global__ void __launch_bounds__(NT, 2) mlp_fwd_kernel(const Args a) {
    using P = Dims<KIN, HID, NOUT>;
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *sW1 = smem, *sW2 = sW1 + P::kW1, *sX = sW2 + P::kW2, *sH = sX + TOK * KIN * 2;
    float *sB1 = reinterpret_cast<float *>(sH + TOK * HID * 2), *sB2 = sB1 + HID;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sB2 + NOUT);
    uint32_t *tslot = reinterpret_cast<uint32_t *>(bars + 2);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = warp & 3, half = warp >> 2, row = q * 32 + lane;
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(tslot, 256);
    for (int i = tid; i < HID; i += NT) sB1[i] = a.b1[i];
    for (int i = tid; i < NOUT; i += NT) sB2[i] = a.b2[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tslot, tlane = tbase + ((uint32_t)(q * 32) << 16);
    if (tid == 0) { mbar_expect_tx(&bars[0], P::kW1 + P::kW2); bulk_g2s(sW1, a.wimg, P::kW1 + P::kW2, &bars[0]); }
    uint32_t ph = 0;
    bool first = true;
    const int ntiles = (a.N + TOK - 1) / TOK;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int t = tile * TOK + row;
        const bool ok = t < a.N;
        if (tile + (int)gridDim.x < ntiles) {                                    // next tile's rows -> L2
            const int t0 = (tile + gridDim.x) * TOK, nr = min(TOK, a.N - t0);
            prefetch_rows48(a.src[0], t0, nr, tid, NT);
            prefetch_rows48(a.src[1], t0, nr, tid, NT);
            if (P::NSRC == 3 && !a.per_episode2) prefetch_rows48(a.src[2], t0, nr, tid, NT);
            if (a.gy) { const char *g = reinterpret_cast<const char *>(a.gy + (size_t)t0 * a.out_cols); for (int o = tid * 128; o < nr * a.out_cols * 2; o += NT * 128) prefetch_l2(g + o); }
        }
        build_x<P::NSRC>(sX, a, t, ok, row, half);
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            if (first) mbar_wait(&bars[0], 0);
            tc_fence_after();
#pragma unroll
            for (int s = 0; s < KIN / 16; ++s)
                mma_bf16_ss(tbase, smem_desc(sX + s * 2 * kCore, kCore, 128), smem_desc(sW1 + s * 2 * HID * 16, HID * 16, 128), idesc_bf16(128, HID), s > 0);
            mma_commit(&bars[1]);
        }
        first = false;
        mbar_wait(&bars[1], ph); ph ^= 1;
        tc_fence_after();
#pragma unroll 2
        for (int c = 0; c < HID / 16; ++c) {                                  // 8-column pieces: this thread takes every other one
            float v[8];
            const int col0 = c * 16 + half * 8;
            tmem_ld8(tlane + col0, v);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float p = v[j] + sB1[col0 + j];
                v[j] = col0 + j < a.n_silu ? p * sigmoidf(p) : p;
            }
            put8(sH, row, col0 / 8, v);
        }
        tc_fence_before();
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
#pragma unroll
            for (int s = 0; s < HID / 16; ++s)
                mma_bf16_ss(tbase + HID, smem_desc(sH + s * 2 * kCore, kCore, 128), smem_desc(sW2 + s * 2 * NOUT * 16, NOUT * 16, 128), idesc_bf16(128, NOUT), s > 0);
            mma_commit(&bars[1]);
        }
        mbar_wait(&bars[1], ph); ph ^= 1;
        tc_fence_after();
#pragma unroll 2
        for (int c = 0; c < NOUT / 16; ++c) {
            float v[8];
            const int col0 = c * 16 + half * 8;
            tmem_ld8(tlane + HID + col0, v);
            if (ok && (a.out_cols & 7) == 0 && col0 + 8 <= a.out_cols) {      // 16-byte aligned piece: one store
                const float *bb = sB2 + col0;
                *reinterpret_cast<uint4 *>(a.y + (size_t)t * a.out_cols + col0) =
                    make_uint4(pack2(v[0] + bb[0], v[1] + bb[1]), pack2(v[2] + bb[2], v[3] + bb[3]), pack2(v[4] + bb[4], v[5] + bb[5]), pack2(v[6] + bb[6], v[7] + bb[7]));
            } else if (ok) {
                __nv_bfloat16 *dst = a.y + (size_t)t * a.out_cols + col0;
#pragma unroll
                for (int j = 0; j < 8; j += 2) {
                    const float o0 = v[j] + sB2[col0 + j], o1 = v[j + 1] + sB2[col0 + j + 1];
                    if (col0 + j + 1 < a.out_cols && (a.out_cols & 1) == 0) *reinterpret_cast<uint32_t *>(dst + j) = pack2(o0, o1);
                    else {
                        if (col0 + j < a.out_cols) dst[j] = __float2bfloat16(o0);
                        if (col0 + j + 1 < a.out_cols) dst[j + 1] = __float2bfloat16(o1);
                    }
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 256);
}

// ----------------------------------------------------------------------------------------- backward
template <int KIN, int HID, int NOUT>
__global__ void __launch_bounds__(NT, 1) mlp_bwd_kernel(const Args a) {
    using P = Dims<KIN, HID, NOUT>;
    constexpr int XW = P::XW, HW = P::HW;
    constexpr uint32_t cPre = 0, cDx = HID, cA = HID + KIN, cB = HID + KIN + XW;      // PRE/dH, dX, dW1|db1, dW2|db2
    static_assert(cB + HW <= 512, "TMEM budget");
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *sDY = smem, *sDP = sDY + P::kDY, *sH = sDP + P::kDP, *sX = sH + P::kH, *sW1 = sX + P::kX, *sW2 = sW1 + P::kW1;
    constexpr bool OST = P::ost;
    float *sOut = reinterpret_cast<float *>(sW2 + P::kW2);                     // [2 sources][128][48] fp32 (OST)
    float *sB1 = sOut + (OST ? 2 * TOK * D : 0);
    uint64_t *bars = reinterpret_cast<uint64_t *>(sB1 + HID);
    uint32_t *tslot = reinterpret_cast<uint32_t *>(bars + 3);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = warp & 3, half = warp >> 2, row = q * 32 + lane;
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1); fence_mbar_init(); }   // 2: d_src tiles to add to
    if (warp == 0) tmem_alloc(tslot, 512);
    for (int i = tid; i < HID; i += NT) sB1[i] = a.b1[i];
    if (half == 0) {                                                          // constant columns: [1, 0 x 7] then zeros
        const float one[8] = {1.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, zero[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        put8(sX, row, KIN / 8, one); put8(sX, row, KIN / 8 + 1, zero);
        put8(sH, row, HID / 8, one); put8(sH, row, HID / 8 + 1, zero);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tslot, tlane = tbase + ((uint32_t)(q * 32) << 16);
    if (tid == 0) { mbar_expect_tx(&bars[0], P::kW1 + P::kW2); bulk_g2s(sW1, a.wimg, P::kW1 + P::kW2, &bars[0]); }
    uint32_t ph = 0, pha = 0;
    bool first = true;
    int pend = -1;                                                            // tile whose dX still sits in sOut
    const int ntiles = (a.N + TOK - 1) / TOK;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int t = tile * TOK + row;
        const bool ok = t < a.N;
        if (tile + (int)gridDim.x < ntiles) {                                    // next tile's rows -> L2
            const int t0 = (tile + gridDim.x) * TOK, nr = min(TOK, a.N - t0);
            prefetch_rows48(a.src[0], t0, nr, tid, NT);
            prefetch_rows48(a.src[1], t0, nr, tid, NT);
            if (P::NSRC == 3 && !a.per_episode2) prefetch_rows48(a.src[2], t0, nr, tid, NT);
            if (a.gy) { const char *g = reinterpret_cast<const char *>(a.gy + (size_t)t0 * a.out_cols); for (int o = tid * 128; o < nr * a.out_cols * 2; o += NT * 128) prefetch_l2(g + o); }
        }
        build_x<P::NSRC>(sX, a, t, ok, row, half);
        {                                                                      // dY image (zero beyond out_cols and beyond N)
            // all of the row's pieces are requested before the first is used: one memory round trip, not NOUT / 16 of them
            const bool fast = (a.out_cols & 3) == 0;                          // rows are 8-byte aligned: two 8-byte loads per piece
            uint2 lo[NOUT / 16], hi[NOUT / 16];
#pragma unroll
            for (int c = 0; c < NOUT / 16; ++c) {
                const int col0 = c * 16 + half * 8;
                const __nv_bfloat16 *src = a.gy + (size_t)t * a.out_cols + col0;
                lo[c] = hi[c] = make_uint2(0u, 0u);
                if (ok && fast && col0 + 8 <= a.out_cols) { lo[c] = __ldg(reinterpret_cast<const uint2 *>(src)); hi[c] = __ldg(reinterpret_cast<const uint2 *>(src) + 1); }
            }
#pragma unroll
            for (int c = 0; c < NOUT / 16; ++c) {
                const int col0 = c * 16 + half * 8;
                float v[8];
                if (fast && (col0 + 8 <= a.out_cols || !ok || col0 >= a.out_cols)) {
                    unpack8(make_uint4(lo[c].x, lo[c].y, hi[c].x, hi[c].y), v);
                } else {
                    const __nv_bfloat16 *src = a.gy + (size_t)t * a.out_cols + col0;
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = (ok && col0 + j < a.out_cols) ? __bfloat162float(src[j]) : 0.f;
                }
                put8(sDY, row, col0 / 8, v);
            }
        }
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {                                                       // PRE = X W1^T
            if (OST && pend >= 0) { rows_store(a.dsrc[0], sOut, pend, a.N); rows_store(a.dsrc[1], sOut + TOK * D, pend, a.N); }
            if (first) mbar_wait(&bars[0], 0);
            tc_fence_after();
#pragma unroll
            for (int s = 0; s < KIN / 16; ++s)
                mma_bf16_ss(tbase + cPre, smem_desc(sX + s * 2 * kCore, kCore, 128), smem_desc(sW1 + s * 2 * HID * 16, HID * 16, 128), idesc_bf16(128, HID), s > 0);
            mma_commit(&bars[1]);
        }
        mbar_wait(&bars[1], ph); ph ^= 1;
        tc_fence_after();
#pragma unroll 2
        for (int c = 0; c < HID / 16; ++c) {                                  // H and act' (parked in dPRE's slot)
            const int col0 = c * 16 + half * 8;
            float v[8], d[8];
            tmem_ld8(tlane + cPre + col0, v);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float p = v[j] + sB1[col0 + j], sg = sigmoidf(p);
                const bool act = col0 + j < a.n_silu;
                v[j] = act ? p * sg : p;
                d[j] = act ? sg * fmaf(p, 1.f - sg, 1.f) : 1.f;
            }
            put8(sH, row, col0 / 8, v);
            put8(sDP, row, col0 / 8, d);
        }
        tc_fence_before();
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            if (OST) {
                if (pend >= 0) bulk_wait_read();                              // the staging buffers are free again
                if (a.dacc) {                                                 // the rows this tile adds to, into the same buffers
                    const int t0 = tile * TOK, rows = min(TOK, a.N - t0);
                    const uint32_t bytes = (uint32_t)rows * D * 4;
                    mbar_expect_tx(&bars[2], ((a.dacc & 1) ? bytes : 0u) + ((a.dacc & 2) ? bytes : 0u));
                    if (a.dacc & 1) bulk_g2s(sOut, a.dsrc[0] + (size_t)t0 * D, bytes, &bars[2]);
                    if (a.dacc & 2) bulk_g2s(sOut + TOK * D, a.dsrc[1] + (size_t)t0 * D, bytes, &bars[2]);
                }
            }
            tc_fence_after();
#pragma unroll
            for (int s = 0; s < NOUT / 16; ++s)                                // dH = dY W2 (W2 image read MN-major)
                mma_bf16_ss(tbase + cPre, smem_desc(sDY + s * 2 * kCore, kCore, 128), smem_desc(sW2 + s * 256, 128, NOUT * 16),
                            idesc_bf16(128, HID) | kIdB, s > 0);
#pragma unroll
            for (int s = 0; s < TOK / 16; ++s)                                 // dW2 | db2 += dY^T [H | 1]
                mma_bf16_ss(tbase + cB, smem_desc(sDY + s * 256, 128, kCore), smem_desc(sH + s * 256, 128, kCore),
                            idesc_bf16(128, HW) | kIdA | kIdB, !first || s > 0);
            mma_commit(&bars[1]);
        }
        mbar_wait(&bars[1], ph); ph ^= 1;
        tc_fence_after();
#pragma unroll 2
        for (int c = 0; c < HID / 16; ++c) {                                  // dPRE = dH . act'
            const int col0 = c * 16 + half * 8;
            float v[8], d[8];
            tmem_ld8(tlane + cPre + col0, v);
            unpack8(*reinterpret_cast<const uint4 *>(sDP + (col0 / 8) * kCore + (row >> 3) * 128 + (row & 7) * 16), d);
#pragma unroll
            for (int j = 0; j < 8; ++j) d[j] *= v[j];
            put8(sDP, row, col0 / 8, d);
        }
        tc_fence_before();
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
#pragma unroll
            for (int s = 0; s < HID / 16; ++s)                                 // dX = dPRE W1 (W1 image read MN-major)
                mma_bf16_ss(tbase + cDx, smem_desc(sDP + s * 2 * kCore, kCore, 128), smem_desc(sW1 + s * 256, 128, HID * 16),
                            idesc_bf16(128, KIN) | kIdB, s > 0);
#pragma unroll
            for (int s = 0; s < TOK / 16; ++s)                                 // dW1 | db1 += dPRE^T [X | 1]
                mma_bf16_ss(tbase + cA, smem_desc(sDP + s * 256, 128, kCore), smem_desc(sX + s * 256, 128, kCore),
                            idesc_bf16(128, XW) | kIdA | kIdB, !first || s > 0);
            mma_commit(&bars[1]);
        }
        first = false;
        mbar_wait(&bars[1], ph); ph ^= 1;
        tc_fence_after();
        if (OST && a.dacc) { mbar_wait(&bars[2], pha); pha ^= 1; }
#pragma unroll
        for (int c = 0; c < KIN / 16; ++c) {                                  // dX, 8 columns at a time, to the source it belongs to
            const int col0 = c * 16 + half * 8, s = col0 / D, off = col0 - s * D;
            float v[8];
            tmem_ld8(tlane + cDx + col0, v);
            if (OST || ok) {
                float4 *dst = reinterpret_cast<float4 *>(OST ? sOut + s * TOK * D + row * D + off : a.dsrc[s] + (size_t)t * D + off);
                const float sc = a.dscale[s];
                float4 o0 = make_float4(0.f, 0.f, 0.f, 0.f), o1 = o0;
                if ((a.dacc >> s) & 1) { o0 = dst[0]; o1 = dst[1]; }
                dst[0] = make_float4(fmaf(sc, v[0], o0.x), fmaf(sc, v[1], o0.y), fmaf(sc, v[2], o0.z), fmaf(sc, v[3], o0.w));
                dst[1] = make_float4(fmaf(sc, v[4], o1.x), fmaf(sc, v[5], o1.y), fmaf(sc, v[6], o1.z), fmaf(sc, v[7], o1.w));
            }
        }
        pend = tile;
        tc_fence_before();
    }
    if (OST && pend >= 0) {                                                   // the last tile's dX
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) { rows_store(a.dsrc[0], sOut, pend, a.N); rows_store(a.dsrc[1], sOut + TOK * D, pend, a.N); }
    }
    if (!first) {                                                             // flush the weight-gradient accumulators (see flush_rows)
        tc_fence_after();
        float *stage = reinterpret_cast<float *>(smem);                       // over the dY / dPRE / H images, which are done with
        static_assert((HID > NOUT ? HID : NOUT) * ((KIN > HID ? KIN : HID) + 1) * 4 <= P::kDY + P::kDP + P::kH, "stage fits the tile images");
        __syncthreads();
#pragma unroll 1
        for (int c = 0; c < XW / 16; ++c) {
            const int col0 = c * 16 + half * 8;
            float v[8];
            tmem_ld8(tlane + cA + col0, v);
            if (row < HID) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (col0 + j < KIN) stage[row * (KIN + 1) + col0 + j] = v[j];
                    else if (col0 + j == KIN) atomicAdd(a.db1 + row, v[j]);
                }
            }
        }
        __syncthreads();
        flush_rows(stage, KIN + 1, HID, KIN, tid, NT, [&](int r, int c) { return a.dW1 + r * KIN + c; });
        __syncthreads();
#pragma unroll 1
        for (int c = 0; c < HW / 16; ++c) {
            const int col0 = c * 16 + half * 8;
            float v[8];
            tmem_ld8(tlane + cB + col0, v);
            if (row < NOUT) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (col0 + j < HID) stage[row * (HID + 1) + col0 + j] = v[j];
                    else if (col0 + j == HID) atomicAdd(a.db2 + row, v[j]);
                }
            }
        }
        __syncthreads();
        flush_rows(stage, HID + 1, NOUT, HID, tid, NT, [&](int r, int c) { return a.dW2 + r * HID + c; });
        tc_fence_before();
    }
    if (OST && tid == 0) bulk_wait_all();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

template <int KIN, int HID, int NOUT>
__global__ void mlp_pack_kernel(const float *__restrict__ w1, const float *__restrict__ w2, unsigned char *__restrict__ img) {
    using P = Dims<KIN, HID, NOUT>;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < HID * KIN) *reinterpret_cast<__nv_bfloat16 *>(img + img_off_bf16(HID, i / KIN, i % KIN)) = __float2bfloat16(w1[i]);
    if (i < NOUT * HID) *reinterpret_cast<__nv_bfloat16 *>(img + P::kW1 + img_off_bf16(NOUT, i / HID, i % HID)) = __float2bfloat16(w2[i]);
}

static bool variant_ok(int in_dim, int hid, int out) {
    return (in_dim == 96 || in_dim == 144) && ((hid == 96 && (out == 112 || out == 16)) || (in_dim == 96 && hid == 64 && out == 32));
}
#define XTRL_MLP_VARIANT(in_dim, hid, out, CALL)                         \
    do {                                                                  \
        if ((in_dim) == 96 && (hid) == 96 && (out) == 112) { CALL(96, 96, 112); }   \
        else if ((in_dim) == 96 && (hid) == 96) { CALL(96, 96, 16); }     \
        else if ((in_dim) == 96) { CALL(96, 64, 32); }                    \
        else if ((out) == 112) { CALL(144, 96, 112); }                    \
        else { CALL(144, 96, 16); }                                       \
    } while (0)

template <typename K>
static int launch(K kernel, const char *name, int smem, int ctas_per_sm, const Args &a, void *stream) {
    static const void *seen[32];
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

static int fill(const xtrl_mlp_io *io, int64_t N, int32_t n, int32_t in_dim, int32_t hid, int32_t out, Args &a) {
    XTRL_REQUIRE(variant_ok(in_dim, hid, out), XTRL_E_UNSUPPORTED, "xtrl_mlp_*: (in_dim, hidden, out) = (%d, %d, %d) is not built", in_dim, hid, out);
    XTRL_REQUIRE(N >= 0 && N <= (int64_t)1 << 24 && n >= 1, XTRL_E_BADARG, "xtrl_mlp_*: bad token count");
    XTRL_REQUIRE(io->out_cols >= 1 && io->out_cols <= out && io->n_silu >= 0 && io->n_silu <= hid, XTRL_E_BADARG, "xtrl_mlp_*: out_cols / n_silu out of range");
    a.src[0] = io->src0; a.src[1] = io->src1; a.src[2] = io->src2; a.per_episode2 = io->src2_per_episode;
    a.b1 = io->b1; a.b2 = io->b2; a.wimg = static_cast<const unsigned char *>(io->w_img);
    a.y = static_cast<__nv_bfloat16 *>(io->y); a.gy = static_cast<const __nv_bfloat16 *>(io->gy);
    a.dsrc[0] = io->d_src0; a.dsrc[1] = io->d_src1; a.dsrc[2] = io->d_src2;
    a.dacc = io->d_src_accumulate & 7;
    for (int i = 0; i < 3; ++i) a.dscale[i] = io->d_src_scale[i] != 0.f ? io->d_src_scale[i] : 1.f;
    a.dW1 = io->d_w1; a.db1 = io->d_b1; a.dW2 = io->d_w2; a.db2 = io->d_b2;
    a.N = (int)N; a.n = n; a.n_silu = io->n_silu; a.out_cols = io->out_cols;
    return XTRL_OK;
}

}  // namespace mlp
}  // namespace xtrl

// ---- dense padded copies of the head parameters <-> the live parameters, one launch per head ----
// The head kernels work on zero-padded dense weights (hidden / out rounded up, several nn.Linear layers stacked as row blocks).
// refresh = copy every parameter block into its place (mode 0); scatter = add every block of the padded gradient scratch to the
// parameter's gradient (mode 1).  Up to kMaxSegs rectangular blocks per launch, passed by value.
namespace xtrl {
constexpr int kMaxSegs = 8;
struct SegArgs { xtrl_seg2d seg[kMaxSegs]; int n, mode; };
__global__ void seg2d_kernel(const SegArgs a) {
    const xtrl_seg2d g = a.seg[blockIdx.y];
    const int total = g.rows * g.cols;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int r = i / g.cols, c = i - r * g.cols;
        const float v = g.src[(size_t)r * g.src_pitch + c];
        float *d = g.dst + (size_t)r * g.dst_pitch + c;
        *d = a.mode ? *d + v : v;
    }
}
}  // namespace xtrl

extern "C" int xtrl_seg2d_copy(const xtrl_seg2d *segs, int32_t nseg, int32_t accumulate, void *stream) {
    using namespace xtrl;
    XTRL_REQUIRE(segs && nseg >= 1 && nseg <= kMaxSegs, XTRL_E_BADARG, "xtrl_seg2d_copy: 1..%d segments (got %d)", kMaxSegs, nseg);
    SegArgs a;
    int most = 0;
    for (int i = 0; i < nseg; ++i) {
        XTRL_REQUIRE(segs[i].src && segs[i].dst && segs[i].rows >= 0 && segs[i].cols >= 0 && segs[i].src_pitch >= segs[i].cols &&
                     segs[i].dst_pitch >= segs[i].cols, XTRL_E_BADARG, "xtrl_seg2d_copy: bad segment %d", i);
        a.seg[i] = segs[i];
        const int t = segs[i].rows * segs[i].cols;
        most = t > most ? t : most;
    }
    if (most == 0) return XTRL_OK;
    a.n = nseg; a.mode = accumulate != 0;
    const int bx = (most + 255) / 256;
    seg2d_kernel<<<dim3(bx < 64 ? bx : 64, nseg), 256, 0, (cudaStream_t)stream>>>(a);
    XTRL_LAUNCH_CHECK("seg2d_kernel");
    return XTRL_OK;
}

extern "C" int64_t xtrl_mlp_img_bytes(int32_t in_dim, int32_t hidden, int32_t out) {
    return xtrl::mlp::variant_ok(in_dim, hidden, out) ? (int64_t)hidden * in_dim * 2 + (int64_t)out * hidden * 2 : -1;
}

extern "C" int xtrl_mlp_pack(const float *w1, const float *w2, void *img, int32_t in_dim, int32_t hidden, int32_t out, void *stream) {
    using namespace xtrl; using namespace xtrl::mlp;
    XTRL_REQUIRE(variant_ok(in_dim, hidden, out), XTRL_E_UNSUPPORTED, "xtrl_mlp_pack: (%d, %d, %d) is not built", in_dim, hidden, out);
    XTRL_REQUIRE(w1 && w2 && img, XTRL_E_BADARG, "xtrl_mlp_pack: null pointer");
    unsigned char *p = static_cast<unsigned char *>(img);
    const int nmax = hidden * in_dim > out * hidden ? hidden * in_dim : out * hidden;
#define CALL(KIN, HID, NOUT) mlp_pack_kernel<KIN, HID, NOUT><<<(nmax + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w1, w2, p)
    XTRL_MLP_VARIANT(in_dim, hidden, out, CALL);
#undef CALL
    XTRL_LAUNCH_CHECK("mlp_pack_kernel");
    return XTRL_OK;
}

extern "C" int xtrl_mlp_fwd(const xtrl_mlp_io *io, int64_t N, int32_t n, int32_t in_dim, int32_t hidden, int32_t out, void *stream) {
    using namespace xtrl; using namespace xtrl::mlp;
    XTRL_REQUIRE(io, XTRL_E_BADARG, "xtrl_mlp_fwd: null io");
    Args a;
    if (int rc = fill(io, N, n, in_dim, hidden, out, a)) return rc;
    if (N == 0) return XTRL_OK;
    XTRL_REQUIRE(io->src0 && io->src1 && (in_dim == 96 || io->src2) && io->b1 && io->b2 && io->w_img && io->y, XTRL_E_BADARG, "xtrl_mlp_fwd: null pointer");
#define CALL(KIN, HID, NOUT) return launch(mlp_fwd_kernel<KIN, HID, NOUT>, "mlp_fwd_kernel", Dims<KIN, HID, NOUT>::fwd_smem, 2, a, stream)
    XTRL_MLP_VARIANT(in_dim, hidden, out, CALL);
#undef CALL
}

extern "C" int xtrl_mlp_bwd(const xtrl_mlp_io *io, int64_t N, int32_t n, int32_t in_dim, int32_t hidden, int32_t out, void *stream) {
    using namespace xtrl; using namespace xtrl::mlp;
    XTRL_REQUIRE(io, XTRL_E_BADARG, "xtrl_mlp_bwd: null io");
    Args a;
    if (int rc = fill(io, N, n, in_dim, hidden, out, a)) return rc;
    if (N == 0) return XTRL_OK;
    XTRL_REQUIRE(io->src0 && io->src1 && (in_dim == 96 || (io->src2 && io->d_src2)) && io->b1 && io->w_img && io->gy && io->d_src0 && io->d_src1 &&
                 io->d_w1 && io->d_b1 && io->d_w2 && io->d_b2, XTRL_E_BADARG, "xtrl_mlp_bwd: null pointer");
    XTRL_REQUIRE((((uintptr_t)io->d_src0 | (uintptr_t)io->d_src1) & 15) == 0, XTRL_E_BADARG,
                 "xtrl_mlp_bwd: d_src0 / d_src1 must be 16-byte aligned (row tiles move by bulk copy)");
#define CALL(KIN, HID, NOUT) return launch(mlp_bwd_kernel<KIN, HID, NOUT>, "mlp_bwd_kernel", Dims<KIN, HID, NOUT>::bwd_smem, 1, a, stream)
    XTRL_MLP_VARIANT(in_dim, hidden, out, CALL);
#undef CALL
}
