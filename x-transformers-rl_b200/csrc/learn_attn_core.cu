// learn_attn_core.cu -- learn-side causal attention core for dim_head = 16, sequences up to 512, forward and
// backward, on tcgen05:      o = dropout(softmax(scale * q k^T + causal)) v       per (episode, head).
//
// Replaces the attention core of the teacher-forced pass inside Agent.learn (x_transformers_rl.py:928-935;
// x-transformers Attention, SURVEY App. A.1: causal softmax(q k^T * dh^-1/2), attention dropout, @ v) and its
// backward.  SURVEY 8f row 1.  With dh = 16 the matmuls are tiny (one K-step per 128 x 128 score tile); the
// work is the per-score softmax arithmetic, so the design keeps that at ~10 (fwd) / ~20 (bwd) instructions per
// score and lets the tensor core do everything else:
//   * one CTA per (episode, head); q, k, v (and dO) live in shared memory as token-major bf16 images for the
//     life of the CTA; 512 threads = 4 threads per query row of a 128 x 128 score tile (TMEM lane = query).
//   * forward: pass A recomputes S tile by tile for the exact row max, pass B recomputes S, writes
//     P = exp(S - max) (dropout applied, bf16) as an operand image and accumulates O += P V in TMEM; nothing is
//     rescaled and every exp is evaluated once.  Saves logsumexp per row.
//   * backward: per tile pair S and dP = dO V^T from the tensor core; the threads form P~ (dropped P) and
//     dS = P (dP~ - delta) scale as operand images; dV += P~^T dO, dK += dS^T Q, dQ += dS K accumulate in TMEM
//     (the transposed operands are the same images read MN-major).  dQ of all four query tiles stays in TMEM
//     until the end: no atomics, deterministic.
// The dropout mask is a counter hash of (call key, step counter, episode*head, query, key), regenerated in the
// backward.
#include "learn_common.cuh"

namespace xtrl {
namespace core {
using namespace umma;
using learn::pack2;

constexpr int R = 512, NTH = 512, DH = 16;
constexpr int kImg = R * DH * 2;              // one q / k / v / dO image: 2 chunks of 8 dims x 512 rows x 16 B
constexpr int kChunk = R * 16;                // bytes between the two 8-dim chunks
constexpr int kTileImg = 128 * 128 * 2;       // P / dS operand image of one tile
constexpr int kTCore = 128 * 16;              // bytes between 8-key chunks of a tile image
constexpr float kLog2e = 1.4426950408889634f;
constexpr uint32_t kIdA = learn::kIdA, kIdB = learn::kIdB;

struct Args {
    const __nv_bfloat16 *q, *k, *v, *o, *d_o;
    __nv_bfloat16 *out, *dq, *dk, *dv;
    float *lse;
    const int64_t *rng_step;
    uint32_t rng_key, drop_thr;
    float drop_scale, scale;
    int n;
};

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t mix32(uint32_t h) {
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return h;
}
__device__ __forceinline__ uint32_t cta_key(const int64_t *rng_step, uint32_t key, uint32_t bh) {
    const uint64_t s = (uint64_t)(rng_step ? *rng_step : 0) * 0x9E3779B97F4A7C15ull;
    return mix32((uint32_t)(s >> 32) ^ (uint32_t)s ^ (key * 0x632BE5ABu) ^ (bh * 0x9E3779B1u + 0x7F4A7C15u));
}
// keep factors of keys (key, key+1), key even, for query qrow
__device__ __forceinline__ void drop2(uint32_t qrow, uint32_t key, uint32_t ckey, uint32_t thr, float scale, float &k0, float &k1) {
    const uint32_t h = mix32((qrow * (uint32_t)(R / 2) + (key >> 1)) * 0x9E3779B1u ^ ckey);
    k0 = (h & 0xffffu) >= thr ? scale : 0.f;
    k1 = (h >> 16) >= thr ? scale : 0.f;
}
// [n][16] bf16 rows of one (episode, head) -> 512-row token-major image (rows >= n zero, up to `rows`)
__device__ __forceinline__ void load_img(unsigned char *img, const __nv_bfloat16 *src, int n, int rows, int tid) {
    for (int i = tid; i < rows * 2; i += NTH) {
        const int r = i >> 1, c = i & 1;
        const uint4 v = r < n ? __ldg(reinterpret_cast<const uint4 *>(src + (size_t)r * DH) + c) : make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4 *>(img + c * kChunk + (r >> 3) * 128 + (r & 7) * 16) = v;
    }
}
// 16 consecutive keys of this thread's query row -> two 16-byte pieces of a tile image
__device__ __forceinline__ void put16(unsigned char *img, int row, int col0, const float (&v)[16]) {
    unsigned char *p = img + (col0 >> 3) * kTCore + (row >> 3) * 128 + (row & 7) * 16;
    *reinterpret_cast<uint4 *>(p) = make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
    *reinterpret_cast<uint4 *>(p + kTCore) = make_uint4(pack2(v[8], v[9]), pack2(v[10], v[11]), pack2(v[12], v[13]), pack2(v[14], v[15]));
}
// token-major 512-row image, 128-row tile t: as a K-major operand [128 x K=16] and as an MN-major operand [N=16 x K=128 tokens]
__device__ __forceinline__ uint64_t desc_rows(const unsigned char *img, int t) { return smem_desc(img + t * 2048, kChunk, 128); }
__device__ __forceinline__ uint64_t desc_cols(const unsigned char *img, int t, int s) { return smem_desc(img + t * 2048 + s * 256, 128, kChunk); }

// -------------------------------------------------------------------------------------------- forward
constexpr int kFwdSmem = 3 * kImg + kTileImg + 4 * 128 * 4 + 64;

__global__ void __launch_bounds__(NTH, 2) attn_core_fwd_kernel(const Args a) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *sQ = smem, *sK = sQ + kImg, *sV = sK + kImg, *sP = sV + kImg;
    float *sRed = reinterpret_cast<float *>(sP + kTileImg);                 // [4][128]
    uint64_t *bar = reinterpret_cast<uint64_t *>(sRed + 4 * 128);
    uint32_t *tslot = reinterpret_cast<uint32_t *>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q4 = warp & 3, c4 = warp >> 2, row = q4 * 32 + lane;
    const int bh = blockIdx.x, n = a.n, nt = (n + 127) >> 7;
    constexpr uint32_t cS = 0, cO = 128;
    if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(tslot, 256);
    const size_t base = (size_t)bh * n * DH;
    load_img(sQ, a.q + base, n, nt * 128, tid);
    load_img(sK, a.k + base, n, nt * 128, tid);
    load_img(sV, a.v + base, n, nt * 128, tid);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tslot, tlane = tbase + ((uint32_t)(q4 * 32) << 16);
    const uint32_t ckey = cta_key(a.rng_step, a.rng_key, (uint32_t)bh);
    const float sl2 = a.scale * kLog2e;
    uint32_t ph = 0;
    for (int qt = 0; qt < nt; ++qt) {
        const int qrow = qt * 128 + row;
        float m = -INFINITY;
        for (int kt = 0; kt <= qt; ++kt) {                                   // pass A: exact row max
            if (tid == 0) {
                tc_fence_after();
                mma_bf16_ss(tbase + cS, desc_rows(sQ, qt), desc_rows(sK, kt), idesc_bf16(128, 128), 0);
                mma_commit(bar);
            }
            mbar_wait(bar, ph); ph ^= 1;
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                float s[16];
                tmem_ld16(tlane + cS + c4 * 32 + c * 16, s);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const bool valid = kt < qt || c4 * 32 + c * 16 + j <= row;
                    m = fmaxf(m, valid ? s[j] : -INFINITY);
                }
            }
            tc_fence_before();
            __syncthreads();
        }
        sRed[c4 * 128 + row] = m;
        __syncthreads();
        m = fmaxf(fmaxf(sRed[row], sRed[128 + row]), fmaxf(sRed[256 + row], sRed[384 + row]));
        const float ml2 = m * sl2;
        float l = 0.f;
        for (int kt = 0; kt <= qt; ++kt) {                                   // pass B: P, row sum, O += P V
            if (tid == 0) {
                tc_fence_after();
                if (kt > 0) {
#pragma unroll
                    for (int s = 0; s < 8; ++s)
                        mma_bf16_ss(tbase + cO, smem_desc(sP + s * 2 * kTCore, kTCore, 128), desc_cols(sV, kt - 1, s),
                                    idesc_bf16(128, DH) | kIdB, kt > 1 || s > 0);
                }
                mma_bf16_ss(tbase + cS, desc_rows(sQ, qt), desc_rows(sK, kt), idesc_bf16(128, 128), 0);
                mma_commit(bar);
            }
            mbar_wait(bar, ph); ph ^= 1;
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int col0 = c4 * 32 + c * 16;
                float s[16];
                tmem_ld16(tlane + cS + col0, s);
#pragma unroll
                for (int j = 0; j < 16; j += 2) {
                    const bool v0 = kt < qt || col0 + j <= row, v1 = kt < qt || col0 + j + 1 <= row;
                    float p0 = v0 ? ex2(fmaf(s[j], sl2, -ml2)) : 0.f, p1 = v1 ? ex2(fmaf(s[j + 1], sl2, -ml2)) : 0.f;
                    l += p0 + p1;
                    if (a.drop_thr) {
                        float k0, k1;
                        drop2((uint32_t)qrow, (uint32_t)(kt * 128 + col0 + j), ckey, a.drop_thr, a.drop_scale, k0, k1);
                        p0 *= k0; p1 *= k1;
                    }
                    s[j] = p0; s[j + 1] = p1;
                }
                put16(sP, row, col0, s);
            }
            fence_proxy_async();
            tc_fence_before();
            __syncthreads();
        }
        if (tid == 0) {
            tc_fence_after();
#pragma unroll
            for (int s = 0; s < 8; ++s)
                mma_bf16_ss(tbase + cO, smem_desc(sP + s * 2 * kTCore, kTCore, 128), desc_cols(sV, qt, s),
                            idesc_bf16(128, DH) | kIdB, qt > 0 || s > 0);
            mma_commit(bar);
        }
        sRed[c4 * 128 + row] = l;
        __syncthreads();
        l = (sRed[row] + sRed[128 + row]) + (sRed[256 + row] + sRed[384 + row]);
        mbar_wait(bar, ph); ph ^= 1;
        tc_fence_after();
        if (c4 < 2) {
            float o[8];
            tmem_ld8(tlane + cO + c4 * 8, o);
            const float inv = 1.f / l;
            if (qrow < n)
                *reinterpret_cast<uint4 *>(a.out + base + (size_t)qrow * DH + c4 * 8) =
                    make_uint4(pack2(o[0] * inv, o[1] * inv), pack2(o[2] * inv, o[3] * inv), pack2(o[4] * inv, o[5] * inv), pack2(o[6] * inv, o[7] * inv));
        } else if (c4 == 2 && qrow < n) {
            a.lse[(size_t)bh * n + qrow] = m * a.scale + __logf(l);
        }
        tc_fence_before();
        __syncthreads();
    }
    if (warp == 0) tmem_dealloc(tbase, 256);
}

// ------------------------------------------------------------------------------------------- backward
constexpr int kBwdSmem = 4 * kImg + 2 * kTileImg + 2 * R * 4 + 64;

__global__ void __launch_bounds__(NTH, 1) attn_core_bwd_kernel(const Args a) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *sQ = smem, *sK = sQ + kImg, *sV = sK + kImg, *sdO = sV + kImg, *sP = sdO + kImg, *sdS = sP + kTileImg;
    float *sL = reinterpret_cast<float *>(sdS + kTileImg), *sDelta = sL + R;
    uint64_t *bar = reinterpret_cast<uint64_t *>(sDelta + R);
    uint32_t *tslot = reinterpret_cast<uint32_t *>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q4 = warp & 3, c4 = warp >> 2, row = q4 * 32 + lane;
    const int bh = blockIdx.x, n = a.n, nt = (n + 127) >> 7;
    constexpr uint32_t cS = 0, cDP = 128, cDQ = 256, cDK = 320, cDV = 336;
    if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(tslot, 512);
    const size_t base = (size_t)bh * n * DH;
    load_img(sQ, a.q + base, n, nt * 128, tid);
    load_img(sK, a.k + base, n, nt * 128, tid);
    load_img(sV, a.v + base, n, nt * 128, tid);
    load_img(sdO, a.d_o + base, n, nt * 128, tid);
    for (int r = tid; r < nt * 128; r += NTH) {                              // delta = rowsum(dO . O); logsumexp in log2 units
        float d = 0.f, L = 1e30f;
        if (r < n) {
            const uint4 *po = reinterpret_cast<const uint4 *>(a.o + base + (size_t)r * DH), *pd = reinterpret_cast<const uint4 *>(a.d_o + base + (size_t)r * DH);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                float x[8], y[8];
                learn::unpack8(__ldg(po + c), x);
                learn::unpack8(__ldg(pd + c), y);
#pragma unroll
                for (int j = 0; j < 8; ++j) d = fmaf(x[j], y[j], d);
            }
            L = a.lse[(size_t)bh * n + r] * kLog2e;
        }
        sDelta[r] = d;
        sL[r] = L;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tslot, tlane = tbase + ((uint32_t)(q4 * 32) << 16);
    const uint32_t ckey = cta_key(a.rng_step, a.rng_key, (uint32_t)bh);
    const float sl2 = a.scale * kLog2e;
    uint32_t ph = 0;
    bool pending = false;                                                    // operand images of the previous pair are waiting for their MMAs
    int pq = 0, pk = 0;
    for (int kt = 0; kt < nt; ++kt) {
        for (int qt = kt; qt < nt; ++qt) {
            if (tid == 0) {
                tc_fence_after();
                if (pending) {
#pragma unroll
                    for (int s = 0; s < 8; ++s) {
                        const uint64_t pT = smem_desc(sP + s * 256, 128, kTCore), dsT = smem_desc(sdS + s * 256, 128, kTCore);
                        mma_bf16_ss(tbase + cDV, pT, desc_cols(sdO, pq, s), idesc_bf16(128, DH) | kIdA | kIdB, pq > pk || s > 0);      // dV += P~^T dO
                        mma_bf16_ss(tbase + cDK, dsT, desc_cols(sQ, pq, s), idesc_bf16(128, DH) | kIdA | kIdB, pq > pk || s > 0);      // dK += dS^T Q
                        mma_bf16_ss(tbase + cDQ + pq * DH, smem_desc(sdS + s * 2 * kTCore, kTCore, 128), desc_cols(sK, pk, s),
                                    idesc_bf16(128, DH) | kIdB, pk > 0 || s > 0);                                                        // dQ += dS K
                    }
                }
                mma_bf16_ss(tbase + cS, desc_rows(sQ, qt), desc_rows(sK, kt), idesc_bf16(128, 128), 0);                                 // S = Q K^T
                mma_bf16_ss(tbase + cDP, desc_rows(sdO, qt), desc_rows(sV, kt), idesc_bf16(128, 128), 0);                               // dP~ = dO V^T
                mma_commit(bar);
            }
            mbar_wait(bar, ph); ph ^= 1;
            tc_fence_after();
            if (pending && pq == nt - 1) {                                   // the previous key tile is complete: write its dK, dV
                if (c4 < 2) {
                    float v[16];
                    tmem_ld16(tlane + (c4 == 0 ? cDK : cDV), v);
                    const int krow = pk * 128 + row;
                    if (krow < n) {
                        uint4 *dst = reinterpret_cast<uint4 *>((c4 == 0 ? a.dk : a.dv) + base + (size_t)krow * DH);
                        dst[0] = make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
                        dst[1] = make_uint4(pack2(v[8], v[9]), pack2(v[10], v[11]), pack2(v[12], v[13]), pack2(v[14], v[15]));
                    }
                }
            }
            const int qrow = qt * 128 + row;
            const float Lq = sL[qrow], delta = sDelta[qrow];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int col0 = c4 * 32 + c * 16;
                float s[16], dp[16];
                tmem_ld16(tlane + cS + col0, s);
                tmem_ld16(tlane + cDP + col0, dp);
#pragma unroll
                for (int j = 0; j < 16; j += 2) {
                    const bool v0 = kt < qt || col0 + j <= row, v1 = kt < qt || col0 + j + 1 <= row;
                    const float p0 = v0 ? ex2(fmaf(s[j], sl2, -Lq)) : 0.f, p1 = v1 ? ex2(fmaf(s[j + 1], sl2, -Lq)) : 0.f;
                    float k0 = 1.f, k1 = 1.f;
                    if (a.drop_thr) drop2((uint32_t)qrow, (uint32_t)(kt * 128 + col0 + j), ckey, a.drop_thr, a.drop_scale, k0, k1);
                    s[j] = p0 * k0; s[j + 1] = p1 * k1;
                    dp[j] = p0 * (dp[j] * k0 - delta) * a.scale;
                    dp[j + 1] = p1 * (dp[j + 1] * k1 - delta) * a.scale;
                }
                put16(sP, row, col0, s);
                put16(sdS, row, col0, dp);
            }
            fence_proxy_async();
            tc_fence_before();
            __syncthreads();
            pending = true; pq = qt; pk = kt;
        }
    }
    if (tid == 0) {                                                          // the last pair's products
        tc_fence_after();
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            const uint64_t pT = smem_desc(sP + s * 256, 128, kTCore), dsT = smem_desc(sdS + s * 256, 128, kTCore);
            mma_bf16_ss(tbase + cDV, pT, desc_cols(sdO, pq, s), idesc_bf16(128, DH) | kIdA | kIdB, pq > pk || s > 0);
            mma_bf16_ss(tbase + cDK, dsT, desc_cols(sQ, pq, s), idesc_bf16(128, DH) | kIdA | kIdB, pq > pk || s > 0);
            mma_bf16_ss(tbase + cDQ + pq * DH, smem_desc(sdS + s * 2 * kTCore, kTCore, 128), desc_cols(sK, pk, s),
                        idesc_bf16(128, DH) | kIdB, pk > 0 || s > 0);
        }
        mma_commit(bar);
    }
    mbar_wait(bar, ph); ph ^= 1;
    tc_fence_after();
    if (c4 < 2) {
        float v[16];
        tmem_ld16(tlane + (c4 == 0 ? cDK : cDV), v);
        const int krow = pk * 128 + row;
        if (krow < n) {
            uint4 *dst = reinterpret_cast<uint4 *>((c4 == 0 ? a.dk : a.dv) + base + (size_t)krow * DH);
            dst[0] = make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
            dst[1] = make_uint4(pack2(v[8], v[9]), pack2(v[10], v[11]), pack2(v[12], v[13]), pack2(v[14], v[15]));
        }
    }
    if (c4 < nt) {                                                           // dQ of query tile c4
        float v[16];
        tmem_ld16(tlane + cDQ + c4 * DH, v);
        const int qrow = c4 * 128 + row;
        if (qrow < n) {
            uint4 *dst = reinterpret_cast<uint4 *>(a.dq + base + (size_t)qrow * DH);
            dst[0] = make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
            dst[1] = make_uint4(pack2(v[8], v[9]), pack2(v[10], v[11]), pack2(v[12], v[13]), pack2(v[14], v[15]));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

static int fill(const xtrl_attn_core_io *io, int32_t n, int32_t dim_head, Args &a) {
    XTRL_REQUIRE(dim_head == DH && n >= 1 && n <= R, XTRL_E_UNSUPPORTED, "xtrl_attn_core_*: built for dim_head=16, 1 <= n <= 512 (got %d, %d)", dim_head, n);
    XTRL_REQUIRE(io->dropout >= 0.f && io->dropout < 1.f, XTRL_E_BADARG, "xtrl_attn_core_*: dropout %f not in [0,1)", io->dropout);
    a.q = static_cast<const __nv_bfloat16 *>(io->q); a.k = static_cast<const __nv_bfloat16 *>(io->k); a.v = static_cast<const __nv_bfloat16 *>(io->v);
    a.o = static_cast<const __nv_bfloat16 *>(io->o); a.d_o = static_cast<const __nv_bfloat16 *>(io->d_o);
    a.out = static_cast<__nv_bfloat16 *>(io->o_out); a.dq = static_cast<__nv_bfloat16 *>(io->dq); a.dk = static_cast<__nv_bfloat16 *>(io->dk);
    a.dv = static_cast<__nv_bfloat16 *>(io->dv); a.lse = io->lse;
    a.rng_step = io->rng_step; a.rng_key = io->rng_key;
    a.drop_thr = (uint32_t)(io->dropout * 65536.f + 0.5f);
    a.drop_scale = 1.f / (1.f - (float)a.drop_thr / 65536.f);
    a.scale = io->scale; a.n = n;
    return XTRL_OK;
}

}  // namespace core
}  // namespace xtrl

extern "C" int xtrl_attn_core_supported(int32_t n, int32_t dim_head) { return dim_head == xtrl::core::DH && n >= 1 && n <= xtrl::core::R; }

extern "C" int xtrl_attn_core_fwd(const xtrl_attn_core_io *io, int32_t bh, int32_t n, int32_t dim_head, void *stream) {
    using namespace xtrl; using namespace xtrl::core;
    XTRL_REQUIRE(io && bh >= 0, XTRL_E_BADARG, "xtrl_attn_core_fwd: bad arguments");
    if (bh == 0) return XTRL_OK;
    XTRL_REQUIRE(io->q && io->k && io->v && io->o_out && io->lse, XTRL_E_BADARG, "xtrl_attn_core_fwd: null pointer");
    Args a;
    if (int rc = fill(io, n, dim_head, a)) return rc;
    static bool attr = false;
    if (!attr) { XTRL_CUDA(cudaFuncSetAttribute(attn_core_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmem)); attr = true; }
    attn_core_fwd_kernel<<<bh, NTH, kFwdSmem, (cudaStream_t)stream>>>(a);
    XTRL_LAUNCH_CHECK("attn_core_fwd_kernel");
    return XTRL_OK;
}

extern "C" int xtrl_attn_core_bwd(const xtrl_attn_core_io *io, int32_t bh, int32_t n, int32_t dim_head, void *stream) {
    using namespace xtrl; using namespace xtrl::core;
    XTRL_REQUIRE(io && bh >= 0, XTRL_E_BADARG, "xtrl_attn_core_bwd: bad arguments");
    if (bh == 0) return XTRL_OK;
    XTRL_REQUIRE(io->q && io->k && io->v && io->o && io->d_o && io->lse && io->dq && io->dk && io->dv, XTRL_E_BADARG,
                 "xtrl_attn_core_bwd: null pointer");
    Args a;
    if (int rc = fill(io, n, dim_head, a)) return rc;
    static bool attr = false;
    if (!attr) { XTRL_CUDA(cudaFuncSetAttribute(attn_core_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem)); attr = true; }
    attn_core_bwd_kernel<<<bh, NTH, kBwdSmem, (cudaStream_t)stream>>>(a);
    XTRL_LAUNCH_CHECK("attn_core_bwd_kernel");
    return XTRL_OK;
}
