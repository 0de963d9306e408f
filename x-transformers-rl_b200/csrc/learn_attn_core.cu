// learn_attn_core.cu -- learn-side causal attention core for dim_head = 16, sequences up to 512, forward and
// backward, on tcgen05:      o = dropout(softmax(scale * q k^T + causal)) v       per (episode, head).
//
// Replaces the attention core of the teacher-forced pass inside Agent.learn (x_transformers_rl.py:928-935;
// x-transformers Attention, SURVEY App. A.1: causal softmax(q k^T * dh^-1/2), attention dropout, @ v) and its
// backward.  SURVEY 8f row 1.  With dh = 16 the matmuls are tiny (one K-step per score tile); the work is the
// per-score softmax arithmetic, so the design is about instructions per score and about never making the threads wait
// for the tensor core:
//   * one CTA per (episode, head); q, k, v (and dO) live in shared memory as token-major bf16 images for the life of
//     the CTA.  16 compute warps work on [128 x 128] score tiles (TMEM lane = query in the forward, = key in the
//     backward), plus one (forward) or two (backward) warps that only issue MMAs.  TMEM score tiles and mbarriers keep the
//     tensor core ahead of and behind the threads; the loops have no __syncthreads.
//   * forward: the softmax shift of a row is the bound scale * |q_row| * max|k| (Cauchy-Schwarz), so each score is visited
//     once: P = exp(s - shift), rounded to bf16 with the dropout mask ANDed in, goes to the tensor core THROUGH TENSOR
//     MEMORY as the A operand of O += P V; the row sum is kept in fp32; logsumexp is saved.  If the bound is loose (> 40)
//     the same loop runs twice per query tile: an exact row-max pass first.
//   * backward: S^T and dP~^T from the tensor core; the threads form P~^T (dropped P) and dS'^T = P~ dP~ - P delta, which
//     feed dV += P~^T dO and dK += dS'^T Q through tensor memory and dQ += dS' K through one shared-memory image; scale is
//     applied when dQ / dK are written.  dQ of all four query tiles stays in TMEM until the end: no atomics, deterministic.
// (The round-1 kernels -- 64-key steps, P / dS handed over as shared-memory images -- are in the history: 190 / 379 us
//  against 165 / 303 us at the c3 minibatch.)
// The dropout mask is a counter hash of (call key, step counter, episode*head, query, key), regenerated in the
// backward; p is quantised to 1/128.
#include "learn_common.cuh"

namespace xtrl {
namespace core {
using namespace umma;
using learn::pack2; using learn::pk2; using learn::upk2; using learn::fma2; using learn::mul2; using learn::add2;

constexpr int R = 512, NCOMP = 512, NTH = NCOMP + 32, DH = 16;      // 16 compute warps + one warp that only issues MMAs
constexpr int kImg = R * DH * 2;              // one q / k / v / dO image: 2 chunks of 8 dims x 512 rows x 16 B
constexpr int kChunk = R * 16;                // bytes between the two 8-dim chunks
constexpr int kTCore = 128 * 16;              // bytes between 8-key chunks of a tile image
constexpr float kLog2e = 1.4426950408889634f;
constexpr uint32_t kIdA = learn::kIdA, kIdB = learn::kIdB;

struct Args {
    const __nv_bfloat16 *q, *k, *v, *o, *d_o;
    __nv_bfloat16 *out, *dq, *dk, *dv;
    float *lse;
    const int64_t *rng_step;
    uint32_t rng_key, drop_add4;              // (128 - threshold) replicated in 4 bytes, see keep4(); 0 = no dropout
    float drop_scale, drop_log2scale, scale;
    int n;
    int ahead;                                // CTAs resident at once: head bh + ahead is the one this SM runs next
};

// Pull the rows of the head this SM will work on next into L2 while this CTA runs (CTAs are handed out in blockIdx order, one
// wave = `ahead` CTAs).  The prologue of a CTA is two dependent rounds of global loads that nothing overlaps -- with one CTA
// per SM (the backward) they are pure dead time, 13.8 waves deep -- so they should at least hit L2 instead of DRAM.
__device__ __forceinline__ void warm_next_head(const Args &a, int bh_next, int nbh, int tid, int nthreads, bool bwd) {
    if (bh_next >= nbh) return;
    const size_t off = (size_t)bh_next * a.n * 16;
    const int lines = (a.n * 16 * 2 + 127) >> 7, na = bwd ? 5 : 3;
    for (int i = tid; i < lines * na; i += nthreads) {
        const int m = i / lines, l = i - m * lines;
        const __nv_bfloat16 *p = m == 0 ? a.q : m == 1 ? a.k : m == 2 ? a.v : m == 3 ? a.d_o : a.o;
        learn::prefetch_l2(reinterpret_cast<const char *>(p + off) + l * 128);
    }
    if (bwd) for (int i = tid; i < (a.n * 4 + 127) >> 7; i += nthreads) learn::prefetch_l2(reinterpret_cast<const char *>(a.lse + (size_t)bh_next * a.n) + i * 128);
}

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
// Dropout decisions of the four keys 4*g4 .. 4*g4+3 of query qrow: one hash -> four 7-bit uniforms u; the sign bit of
// byte j of the result is set <=> u_j >= threshold <=> keep.  hseed = (qrow * R/4 + first g4 of the chunk) * golden + CTA key;
// add4 = (128 - threshold) in every byte.  PRMT's sign-replicate selectors (8 | byte) turn the sign bits into masks.
constexpr uint32_t kGolden = 0x9E3779B1u;
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t sel) {          // PTX prmt (default mode): selector bit 3 = replicate the sign
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(0u), "r"(sel));
    return d;
}
__device__ __forceinline__ uint32_t keep4(uint32_t hseed, int g, uint32_t add4) {
    uint32_t h = hseed + (uint32_t)g * kGolden;
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 15;
    return (h & 0x7f7f7f7fu) + add4;
}
// NA arrays at once: every thread first issues all its 16-byte loads (2 per array), then stores them
template <int NA>
__device__ __forceinline__ void load_imgs(unsigned char *const (&img)[NA], const __nv_bfloat16 *const (&src)[NA], int n, int rows, int tid) {
    uint4 v[NA][2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int i = tid + k * NTH, r = i >> 1, c = i & 1;
#pragma unroll
        for (int a = 0; a < NA; ++a)
            v[a][k] = (i < rows * 2 && r < n) ? __ldg(reinterpret_cast<const uint4 *>(src[a] + (size_t)r * DH) + c) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int i = tid + k * NTH, r = i >> 1, c = i & 1;
        if (i < rows * 2) {
#pragma unroll
            for (int a = 0; a < NA; ++a) *reinterpret_cast<uint4 *>(img[a] + c * kChunk + (r >> 3) * 128 + (r & 7) * 16) = v[a][k];
        }
    }
}
static_assert(2 * NTH >= 2 * R, "load_imgs covers an image in two passes");
__device__ __forceinline__ float row_norm2(const unsigned char *img, int r) {
    float x[8], y[8], s = 0.f;
    learn::unpack8(*reinterpret_cast<const uint4 *>(img + (r >> 3) * 128 + (r & 7) * 16), x);
    learn::unpack8(*reinterpret_cast<const uint4 *>(img + kChunk + (r >> 3) * 128 + (r & 7) * 16), y);
#pragma unroll
    for (int j = 0; j < 8; ++j) s = fmaf(x[j], x[j], fmaf(y[j], y[j], s));
    return s;
}
// Descriptors are built once per operand; every MMA then only adds a start-address offset (in 16-byte units) -- a
// tcgen05.mma of these shapes costs ~48 cycles of issue, building a descriptor from a pointer costs about as much again.
// token-major 512-row image: rows(t) = K-major operand [128 rows of tile t x K=16] (64-row key tiles: + kt * 64);
// cols(t, s) = MN-major operand [N=16 x K = 16 tokens of step s of tile t]
struct ImgDesc {
    uint64_t rows, cols;
    __device__ __forceinline__ explicit ImgDesc(const unsigned char *img) : rows(smem_desc(img, kChunk, 128)), cols(smem_desc(img, 128, kChunk)) {}
};
// step image [128 queries x 64 keys]: kmaj(s) = A operand [128 q x K = 16 keys of step s]; mnmaj(s) = A operand [M = keys x K = 16 queries of step s]
struct StepDesc {
    uint64_t kmaj, mnmaj;
    __device__ __forceinline__ explicit StepDesc(const unsigned char *img) : kmaj(smem_desc(img, kTCore, 128)), mnmaj(smem_desc(img, 128, kTCore)) {}
};

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void csync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }      // the 16 compute warps
// this warp's part of a step image is written and its TMEM tile is read: tell the MMA warp (one arrival per warp)
__device__ __forceinline__ void images_ready(uint64_t *bar, int lane) {
    fence_proxy_async();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
}

constexpr float kBoundMax = 40.f;             // use the norm bound instead of the exact row max while it is below this

// ---------------------------------------------------------------------------------------- forward, version 2
// Steps of one [128 queries x 128 keys] tile: a thread turns 32 scores per step into probabilities (two chunks of 16 keys), so
// the per-step barrier / fence / geometry cost is spread over twice the work, and P goes to the tensor core THROUGH TENSOR
// MEMORY (tcgen05.st + the A-from-TMEM form of tcgen05.mma) instead of a shared-memory image: no 16-byte scatter, no
// generic->async proxy fence, no 4 KB operand read per product.  One S tile lives in TMEM at a time; the next one is issued as
// soon as every warp has READ the current one (barrier bR, half way through the step), so it is ready when the step ends;
// two CTAs per SM cover what is left.  The exact-row-max case (norm bound too loose) is the same loop run twice per query
// tile: a pass that only takes maxima, then the pass that forms P.
constexpr int kFwd2Smem = 3 * kImg + 2 * 4 * 128 * 4 + R * 4 + 128;
#ifndef XTRL_WAIT_NS
#define XTRL_WAIT_NS 2000
#endif
constexpr uint32_t kWaitNs = XTRL_WAIT_NS;   // suspend-time hint of the compute warps' barrier waits (a sleeping warp costs no issue slots)

template <bool DIAG, bool DROP>
__device__ __forceinline__ float fwd2_chunk(uint32_t taddr, int row, int keyb, float sl2, float off, uint32_t qrow, uint32_t g4, uint32_t ckey,
                                            uint32_t add4, uint32_t (&w)[8]) {
    uint32_t r[16];
    tmem_ld16_issue(taddr, r);
    tmem_ld_wait();
    tmem_pin16(r);
    const uint64_t sl22 = pk2(sl2, sl2), off2 = pk2(off, off);
    uint64_t l2 = pk2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
        float x0, x1;
        upk2(fma2(pk2(__uint_as_float(r[j]), __uint_as_float(r[j + 1])), sl22, off2), x0, x1);
        float p0 = ex2(x0), p1 = ex2(x1);
        if (DIAG) { p0 = keyb + j <= row ? p0 : 0.f; p1 = keyb + j + 1 <= row ? p1 : 0.f; }
        l2 = add2(l2, pk2(p0, p1));
        w[j >> 1] = pack2(p0, p1);
    }
    if (DROP) {
        const uint32_t hseed = (qrow * (uint32_t)(R / 4) + g4) * kGolden + ckey;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const uint32_t t = keep4(hseed, g, add4);
            w[2 * g] &= prmt(t, 0x9988);
            w[2 * g + 1] &= prmt(t, 0xbbaa);
        }
    }
    float l, lh;
    upk2(l2, l, lh);
    return l + lh;
}

template <bool DROP>
__global__ void __launch_bounds__(NTH, 2) attn_core_fwd2_kernel(const Args a) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *sQ = smem, *sK = sQ + kImg, *sV = sK + kImg;
    float *sRed = reinterpret_cast<float *>(sV + kImg);                     // [2][4][128]
    float *sQn = sRed + 2 * 4 * 128;                                         // |q_row|^2
    uint64_t *bS = reinterpret_cast<uint64_t *>(sQn + R), *bR = bS + 1, *bI = bR + 1, *bPV = bI + 1;
    uint32_t *tslot = reinterpret_cast<uint32_t *>(bPV + 1);
    int *sMax = reinterpret_cast<int *>(tslot + 1);                          // [0] max |k|^2, [1] max |q|^2 (float bits)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q4 = warp & 3, c4 = warp >> 2, row = q4 * 32 + lane;           // c4 == 4: the MMA warp
    const int bh = blockIdx.x, n = a.n, nt = (n + 127) >> 7;
    constexpr uint32_t cS = 0, cP = 128, cO = 192;
    if (tid == 0) {
        mbar_init(bS, 1); mbar_init(bR, NCOMP / 32); mbar_init(bI, NCOMP / 32); mbar_init(bPV, 1);
        fence_mbar_init(); sMax[0] = 0; sMax[1] = 0;
    }
    if (warp == 0) tmem_alloc(tslot, 256);
    const size_t base = (size_t)bh * n * DH;
    {
        unsigned char *const imgs[3] = {sQ, sK, sV};
        const __nv_bfloat16 *const srcs[3] = {a.q + base, a.k + base, a.v + base};
        load_imgs<3>(imgs, srcs, n, nt * 128, tid);
    }
    __syncthreads();
    for (int r = tid; r < nt * 128 + 32; r += NTH) {                         // norms for the softmax shift bound (whole warps: shuffles)
        const bool in = r < nt * 128;
        const float qn = in ? row_norm2(sQ, r) : 0.f, kn = learn::warp_max_f(in ? row_norm2(sK, r) : 0.f), qm = learn::warp_max_f(qn);
        if (in) sQn[r] = qn;
        if (lane == 0) { atomicMax(&sMax[0], __float_as_int(kn)); atomicMax(&sMax[1], __float_as_int(qm)); }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tslot, tlane = tbase + ((uint32_t)(q4 * 32) << 16);
    const float sl2 = a.scale * kLog2e;
    const float kmax2 = __int_as_float(sMax[0]);
    const bool bounded = a.scale * sqrtf(kmax2 * __int_as_float(sMax[1])) <= kBoundMax;          // CTA-uniform
    const int npass = bounded ? 1 : 2;
    const int total = npass * (nt * (nt + 1) / 2);                           // steps: (query tile, pass, key tile <= query tile)
    auto advance = [&](int &qt, int &ps, int &kt) { if (++kt > qt) { kt = 0; if (++ps >= npass) { ps = 0; ++qt; } } };

    if (warp == NCOMP / 32) {
        // ---- MMA warp: S of the next step as soon as the current S tile has been read; O += P V as soon as P is written ----
        if (lane == 0) {
            const ImgDesc dQ_(sQ), dK_(sK), dV_(sV);
            int qt = 0, ps = 0, kt = 0, iq = 0, ip = 0, ik = 0;
            tc_fence_after();
            mma_bf16_ss(tbase + cS, dQ_.rows + 0, dK_.rows + 0, idesc_bf16(128, 128), 0);
            mma_commit(bS);
            advance(iq, ip, ik);
            for (int i = 0; i < total; ++i) {
                mbar_wait_sleep(bR, i & 1, 400);
                tc_fence_after();
                if (i + 1 < total) {
                    mma_bf16_ss(tbase + cS, dQ_.rows + iq * 128, dK_.rows + ik * 128, idesc_bf16(128, 128), 0);
                    mma_commit(bS);
                    advance(iq, ip, ik);
                }
                mbar_wait_sleep(bI, i & 1, 400);
                tc_fence_after();
                if (ps == npass - 1) {
                    const uint64_t vb = dV_.cols + kt * 128;
#pragma unroll
                    for (int s = 0; s < 8; ++s)                               // K = 128 keys, A = P from tensor memory
                        mma_bf16_ts(tbase + cO + (qt & 1) * DH, tbase + cP + s * 8, vb + s * 16, idesc_bf16(128, DH) | kIdB, kt > 0 || s > 0);
                }
                mma_commit(bPV);
                advance(qt, ps, kt);
            }
        }
    } else {
        // ---- compute warps: row = query, chunks c4 and c4 + 4 of the step's eight 16-key chunks ----
        const uint32_t ckey = cta_key(a.rng_step, a.rng_key, (uint32_t)bh);
        auto arrive = [&](uint64_t *bar) { tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive(bar); };
        auto finalize = [&](int qt, float m) {                               // O / l and logsumexp of query tile qt (its last P.V is complete)
            const int qrow = qt * 128 + row;
            const float *red = sRed + (qt & 1) * 512;
            const float l = ((red[row] + red[128 + row]) + (red[256 + row] + red[384 + row])) / a.drop_scale;
            if (c4 < 2) {
                float o[8];
                tmem_ld8(tlane + cO + (qt & 1) * DH + c4 * 8, o);
                const float inv = 1.f / l;
                if (qrow < n)
                    *reinterpret_cast<uint4 *>(a.out + base + (size_t)qrow * DH + c4 * 8) =
                        make_uint4(pack2(o[0] * inv, o[1] * inv), pack2(o[2] * inv, o[3] * inv), pack2(o[4] * inv, o[5] * inv), pack2(o[6] * inv, o[7] * inv));
            } else if (c4 == 2 && qrow < n) {
                a.lse[(size_t)bh * n + qrow] = m + __logf(l);
            }
            tc_fence_before();
        };
        int qt = 0, ps = 0, kt = 0;
        float m = 0.f, m_prev = 0.f, l = 0.f, mx = -INFINITY;
        if (bounded) m = a.scale * sqrtf(kmax2 * sQn[row]) * 1.0002f + 1e-4f;                      // >= every scaled score of the row (Cauchy-Schwarz)
        float off = a.drop_log2scale - m * kLog2e;
        for (int i = 0; i < total; ++i) {
            mbar_wait_sleep(bS, i & 1, kWaitNs);
            tc_fence_after();
            const bool diag = kt == qt, main_pass = ps == npass - 1;
            const uint32_t qabs = (uint32_t)(qt * 128 + row);
            uint32_t w[2][8];
            if (!main_pass) {
                // exact row maximum over this key tile (scaled later)
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    const int keyb = (c4 + 4 * cc) * 16;
                    if (!diag || keyb <= q4 * 32 + 31) {
                        float sv[16];
                        tmem_ld16(tlane + cS + keyb, sv);
#pragma unroll
                        for (int j = 0; j < 16; ++j) mx = fmaxf(mx, (!diag || keyb + j <= row) ? sv[j] : -INFINITY);
                    }
                }
                arrive(bR);
                arrive(bI);
                if (kt == qt) {                                              // last key tile of the max pass: the four column groups meet
                    sRed[(qt & 1) * 512 + c4 * 128 + row] = mx;
                    csync();
                    const float *red = sRed + (qt & 1) * 512;
                    m = fmaxf(fmaxf(red[row], red[128 + row]), fmaxf(red[256 + row], red[384 + row])) * a.scale;
                    off = a.drop_log2scale - m * kLog2e;
                    mx = -INFINITY;
                    csync();
                }
            } else {
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    const int keyb = (c4 + 4 * cc) * 16;                     // first key of the chunk inside the key tile
                    const uint32_t g4 = (uint32_t)(kt * 128 + keyb) >> 2;
                    if (!diag) l += fwd2_chunk<false, DROP>(tlane + cS + keyb, row, 0, sl2, off, qabs, g4, ckey, a.drop_add4, w[cc]);
                    else if (keyb <= q4 * 32 + 31) l += fwd2_chunk<true, DROP>(tlane + cS + keyb, row, keyb, sl2, off, qabs, g4, ckey, a.drop_add4, w[cc]);
                    else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) w[cc][j] = 0u;
                    }
                }
                arrive(bR);                                                  // the S tile has been read: the next one may be issued
                if (i > 0) { mbar_wait_sleep(bPV, (i - 1) & 1, kWaitNs); tc_fence_after(); }      // the P.V that read the P columns is complete
                if (kt == 0 && qt > 0) finalize(qt - 1, m_prev);             // ... and so is the previous query tile's last one
                tmem_st8(tlane + cP + c4 * 8, w[0]);
                tmem_st8(tlane + cP + (c4 + 4) * 8, w[1]);
                tmem_st_wait();
                arrive(bI);
                if (kt == qt) {                                              // last key tile of the query tile: row sums of the four column groups
                    sRed[(qt & 1) * 512 + c4 * 128 + row] = l;
                    csync();
                    m_prev = m; l = 0.f;
                    if (bounded && qt + 1 < nt) { m = a.scale * sqrtf(kmax2 * sQn[(qt + 1) * 128 + row]) * 1.0002f + 1e-4f; off = a.drop_log2scale - m * kLog2e; }
                }
            }
            advance(qt, ps, kt);
        }
        mbar_wait(bPV, (total - 1) & 1);
        tc_fence_after();
        finalize(nt - 1, m_prev);
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 256);
}

// ---- shared by the backward ----
__device__ __forceinline__ void store_row16(__nv_bfloat16 *p, const float (&v)[16], float f) {
    uint4 *dst = reinterpret_cast<uint4 *>(p);
    dst[0] = make_uint4(pack2(v[0] * f, v[1] * f), pack2(v[2] * f, v[3] * f), pack2(v[4] * f, v[5] * f), pack2(v[6] * f, v[7] * f));
    dst[1] = make_uint4(pack2(v[8] * f, v[9] * f), pack2(v[10] * f, v[11] * f), pack2(v[12] * f, v[13] * f), pack2(v[14] * f, v[15] * f));
}

// ------------------------------------------------------------------------------------ backward, version 2
// Key-row orientation: the tensor core produces S^T = K Q^T and dP~^T = V dO^T, so TMEM lane = KEY and a thread walks along
// queries.  P~^T and dS'^T are then exactly the A operands of dV += P~^T dO and dK += dS'^T Q with M = keys, and are handed to
// the tensor core THROUGH TENSOR MEMORY (tcgen05.st + the A-from-TMEM form of tcgen05.mma): no shared-memory image and no
// 4 KB operand read per product (measured, scripts/umma_issue_probe.cu: N = 16 products cost <= 14 cycles of tensor time from
// TMEM against 39 from shared memory).  Only dS' still crosses shared memory (read MN-major with M = queries for dQ += dS' K).
// One issuing warp sustains one tcgen05.mma per ~55 cycles whatever its size (same probe), and a [128 keys x 128 queries]
// tile needs 28 of them, so TWO warps issue: warp A the S^T / dP~^T tiles and the dQ products, warp B the dV / dK products.
// A tile is processed as two halves of 64 queries (TMEM budget: 2 x (S^T 64 + dP~^T 64 + P~^T 32 + dS'^T 32) + dV, dK 2 x 16
// each + dQ 4 x 16 = 512 columns); 16 compute warps = 4 lane quarters (keys) x 4 groups of 16 queries per half.
// Per-query statistics (logsumexp, delta) are warp-uniform here and come from shared memory as broadcast loads; the softmax /
// dS' arithmetic runs on packed fp32 pairs (FFMA2 / FMUL2).  The dropout mask is the forward's: one hash per (query, 4 keys);
// the four lanes that own those keys compute a quarter of the hashes each and exchange them by shuffle.
constexpr int NTH2 = NCOMP + 64;              // 16 compute warps + two MMA-issuing warps
constexpr int kTileImg2 = 128 * 128 * 2;      // dS' image of one [128 keys x 128 queries] tile (rows = keys)
constexpr int kBwd2Smem = 2 * kTileImg2 + 4 * kImg + 2 * R * 4 + 128;

__device__ __forceinline__ uint32_t keep_word(uint32_t h, uint32_t add4) {       // keep4()'s mixing of an already-seeded word
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 15;
    return (h & 0x7f7f7f7fu) + add4;
}

// one chunk of the backward in key-row orientation: my key, the 16 queries q0 .. q0 + 15 of this half.
//   ts / tdp: TMEM addresses of my 16 S^T / dP~^T columns; stat: the 16 queries' statistics in blocks of two queries
//   [off_q, off_q+1, ndelta_q, ndelta_q+1] (shared memory, broadcast loads that land as ready-made fp32 pairs)
//   returns the packed P~^T (wp) and dS'^T (wd) rows: 8 words each.
template <bool DIAG, bool DROP>
__device__ __forceinline__ void bwd2_chunk(uint32_t ts, uint32_t tdp, const float4 *stat, int key_local, int q_local0, float sl2,
                                           uint32_t hbase, uint32_t q_abs0, uint32_t add4, int lane, uint32_t (&wp)[8], uint32_t (&wd)[8]) {
    uint32_t rs[16], rd[16];
    tmem_ld16_issue(ts, rs);
    tmem_ld16_issue(tdp, rd);
    // the quad's hashes: lane (l & 3) seeds queries e = (l & 3) + 4 k
    uint32_t town[4];
    if (DROP) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            town[k] = keep_word((q_abs0 + (uint32_t)(4 * k + (lane & 3))) * ((uint32_t)(R / 4) * kGolden) + hbase, add4);
    }
    const uint32_t sel = 0x8888u + 0x1111u * (uint32_t)(lane & 3);           // replicate the sign of MY key's byte
    const uint64_t sl22 = pk2(sl2, sl2);
    tmem_ld_wait();
    tmem_pin16(rs);
    tmem_pin16(rd);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float4 st01 = stat[2 * k], st23 = stat[2 * k + 1];
        const float offv[4] = {st01.x, st01.y, st23.x, st23.y}, ndv[4] = {st01.z, st01.w, st23.z, st23.w};
        float ps[4], pd[4], ds[4];
        uint32_t msk[4];
        if (DROP) {
#pragma unroll
            for (int d = 0; d < 4; ++d) msk[d] = prmt(__shfl_sync(0xffffffffu, town[k], (lane & ~3) + d), sel);
        }
#pragma unroll
        for (int d = 0; d < 4; d += 2) {
            const int j = 4 * k + d;
            const uint64_t x = fma2(pk2(__uint_as_float(rs[j]), __uint_as_float(rs[j + 1])), sl22, pk2(offv[d], offv[d + 1]));
            float x0, x1;
            upk2(x, x0, x1);
            ps[d] = ex2(x0); ps[d + 1] = ex2(x1);                                 // P * drop_scale
            if (DIAG) { ps[d] = key_local <= q_local0 + j ? ps[d] : 0.f; ps[d + 1] = key_local <= q_local0 + j + 1 ? ps[d + 1] : 0.f; }
            pd[d] = DROP ? __uint_as_float(__float_as_uint(ps[d]) & msk[d]) : ps[d];
            pd[d + 1] = DROP ? __uint_as_float(__float_as_uint(ps[d + 1]) & msk[d + 1]) : ps[d + 1];
            const uint64_t pn = mul2(pk2(ps[d], ps[d + 1]), pk2(ndv[d], ndv[d + 1]));
            const uint64_t dd = fma2(pk2(pd[d], pd[d + 1]), pk2(__uint_as_float(rd[j]), __uint_as_float(rd[j + 1])), pn);
            upk2(dd, ds[d], ds[d + 1]);
        }
        wp[2 * k] = pack2(pd[0], pd[1]); wp[2 * k + 1] = pack2(pd[2], pd[3]);
        wd[2 * k] = pack2(ds[0], ds[1]); wd[2 * k + 1] = pack2(ds[2], ds[3]);
    }
}

// CQ = chunks of 16 queries a compute warp handles per half (1: 16 compute warps, 2: 8 compute warps with twice the work each,
// i.e. the per-half barrier / fence / geometry cost amortised over 32 scores per thread)
template <bool DROP, int CQ>
__global__ void __launch_bounds__((16 / CQ + 2) * 32, 1) attn_core_bwd2_kernel(const Args a) {
    constexpr int NW = 16 / CQ, NTHK = (NW + 2) * 32;                        // compute warps; threads
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *sImg = smem;                                              // [2] dS' tile images (rows = keys, columns = queries)
    unsigned char *sQ = sImg + 2 * kTileImg2, *sK = sQ + kImg, *sV = sK + kImg, *sdO = sV + kImg;
    float *sStat = reinterpret_cast<float *>(sdO + kImg);                    // per pair of queries: [off, off, ndelta, ndelta], off = drop_log2scale
    uint64_t *bS = reinterpret_cast<uint64_t *>(sStat + 2 * R), *bI = bS + 2, *bP = bI + 2, *bF = bP + 2;      // - lse * log2e, ndelta = -delta / drop_scale
    uint32_t *tslot = reinterpret_cast<uint32_t *>(bF + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q4 = warp & 3, c4 = warp >> 2;                                 // c4 == NW / 4: the MMA warps (A, then B)
    const int bh = blockIdx.x, n = a.n, nt = (n + 127) >> 7;
    constexpr uint32_t cST = 0, cDPT = 128, cPT = 256, cDST = 320, cDV = 384, cDK = 416, cDQ = 448;
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(&bS[i], 1); mbar_init(&bI[i], NW); mbar_init(&bP[i], 1); }
        mbar_init(bF, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(tslot, 512);
    const size_t base = (size_t)bh * n * DH;
    {
        unsigned char *const imgs[4] = {sQ, sK, sV, sdO};
        const __nv_bfloat16 *const srcs[4] = {a.q + base, a.k + base, a.v + base, a.d_o + base};
        for (int i = tid; i < nt * 128 * 2; i += NTHK) {
            const int r = i >> 1, c = i & 1;
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                const uint4 v = r < n ? __ldg(reinterpret_cast<const uint4 *>(srcs[m] + (size_t)r * DH) + c) : make_uint4(0, 0, 0, 0);
                *reinterpret_cast<uint4 *>(imgs[m] + c * kChunk + (r >> 3) * 128 + (r & 7) * 16) = v;
            }
        }
    }
    for (int r = tid; r < nt * 128; r += NTHK) {                             // delta = rowsum(dO . O); logsumexp in log2 units
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
        sStat[(r >> 1) * 4 + (r & 1)] = a.drop_log2scale - L;
        sStat[(r >> 1) * 4 + 2 + (r & 1)] = -d / a.drop_scale;
    }
    warm_next_head(a, bh + a.ahead, gridDim.x, tid, NTHK, true);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tslot, tlane = tbase + ((uint32_t)(q4 * 32) << 16);
    const int total = nt * (nt + 1);                                         // half-steps: tiles (kt, qt >= kt) x 2 halves of 64 queries
    auto next_tile = [&](int &kt, int &qt) { if (++qt >= nt) { ++kt; qt = kt; } };

    if (warp == NW) {
        // ---- MMA warp A: S^T / dP~^T of every half (two halves ahead of the threads), dQ += dS' K per finished tile ----
        if (lane == 0) {
            const ImgDesc dQ_(sQ), dK_(sK), dV_(sV), dO_(sdO);
            const StepDesc im0(sImg);
            auto issue_s = [&](int kt, int qt, int h) {
                mma_bf16_ss(tbase + cST + h * 64, dK_.rows + kt * 128, dQ_.rows + qt * 128 + h * 64, idesc_bf16(128, 64), 0);
                mma_bf16_ss(tbase + cDPT + h * 64, dV_.rows + kt * 128, dO_.rows + qt * 128 + h * 64, idesc_bf16(128, 64), 0);
                mma_commit(&bS[h]);
            };
            int kt = 0, qt = 0, ikt = 0, iqt = 0;                            // tile of half i; tile of the half being issued (i + 2)
            tc_fence_after();
            issue_s(0, 0, 0);
            issue_s(0, 0, 1);
            next_tile(ikt, iqt);
            for (int i = 0; i < total; ++i) {
                const int h = i & 1;
                mbar_wait_sleep(&bI[h], (i >> 1) & 1, 1000);
                tc_fence_after();
                if (h == 1) {                                                // both halves of the tile are in the image
                    const uint64_t am = im0.mnmaj + ((i >> 1) & 1) * (kTileImg2 / 16), kc = dK_.cols + kt * 128;
#pragma unroll
                    for (int s = 0; s < 8; ++s)                               // K = 128 keys
                        mma_bf16_ss(tbase + cDQ + qt * DH, am + s * 16, kc + s * 16, idesc_bf16(128, DH) | kIdA | kIdB, kt > 0 || s > 0);
                }
                if (i + 2 < total) issue_s(ikt, iqt, h);
                if (h == 1) { next_tile(kt, qt); next_tile(ikt, iqt); }
            }
            mma_commit(bF);
        }
    } else if (warp == NW + 1) {
        // ---- MMA warp B: dV += P~^T dO and dK += dS'^T Q of every half, A operands from tensor memory ----
        if (lane == 0) {
            const ImgDesc dQ_(sQ), dO_(sdO);
            int kt = 0, qt = 0;
            tc_fence_after();
            for (int i = 0; i < total; ++i) {
                const int h = i & 1;
                mbar_wait_sleep(&bI[h], (i >> 1) & 1, 1000);
                tc_fence_after();
                const bool first = qt == kt && h == 0;
                const uint64_t doc = dO_.cols + qt * 128 + h * 64, qc = dQ_.cols + qt * 128 + h * 64;
                const uint32_t tdv = tbase + cDV + (kt & 1) * DH, tdk = tbase + cDK + (kt & 1) * DH;
#pragma unroll
                for (int s = 0; s < 4; ++s) {                                 // K = 64 queries
                    mma_bf16_ts(tdv, tbase + cPT + h * 32 + s * 8, doc + s * 16, idesc_bf16(128, DH) | kIdB, !first || s > 0);
                    mma_bf16_ts(tdk, tbase + cDST + h * 32 + s * 8, qc + s * 16, idesc_bf16(128, DH) | kIdB, !first || s > 0);
                }
                mma_commit(&bP[h]);
                if (h == 1) next_tile(kt, qt);
            }
        }
    } else {
        // ---- compute warps: lane = key row (q4 * 32 + lane of the key tile), c4 = group of 16 queries of the half ----
        const uint32_t ckey = cta_key(a.rng_step, a.rng_key, (uint32_t)bh);
        const float sl2 = a.scale * kLog2e;
        const int key_local = q4 * 32 + lane;
        auto write_dkv = [&](int kt) {                                       // dK (c4 == 0) and dV (c4 == 1) rows of key tile kt
            if (c4 < 2) {
                float v[16];
                tmem_ld16(tlane + (c4 == 0 ? cDK : cDV) + (kt & 1) * DH, v);
                const int krow = kt * 128 + key_local;
                if (krow < n) store_row16((c4 == 0 ? a.dk : a.dv) + base + (size_t)krow * DH, v, c4 == 0 ? a.scale : 1.f);
            }
        };
        static_assert(NW / 4 >= 2, "dK / dV leave through two column groups of warps");
        int kt = 0, qt = 0;
        for (int i = 0; i < total; ++i) {
            const int h = i & 1;
            mbar_wait_sleep(&bS[h], (i >> 1) & 1, kWaitNs);
            tc_fence_after();
            if (h == 0 && qt == kt && kt > 0) {                              // a key tile has finished: its dK, dV leave tensor memory
                mbar_wait(&bP[1], ((i - 1) >> 1) & 1);
                tc_fence_after();
                write_dkv(kt - 1);
            }
            const bool diag = qt == kt;
            uint32_t wp[CQ][8], wd[CQ][8];
            const uint32_t hbase = (uint32_t)(kt * 32 + q4 * 8 + (lane >> 2)) * kGolden + ckey;
#pragma unroll
            for (int cc = 0; cc < CQ; ++cc) {
                const int q_local0 = h * 64 + (c4 * CQ + cc) * 16;           // first query of the chunk inside the query tile
                const uint32_t q_abs0 = (uint32_t)(qt * 128 + q_local0);
                const uint32_t ts = tlane + cST + h * 64 + (c4 * CQ + cc) * 16, tdp = tlane + cDPT + h * 64 + (c4 * CQ + cc) * 16;
                const float4 *stat = reinterpret_cast<const float4 *>(sStat) + (q_abs0 >> 1);
                if (!diag || q_local0 >= q4 * 32 + 31) bwd2_chunk<false, DROP>(ts, tdp, stat, key_local, q_local0, sl2, hbase, q_abs0, a.drop_add4, lane, wp[cc], wd[cc]);
                else if (q_local0 + 15 >= q4 * 32) bwd2_chunk<true, DROP>(ts, tdp, stat, key_local, q_local0, sl2, hbase, q_abs0, a.drop_add4, lane, wp[cc], wd[cc]);
                else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) { wp[cc][j] = 0u; wd[cc][j] = 0u; }
                }
            }
            if (i >= 2) { mbar_wait_sleep(&bP[h], ((i - 2) >> 1) & 1, kWaitNs); tc_fence_after(); }      // the products that read these TMEM operands are done
#pragma unroll
            for (int cc = 0; cc < CQ; ++cc) {
                const int q_local0 = h * 64 + (c4 * CQ + cc) * 16;
                tmem_st8(tlane + cPT + h * 32 + (c4 * CQ + cc) * 8, wp[cc]);
                tmem_st8(tlane + cDST + h * 32 + (c4 * CQ + cc) * 8, wd[cc]);
                // dS' also goes to the tile image (rows = keys, columns = queries) for dQ += dS' K
                unsigned char *img = sImg + ((i >> 1) & 1) * kTileImg2 + (q_local0 >> 3) * kTCore + (key_local >> 3) * 128 + (key_local & 7) * 16;
                *reinterpret_cast<uint4 *>(img) = make_uint4(wd[cc][0], wd[cc][1], wd[cc][2], wd[cc][3]);
                *reinterpret_cast<uint4 *>(img + kTCore) = make_uint4(wd[cc][4], wd[cc][5], wd[cc][6], wd[cc][7]);
            }
            tmem_st_wait();
            images_ready(&bI[h], lane);
            if (h == 1) next_tile(kt, qt);
        }
        mbar_wait(&bP[1], ((total - 1) >> 1) & 1);                            // the last products (dK, dV of the last key tile)
        mbar_wait(bF, 0);                                                     // the last dQ products
        tc_fence_after();
        write_dkv(nt - 1);
        for (int t = c4; t < nt; t += NW / 4) {                               // dQ of query tile t (lane = query row)
            float v[16];
            tmem_ld16(tlane + cDQ + t * DH, v);
            const int qrow = t * 128 + key_local;
            if (qrow < n) store_row16(a.dq + base + (size_t)qrow * DH, v, a.scale);
        }
        tc_fence_before();
    }
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
    const uint32_t thr = (uint32_t)(io->dropout * 128.f + 0.5f);              // the drop probability is quantised to 1/128
    a.drop_add4 = thr ? (128u - thr) * 0x01010101u : 0u;
    a.drop_scale = 1.f / (1.f - (float)thr / 128.f);
    a.drop_log2scale = log2f(a.drop_scale);
    a.scale = io->scale; a.n = n; a.ahead = num_sms();
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
    if (!attr) {
        XTRL_CUDA(cudaFuncSetAttribute(attn_core_fwd2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwd2Smem));
        XTRL_CUDA(cudaFuncSetAttribute(attn_core_fwd2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwd2Smem));
        attr = true;
    }
    if (a.drop_add4) attn_core_fwd2_kernel<true><<<bh, NTH, kFwd2Smem, (cudaStream_t)stream>>>(a);
    else attn_core_fwd2_kernel<false><<<bh, NTH, kFwd2Smem, (cudaStream_t)stream>>>(a);
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
    if (!attr) {
        XTRL_CUDA(cudaFuncSetAttribute(attn_core_bwd2_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwd2Smem));
        XTRL_CUDA(cudaFuncSetAttribute(attn_core_bwd2_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwd2Smem));
        attr = true;
    }
    if (a.drop_add4) attn_core_bwd2_kernel<true, 1><<<bh, NTH2, kBwd2Smem, (cudaStream_t)stream>>>(a);
    else attn_core_bwd2_kernel<false, 1><<<bh, NTH2, kBwd2Smem, (cudaStream_t)stream>>>(a);
    XTRL_LAUNCH_CHECK("attn_core_bwd_kernel");
    return XTRL_OK;
}
