// ppo_loss.cu -- fused PPO clipped-surrogate + entropy + HL-Gauss clipped value loss + world-model
// Gaussian NLL + done BCE: forward values AND gradients w.r.t. every head output, one pass.
//
// Replaces, on the rows of one minibatch (x_transformers_rl/x_transformers_rl.py):
//   compute_actor_loss :413-444 (Discrete :197-222, normalize :103-112),
//   compute_critic_loss :446-477 (HLGaussLoss; per-element, U3),
//   compute_autoregressive_loss :398-404 (+ Continuous.mean_variance :232-241, softclamp :124),
//   compute_done_loss :406-411, and the assembly + backward of Agent.learn :939-983.
//
// HBM-bound: per timestep 8A+8K+21 B (+20(S+1)+9 B for the world-model terms), each byte touched
// once.  One warp per timestep row: the K critic logits and their gradients move as coalesced
// 128 B rows (lane <-> bin, stride 32), bin-edge erf values are shared between neighbouring lanes
// by shuffle, and erf is only evaluated where fp32 erf is not saturated.
// Launch sequence (all on the caller's stream): memset workspace -> stats (masked advantage
// mean / unbiased variance, counts; fp64 accumulators) -> main -> finalize (8 scalars).
#include "common.cuh"
#include <math.h>

namespace xtrl {

enum { WS_ADV_SUM = 0, WS_ADV_SQ, WS_N_MASK, WS_N_WM, WS_AC_MASKED, WS_ACTOR_ALL, WS_CRITIC_ALL,
       WS_NLL, WS_DONE, WS_COUNT };

template <typename T> __device__ __forceinline__ float ld_as_float(const T *p);
template <> __device__ __forceinline__ float ld_as_float<float>(const float *p) { return __ldg(p); }
template <> __device__ __forceinline__ float ld_as_float<__nv_bfloat16>(const __nv_bfloat16 *p) {
    return __bfloat162float(*p);
}
template <typename T> __device__ __forceinline__ void st_from_float(T *p, float v);
template <> __device__ __forceinline__ void st_from_float<float>(float *p, float v) { *p = v; }
template <> __device__ __forceinline__ void st_from_float<__nv_bfloat16>(__nv_bfloat16 *p, float v) {
    *p = __float2bfloat16_rn(v);
}

struct HlSupport {           // torch.linspace(lo, hi, K+1) in fp32 (two-sided formula)
    float lo, hi, step, denom;   // denom = sqrt(2) * sigma, sigma = 2 * (edge1 - edge0)
    int K;
    __device__ __forceinline__ float edge(int i) const {
        return (i < (K + 1) / 2) ? lo + step * (float)i : hi - step * (float)(K - i);
    }
    __device__ __forceinline__ float centre(int k) const { return (edge(k) + edge(k + 1)) * 0.5f; }
};

static HlSupport make_support(float lo, float hi, int K) {
    HlSupport s;
    s.lo = lo; s.hi = hi; s.K = K;
    s.step = (hi - lo) / (float)K;
    const float e1 = (1 < (K + 1) / 2) ? lo + s.step * 1.f : hi - s.step * (float)(K - 1);
    const double sigma = 2.0 * (double)(e1 - lo);
    s.denom = (float)(sqrt(2.0) * sigma);
    return s;
}

// ---- stats: masked sum / sum of squares of the advantages, counts --------------------------
__global__ void __launch_bounds__(256)
ppo_stats_kernel(const float *__restrict__ returns, const float *__restrict__ old_values,
                 const int32_t *__restrict__ lens, int b, int T, double *__restrict__ ws) {
    double s = 0., q = 0., n = 0., nw = 0.;
    const int64_t N = (int64_t)b * T;
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < N; r += (int64_t)gridDim.x * blockDim.x) {
        const int e = (int)(r / T), t = (int)(r - (int64_t)e * T);
        if (t < __ldg(lens + e)) {
            const double a = (double)(__ldg(returns + r) - __ldg(old_values + r));
            s += a; q += a * a; n += 1.;
            if (t < T - 1) nw += 1.;
        }
    }
    __shared__ double sh[4][8];
    s = warp_sum(s); q = warp_sum(q); n = warp_sum(n); nw = warp_sum(nw);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { sh[0][w] = s; sh[1][w] = q; sh[2][w] = n; sh[3][w] = nw; }
    __syncthreads();
    if (threadIdx.x < 4) {
        double a = 0.;
        for (int i = 0; i < 8; ++i) a += sh[threadIdx.x][i];
        const int slot = threadIdx.x == 0 ? WS_ADV_SUM : threadIdx.x == 1 ? WS_ADV_SQ
                       : threadIdx.x == 2 ? WS_N_MASK : WS_N_WM;
        atomicAdd(ws + slot, a);
    }
}

// ---- main ----------------------------------------------------------------------------------
struct LossArgs {
    xtrl_loss_cfg cfg;
    xtrl_loss_io io;
    HlSupport sup;
    int b, T;
};

constexpr int kLossWarps = 8;
constexpr int kTileRows = 64;            // rows (timesteps) staged per tile; 8 per warp

// ---- asynchronous tile staging (cp.async 16 B, zero-fill past the end of the source) -----------
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc, int src_bytes) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

// copies bytes [start, start+want) of a buffer of `total` bytes into smem (16-byte chunks; everything
// past `total` reads as zero).  start is a multiple of 16 by construction (tile starts are 64 rows).
__device__ __forceinline__ void stage_region(unsigned char *dst, const void *base, int64_t start, int want, int64_t total) {
    if (base == nullptr) return;
    const unsigned char *src = (const unsigned char *)base;
    for (int c = threadIdx.x; c * 16 < want; c += kLossWarps * 32) {
        const int64_t off = start + (int64_t)c * 16;
        const int64_t left = total - off;
        const int valid = left >= 16 ? 16 : (left > 0 ? (int)left : 0);
        cp_async16(dst + c * 16, valid > 0 ? src + off : src, valid);
    }
}

struct TileLayout {                      // byte offsets inside one stage of shared memory
    int vl, ra, sp, dl, act, olp, ret, ov, tgt, dn, total;
};
__host__ __device__ static inline TileLayout tile_layout(int AW, int AL, int K, int S1, int sz) {   // AW: actor head width, AL: actions per row
    TileLayout t; int o = 0;
    auto take = [&](int n) { int at = o; o += (n + 15) & ~15; return at; };
    t.vl = take(kTileRows * K * sz);
    t.ra = take(kTileRows * AW * sz);
    t.sp = take(kTileRows * 2 * S1 * sz);
    t.dl = take(kTileRows * sz);
    t.act = take(kTileRows * 4 * AL);
    t.olp = take(kTileRows * 4 * AL);
    t.ret = take(kTileRows * 4);
    t.ov = take(kTileRows * 4);
    t.tgt = take((kTileRows + 1) * S1 * 4);
    t.dn = take(kTileRows);
    t.total = o;
    return t;
}

template <typename T> __device__ __forceinline__ float sm_as_float(const T *p);
template <> __device__ __forceinline__ float sm_as_float<float>(const float *p) { return *p; }
template <> __device__ __forceinline__ float sm_as_float<__nv_bfloat16>(const __nv_bfloat16 *p) { return __bfloat162float(*p); }

// erf with the fp32-saturated region (|z| >= 4 rounds to +-1 exactly) skipped warp-uniformly
__device__ __forceinline__ float erf_window(float z) {
    const bool need = fabsf(z) < 4.f;
    float c = copysignf(1.f, z);
    if (__any_sync(0xffffffffu, need)) { const float e = erff(z); c = need ? e : c; }
    return c;
}

template <typename TIO, int KJ>
__global__ void __launch_bounds__(kLossWarps * 32, 2)
ppo_loss_kernel(const LossArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int A = a.cfg.num_actions, K = a.cfg.num_bins, S1 = a.cfg.state_dim + 1, T = a.T;
    const int64_t N = (int64_t)a.b * T;
    const double *ws = a.io.workspace;
    const bool cont = a.cfg.continuous != 0;
    const int AW = cont ? 2 * A : A, AL = cont ? A : 1;
    const TileLayout tl = tile_layout(AW, AL, K, S1, (int)sizeof(TIO));

    // masked advantage statistics (normalize(), :103-112)
    const double n_mask = ws[WS_N_MASK], n_wm = ws[WS_N_WM];
    float adv_mean = 0.f, adv_inv_std = 1.f;
    if (a.cfg.normalize_advantages && n_mask > 0.) {
        const double mean = ws[WS_ADV_SUM] / n_mask;
        const double var = (ws[WS_ADV_SQ] - n_mask * mean * mean) / (n_mask - 1.);   // unbiased
        adv_mean = (float)mean;
        adv_inv_std = 1.f / sqrtf(fmaxf((float)var, 1e-5f));
    }
    const float g_ac = n_mask > 0. ? (float)(1. / n_mask) : 0.f;                      // d mean / d elem
    const float g_done = g_ac * a.cfg.autoreg_weight;
    const float g_wm = n_wm > 0. ? a.cfg.autoreg_weight * (float)(1. / (n_wm * S1)) : 0.f;
    const float feps = 1.1920928955078125e-07f;                                      // finfo(fp32).eps

    TIO *g_ra = (TIO *)a.io.g_raw_actions;
    TIO *g_vl = (TIO *)a.io.g_value_logits;
    TIO *g_sp = (TIO *)a.io.g_state_pred_raw;
    TIO *g_dl = (TIO *)a.io.g_done_logit;
    const bool has_sp = a.io.state_pred_raw != nullptr, has_dl = a.io.done_logit != nullptr;

    const int64_t ntiles = (N + kTileRows - 1) / kTileRows;
    auto stage = [&](int64_t tile, int buf) {
        unsigned char *st = smem + (size_t)buf * tl.total;
        const int64_t r0 = tile * kTileRows;
        const int sz = (int)sizeof(TIO);
        stage_region(st + tl.vl, a.io.value_logits, r0 * K * sz, kTileRows * K * sz, N * K * sz);
        stage_region(st + tl.ra, a.io.raw_actions, r0 * AW * sz, kTileRows * AW * sz, N * AW * sz);
        stage_region(st + tl.sp, a.io.state_pred_raw, r0 * 2 * S1 * sz, kTileRows * 2 * S1 * sz, N * 2 * S1 * sz);
        stage_region(st + tl.dl, a.io.done_logit, r0 * sz, kTileRows * sz, N * sz);
        stage_region(st + tl.act, cont ? (const void *)a.io.actions_f : (const void *)a.io.actions, r0 * 4 * AL, kTileRows * 4 * AL, N * 4 * AL);
        stage_region(st + tl.olp, a.io.old_log_probs, r0 * 4 * AL, kTileRows * 4 * AL, N * 4 * AL);
        stage_region(st + tl.ret, a.io.returns, r0 * 4, kTileRows * 4, N * 4);
        stage_region(st + tl.ov, a.io.old_values, r0 * 4, kTileRows * 4, N * 4);
        stage_region(st + tl.tgt, a.io.target, r0 * S1 * 4, (kTileRows + 1) * S1 * 4, N * S1 * 4);
        stage_region(st + tl.dn, a.io.dones, r0, kTileRows, N);
        cp_async_commit();
    };

    float acc_ac = 0.f, acc_actor = 0.f, acc_critic = 0.f, acc_nll = 0.f, acc_done = 0.f;   // per THREAD partial sums
    __shared__ float s_crit[kTileRows];

    // per-lane constants of my KJ bins: lower edge and centre (lane <-> bin k = j*32 + lane)
    float edge_lo[KJ], centre[KJ];
#pragma unroll
    for (int j = 0; j < KJ; ++j) {
        const int k = j * 32 + lane;
        edge_lo[j] = a.sup.edge(min(k, K));
        centre[j] = k < K ? a.sup.centre(k) : 0.f;
    }
    const float inv_denom = 1.f / a.sup.denom;

    int buf = 0;
    if ((int64_t)blockIdx.x < ntiles) stage(blockIdx.x, 0);
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, buf ^= 1) {
        const int64_t next = tile + gridDim.x;
        if (next < ntiles) { stage(next, buf ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
        __syncthreads();
        const unsigned char *st = smem + (size_t)buf * tl.total;
        const TIO *s_vl = (const TIO *)(st + tl.vl), *s_ra = (const TIO *)(st + tl.ra);
        const TIO *s_sp = (const TIO *)(st + tl.sp), *s_dl = (const TIO *)(st + tl.dl);
        const int32_t *s_act = (const int32_t *)(st + tl.act);
        const float *s_olp = (const float *)(st + tl.olp), *s_ret = (const float *)(st + tl.ret);
        const float *s_ov = (const float *)(st + tl.ov), *s_tgt = (const float *)(st + tl.tgt);
        const uint8_t *s_dn = (const uint8_t *)(st + tl.dn);

        for (int lr = warp; lr < kTileRows; lr += kLossWarps) {
            const int64_t r = tile * kTileRows + lr;
            if (r >= N) break;
            const int e = (int)(r / T), t = (int)(r - (int64_t)e * T);
            const bool masked = t < __ldg(a.io.episode_lens + e);
            const float ret = s_ret[lr];
            const float v_old = s_ov[lr];

            // ---------------- critic: HL-Gauss CE against returns and clamped returns ------------
            float lg[KJ], ex[KJ];
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < KJ; ++j) {
                const int k = j * 32 + lane;
                lg[j] = k < K ? sm_as_float(s_vl + lr * K + k) : -INFINITY;
                mx = fmaxf(mx, lg[j]);
            }
            mx = warp_max(mx);
            float se = 0.f, sc = 0.f;
#pragma unroll
            for (int j = 0; j < KJ; ++j) {
                const int k = j * 32 + lane;
                ex[j] = k < K ? expf(lg[j] - mx) : 0.f;
                se += ex[j];
                sc = fmaf(ex[j], centre[j], sc);
            }
            se = warp_sum(se); sc = warp_sum(sc);
            const float v_new = sc / se;                       // hl_gauss(values), :455
            const float lse = logf(se) + mx;

            const float y0 = fminf(fmaxf(ret, a.sup.lo), a.sup.hi);                                     // clamp_to_range
            const float y1 = fminf(fmaxf(fminf(fmaxf(ret, -a.cfg.value_clip), a.cfg.value_clip), a.sup.lo), a.sup.hi);  // :459
            const bool same = y0 == y1;                        // returns inside +-value_clip: one histogram serves both
            // normalisers: lanes 0..3 evaluate erf at (hi,y0) (lo,y0) (hi,y1) (lo,y1), then broadcast
            const float zarg = ((lane & 1) ? a.sup.lo : a.sup.hi) - ((lane & 2) ? y1 : y0);
            const float zerf = erff(zarg * inv_denom);
            const float iz0 = 1.f / fmaxf(__shfl_sync(0xffffffffu, zerf, 0) - __shfl_sync(0xffffffffu, zerf, 1), 1e-10f);
            const float iz1 = 1.f / fmaxf(__shfl_sync(0xffffffffu, zerf, 2) - __shfl_sync(0xffffffffu, zerf, 3), 1e-10f);
            // cdf at the lower edge of each of my bins; the upper edge comes from the next lane
            float c0[KJ], c1[KJ];
#pragma unroll
            for (int j = 0; j < KJ; ++j) {
                c0[j] = erf_window((edge_lo[j] - y0) * inv_denom);
                c1[j] = same ? c0[j] : erf_window((edge_lo[j] - y1) * inv_denom);
            }
            float p0[KJ], p1[KJ];
            float ce0 = 0.f, ce1 = 0.f, sp0 = 0.f, sp1 = 0.f;
#pragma unroll
            for (int j = 0; j < KJ; ++j) {
                const int k = j * 32 + lane;
                // upper edge k+1: lane+1 of this j, or lane 0 of j+1
                float u0 = __shfl_down_sync(0xffffffffu, c0[j], 1);
                float u1 = __shfl_down_sync(0xffffffffu, c1[j], 1);
                const float n0 = __shfl_sync(0xffffffffu, (j + 1 < KJ) ? c0[(j + 1 < KJ) ? j + 1 : j] : 0.f, 0);
                const float n1 = __shfl_sync(0xffffffffu, (j + 1 < KJ) ? c1[(j + 1 < KJ) ? j + 1 : j] : 0.f, 0);
                if (lane == 31) { u0 = n0; u1 = n1; }
                if (k + 1 == K && (K & 31) == 0) {               // the last edge has no lane of its own
                    u0 = erff((a.sup.hi - y0) * inv_denom);
                    u1 = erff((a.sup.hi - y1) * inv_denom);
                }
                p0[j] = k < K ? (u0 - c0[j]) * iz0 : 0.f;
                p1[j] = k < K ? (u1 - c1[j]) * iz1 : 0.f;
                const float lsm = k < K ? lg[j] - lse : 0.f;
                ce0 -= p0[j] * lsm; ce1 -= p1[j] * lsm;
                sp0 += p0[j]; sp1 += p1[j];
            }
            ce0 = warp_sum(ce0); sp0 = warp_sum(sp0);
            if (same) { ce1 = ce0; sp1 = sp0; } else { ce1 = warp_sum(ce1); sp1 = warp_sum(sp1); }
            const float lo_old = v_old - a.cfg.value_clip, hi_old = v_old + a.cfg.value_clip;
            const bool zero = (ret < v_new && v_new < lo_old) || (hi_old < v_new && v_new < ret);   // :470-472
            const float critic = zero ? 0.f : fminf(ce0, ce1);
            if (g_vl) {
                // d min(ce0, ce1): the smaller one; an exact tie splits evenly (torch minimum)
                const float w0 = zero ? 0.f : (ce0 < ce1 ? 1.f : (ce0 > ce1 ? 0.f : 0.5f));
                const float w1 = zero ? 0.f : 1.f - w0;
                const float gc = masked ? g_ac * a.cfg.critic_weight : 0.f;
                const float inv_se = 1.f / se;
#pragma unroll
                for (int j = 0; j < KJ; ++j) {
                    const int k = j * 32 + lane;
                    if (k < K) {
                        const float sm = ex[j] * inv_se;
                        const float g = gc * (w0 * (sm * sp0 - p0[j]) + w1 * (sm * sp1 - p1[j]));
                        st_from_float(g_vl + r * K + k, g);
                    }
                }
            }

            if (lane == 0) s_crit[lr] = critic;
        }
        __syncthreads();

        // ---- phase B: the narrow terms, one THREAD per (row, part) instead of one warp per row -------------
        // part 0: actor (A logits), part 1: world-model NLL over the S+1 dims, part 2: done BCE.  A warp now
        // finishes 32 rows of a part per pass; these terms cost ~16 warp-instructions per row instead of ~430.
        for (int item = threadIdx.x; item < kTileRows * 3; item += kLossWarps * 32) {
            const int part = item / kTileRows, lr = item - part * kTileRows;
            const int64_t r = tile * kTileRows + lr;
            if (r >= N) continue;
            const int e = (int)(r / T), t = (int)(r - (int64_t)e * T);
            const bool masked = t < __ldg(a.io.episode_lens + e);
            if (part == 0 && cont) {
                // actor, Continuous (:224-277, :413-444): per action dim, summed over dims (:442)
                const float ret = s_ret[lr], v_old = s_ov[lr];
                const float adv = ((ret - v_old) - adv_mean) * adv_inv_std;
                const float ga = masked ? g_ac * a.cfg.actor_weight : 0.f;
                const float *s_af = (const float *)s_act;
                float actor = 0.f;
                for (int i = 0; i < A; ++i) {
                    const float mean = sm_as_float(s_ra + (lr * A + i) * 2), lv = sm_as_float(s_ra + (lr * A + i) * 2 + 1);
                    const float th = tanhf(lv / 3.f), var = expf(th * 3.f);
                    const bool inside = var > 1e-5f;
                    const float sd = sqrtf(fmaxf(var, 1e-5f));
                    const float v = s_af[lr * A + i], df = v - mean;
                    float logp = -(df * df) / (2.f * sd * sd) - logf(sd) - 0.91893853320467274178f;
                    if (a.cfg.squash) logp -= logf(fmaxf(1.f - v * v, 1e-20f));
                    const float ent = a.cfg.squash ? -logp : 1.41893853320467274178f + logf(sd);   // 0.5 + 0.5 log(2 pi) + log sd
                    const float ratio = expf(logp - s_olp[lr * A + i]);
                    const float clipped = fminf(fmaxf(ratio, 1.f - a.cfg.eps_clip), 1.f + a.cfg.eps_clip);
                    const float surr1 = ratio * adv, surr2 = clipped * adv;
                    actor += -fminf(surr1, surr2) - a.cfg.entropy_weight * ent;
                    if (g_ra) {
                        const bool in_range = ratio >= 1.f - a.cfg.eps_clip && ratio <= 1.f + a.cfg.eps_clip;
                        const float dmin = in_range ? adv : (surr1 < surr2 ? adv : 0.f);
                        const float g_lp = ga * (-dmin * ratio + (a.cfg.squash ? a.cfg.entropy_weight : 0.f));
                        const float g_mean = g_lp * df / (sd * sd);
                        float g_sd = g_lp * (df * df / (sd * sd * sd) - 1.f / sd);
                        if (!a.cfg.squash) g_sd -= ga * a.cfg.entropy_weight / sd;
                        const float g_lv = inside ? g_sd * (0.5f / sd) * var * (1.f - th * th) : 0.f;
                        st_from_float(g_ra + (r * A + i) * 2, g_mean);
                        st_from_float(g_ra + (r * A + i) * 2 + 1, g_lv);
                    }
                }
                const float critic = s_crit[lr];
                if (a.io.per_step) { a.io.per_step[r * 2] = actor; a.io.per_step[r * 2 + 1] = critic; }
                acc_actor += actor; acc_critic += critic;
                if (masked) acc_ac += a.cfg.actor_weight * actor + a.cfg.critic_weight * critic;
            } else if (part == 0) {
                // actor: clipped surrogate + entropy (:413-444, Discrete :197-222)
                const int act = s_act[lr];
                const float ret = s_ret[lr], v_old = s_ov[lr];
                float amx = -INFINITY;
                for (int i = 0; i < A; ++i) amx = fmaxf(amx, sm_as_float(s_ra + lr * A + i));
                float qs = 0.f;
                for (int i = 0; i < A; ++i) qs += expf(sm_as_float(s_ra + lr * A + i) - amx);
                float ssum = 0.f;
                for (int i = 0; i < A; ++i) ssum += expf(sm_as_float(s_ra + lr * A + i) - amx) / qs;   // Categorical renormalises
                float ent = 0.f, logp = 0.f, p_act = 1.f, dot_p = 0.f;
                bool in_act = false;
                for (int i = 0; i < A; ++i) {
                    const float sm = expf(sm_as_float(s_ra + lr * A + i) - amx) / qs, p = sm / ssum;
                    const float L = logf(fminf(fmaxf(p, feps), 1.f - feps));
                    ent -= p * L;
                    if (i == act) { logp = L; p_act = p; in_act = p >= feps && p <= 1.f - feps; }
                }
                const float ratio = expf(logp - s_olp[lr]);
                const float clipped = fminf(fmaxf(ratio, 1.f - a.cfg.eps_clip), 1.f + a.cfg.eps_clip);
                const float adv = ((ret - v_old) - adv_mean) * adv_inv_std;
                const float surr1 = ratio * adv, surr2 = clipped * adv;
                const float actor = -fminf(surr1, surr2) - a.cfg.entropy_weight * ent;
                const float critic = s_crit[lr];
                if (a.io.per_step) { a.io.per_step[r * 2] = actor; a.io.per_step[r * 2 + 1] = critic; }
                acc_actor += actor; acc_critic += critic;
                if (masked) acc_ac += a.cfg.actor_weight * actor + a.cfg.critic_weight * critic;
                if (g_ra) {
                    const bool in_range = ratio >= 1.f - a.cfg.eps_clip && ratio <= 1.f + a.cfg.eps_clip;
                    const float dmin = in_range ? adv : (surr1 < surr2 ? adv : 0.f);
                    const float ga = masked ? g_ac * a.cfg.actor_weight : 0.f;
                    const float g_logp = -ga * dmin * ratio;
                    // gp_i = d/dp_i of (-beta*ent + logp term);  p = sm / sum(sm);  sm = softmax(logits)
                    auto gp_of = [&](int i, float p, float L) {
                        const bool inside = p >= feps && p <= 1.f - feps;
                        float gp = ga * a.cfg.entropy_weight * (L + (inside ? 1.f : 0.f));
                        if (i == act && in_act) gp += g_logp / p_act;
                        return gp;
                    };
                    float dot_q = 0.f;
                    for (int i = 0; i < A; ++i) {
                        const float sm = expf(sm_as_float(s_ra + lr * A + i) - amx) / qs, p = sm / ssum;
                        dot_p += gp_of(i, p, logf(fminf(fmaxf(p, feps), 1.f - feps))) * p;
                    }
                    for (int i = 0; i < A; ++i) {
                        const float sm = expf(sm_as_float(s_ra + lr * A + i) - amx) / qs, p = sm / ssum;
                        const float gq = (gp_of(i, p, logf(fminf(fmaxf(p, feps), 1.f - feps))) - dot_p) / ssum;
                        dot_q += gq * sm;
                    }
                    for (int i = 0; i < A; ++i) {
                        const float sm = expf(sm_as_float(s_ra + lr * A + i) - amx) / qs, p = sm / ssum;
                        const float gq = (gp_of(i, p, logf(fminf(fmaxf(p, feps), 1.f - feps))) - dot_p) / ssum;
                        st_from_float(g_ra + r * A + i, sm * (gq - dot_q));
                    }
                }
            } else if (part == 1) {
                // world model: Gaussian NLL of next (state || reward) (:398-404, :232-241)
                if (has_sp) {
                    const bool valid = masked && t < T - 1;
                    float nll = 0.f;
                    for (int dd = 0; dd < S1; ++dd) {
                        const float mean = sm_as_float(s_sp + (lr * S1 + dd) * 2);
                        const float lv = sm_as_float(s_sp + (lr * S1 + dd) * 2 + 1);
                        float gm = 0.f, gl = 0.f;
                        if (valid) {
                            const float th = tanhf(lv / 3.f);
                            const float var = fmaxf(expf(th * 3.f), 1e-6f);
                            const float d = mean - s_tgt[(lr + 1) * S1 + dd];
                            nll += 0.5f * (logf(var) + d * d / var);
                            gm = g_wm * d / var;
                            gl = g_wm * 0.5f * (1.f / var - d * d / (var * var)) * var * (1.f - th * th);
                        }
                        if (g_sp) {
                            st_from_float(g_sp + (r * S1 + dd) * 2, gm);
                            st_from_float(g_sp + (r * S1 + dd) * 2 + 1, gl);
                        }
                    }
                    acc_nll += nll;
                }
            } else if (has_dl) {
                // done head: BCE on sigmoid(logit) (:406-411)
                const float z = sm_as_float(s_dl + lr);
                const float pd = 1.f / (1.f + expf(-z));
                const float y = s_dn[lr] ? 1.f : 0.f;
                const float l1 = logf(pd), l0 = logf(1.f - pd);
                const float bce = -(y * fmaxf(l1, -100.f) + (1.f - y) * fmaxf(l0, -100.f));
                float g = 0.f;
                if (masked) {
                    acc_done += bce;
                    const float dp = -((l1 > -100.f ? y / pd : 0.f) - (l0 > -100.f ? (1.f - y) / (1.f - pd) : 0.f));
                    g = g_done * dp * pd * (1.f - pd);
                    if (!isfinite(g)) g = 0.f;
                }
                if (g_dl) st_from_float(g_dl + r, g);
            }
        }
        __syncthreads();                                           // this stage is refilled two tiles from now
    }

    acc_ac = warp_sum(acc_ac); acc_actor = warp_sum(acc_actor); acc_critic = warp_sum(acc_critic);
    acc_nll = warp_sum(acc_nll); acc_done = warp_sum(acc_done);
    __shared__ float sh[5][kLossWarps];
    if (lane == 0) { sh[0][warp] = acc_ac; sh[1][warp] = acc_actor; sh[2][warp] = acc_critic; sh[3][warp] = acc_nll; sh[4][warp] = acc_done; }
    __syncthreads();
    if (threadIdx.x < 5) {
        double s = 0.;
        for (int i = 0; i < kLossWarps; ++i) s += (double)sh[threadIdx.x][i];
        atomicAdd(a.io.workspace + WS_AC_MASKED + threadIdx.x, s);
    }
}

__global__ void ppo_finalize_kernel(const double *__restrict__ ws, float *__restrict__ scalars,
                                    int64_t N, int S1, float w_ar, int normalize) {
    const double n = ws[WS_N_MASK], nw = ws[WS_N_WM] * S1;
    const double ac = n > 0 ? ws[WS_AC_MASKED] / n : NAN;
    const double nll = nw > 0 ? ws[WS_NLL] / nw : NAN;
    const double dn = n > 0 ? ws[WS_DONE] / n : NAN;
    scalars[0] = (float)(ac + (nll + dn) * (double)w_ar);
    scalars[1] = (float)(ws[WS_ACTOR_ALL] / (double)N);
    scalars[2] = (float)(ws[WS_CRITIC_ALL] / (double)N);
    scalars[3] = (float)nll;
    scalars[4] = (float)dn;
    scalars[5] = (float)n;
    const double mean = n > 0 ? ws[WS_ADV_SUM] / n : 0.;
    scalars[6] = (float)mean;
    scalars[7] = (float)(n > 1 ? (ws[WS_ADV_SQ] - n * mean * mean) / (n - 1.) : NAN);
}

// ---- HL-Gauss decode -----------------------------------------------------------------------
__global__ void __launch_bounds__(256)
hl_decode_kernel(const float *__restrict__ logits, float *__restrict__ out, int64_t n, HlSupport sup) {
    const int lane = threadIdx.x & 31;
    const int K = sup.K;
    for (int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); r < n; r += (int64_t)gridDim.x * 8) {
        float mx = -INFINITY;
        for (int k = lane; k < K; k += 32) mx = fmaxf(mx, __ldg(logits + r * K + k));
        mx = warp_max(mx);
        float se = 0.f, sc = 0.f;
        for (int k = lane; k < K; k += 32) {
            const float ex = expf(__ldg(logits + r * K + k) - mx);
            se += ex; sc += ex * sup.centre(k);
        }
        se = warp_sum(se); sc = warp_sum(sc);
        if (lane == 0) out[r] = sc / se;
    }
}

template <typename TIO, int KJ>
static int launch_loss_kj(const LossArgs &la, int blocks, size_t smem, cudaStream_t s) {
    XTRL_CUDA(cudaFuncSetAttribute(ppo_loss_kernel<TIO, KJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ppo_loss_kernel<TIO, KJ><<<blocks, kLossWarps * 32, smem, s>>>(la);
    XTRL_LAUNCH_CHECK("ppo_loss_kernel");
    return XTRL_OK;
}

template <typename TIO>
static int launch_loss(const LossArgs &la, int KJ, int blocks, size_t smem, cudaStream_t s) {
    switch (KJ) {
        case 1: return launch_loss_kj<TIO, 1>(la, blocks, smem, s);
        case 2: return launch_loss_kj<TIO, 2>(la, blocks, smem, s);
        case 4: return launch_loss_kj<TIO, 4>(la, blocks, smem, s);
        default: return launch_loss_kj<TIO, 8>(la, blocks, smem, s);
    }
}

}  // namespace xtrl

extern "C" int64_t xtrl_ppo_loss_workspace_bytes(void) { return (int64_t)sizeof(double) * 16; }

extern "C" int xtrl_ppo_loss_fwd_bwd(const xtrl_loss_cfg *cfg, const xtrl_loss_io *io,
                                     int32_t batch, int32_t num_timesteps, void *stream) {
    using namespace xtrl;
    XTRL_REQUIRE(cfg && io, XTRL_E_BADARG, "xtrl_ppo_loss_fwd_bwd: null cfg/io");
    XTRL_REQUIRE(batch >= 0 && num_timesteps >= 0, XTRL_E_BADARG, "xtrl_ppo_loss_fwd_bwd: negative shape");
    XTRL_REQUIRE(cfg->num_actions >= 1 && cfg->num_actions <= 32, XTRL_E_UNSUPPORTED,
                 "xtrl_ppo_loss_fwd_bwd: num_actions=%d outside [1,32]", cfg->num_actions);
    XTRL_REQUIRE(cfg->num_bins >= 2 && cfg->num_bins <= 256, XTRL_E_UNSUPPORTED,
                 "xtrl_ppo_loss_fwd_bwd: num_bins=%d outside [2,256]", cfg->num_bins);
    XTRL_REQUIRE(cfg->state_dim >= 1 && cfg->state_dim + 1 <= 32, XTRL_E_UNSUPPORTED,
                 "xtrl_ppo_loss_fwd_bwd: state_dim=%d outside [1,31]", cfg->state_dim);
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t N = (int64_t)batch * num_timesteps;
    if (N == 0) {                                           // empty minibatch: zero scalars, nothing else (buffers may be empty)
        XTRL_REQUIRE(io->scalars, XTRL_E_BADARG, "xtrl_ppo_loss_fwd_bwd: null scalars");
        XTRL_CUDA(cudaMemsetAsync(io->scalars, 0, sizeof(float) * 8, s));
        return XTRL_OK;
    }
    XTRL_REQUIRE(io->raw_actions && io->value_logits && (cfg->continuous ? (const void *)io->actions_f : (const void *)io->actions) && io->old_log_probs && io->returns &&
                 io->old_values && io->episode_lens && io->scalars && io->workspace, XTRL_E_BADARG,
                 "xtrl_ppo_loss_fwd_bwd: null required pointer");
    XTRL_REQUIRE(!io->state_pred_raw || io->target, XTRL_E_BADARG, "xtrl_ppo_loss_fwd_bwd: state_pred_raw without target");
    XTRL_REQUIRE(!io->done_logit || io->dones, XTRL_E_BADARG, "xtrl_ppo_loss_fwd_bwd: done_logit without dones");
    XTRL_CUDA(cudaMemsetAsync(io->workspace, 0, sizeof(double) * 16, s));
    const int sblocks = (int)((N + 255) / 256 < 4 * num_sms() ? (N + 255) / 256 : 4 * num_sms());
    ppo_stats_kernel<<<sblocks, 256, 0, s>>>(io->returns, io->old_values, io->episode_lens, batch, num_timesteps, io->workspace);
    XTRL_LAUNCH_CHECK("ppo_stats_kernel");
    LossArgs la;
    la.cfg = *cfg; la.io = *io; la.b = batch; la.T = num_timesteps;
    la.sup = make_support(cfg->reward_min, cfg->reward_max, cfg->num_bins);
    const int kj = (cfg->num_bins + 31) / 32;
    const int KJ = kj <= 1 ? 1 : kj <= 2 ? 2 : kj <= 4 ? 4 : 8;
    const int sz = cfg->bf16_io ? 2 : 4;
    const int AWh = cfg->continuous ? 2 * cfg->num_actions : cfg->num_actions, ALh = cfg->continuous ? cfg->num_actions : 1;
    const TileLayout tl = tile_layout(AWh, ALh, cfg->num_bins, cfg->state_dim + 1, sz);
    const size_t smem = 2 * (size_t)tl.total;                              // double-buffered tile stages
    XTRL_REQUIRE(smem <= 200 * 1024, XTRL_E_UNSUPPORTED, "xtrl_ppo_loss_fwd_bwd: tile needs %zu B shared memory", smem);
    const void *ptrs[] = {io->raw_actions, io->value_logits, io->state_pred_raw, io->done_logit, io->actions, io->old_log_probs,
                          io->returns, io->old_values, io->target, io->dones, io->actions_f};
    for (const void *p : ptrs)
        XTRL_REQUIRE(((uintptr_t)p & 15) == 0, XTRL_E_BADARG, "xtrl_ppo_loss_fwd_bwd: input pointers must be 16-byte aligned");
    const int64_t ntiles = (N + kTileRows - 1) / kTileRows;
    const int per_sm = smem <= 100 * 1024 ? 2 : 1;
    const int blocks = (int)(ntiles < (int64_t)per_sm * num_sms() ? ntiles : (int64_t)per_sm * num_sms());   // persistent CTAs
    int rc = cfg->bf16_io ? launch_loss<__nv_bfloat16>(la, KJ, blocks, smem, s) : launch_loss<float>(la, KJ, blocks, smem, s);
    if (rc) return rc;
    ppo_finalize_kernel<<<1, 1, 0, s>>>(io->workspace, io->scalars, N, cfg->state_dim + 1, cfg->autoreg_weight,
                                        cfg->normalize_advantages);
    XTRL_LAUNCH_CHECK("ppo_finalize_kernel");
    return XTRL_OK;
}

extern "C" int xtrl_hl_decode(const float *logits, float *out, int64_t n, int32_t num_bins,
                              float reward_min, float reward_max, void *stream) {
    using namespace xtrl;
    XTRL_REQUIRE(n >= 0 && num_bins >= 2, XTRL_E_BADARG, "xtrl_hl_decode: bad shape");
    if (n == 0) return XTRL_OK;
    XTRL_REQUIRE(logits && out, XTRL_E_BADARG, "xtrl_hl_decode: null pointer");
    const int64_t want = (n + 7) / 8;
    const int blocks = (int)(want < 8 * num_sms() ? want : 8 * num_sms());
    hl_decode_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(logits, out, n, make_support(reward_min, reward_max, num_bins));
    XTRL_LAUNCH_CHECK("hl_decode_kernel");
    return XTRL_OK;
}
