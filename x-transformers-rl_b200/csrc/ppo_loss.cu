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
// once.  One THREAD per timestep row over shared-memory tiles of 128 rows (see ppo_loss_rows_kernel).
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

// ---- main: one THREAD per timestep row ----------------------------------------------------------------------------
// A warp-per-row layout (round 1) spent ~600 warp instructions per row on the K-wide critic term: five butterfly reductions,
// 4 x 32 exp slots for 100 bins, and erf evaluated by whole warps although only the ~24 bin edges within 4 sigma-widths of the
// target have a non-saturated erf.  Here a thread owns a row: max / sum-exp / expectation are register recurrences over the K
// logits (read as 16-byte pieces from a shared-memory tile whose row stride is an odd number of 16-byte units: conflict-free),
// the HL-Gauss histogram is evaluated only over its 26-bin window with a rolling erf (every other bin is exactly 0 in fp32,
// as in the oracle, because erff saturates to +-1 beyond |z| = 4), and the gradient row is written in place over the logits
// and leaves the tile as coalesced row copies.  ~100 warp-instruction equivalents per row.  Narrow per-row fields (actor
// logits, state prediction, targets, returns ...) are read straight from global memory by the row's thread.
constexpr int kRowThreads = 128;         // rows per tile = threads per CTA
constexpr int kWin = 26;                 // bins a histogram window can touch: 2 * 4 / (step / denom) + margins = 22.6 + 3

template <typename TIO> struct RowIO;
template <> struct RowIO<float> {
    static constexpr int VE = 4;
    static __device__ __forceinline__ void load(const unsigned char *row, int piece, float (&v)[4]) {
        const float4 t = *reinterpret_cast<const float4 *>(row + piece * 16);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    static __device__ __forceinline__ void store(unsigned char *row, int piece, const float (&v)[4]) {
        *reinterpret_cast<float4 *>(row + piece * 16) = make_float4(v[0], v[1], v[2], v[3]);
    }
    static __device__ __forceinline__ float ld(const unsigned char *row, int k) { return reinterpret_cast<const float *>(row)[k]; }
    static __device__ __forceinline__ void st(unsigned char *row, int k, float v) { reinterpret_cast<float *>(row)[k] = v; }
    static __device__ __forceinline__ float ex(float x) { return expf(x); }
};
template <> struct RowIO<__nv_bfloat16> {
    static constexpr int VE = 8;
    static __device__ __forceinline__ void load(const unsigned char *row, int piece, float (&v)[8]) {
        const uint4 t = *reinterpret_cast<const uint4 *>(row + piece * 16);
        const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
    }
    static __device__ __forceinline__ void store(unsigned char *row, int piece, const float (&v)[8]) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]); w[i] = *reinterpret_cast<const uint32_t *>(&h); }
        *reinterpret_cast<uint4 *>(row + piece * 16) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    static __device__ __forceinline__ float ld(const unsigned char *row, int k) { return __bfloat162float(reinterpret_cast<const __nv_bfloat16 *>(row)[k]); }
    static __device__ __forceinline__ void st(unsigned char *row, int k, float v) { reinterpret_cast<__nv_bfloat16 *>(row)[k] = __float2bfloat16_rn(v); }
    static __device__ __forceinline__ float ex(float x) { return __expf(x); }      // the 1e-2 class
};

// cooperative copy of one row of `rowbytes` bytes in units of `unit` bytes (16 / 8 / 4 / 2: the largest that divides rowbytes)
__device__ __forceinline__ void copy_row(unsigned char *dst, const unsigned char *src, int rowbytes, int unit, int lane) {
    if (unit == 16) { for (int o = lane * 16; o < rowbytes; o += 512) *reinterpret_cast<uint4 *>(dst + o) = *reinterpret_cast<const uint4 *>(src + o); }
    else if (unit == 8) { for (int o = lane * 8; o < rowbytes; o += 256) *reinterpret_cast<uint2 *>(dst + o) = *reinterpret_cast<const uint2 *>(src + o); }
    else if (unit == 4) { for (int o = lane * 4; o < rowbytes; o += 128) *reinterpret_cast<uint32_t *>(dst + o) = *reinterpret_cast<const uint32_t *>(src + o); }
    else { for (int o = lane * 2; o < rowbytes; o += 64) *reinterpret_cast<uint16_t *>(dst + o) = *reinterpret_cast<const uint16_t *>(src + o); }
}

struct HistWindow { int k0, k1; };       // bins [k0, k1) can be non-zero
__device__ __forceinline__ HistWindow hist_window(const HlSupport &sup, float y, int K) {
    const float pos = (y - 4.f * sup.denom - sup.lo) / sup.step;             // edges below this index have erf == -1 exactly
    int k0 = (int)floorf(pos) - 1;
    k0 = k0 < 0 ? 0 : (k0 > K ? K : k0);
    HistWindow w;
    w.k0 = k0; w.k1 = k0 + kWin < K ? k0 + kWin : K;
    return w;
}

template <typename TIO>
__global__ void __launch_bounds__(kRowThreads)
ppo_loss_rows_kernel(const LossArgs a, int row_stride, int unit) {
    using IO = RowIO<TIO>;
    constexpr int VE = IO::VE;
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int A = a.cfg.num_actions, K = a.cfg.num_bins, S1 = a.cfg.state_dim + 1, T = a.T;
    const int64_t N = (int64_t)a.b * T;
    const double *ws = a.io.workspace;
    const bool cont = a.cfg.continuous != 0;
    float *s_edge = reinterpret_cast<float *>(smem + (size_t)kRowThreads * row_stride);      // [K + 1] bin edges, [K] centres
    float *s_centre = s_edge + K + 1;
    for (int k = tid; k <= K; k += kRowThreads) { s_edge[k] = a.sup.edge(k); if (k < K) s_centre[k] = a.sup.centre(k); }

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
    const float feps = 1.1920928955078125e-07f;
    const float inv_denom = 1.f / a.sup.denom;
    const TIO *in_vl = (const TIO *)a.io.value_logits, *in_ra = (const TIO *)a.io.raw_actions;
    const TIO *in_sp = (const TIO *)a.io.state_pred_raw, *in_dl = (const TIO *)a.io.done_logit;
    TIO *g_ra = (TIO *)a.io.g_raw_actions, *g_vl = (TIO *)a.io.g_value_logits, *g_sp = (TIO *)a.io.g_state_pred_raw, *g_dl = (TIO *)a.io.g_done_logit;
    const int rowbytes = K * (int)sizeof(TIO), npieces = (K + VE - 1) / VE, nfull = K / VE;
    float acc_ac = 0.f, acc_actor = 0.f, acc_critic = 0.f, acc_nll = 0.f, acc_done = 0.f;

    const int64_t ntiles = (N + kRowThreads - 1) / kRowThreads;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t r0 = tile * kRowThreads;
        const int nrows = (int)(N - r0 < kRowThreads ? N - r0 : kRowThreads);
        __syncthreads();                                                     // the previous tile has left shared memory (and the tables are written)
        for (int row = warp; row < nrows; row += kRowThreads / 32)
            copy_row(smem + (size_t)row * row_stride, reinterpret_cast<const unsigned char *>(in_vl) + (r0 + row) * rowbytes, rowbytes, unit, lane);
        __syncthreads();
        if (tid < nrows) {
            const int64_t r = r0 + tid;
            const int e = (int)(r / T), t = (int)(r - (int64_t)e * T);
            const bool masked = t < __ldg(a.io.episode_lens + e);
            const float ret = __ldg(a.io.returns + r), v_old = __ldg(a.io.old_values + r);
            unsigned char *xr = smem + (size_t)tid * row_stride;

            // ---------------- critic: HL-Gauss CE against returns and clamped returns (:446-477) ------------
            float mx = -INFINITY;
            for (int p = 0; p < npieces; ++p) {
                float v[VE];
                IO::load(xr, p, v);
#pragma unroll
                for (int j = 0; j < VE; ++j) if (p < nfull || p * VE + j < K) mx = fmaxf(mx, v[j]);
            }
            float se = 0.f, sc = 0.f;
            for (int p = 0; p < npieces; ++p) {
                float v[VE];
                IO::load(xr, p, v);
#pragma unroll
                for (int j = 0; j < VE; ++j) {
                    if (p < nfull || p * VE + j < K) {
                        const float ex = IO::ex(v[j] - mx);
                        se += ex;
                        sc = fmaf(ex, s_centre[p * VE + j], sc);
                    }
                }
            }
            const float v_new = sc / se;                                      // hl_gauss(values), :455
            const float lse = logf(se) + mx, inv_se = 1.f / se;
            const float y0 = fminf(fmaxf(ret, a.sup.lo), a.sup.hi);                                     // clamp_to_range
            const float y1 = fminf(fmaxf(fminf(fmaxf(ret, -a.cfg.value_clip), a.cfg.value_clip), a.sup.lo), a.sup.hi);  // :459
            const bool same = y0 == y1;
            const float iz0 = 1.f / fmaxf(erff((a.sup.hi - y0) * inv_denom) - erff((a.sup.lo - y0) * inv_denom), 1e-10f);
            const float iz1 = same ? iz0 : 1.f / fmaxf(erff((a.sup.hi - y1) * inv_denom) - erff((a.sup.lo - y1) * inv_denom), 1e-10f);
            auto ce_of = [&](float y, float iz, float &ce, float &sp) {      // cross entropy and histogram mass over the window
                const HistWindow w = hist_window(a.sup, y, K);
                ce = 0.f; sp = 0.f;
                float c_prev = erff((s_edge[w.k0] - y) * inv_denom);
                for (int k = w.k0; k < w.k1; ++k) {
                    const float c_next = erff((s_edge[k + 1] - y) * inv_denom);
                    const float p = (c_next - c_prev) * iz;
                    ce -= p * (IO::ld(xr, k) - lse);
                    sp += p;
                    c_prev = c_next;
                }
            };
            float ce0, sp0, ce1, sp1;
            ce_of(y0, iz0, ce0, sp0);
            if (same) { ce1 = ce0; sp1 = sp0; } else ce_of(y1, iz1, ce1, sp1);
            const float lo_old = v_old - a.cfg.value_clip, hi_old = v_old + a.cfg.value_clip;
            const bool zero = (ret < v_new && v_new < lo_old) || (hi_old < v_new && v_new < ret);   // :470-472
            const float critic = zero ? 0.f : fminf(ce0, ce1);
            if (g_vl) {
                // d min(ce0, ce1): the smaller one; an exact tie splits evenly (torch minimum)
                const float w0 = zero ? 0.f : (ce0 < ce1 ? 1.f : (ce0 > ce1 ? 0.f : 0.5f));
                const float w1 = zero ? 0.f : 1.f - w0;
                const float gc = masked ? g_ac * a.cfg.critic_weight : 0.f;
                const float Asm = gc * (w0 * sp0 + w1 * sp1) * inv_se;        // coefficient of exp(x_k - mx) in every bin
                // histogram term, window of the selected target first (reads x_k, writes g_k in place)
                const bool use0 = same || w0 >= 0.5f;
                const float ysel = use0 ? y0 : y1, izsel = use0 ? iz0 : iz1;
                const float wsel = same ? gc * (w0 + w1) : gc * (use0 ? w0 : w1);
                const HistWindow w = hist_window(a.sup, ysel, K);
                {
                    float c_prev = erff((s_edge[w.k0] - ysel) * inv_denom);
                    for (int k = w.k0; k < w.k1; ++k) {
                        const float c_next = erff((s_edge[k + 1] - ysel) * inv_denom);
                        const float p = (c_next - c_prev) * izsel;
                        IO::st(xr, k, Asm * IO::ex(IO::ld(xr, k) - mx) - wsel * p);
                        c_prev = c_next;
                    }
                }
                for (int p = 0; p < npieces; ++p) {                           // every other bin: A * softmax_k
                    const int kb = p * VE;
                    if (p < nfull && (kb + VE <= w.k0 || kb >= w.k1)) {
                        float v[VE];
                        IO::load(xr, p, v);
#pragma unroll
                        for (int j = 0; j < VE; ++j) v[j] = Asm * IO::ex(v[j] - mx);
                        IO::store(xr, p, v);
                    } else {
                        for (int k = kb; k < kb + VE && k < K; ++k)
                            if (k < w.k0 || k >= w.k1) IO::st(xr, k, Asm * IO::ex(IO::ld(xr, k) - mx));
                    }
                }
                if (!same && w0 == 0.5f) {                                    // exact tie of two different targets: the other half
                    const HistWindow w2 = hist_window(a.sup, y1, K);
                    float c_prev = erff((s_edge[w2.k0] - y1) * inv_denom);
                    for (int k = w2.k0; k < w2.k1; ++k) {
                        const float c_next = erff((s_edge[k + 1] - y1) * inv_denom);
                        IO::st(xr, k, IO::ld(xr, k) - gc * w1 * (c_next - c_prev) * iz1);
                        c_prev = c_next;
                    }
                }
            }

            // ---------------- actor: clipped surrogate + entropy (:413-444) ------------------------------------
            const float adv = ((ret - v_old) - adv_mean) * adv_inv_std;
            const float ga = masked ? g_ac * a.cfg.actor_weight : 0.f;
            float actor = 0.f;
            if (cont) {                                                      // Continuous (:224-277): per action dim, summed (:442)
                for (int i = 0; i < A; ++i) {
                    const float mean = ld_as_float(in_ra + (r * A + i) * 2), lv = ld_as_float(in_ra + (r * A + i) * 2 + 1);
                    const float th = tanhf(lv / 3.f), var = expf(th * 3.f);
                    const bool inside = var > 1e-5f;
                    const float sd = sqrtf(fmaxf(var, 1e-5f));
                    const float v = __ldg(a.io.actions_f + r * A + i), df = v - mean;
                    float logp = -(df * df) / (2.f * sd * sd) - logf(sd) - 0.91893853320467274178f;
                    if (a.cfg.squash) logp -= logf(fmaxf(1.f - v * v, 1e-20f));
                    const float ent = a.cfg.squash ? -logp : 1.41893853320467274178f + logf(sd);
                    const float ratio = expf(logp - __ldg(a.io.old_log_probs + r * A + i));
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
            } else {                                                         // Discrete (:197-222)
                const int act = __ldg(a.io.actions + r);
                float amx = -INFINITY;
                for (int i = 0; i < A; ++i) amx = fmaxf(amx, ld_as_float(in_ra + r * A + i));
                float qs = 0.f;
                for (int i = 0; i < A; ++i) qs += expf(ld_as_float(in_ra + r * A + i) - amx);
                float ssum = 0.f;
                for (int i = 0; i < A; ++i) ssum += expf(ld_as_float(in_ra + r * A + i) - amx) / qs;   // Categorical renormalises
                float ent = 0.f, logp = 0.f, p_act = 1.f, dot_p = 0.f;
                bool in_act = false;
                for (int i = 0; i < A; ++i) {
                    const float sm = expf(ld_as_float(in_ra + r * A + i) - amx) / qs, p = sm / ssum;
                    const float L = logf(fminf(fmaxf(p, feps), 1.f - feps));
                    ent -= p * L;
                    if (i == act) { logp = L; p_act = p; in_act = p >= feps && p <= 1.f - feps; }
                }
                const float ratio = expf(logp - __ldg(a.io.old_log_probs + r));
                const float clipped = fminf(fmaxf(ratio, 1.f - a.cfg.eps_clip), 1.f + a.cfg.eps_clip);
                const float surr1 = ratio * adv, surr2 = clipped * adv;
                actor = -fminf(surr1, surr2) - a.cfg.entropy_weight * ent;
                if (g_ra) {
                    const bool in_range = ratio >= 1.f - a.cfg.eps_clip && ratio <= 1.f + a.cfg.eps_clip;
                    const float dmin = in_range ? adv : (surr1 < surr2 ? adv : 0.f);
                    const float g_logp = -ga * dmin * ratio;
                    auto gp_of = [&](int i, float p, float L) {              // d/dp_i of (-beta * ent + logp term)
                        const bool inside = p >= feps && p <= 1.f - feps;
                        float gp = ga * a.cfg.entropy_weight * (L + (inside ? 1.f : 0.f));
                        if (i == act && in_act) gp += g_logp / p_act;
                        return gp;
                    };
                    float dot_q = 0.f;
                    for (int i = 0; i < A; ++i) {
                        const float sm = expf(ld_as_float(in_ra + r * A + i) - amx) / qs, p = sm / ssum;
                        dot_p += gp_of(i, p, logf(fminf(fmaxf(p, feps), 1.f - feps))) * p;
                    }
                    for (int i = 0; i < A; ++i) {
                        const float sm = expf(ld_as_float(in_ra + r * A + i) - amx) / qs, p = sm / ssum;
                        dot_q += (gp_of(i, p, logf(fminf(fmaxf(p, feps), 1.f - feps))) - dot_p) / ssum * sm;
                    }
                    for (int i = 0; i < A; ++i) {
                        const float sm = expf(ld_as_float(in_ra + r * A + i) - amx) / qs, p = sm / ssum;
                        const float gq = (gp_of(i, p, logf(fminf(fmaxf(p, feps), 1.f - feps))) - dot_p) / ssum;
                        st_from_float(g_ra + r * A + i, sm * (gq - dot_q));
                    }
                }
            }
            if (a.io.per_step) { a.io.per_step[r * 2] = actor; a.io.per_step[r * 2 + 1] = critic; }
            acc_actor += actor; acc_critic += critic;
            if (masked) acc_ac += a.cfg.actor_weight * actor + a.cfg.critic_weight * critic;

            // ---------------- world model: Gaussian NLL of next (state || reward) (:398-404, :232-241) ----------
            if (in_sp) {
                const bool valid = masked && t < T - 1;
                float nll = 0.f;
                for (int dd = 0; dd < S1; ++dd) {
                    float gm = 0.f, gl = 0.f;
                    if (valid) {
                        const float mean = ld_as_float(in_sp + (r * S1 + dd) * 2), lv = ld_as_float(in_sp + (r * S1 + dd) * 2 + 1);
                        const float th = tanhf(lv / 3.f);
                        const float var = fmaxf(expf(th * 3.f), 1e-6f);
                        const float d = mean - __ldg(a.io.target + (r + 1) * S1 + dd);
                        nll += 0.5f * (logf(var) + d * d / var);
                        gm = g_wm * d / var;
                        gl = g_wm * 0.5f * (1.f / var - d * d / (var * var)) * var * (1.f - th * th);
                    }
                    if (g_sp) { st_from_float(g_sp + (r * S1 + dd) * 2, gm); st_from_float(g_sp + (r * S1 + dd) * 2 + 1, gl); }
                }
                acc_nll += nll;
            }
            // ---------------- done head: BCE on sigmoid(logit) (:406-411) ---------------------------------------
            if (in_dl) {
                const float z = ld_as_float(in_dl + r);
                const float pd = 1.f / (1.f + expf(-z));
                const float y = __ldg(a.io.dones + r) ? 1.f : 0.f;
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
        __syncthreads();
        if (g_vl)
            for (int row = warp; row < nrows; row += kRowThreads / 32)
                copy_row(reinterpret_cast<unsigned char *>(g_vl) + (r0 + row) * rowbytes, smem + (size_t)row * row_stride, rowbytes, unit, lane);
    }

    acc_ac = warp_sum(acc_ac); acc_actor = warp_sum(acc_actor); acc_critic = warp_sum(acc_critic);
    acc_nll = warp_sum(acc_nll); acc_done = warp_sum(acc_done);
    __shared__ float sh[5][kRowThreads / 32];
    if (lane == 0) { sh[0][warp] = acc_ac; sh[1][warp] = acc_actor; sh[2][warp] = acc_critic; sh[3][warp] = acc_nll; sh[4][warp] = acc_done; }
    __syncthreads();
    if (threadIdx.x < 5) {
        double s = 0.;
        for (int i = 0; i < kRowThreads / 32; ++i) s += (double)sh[threadIdx.x][i];
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
    const int sz = cfg->bf16_io ? 2 : 4;
    const void *ptrs[] = {io->raw_actions, io->value_logits, io->state_pred_raw, io->done_logit, io->actions, io->old_log_probs,
                          io->returns, io->old_values, io->target, io->dones, io->actions_f, io->g_value_logits};
    for (const void *p : ptrs)
        XTRL_REQUIRE(((uintptr_t)p & 15) == 0, XTRL_E_BADARG, "xtrl_ppo_loss_fwd_bwd: input pointers must be 16-byte aligned");
    // one thread per row; the tile's row stride is an odd number of 16-byte units (16-byte row pieces are then conflict-free)
    const int rowbytes = cfg->num_bins * sz;
    int stride = (rowbytes + 15) & ~15;
    if (((stride / 16) & 1) == 0) stride += 16;
    const int unit = rowbytes % 16 == 0 ? 16 : rowbytes % 8 == 0 ? 8 : rowbytes % 4 == 0 ? 4 : 2;
    const size_t smem = (size_t)kRowThreads * stride + (size_t)(2 * cfg->num_bins + 1) * sizeof(float) + 16;
    XTRL_REQUIRE(smem <= 200 * 1024, XTRL_E_UNSUPPORTED, "xtrl_ppo_loss_fwd_bwd: tile needs %zu B shared memory", smem);
    const int64_t ntiles = (N + kRowThreads - 1) / kRowThreads;
    int per_sm = (int)((220 * 1024) / (smem + 1024));
    per_sm = per_sm > 12 ? 12 : (per_sm < 1 ? 1 : per_sm);
    const int blocks = (int)(ntiles < (int64_t)per_sm * num_sms() ? ntiles : (int64_t)per_sm * num_sms());   // persistent CTAs
    if (cfg->bf16_io) {
        XTRL_CUDA(cudaFuncSetAttribute(ppo_loss_rows_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ppo_loss_rows_kernel<__nv_bfloat16><<<blocks, kRowThreads, smem, s>>>(la, stride, unit);
    } else {
        XTRL_CUDA(cudaFuncSetAttribute(ppo_loss_rows_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ppo_loss_rows_kernel<float><<<blocks, kRowThreads, smem, s>>>(la, stride, unit);
    }
    XTRL_LAUNCH_CHECK("ppo_loss_rows_kernel");
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
