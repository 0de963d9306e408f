// optim.cu -- clip_grad_norm_ + AdoptAtan2 + EMA over the flat fp32 parameter buffer, ONE launch.
//
// Replaces the per-tensor python loops of nn.utils.clip_grad_norm_ (x_transformers_rl.py:987),
// AdoptAtan2.step (:749, :991) and the EMA post-step hook (:753; App. A.5), plus the
// 1/world scaling that follows DDP's gradient all-reduce (:885, :981), by one cooperative kernel over
// contiguous buffers with two grid-wide barriers:
//   phase 1  fp64 sum of squares of the (scaled) gradient                      reads g
//   phase 2  first-moment update + per-tensor sums of the "cautious" scale     reads g, m, v; writes m
//   phase 3  regen, parameter, second moment, EMA, optional EMA -> online copy reads g, m, p, p0, v, ema; writes p, v, ema
// HBM-bound: 40 algorithmic bytes per parameter (g; m, v, p, ema read+write; p0), 56 as built when the
// buffers exceed L2 (the gradient is read once per phase).  Each CTA owns a contiguous span of the buffer
// and every thread walks it forward in 16-byte units, so the parameter-tensor (segment) of an element is
// found by a monotone cursor instead of a search, and the per-tensor sums are aggregated per warp.
#include "common.cuh"
#include <cooperative_groups.h>

namespace xtrl {
namespace cg = cooperative_groups;

constexpr int kOptThreads = 256;

struct OptArgs {
    float *p, *m, *v, *p_init, *ema;
    const float *g;
    const int64_t *seg;
    int nt;
    int64_t n;
    xtrl_opt_cfg c;
    double *ws;
    float *grad_norm_out;
};

__device__ __forceinline__ int find_segment(const int64_t *__restrict__ off, int nt, int64_t i) {
    int lo = 0, hi = nt;                       // off[lo] <= i < off[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(off + mid) <= i) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ float lerp_(float a, float b, float w) {
    const float d = b - a;
    return w < 0.5f ? fmaf(w, d, a) : fmaf(-d, 1.f - w, b);
}

// V consecutive floats of a buffer (V = 4: one 16-byte access when the unit is complete)
template <int V>
__device__ __forceinline__ void load_unit(const float *base, int64_t i0, int64_t n, float (&x)[V]) {
    if constexpr (V == 4) {
        if (i0 + 4 <= n) {
            const float4 t = *reinterpret_cast<const float4 *>(base + i0);
            x[0] = t.x; x[1] = t.y; x[2] = t.z; x[3] = t.w;
            return;
        }
    }
    {
#pragma unroll
        for (int e = 0; e < V; ++e) x[e] = i0 + e < n ? base[i0 + e] : 0.f;
    }
}
template <int V>
__device__ __forceinline__ void store_unit(float *base, int64_t i0, int64_t n, const float (&x)[V]) {
    if constexpr (V == 4) {
        if (i0 + 4 <= n) {
            *reinterpret_cast<float4 *>(base + i0) = make_float4(x[0], x[1], x[2], x[3]);
            return;
        }
    }
    {
#pragma unroll
        for (int e = 0; e < V; ++e) if (i0 + e < n) base[i0 + e] = x[e];
    }
}

// monotone cursor over the tensor boundaries: seg[s] <= i < seg[s + 1]
struct SegCursor {
    int s; int64_t hi;
    __device__ __forceinline__ void seek(const int64_t *seg, int nt, int64_t i) { s = find_segment(seg, nt, i); hi = __ldg(seg + s + 1); }
    __device__ __forceinline__ void advance(const int64_t *seg, int nt, int64_t i) {
        while (i >= hi && s + 1 < nt) { ++s; hi = __ldg(seg + s + 1); }
    }
};

template <int V>
__global__ void __launch_bounds__(kOptThreads)
opt_fused_kernel(const OptArgs a) {
    cg::grid_group grid = cg::this_grid();
    const xtrl_opt_cfg c = a.c;
    const int64_t n = a.n, U = (n + V - 1) / V;
    int64_t per_cta = (U + gridDim.x - 1) / gridDim.x;
    per_cta = (per_cta + kOptThreads - 1) / kOptThreads * kOptThreads;
    const int64_t u_lo = (int64_t)blockIdx.x * per_cta, u_hi = u_lo + per_cta < U ? u_lo + per_cta : U;
    const int tid = threadIdx.x, lane = tid & 31;
    const float gs = c.grad_scale;
    __shared__ double sh[kOptThreads / 32];

    // ---- phase 1: squared norm of the gradient (fp32 within a unit, fp64 across units) ----
    double s = 0.;
    for (int64_t u = u_lo + tid; u < u_hi; u += kOptThreads) {
        float x[V];
        load_unit<V>(a.g, u * V, n, x);
        float q = 0.f;
#pragma unroll
        for (int e = 0; e < V; ++e) { const float t = x[e] * gs; q = fmaf(t, t, q); }
        s += (double)q;
    }
    s = warp_sum(s);
    if (lane == 0) sh[tid >> 5] = s;
    __syncthreads();
    if (tid == 0) {
        double t = 0.;
        for (int i = 0; i < kOptThreads / 32; ++i) t += sh[i];
        if (t != 0.) atomicAdd(a.ws, t);
    }
    grid.sync();
    const float total = (float)sqrt(__ldcg(a.ws));
    const float coef = (c.max_grad_norm > 0.f ? fminf(c.max_grad_norm / (total + 1e-6f), 1.f) : 1.f) * gs;
    if (a.grad_norm_out && blockIdx.x == 0 && tid == 0) *a.grad_norm_out = total;
    const bool cautious = c.cautious_factor < 1.f;

    // ---- phase 2: first moment, per-tensor sums of the cautious scale ----
    if (!c.first_step) {
        SegCursor cur;
        int acc_seg = -1, acc_ones = 0, acc_cnt = 0;          // lane 0's running sums for the tensor the warp is in
        auto flush = [&]() {
            if (acc_seg >= 0 && acc_cnt > 0)
                atomicAdd(a.ws + 2 + acc_seg, (double)acc_ones + (double)c.cautious_factor * (double)(acc_cnt - acc_ones));
            acc_ones = acc_cnt = 0;
        };
        bool seeked = false;
        for (int64_t u0 = u_lo + (tid & ~31); u0 < u_hi; u0 += kOptThreads) {      // warp-uniform loop: u0 = the warp's first unit
            const int64_t u = u0 + lane, i0 = u * V;
            const bool live = u < u_hi;
            float gx[V], mx[V], vx[V];
            bool al[V] = {};
            if (live) {
                load_unit<V>(a.g, i0, n, gx); load_unit<V>(a.m, i0, n, mx); load_unit<V>(a.v, i0, n, vx);
#pragma unroll
                for (int e = 0; e < V; ++e) {
                    const float g = gx[e] * coef;
                    mx[e] = lerp_(mx[e], atan2f(g, c.b * sqrtf(vx[e])), 1.f - c.beta1);
                    al[e] = g * mx[e] > 0.f;
                }
                store_unit<V>(a.m, i0, n, mx);
            }
            if (cautious) {
                const int64_t w_first = u0 * V, w_last = (u0 + 31) * V + V - 1 < n ? (u0 + 31) * V + V - 1 : n - 1;
                if (!seeked) { cur.seek(a.seg, a.nt, w_first); seeked = true; } else cur.advance(a.seg, a.nt, w_first);
                if (w_last < cur.hi) {                           // the warp's 32 units lie in one tensor: ballots
                    int ones = 0, cnt = 0;
#pragma unroll
                    for (int e = 0; e < V; ++e) {
                        const bool in = live && i0 + e < n;
                        ones += __popc(__ballot_sync(0xffffffffu, in && al[e]));
                        cnt += __popc(__ballot_sync(0xffffffffu, in));
                    }
                    if (lane == 0) {
                        if (acc_seg != cur.s) { flush(); acc_seg = cur.s; }
                        acc_ones += ones; acc_cnt += cnt;
                    }
                } else if (live) {                               // the warp straddles a tensor boundary: per element
                    SegCursor mine = cur;
#pragma unroll
                    for (int e = 0; e < V; ++e) {
                        if (i0 + e < n) {
                            mine.advance(a.seg, a.nt, i0 + e);
                            atomicAdd(a.ws + 2 + mine.s, al[e] ? 1. : (double)c.cautious_factor);
                        }
                    }
                }
            }
        }
        if (lane == 0) flush();
    }
    if (cautious && !c.first_step) grid.sync();

    // ---- phase 3: parameters, second moment, EMA ----
    {
        SegCursor cur;
        bool seeked = false;
        const float regen_w = c.lr / c.init_lr * c.regen_rate, step = -(c.lr * c.a);
        for (int64_t u = u_lo + tid; u < u_hi; u += kOptThreads) {
            const int64_t i0 = u * V;
            float gx[V], px[V], vx[V], mx[V], ix[V], ex[V];
            load_unit<V>(a.g, i0, n, gx); load_unit<V>(a.p, i0, n, px);
            if (c.first_step) {
#pragma unroll
                for (int e = 0; e < V; ++e) { const float g = gx[e] * coef; mx[e] = 0.f; vx[e] = g * g; ix[e] = px[e]; }
                store_unit<V>(a.m, i0, n, mx); store_unit<V>(a.v, i0, n, vx); store_unit<V>(a.p_init, i0, n, ix);
            } else {
                load_unit<V>(a.v, i0, n, vx); load_unit<V>(a.m, i0, n, mx);
                if (c.regen_rate > 0.f) load_unit<V>(a.p_init, i0, n, ix);
                if (cautious) { if (!seeked) { cur.seek(a.seg, a.nt, i0); seeked = true; } }
#pragma unroll
                for (int e = 0; e < V; ++e) {
                    const float g = gx[e] * coef;
                    float pi = px[e];
                    if (c.regen_rate > 0.f) pi = lerp_(pi, ix[e], regen_w);
                    float scale = 1.f;
                    if (cautious && i0 + e < n) {
                        cur.advance(a.seg, a.nt, i0 + e);
                        const float mean = (float)(__ldcg(a.ws + 2 + cur.s) / (double)(cur.hi - __ldg(a.seg + cur.s)));
                        scale = ((g * mx[e] > 0.f) ? 1.f : c.cautious_factor) / fmaxf(mean, 1e-5f);
                    }
                    px[e] = fmaf(mx[e] * scale, step, pi);
                    vx[e] = lerp_(vx[e], g * g, 1.f - c.beta2);
                }
                store_unit<V>(a.v, i0, n, vx);
            }
            if (a.ema) {
                if (c.ema_mode == 2.f || c.copy_ema_to_online) load_unit<V>(a.ema, i0, n, ex);
#pragma unroll
                for (int e = 0; e < V; ++e) {
                    float ei = (c.ema_mode == 2.f || c.copy_ema_to_online) ? ex[e] : 0.f;
                    if (c.ema_mode == 1.f) ei = px[e];
                    else if (c.ema_mode == 2.f) ei = lerp_(ei, px[e], 1.f - c.ema_decay);
                    ex[e] = ei;
                    if (c.copy_ema_to_online) px[e] = ei;
                }
                if (c.ema_mode != 0.f) store_unit<V>(a.ema, i0, n, ex);
            }
            store_unit<V>(a.p, i0, n, px);
        }
    }
}

template <int V>
static int launch_opt(const OptArgs &a, cudaStream_t s) {
    static int per_sm = 0;
    if (per_sm == 0) {
        XTRL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, opt_fused_kernel<V>, kOptThreads, 0));
        if (per_sm < 1) per_sm = 1;
        if (per_sm > 4) per_sm = 4;
    }
    const int64_t U = (a.n + V - 1) / V, want = (U + kOptThreads - 1) / kOptThreads;
    const int64_t cap = (int64_t)per_sm * num_sms();
    const int blocks = (int)(want < cap ? want : cap);
    void *args[] = {(void *)&a};
    XTRL_CUDA(cudaLaunchCooperativeKernel((const void *)opt_fused_kernel<V>, dim3(blocks), dim3(kOptThreads), args, 0, s));
    return XTRL_OK;
}

}  // namespace xtrl

extern "C" int64_t xtrl_opt_workspace_bytes(int32_t num_tensors) {
    return (int64_t)sizeof(double) * (2 + (num_tensors > 0 ? num_tensors : 0));
}

extern "C" int xtrl_adopt_atan2_step(float *params, const float *grads, float *m, float *v, float *param_init,
                                     float *ema, const int64_t *seg_offsets, int32_t num_tensors, int64_t n,
                                     const xtrl_opt_cfg *cfg, double *workspace, float *grad_norm_out,
                                     void *stream) {
    using namespace xtrl;
    XTRL_REQUIRE(cfg, XTRL_E_BADARG, "xtrl_adopt_atan2_step: null cfg");
    XTRL_REQUIRE(n >= 0 && num_tensors >= 1, XTRL_E_BADARG, "xtrl_adopt_atan2_step: bad sizes");
    if (n == 0) return XTRL_OK;
    XTRL_REQUIRE(params && grads && m && v && param_init && seg_offsets && workspace, XTRL_E_BADARG,
                 "xtrl_adopt_atan2_step: null pointer");
    XTRL_REQUIRE(cfg->grad_scale > 0.f, XTRL_E_BADARG, "xtrl_adopt_atan2_step: grad_scale must be > 0 (1 / world size; 1 on one GPU)");
    cudaStream_t s = (cudaStream_t)stream;
    XTRL_CUDA(cudaMemsetAsync(workspace, 0, (size_t)xtrl_opt_workspace_bytes(num_tensors), s));
    OptArgs a;
    a.p = params; a.g = grads; a.m = m; a.v = v; a.p_init = param_init; a.ema = ema; a.seg = seg_offsets; a.nt = num_tensors; a.n = n;
    a.c = *cfg; a.ws = workspace; a.grad_norm_out = grad_norm_out;
    const uintptr_t bits = (uintptr_t)params | (uintptr_t)grads | (uintptr_t)m | (uintptr_t)v | (uintptr_t)param_init | (uintptr_t)ema;
    return (bits & 15) == 0 ? launch_opt<4>(a, s) : launch_opt<1>(a, s);      // 16-byte units when every buffer allows it
}
