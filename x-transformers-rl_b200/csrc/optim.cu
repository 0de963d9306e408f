// optim.cu -- clip_grad_norm_ + AdoptAtan2 + EMA over the flat fp32 parameter buffer.
//
// Replaces the per-tensor python loops of nn.utils.clip_grad_norm_ (x_transformers_rl.py:987),
// AdoptAtan2.step (:749, :991) and the EMA post-step hook (:753; App. A.5) by three
// launches over one contiguous buffer: (1) fp64 sum of squares, (2) first-moment update + the
// per-tensor sum of the "cautious" scale, (3) parameter / second-moment / EMA update.
// HBM-bound: 8 fp32 streams of n elements; n <= ~3e5 here, so this is launch-latency territory --
// the point is 3 launches instead of ~50 x 12.
#include "common.cuh"

namespace xtrl {

constexpr int kOptThreads = 256;

__global__ void __launch_bounds__(kOptThreads)
opt_sumsq_kernel(const float *__restrict__ g, int64_t n, double *__restrict__ ws) {
    double s = 0.;
    for (int64_t i = blockIdx.x * (int64_t)kOptThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kOptThreads) {
        const double x = (double)g[i];
        s += x * x;
    }
    s = warp_sum(s);
    __shared__ double sh[kOptThreads / 32];
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.;
        for (int i = 0; i < kOptThreads / 32; ++i) t += sh[i];
        atomicAdd(ws, t);
    }
}

__device__ __forceinline__ int find_segment(const int64_t *__restrict__ off, int nt, int64_t i) {
    int lo = 0, hi = nt;                       // off[lo] <= i < off[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(off + mid) <= i) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ float clip_coef(const double *ws, float max_norm) {
    const float total = (float)sqrt(ws[0]);
    return fminf(max_norm / (total + 1e-6f), 1.f);
}

__device__ __forceinline__ float lerp_(float a, float b, float w) {
    const float d = b - a;
    return w < 0.5f ? fmaf(w, d, a) : fmaf(-d, 1.f - w, b);
}

__global__ void __launch_bounds__(kOptThreads)
opt_moment_kernel(const float *__restrict__ grads, float *__restrict__ m, const float *__restrict__ v,
                  const int64_t *__restrict__ seg, int nt, int64_t n, xtrl_opt_cfg c, double *__restrict__ ws) {
    const float coef = c.max_grad_norm > 0.f ? clip_coef(ws, c.max_grad_norm) : 1.f;
    for (int64_t i = blockIdx.x * (int64_t)kOptThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kOptThreads) {
        const float g = grads[i] * coef;
        const float mi = lerp_(m[i], atan2f(g, c.b * sqrtf(v[i])), 1.f - c.beta1);
        m[i] = mi;
        if (c.cautious_factor < 1.f) {         // one atomic per (warp, tensor) instead of per element
            const int sgm = find_segment(seg, nt, i);
            const unsigned active = __activemask();
            const unsigned peers = __match_any_sync(active, sgm);
            const unsigned aligned = __ballot_sync(active, g * mi > 0.f) & peers;
            if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) {
                const int ones = __popc(aligned), rest = __popc(peers) - ones;
                atomicAdd(ws + 2 + sgm, (double)ones + (double)c.cautious_factor * (double)rest);
            }
        }
    }
}

__global__ void __launch_bounds__(kOptThreads)
opt_apply_kernel(float *__restrict__ p, const float *__restrict__ grads, float *__restrict__ m,
                 float *__restrict__ v, float *__restrict__ p_init, float *__restrict__ ema,
                 const int64_t *__restrict__ seg, int nt, int64_t n, xtrl_opt_cfg c,
                 const double *__restrict__ ws, float *__restrict__ grad_norm_out) {
    const float coef = c.max_grad_norm > 0.f ? clip_coef(ws, c.max_grad_norm) : 1.f;
    if (grad_norm_out && blockIdx.x == 0 && threadIdx.x == 0) *grad_norm_out = (float)sqrt(ws[0]);
    for (int64_t i = blockIdx.x * (int64_t)kOptThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kOptThreads) {
        const float g = grads[i] * coef;
        float pi = p[i];
        if (c.first_step) {                    // ADOPT: the first step only seeds v (and param_init)
            m[i] = 0.f;
            v[i] = g * g;
            p_init[i] = pi;
        } else {
            if (c.regen_rate > 0.f) pi = lerp_(pi, p_init[i], c.lr / c.init_lr * c.regen_rate);
            const float mi = m[i];
            float scale = 1.f;
            if (c.cautious_factor < 1.f) {
                const int s = find_segment(seg, nt, i);
                const float mean = (float)(ws[2 + s] / (double)(__ldg(seg + s + 1) - __ldg(seg + s)));
                scale = ((g * mi > 0.f) ? 1.f : c.cautious_factor) / fmaxf(mean, 1e-5f);
            }
            pi = fmaf(mi * scale, -(c.lr * c.a), pi);
            v[i] = lerp_(v[i], g * g, 1.f - c.beta2);
        }
        if (ema) {
            float ei = ema[i];
            if (c.ema_mode == 1.f) ei = pi;
            else if (c.ema_mode == 2.f) ei = lerp_(ei, pi, 1.f - c.ema_decay);
            ema[i] = ei;
            if (c.copy_ema_to_online) pi = ei;
        }
        p[i] = pi;
    }
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
    cudaStream_t s = (cudaStream_t)stream;
    XTRL_CUDA(cudaMemsetAsync(workspace, 0, (size_t)xtrl_opt_workspace_bytes(num_tensors), s));
    const int64_t want = (n + kOptThreads - 1) / kOptThreads;
    const int blocks = (int)(want < 2 * kNumSMs ? want : 2 * kNumSMs);
    opt_sumsq_kernel<<<blocks, kOptThreads, 0, s>>>(grads, n, workspace);
    XTRL_LAUNCH_CHECK("opt_sumsq_kernel");
    if (!cfg->first_step) {
        opt_moment_kernel<<<blocks, kOptThreads, 0, s>>>(grads, m, v, seg_offsets, num_tensors, n, *cfg, workspace);
        XTRL_LAUNCH_CHECK("opt_moment_kernel");
    }
    opt_apply_kernel<<<blocks, kOptThreads, 0, s>>>(params, grads, m, v, param_init, ema, seg_offsets, num_tensors, n,
                                                    *cfg, workspace, grad_norm_out);
    XTRL_LAUNCH_CHECK("opt_apply_kernel");
    return XTRL_OK;
}
