// gae.cu -- GAE / returns as a reverse first-order linear scan over time.
//
// Replaces calc_gae (x_transformers_rl/x_transformers_rl.py:616-640):
//     v_T = 0;  delta_t = r_t + gamma*v_{t+1}*m_t - v_t;  gate_t = gamma*lam*m_t;
//     gae_t = delta_t + gate_t*gae_{t+1}   (AssocScan(reverse=True), :634-636);  returns = gae + v
//
// HBM-bound: 13 algorithmic bytes per timestep (4 r + 4 v + 1 boundary read, 4 returns written),
// every input read from DRAM exactly once.  Two layouts, chosen by the strides:
//   episode-major (E,T), stride_t==1 -- the reference's layout.  One warp per episode, lanes along
//       time (coalesced 128 B rows), a 5-step shuffle scan of the affine maps (a,b): h -> b + a*h
//       per 32-step block, blocks walked from the end with the carry in a register; PF blocks are
//       prefetched into registers before they are scanned to keep 3*PF loads per lane in flight.
//   time-major (T,B), stride_e==1 -- the rollout buffer's layout.  A CTA owns 32 envs (lanes) x W
//       time chunks (warps); each thread folds its chunk into one affine map, the W maps of an env
//       are composed through shared memory, and a second sweep (L2-resident re-read) writes.
#include "common.cuh"

namespace xtrl {

constexpr int kGaeWarpsPerBlock = 8;    // episode-major: 8 episodes per CTA
constexpr int kGaePrefetch = 8;         // blocks of 32 timesteps held in registers at once

__global__ void __launch_bounds__(kGaeWarpsPerBlock * 32)
gae_episode_major_kernel(const float *__restrict__ rewards, const float *__restrict__ values,
                         const uint8_t *__restrict__ boundary, float *__restrict__ returns,
                         int E, int T, int64_t stride_e, float gamma, float gl) {
    const int lane = threadIdx.x & 31;
    const int e = blockIdx.x * kGaeWarpsPerBlock + (threadIdx.x >> 5);
    if (e >= E) return;
    const float *r = rewards + e * stride_e;
    const float *v = values + e * stride_e;
    const uint8_t *bd = boundary + e * stride_e;
    float *out = returns + e * stride_e;

    const int nblk = (T + 31) >> 5;
    float carry = 0.f;       // gae of the first timestep of the block after this one
    float v_after = 0.f;     // value of that timestep (v_T = 0)
    for (int hi = nblk; hi > 0; hi -= kGaePrefetch) {
        float rr[kGaePrefetch], vv[kGaePrefetch];
        uint8_t bb[kGaePrefetch];
#pragma unroll
        for (int j = 0; j < kGaePrefetch; ++j) {          // block hi-1-j
            const int t = ((hi - 1 - j) << 5) + lane;
            const bool ok = (hi - 1 - j) >= 0 && t < T;
            rr[j] = ok ? __ldg(r + t) : 0.f;
            vv[j] = ok ? __ldg(v + t) : 0.f;
            bb[j] = ok ? __ldg(bd + t) : (uint8_t)0;
        }
#pragma unroll
        for (int j = 0; j < kGaePrefetch; ++j) {
            if (hi - 1 - j < 0) break;                    // warp-uniform
            const int t = ((hi - 1 - j) << 5) + lane;
            const bool ok = t < T;
            float v_next = __shfl_down_sync(0xffffffffu, vv[j], 1);
            if (lane == 31) v_next = v_after;
            const float m = bb[j] ? 0.f : 1.f;
            // identity map (a=1, b=0) past the end so the (zero) carry passes through
            float a = ok ? gl * m : 1.f;
            float b = ok ? (rr[j] + gamma * v_next * m - vv[j]) : 0.f;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {            // suffix composition over lanes
                const float a2 = __shfl_down_sync(0xffffffffu, a, d);
                const float b2 = __shfl_down_sync(0xffffffffu, b, d);
                if (lane + d < 32) { b = fmaf(a, b2, b); a *= a2; }
            }
            const float h = fmaf(a, carry, b);
            if (ok) out[t] = h + vv[j];
            carry = __shfl_sync(0xffffffffu, h, 0);
            v_after = __shfl_sync(0xffffffffu, vv[j], 0);
        }
    }
}

constexpr int kGaeChunks = 32;          // time-major: warps (time chunks) per CTA
constexpr int kGaeRegChunk = 16;        // chunk length held in registers (T <= 512)

// REG: the thread's whole chunk (<= 32 timesteps) is loaded into registers up front -- all 3*C loads of
// a thread are in flight together, inputs cross the memory system exactly once -- then folded, carried
// through shared memory, and replayed from registers.  !REG (T > 512): two sweeps, the second from L2.
template <bool REG>
__global__ void __launch_bounds__(kGaeChunks * 32)
gae_time_major_kernel(const float *__restrict__ rewards, const float *__restrict__ values,
                      const uint8_t *__restrict__ boundary, float *__restrict__ returns,
                      int E, int T, int64_t stride_t, float gamma, float gl) {
    __shared__ float sA[kGaeChunks][32];
    __shared__ float sB[kGaeChunks][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int e = blockIdx.x * 32 + lane;
    const bool live = e < E;
    const int C = (T + kGaeChunks - 1) / kGaeChunks;
    const int t_lo = w * C, t_hi = min(T, t_lo + C);      // [t_lo, t_hi)

    if constexpr (REG) {
        float dl[kGaeRegChunk], gt[kGaeRegChunk], vv[kGaeRegChunk + 1];
#pragma unroll
        for (int i = 0; i <= kGaeRegChunk; ++i) {          // values, one extra for v_{t+1} of the last step
            const int t = t_lo + i;
            vv[i] = (live && i <= C && t < T) ? __ldg(values + (int64_t)t * stride_t + e) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < kGaeRegChunk; ++i) {
            const int t = t_lo + i;
            const bool ok = live && i < C && t < t_hi;
            const float rv = ok ? __ldg(rewards + (int64_t)t * stride_t + e) : 0.f;
            const float m = (ok && __ldg(boundary + (int64_t)t * stride_t + e)) ? 0.f : 1.f;
            gt[i] = ok ? gl * m : 1.f;                      // identity map outside the chunk
            dl[i] = ok ? rv + gamma * vv[i + 1] * m - vv[i] : 0.f;
        }
        float A = 1.f, Bc = 0.f;
#pragma unroll
        for (int i = kGaeRegChunk - 1; i >= 0; --i) { Bc = fmaf(gt[i], Bc, dl[i]); A *= gt[i]; }
        sA[w][lane] = A;
        sB[w][lane] = Bc;
        __syncthreads();
        float h = 0.f;                                     // gae at t_hi
        for (int c = kGaeChunks - 1; c > w; --c) h = fmaf(sA[c][lane], h, sB[c][lane]);
#pragma unroll
        for (int i = kGaeRegChunk - 1; i >= 0; --i) {
            h = fmaf(gt[i], h, dl[i]);
            const int t = t_lo + i;
            if (live && i < C && t < t_hi) returns[(int64_t)t * stride_t + e] = h + vv[i];
        }
        return;
    }

    // sweep 1: fold the chunk into h -> B + A*h
    float A = 1.f, Bc = 0.f;
    if (live && t_lo < t_hi) {
        float v_next = (t_hi < T) ? __ldg(values + (int64_t)t_hi * stride_t + e) : 0.f;
        for (int t = t_hi - 1; t >= t_lo; --t) {
            const int64_t i = (int64_t)t * stride_t + e;
            const float rv = __ldg(rewards + i), vv = __ldg(values + i);
            const float m = __ldg(boundary + i) ? 0.f : 1.f;
            const float g = gl * m;
            Bc = fmaf(g, Bc, rv + gamma * v_next * m - vv);
            A *= g;
            v_next = vv;
        }
    }
    sA[w][lane] = A;
    sB[w][lane] = Bc;
    __syncthreads();
    if (!live || t_lo >= t_hi) return;
    float h = 0.f;                                        // gae at t_hi
    for (int c = kGaeChunks - 1; c > w; --c) h = fmaf(sA[c][lane], h, sB[c][lane]);

    // sweep 2: replay the chunk with the true carry (inputs now come from L2)
    float v_next = (t_hi < T) ? __ldg(values + (int64_t)t_hi * stride_t + e) : 0.f;
    for (int t = t_hi - 1; t >= t_lo; --t) {
        const int64_t i = (int64_t)t * stride_t + e;
        const float rv = __ldg(rewards + i), vv = __ldg(values + i);
        const float m = __ldg(boundary + i) ? 0.f : 1.f;
        h = fmaf(gl * m, h, rv + gamma * v_next * m - vv);
        returns[i] = h + vv;
        v_next = vv;
    }
}

}  // namespace xtrl

extern "C" int xtrl_gae(const float *rewards, const float *values, const uint8_t *is_boundary,
                        float *returns, int32_t E, int32_t T, int64_t stride_e, int64_t stride_t,
                        float gamma, float lam, void *stream) {
    using namespace xtrl;
    XTRL_REQUIRE(E >= 0 && T >= 0, XTRL_E_BADARG, "xtrl_gae: negative shape (E=%d, T=%d)", E, T);
    if (E == 0 || T == 0) return XTRL_OK;
    XTRL_REQUIRE(rewards && values && is_boundary && returns, XTRL_E_BADARG, "xtrl_gae: null pointer");
    const float gl = (float)((double)gamma * (double)lam);
    cudaStream_t s = (cudaStream_t)stream;
    if (stride_t == 1 && stride_e >= T) {
        const int blocks = (E + kGaeWarpsPerBlock - 1) / kGaeWarpsPerBlock;
        gae_episode_major_kernel<<<blocks, kGaeWarpsPerBlock * 32, 0, s>>>(
            rewards, values, is_boundary, returns, E, T, stride_e, gamma, gl);
        XTRL_LAUNCH_CHECK("gae_episode_major_kernel");
    } else if (stride_e == 1 && stride_t >= E) {
        const int blocks = (E + 31) / 32;
        if (T <= kGaeChunks * kGaeRegChunk)
            gae_time_major_kernel<true><<<blocks, kGaeChunks * 32, 0, s>>>(rewards, values, is_boundary, returns, E, T, stride_t, gamma, gl);
        else
            gae_time_major_kernel<false><<<blocks, kGaeChunks * 32, 0, s>>>(rewards, values, is_boundary, returns, E, T, stride_t, gamma, gl);
        XTRL_LAUNCH_CHECK("gae_time_major_kernel");
    } else {
        return fail(XTRL_E_UNSUPPORTED,
                    "xtrl_gae: layout must be episode-major (stride_t==1) or time-major (stride_e==1); "
                    "got stride_e=%lld stride_t=%lld", (long long)stride_e, (long long)stride_t);
    }
    return XTRL_OK;
}
