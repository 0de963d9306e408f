// learn_ln.cu -- LayerNorm over dim = 48 rows with a gain (x-transformers' LayerNorm: F.layer_norm(x) * gamma),
// forward and backward, for the final norm of the learn-side sequence model (x_transformers_rl.py:505-512 -> Decoder's
// final norm).  HBM-bound: one thread per row keeps the 48 values in registers (torch's generic
// kernel takes 187 us for 128 k such rows; this is a 50 MB pass).  The sub-layer kernels carry their own LayerNorms.
#include "learn_common.cuh"

namespace xtrl {
namespace ln {
using namespace learn;

__global__ void __launch_bounds__(128) ln_fwd_kernel(const float *__restrict__ x, const float *__restrict__ gamma, float *__restrict__ y, int N) {
    __shared__ float sG[D];
    if (threadIdx.x < D) sG[threadIdx.x] = gamma[threadIdx.x];
    __syncthreads();
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < N; t += gridDim.x * blockDim.x) {
        float xr[D], mean, rstd;
        load_row(x + (size_t)t * D, true, xr);
        ln_stats(xr, mean, rstd);
        float4 *dst = reinterpret_cast<float4 *>(y + (size_t)t * D);
#pragma unroll
        for (int i = 0; i < D / 4; ++i)
            dst[i] = make_float4((xr[4 * i] - mean) * rstd * sG[4 * i], (xr[4 * i + 1] - mean) * rstd * sG[4 * i + 1],
                                 (xr[4 * i + 2] - mean) * rstd * sG[4 * i + 2], (xr[4 * i + 3] - mean) * rstd * sG[4 * i + 3]);
    }
}

__global__ void __launch_bounds__(128) ln_bwd_kernel(const float *__restrict__ x, const float *__restrict__ dy, const float *__restrict__ gamma,
                                                     float *__restrict__ dx, float *__restrict__ dgamma, int N) {
    __shared__ float sG[D], sAcc[D];
    if (threadIdx.x < D) { sG[threadIdx.x] = gamma[threadIdx.x]; sAcc[threadIdx.x] = 0.f; }
    __syncthreads();
    float acc[D];
#pragma unroll
    for (int i = 0; i < D; ++i) acc[i] = 0.f;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < N; t += gridDim.x * blockDim.x) {
        float xr[D], g[D], mean, rstd;
        load_row(x + (size_t)t * D, true, xr);
        load_row(dy + (size_t)t * D, true, g);
        ln_stats(xr, mean, rstd);
        float c1 = 0.f, c2 = 0.f;
#pragma unroll
        for (int i = 0; i < D; ++i) {
            xr[i] = (xr[i] - mean) * rstd;
            acc[i] = fmaf(g[i], xr[i], acc[i]);
            g[i] *= sG[i];
            c1 += g[i]; c2 = fmaf(g[i], xr[i], c2);
        }
        c1 *= (1.f / D); c2 *= (1.f / D);
        float4 *dst = reinterpret_cast<float4 *>(dx + (size_t)t * D);
#pragma unroll
        for (int i = 0; i < D / 4; ++i)
            dst[i] = make_float4(rstd * (g[4 * i] - c1 - xr[4 * i] * c2), rstd * (g[4 * i + 1] - c1 - xr[4 * i + 1] * c2),
                                 rstd * (g[4 * i + 2] - c1 - xr[4 * i + 2] * c2), rstd * (g[4 * i + 3] - c1 - xr[4 * i + 3] * c2));
    }
#pragma unroll
    for (int i = 0; i < D; ++i) {
        const float s = warp_sum(acc[i]);
        if ((threadIdx.x & 31) == 0) atomicAdd(&sAcc[i], s);
    }
    __syncthreads();
    if (threadIdx.x < D && sAcc[threadIdx.x] != 0.f) atomicAdd(dgamma + threadIdx.x, sAcc[threadIdx.x]);
}

// Weight gradient of a skinny input layer: dW[k][j] = sum_t f[t][k] * dy[t][j] for K <= 16 input features and 48
// outputs, where f is either a dense [N][K] fp32 matrix (to_state_embed: state -> 48, with the bias as sum_t dy) or the
// one-hot of an id in [-1, K) (the action embedding; -1 = no row).  A 128 k x 8 by 128 k x 48 product is one 25 MB pass
// over dy; cuBLAS' fp32 SIMT kernel takes 180 us for it.  Block = 5 token groups x 48 output columns.
constexpr int kSmallK = 16;
template <bool IDS>
__global__ void __launch_bounds__(240) small_dw_kernel(const float *__restrict__ f, const int64_t *__restrict__ ids, const float *__restrict__ dy,
                                                       float *__restrict__ dW, float *__restrict__ db, int N, int K, int out_kj) {
    __shared__ float sAcc[kSmallK + 1][D];
    for (int i = threadIdx.x; i < (kSmallK + 1) * D; i += blockDim.x) (&sAcc[0][0])[i] = 0.f;
    __syncthreads();
    const int j = threadIdx.x % D, g = threadIdx.x / D;
    float acc[kSmallK], accb = 0.f;
#pragma unroll
    for (int k = 0; k < kSmallK; ++k) acc[k] = 0.f;
    for (int t = blockIdx.x * 5 + g; t < N; t += gridDim.x * 5) {
        const float d = __ldg(dy + (size_t)t * D + j);
        accb += d;
        if (IDS) {
            const int id = (int)__ldg(ids + t);
#pragma unroll
            for (int k = 0; k < kSmallK; ++k) acc[k] += id == k ? d : 0.f;
        } else {
#pragma unroll
            for (int k = 0; k < kSmallK; ++k)
                if (k < K) acc[k] = fmaf(__ldg(f + (size_t)t * K + k), d, acc[k]);
        }
    }
#pragma unroll
    for (int k = 0; k < kSmallK; ++k)
        if (k < K) atomicAdd(&sAcc[k][j], acc[k]);
    atomicAdd(&sAcc[kSmallK][j], accb);
    __syncthreads();
    for (int i = threadIdx.x; i < K * D; i += blockDim.x) {
        const int k = i / D, jj = i % D;
        atomicAdd(dW + (out_kj ? k * D + jj : jj * K + k), sAcc[k][jj]);
    }
    if (db && threadIdx.x < D) atomicAdd(db + threadIdx.x, sAcc[kSmallK][threadIdx.x]);
}

}  // namespace ln
}  // namespace xtrl

extern "C" int xtrl_small_dw(const float *feats, const int64_t *ids, const float *dy, float *d_w, float *d_b, int64_t N, int32_t K,
                             int32_t dim, int32_t out_kj, void *stream) {
    using namespace xtrl;
    XTRL_REQUIRE(dim == learn::D && K >= 1 && K <= ln::kSmallK, XTRL_E_UNSUPPORTED, "xtrl_small_dw: built for dim=48, K <= 16 (got %d, %d)", dim, K);
    XTRL_REQUIRE(N >= 0 && N <= (int64_t)1 << 30, XTRL_E_BADARG, "xtrl_small_dw: bad row count");
    if (N == 0) return XTRL_OK;
    XTRL_REQUIRE((feats != nullptr) != (ids != nullptr) && dy && d_w, XTRL_E_BADARG, "xtrl_small_dw: exactly one of feats / ids, and dy, d_w");
    const int64_t want = (N + 5 * 64 - 1) / (5 * 64);
    const int blocks = (int)(want < 4 * num_sms() ? want : 4 * num_sms());
    if (ids) ln::small_dw_kernel<true><<<blocks, 240, 0, (cudaStream_t)stream>>>(feats, ids, dy, d_w, d_b, (int)N, K, out_kj);
    else ln::small_dw_kernel<false><<<blocks, 240, 0, (cudaStream_t)stream>>>(feats, ids, dy, d_w, d_b, (int)N, K, out_kj);
    XTRL_LAUNCH_CHECK("small_dw_kernel");
    return XTRL_OK;
}

extern "C" int xtrl_ln_fwd(const float *x, const float *gamma, float *y, int64_t N, int32_t dim, void *stream) {
    using namespace xtrl;
    XTRL_REQUIRE(dim == learn::D, XTRL_E_UNSUPPORTED, "xtrl_ln_fwd: built for dim=%d (got %d)", learn::D, dim);
    XTRL_REQUIRE(N >= 0 && N <= (int64_t)1 << 30, XTRL_E_BADARG, "xtrl_ln_fwd: bad row count");
    if (N == 0) return XTRL_OK;
    XTRL_REQUIRE(x && gamma && y, XTRL_E_BADARG, "xtrl_ln_fwd: null pointer");
    const int blocks = (int)((N + 127) / 128);
    ln::ln_fwd_kernel<<<blocks < 8 * num_sms() ? blocks : 8 * num_sms(), 128, 0, (cudaStream_t)stream>>>(x, gamma, y, (int)N);
    XTRL_LAUNCH_CHECK("ln_fwd_kernel");
    return XTRL_OK;
}

extern "C" int xtrl_ln_bwd(const float *x, const float *dy, const float *gamma, float *dx, float *d_gamma, int64_t N, int32_t dim, void *stream) {
    using namespace xtrl;
    XTRL_REQUIRE(dim == learn::D, XTRL_E_UNSUPPORTED, "xtrl_ln_bwd: built for dim=%d (got %d)", learn::D, dim);
    XTRL_REQUIRE(N >= 0 && N <= (int64_t)1 << 30, XTRL_E_BADARG, "xtrl_ln_bwd: bad row count");
    if (N == 0) return XTRL_OK;
    XTRL_REQUIRE(x && dy && gamma && dx && d_gamma, XTRL_E_BADARG, "xtrl_ln_bwd: null pointer");
    const int blocks = (int)((N + 127) / 128);
    ln::ln_bwd_kernel<<<blocks < 4 * num_sms() ? blocks : 4 * num_sms(), 128, 0, (cudaStream_t)stream>>>(x, dy, gamma, dx, d_gamma, (int)N);
    XTRL_LAUNCH_CHECK("ln_bwd_kernel");
    return XTRL_OK;
}
