// wm_step.cu -- one rollout timestep of the world-model / actor-critic for a batch of envs.
//
// Replaces, for every env at once (reference is batch 1, ~45 ATen launches per layer, >=9 host
// syncs per step -- x_transformers_rl/x_transformers_rl.py:1250-1289):
//   RSNorm.eval (:588-594) -> project_in + SafeEmbedding(prev action) + reward conditioning
//   (:492-503) -> x-transformers Decoder with a KV cache (call site :721-734 / :505-512; SURVEY
//   App. A.1: pre-LN, q/k/v/out without bias, dim_head 16, optional sigmoid value gate and
//   value-residual mix, GELU feed-forward, final LN) -> cat(embed, state_embed[, latent])
//   (:537-549) -> action_head / critic_head (:348-369) -> Discrete (:197-222): inverse-CDF sample
//   + clamped log-prob -> HL-Gauss expectation of the critic logits (:843).
//
// One launch per timestep.  A CTA owns kTE envs; their activations never leave shared memory.
// Bound: HBM, by the KV-cache stream -- L*2*I*(t+1) elements per env-step against ~0.9 MFLOP of
// dense math.  Layout [k|v][layer][env][head][Tmax][16]: the stream of one
// (env, head) is contiguous, a warp owns one (env, head), lane j reads key j's 64 B row
// (4 x LDG.128, no L1 allocation), keeps a private online softmax, and the 32 partial
// (max, sum, acc[16]) states are merged once by shuffles.  Dense parts: fp32 FFMA with weights
// stored K-major so a warp reads consecutive output features (coalesced, L1/L2 resident).
#include "common.cuh"
#include "umma.cuh"
#include <math.h>

namespace xtrl {

#ifndef XTRL_TE
#define XTRL_TE 16
#endif
#ifndef XTRL_MINBLOCKS
#define XTRL_MINBLOCKS 2
#endif
#ifndef XTRL_KVBUF
#define XTRL_KVBUF 2
#endif
constexpr int kTE = XTRL_TE;       // envs per CTA
constexpr int kMinBlocks = XTRL_MINBLOCKS;   // resident CTAs per SM the register budget is sized for
constexpr int kKvBuf = XTRL_KVBUF; // 32-key chunks in flight per warp in the attention stream
constexpr int kThreads = 256;
constexpr int kDH = 16;            // dim_head the kernels are built for
constexpr int kSplitK = 64 / XTRL_TE;   // K slices of the narrow (N = D) residual layers (4 at kTE=16, 8 at kTE=8)

static inline int64_t align4(int64_t x) { return (x + 3) & ~(int64_t)3; }

struct StepDims {
    int S, Spad, A, Apad, D, H, I, L, K, Kpad, G, NQ, HIN, Tmax;
    int gate, vres, mix, rot, kvbf16, tc, cont, squash;
};

static int make_dims(const xtrl_model_desc *d, StepDims *o) {
    o->S = d->state_dim; o->Spad = (int)align4(d->state_dim + 0); o->A = d->num_actions;
    o->cont = d->continuous_actions != 0; o->squash = d->squash_continuous != 0;
    o->Apad = ((o->cont ? 2 : 1) * d->num_actions + 3) & ~3;      /* actor head width: A logits, or A (mean, log-variance) pairs */ o->D = d->dim; o->H = d->heads; o->I = d->heads * d->dim_head;
    o->L = d->depth; o->K = d->num_bins; o->Kpad = (d->num_bins + 3) & ~3; o->G = d->dim_gene;
    o->gate = d->gate_values != 0; o->vres = d->value_residual != 0;
    o->mix = o->vres && d->learned_mix != 0;
    o->NQ = 3 * o->I + (o->gate ? o->I : 0) + (o->mix ? ((o->H + 3) & ~3) : 0);
    o->HIN = (d->dim_gene > 0 ? 3 : 2) * d->dim; o->Tmax = d->max_timesteps;
    o->rot = d->rotary_true_pos != 0; o->kvbf16 = d->kv_bf16 != 0; o->tc = d->dense_bf16 != 0;
    return 0;
}

static int check_desc(const xtrl_model_desc *d) {
    XTRL_REQUIRE(d, XTRL_E_BADARG, "null model desc");
    XTRL_REQUIRE(d->dim_head == kDH, XTRL_E_UNSUPPORTED, "dim_head=%d: kernels are built for dim_head=16", d->dim_head);
    XTRL_REQUIRE(d->dim >= 16 && d->dim <= 64 && d->dim % 4 == 0, XTRL_E_UNSUPPORTED,
                 "dim=%d: supported hidden_dim is a multiple of 4 in [16,64] (reference default 48)", d->dim);
    XTRL_REQUIRE(d->heads >= 1 && d->heads <= 16, XTRL_E_UNSUPPORTED, "heads=%d outside [1,16]", d->heads);
    XTRL_REQUIRE(d->depth >= 1 && d->depth <= XTRL_MAX_DEPTH, XTRL_E_UNSUPPORTED, "depth=%d outside [1,%d]", d->depth, XTRL_MAX_DEPTH);
    XTRL_REQUIRE(d->state_dim >= 1 && d->state_dim <= 63, XTRL_E_UNSUPPORTED, "state_dim=%d outside [1,63]", d->state_dim);
    XTRL_REQUIRE(d->num_actions >= 1 && d->num_actions <= 32, XTRL_E_UNSUPPORTED, "num_actions=%d outside [1,32]", d->num_actions);
    XTRL_REQUIRE(d->num_bins >= 2 && d->num_bins <= 256, XTRL_E_UNSUPPORTED, "num_bins=%d outside [2,256]", d->num_bins);
    XTRL_REQUIRE(d->dim_gene >= 0 && d->dim_gene <= 256 && d->dim_gene % 4 == 0, XTRL_E_UNSUPPORTED,
                 "dim_gene=%d: must be a multiple of 4 in [0,256]", d->dim_gene);
    XTRL_REQUIRE(d->max_timesteps >= 1, XTRL_E_BADARG, "max_timesteps=%d", d->max_timesteps);
    XTRL_REQUIRE(!d->learned_mix || d->value_residual, XTRL_E_BADARG, "learned_mix without value_residual");
    XTRL_REQUIRE(!d->dense_bf16 || d->dim % 16 == 0, XTRL_E_UNSUPPORTED,
                 "dense_bf16: hidden dim %d must be a multiple of 16 (UMMA K step)", d->dim);
    return XTRL_OK;
}

}  // namespace xtrl

extern "C" int xtrl_wm_weight_layout(const xtrl_model_desc *desc, xtrl_weight_layout *out) {
    using namespace xtrl;
    int rc = check_desc(desc);
    if (rc) return rc;
    XTRL_REQUIRE(out, XTRL_E_BADARG, "xtrl_wm_weight_layout: null out");
    StepDims d;
    make_dims(desc, &d);
    int64_t o = 0;
    auto take = [&](int64_t n) { int64_t at = o; o += align4(n); return at; };
    out->s_pad = d.Spad; out->n_qkv = d.NQ; out->head_in = d.HIN;
    out->in_proj = take((int64_t)d.Spad * 2 * d.D);
    out->state_embed_bias = take(d.D);
    out->action_embed = take((int64_t)d.A * d.D);
    out->reward_embed = take(d.D);
    out->latent_w = take((int64_t)d.G * d.D);
    out->latent_b = take(d.D);
    out->final_gamma = take(d.D);
    out->head1_w = take((int64_t)d.HIN * 4 * d.D);
    out->head1_b = take(4 * d.D);
    out->actor2_w = take((int64_t)2 * d.D * d.Apad);
    out->actor2_b = take(d.Apad);
    out->critic2_w = take((int64_t)2 * d.D * d.Kpad);
    out->critic2_b = take(d.Kpad);
    for (int l = 0; l < XTRL_MAX_DEPTH; ++l) {
        if (l >= d.L) { out->layer[l] = {-1, -1, -1, -1, -1, -1, -1, -1, -1}; continue; }
        out->layer[l].ln1_gamma = take(d.D);
        out->layer[l].qkv_w = take((int64_t)d.D * d.NQ);
        out->layer[l].qkv_b = take(d.NQ);
        out->layer[l].out_w = take((int64_t)d.I * d.D);
        out->layer[l].ln2_gamma = take(d.D);
        out->layer[l].ff1_w = take((int64_t)d.D * 4 * d.D);
        out->layer[l].ff1_b = take(4 * d.D);
        out->layer[l].ff2_w = take((int64_t)4 * d.D * d.D);
        out->layer[l].ff2_b = take(d.D);
    }
    out->action_embed_bias = d.cont ? take(d.D) : -1;
    out->total = o;
    return XTRL_OK;
}

extern "C" int64_t xtrl_wm_kv_cache_bytes(const xtrl_model_desc *desc, int64_t num_envs) {
    using namespace xtrl;
    if (check_desc(desc) || num_envs < 0) return -1;
    const int64_t elems = 2LL * desc->depth * num_envs * desc->heads * desc->max_timesteps * desc->dim_head;
    return elems * (desc->kv_bf16 ? 2 : 4);
}

namespace xtrl {

// ---- tensor-core path: where each GEMM's bf16 operand image lives in the tc blob (byte offsets) ----------
constexpr int kWbufBytes = 49152;          // shared-memory weight stage: one 128 x 192 bf16 tile (FF2), 4 tiles at K=48
struct TcLayout {
    int64_t total;
    int tiles_q, tiles_f;
    int64_t qkv[XTRL_MAX_DEPTH], out[XTRL_MAX_DEPTH], ff1[XTRL_MAX_DEPTH], ff2[XTRL_MAX_DEPTH];
    int64_t head_a, head_c;
};
static TcLayout tc_layout(const StepDims &d) {
    TcLayout t{};
    int64_t o = 0;
    auto take = [&](int tiles, int K) { int64_t at = o; o += (int64_t)tiles * 128 * K * 2; return at; };
    t.tiles_q = (d.NQ + 127) / 128; t.tiles_f = (4 * d.D + 127) / 128;
    for (int l = 0; l < d.L; ++l) {
        t.qkv[l] = take(t.tiles_q, d.D); t.out[l] = take(1, d.I);
        t.ff1[l] = take(t.tiles_f, d.D); t.ff2[l] = take(1, 4 * d.D);
    }
    t.head_a = take(1, d.HIN); t.head_c = take(1, d.HIN);
    t.total = o;
    return t;
}

// fp32 K-major W[K][ld] (columns coff .. coff+nvalid) -> bf16 canonical 128-row tile images
__global__ void pack_tc_kernel(const float *__restrict__ src, int ld, int coff, int nvalid, int K, int tiles,
                               __nv_bfloat16 *__restrict__ dst) {
    const int64_t total = (int64_t)tiles * 128 * K;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int tile = (int)(i / (128 * K)), rem = (int)(i - (int64_t)tile * 128 * K), r = rem / K, k = rem - r * K;
        const int n = tile * 128 + r;
        const float v = n < nvalid ? src[(int64_t)k * ld + coff + n] : 0.f;
        dst[(int64_t)tile * 128 * K + umma::img_off_bf16(128, r, k) / 2] = __float2bfloat16_rn(v);
    }
}

struct StepArgs {
    StepDims d;
    xtrl_weight_layout w;
    xtrl_step_io io;
    const float *weights;
    void *kv;
    int B;
    float reward_lo, reward_hi;
    float hl_step;
    TcLayout tc;
    const unsigned char *tcw;
    int epc;                       // envs owned per CTA (<= kTE)
};

struct Smem {                       // offsets in floats into the dynamic shared-memory block
    int x, xn, qkv, v1, ao, ff, st, hin, gene, logit, total;
    int wbuf, xn_img, ao_img, ff_img, hin_img, bars;       // tensor-core path only
};

__host__ __device__ static inline Smem smem_layout(const StepDims &d) {
    Smem s{}; int o = 0;
    auto take = [&](int n) { int at = o; o += (n + 3) & ~3; return at; };
    if (d.tc) {
        // operands of the GEMMs live only as bf16 UMMA images; fp32 copies exist only where SIMT code reads them
        s.wbuf = take(kWbufBytes / 4);                                   // first: 128-byte aligned with the block
        s.xn_img = take(kTE * d.D / 2); s.ao_img = take(kTE * d.I / 2);
        s.ff_img = take(kTE * 4 * d.D / 2); s.hin_img = take(kTE * d.HIN / 2);
        s.bars = take(8);
        s.x = take(kTE * d.D);
        s.qkv = take(kTE * d.NQ > kTE * 4 * d.D ? kTE * d.NQ : kTE * 4 * d.D);   // later reused as the fp32 head hidden
        s.ff = s.qkv;
        s.v1 = take(d.vres ? kTE * d.I : 0);
        s.st = take(kTE * (d.Spad + 4));
        s.gene = take(kTE * d.G);
        s.logit = take(kTE * (d.Apad + d.Kpad));
        s.xn = s.ao = s.hin = 0;
        s.total = o;
        return s;
    }
    s.x = take(kTE * d.D);
    s.xn = take(kTE * d.D);
    s.qkv = take(kTE * d.NQ > kSplitK * kTE * d.D ? kTE * d.NQ : kSplitK * kTE * d.D);   // also split-K scratch
    s.v1 = take(d.vres ? kTE * d.I : 0);
    s.ao = take(kTE * d.I);
    s.ff = take(kTE * 4 * d.D);
    s.st = take(kTE * (d.Spad + 4));          // normalised state (padded) + normalised reward
    s.hin = take(kTE * d.HIN);
    s.gene = take(kTE * d.G);
    s.logit = take(kTE * (d.Apad + d.Kpad));
    s.total = o;
    return s;
}

// Y[e][n] = sum_k X[e][k] * Wt[k][n] for the CTA's kTE envs.  A work item is (4 adjacent columns, a
// group of ES envs, one of KS slices of K); items are spread over the CTA so narrow layers still use
// every warp.  W moves as coalesced LDG.128 rows (two 4-row blocks in flight), X as LDS.128.
// epi(e, n, ks, acc) consumes each (partial, when KS > 1) output.
template <int ES, int KS, class Epi>
__device__ __forceinline__ void dense_tile(const float *__restrict__ Wt, int K, int N,
                                           const float *__restrict__ sX, int ldx, Epi epi) {
    const int quads = N >> 2, groups = kTE / ES, kper = K / KS;
    for (int item = threadIdx.x; item < quads * groups * KS; item += kThreads) {
        const int q = item % quads, rest = item / quads, e0 = (rest % groups) * ES, ks = rest / groups;
        const float *w = Wt + 4 * q;
        float acc[ES][4];
#pragma unroll
        for (int e = 0; e < ES; ++e) { acc[e][0] = 0.f; acc[e][1] = 0.f; acc[e][2] = 0.f; acc[e][3] = 0.f; }
#pragma unroll 2
        for (int k = ks * kper; k < (ks + 1) * kper; k += 4) {
            const float4 w0 = __ldg((const float4 *)(w + (int64_t)(k + 0) * N));
            const float4 w1 = __ldg((const float4 *)(w + (int64_t)(k + 1) * N));
            const float4 w2 = __ldg((const float4 *)(w + (int64_t)(k + 2) * N));
            const float4 w3 = __ldg((const float4 *)(w + (int64_t)(k + 3) * N));
#pragma unroll
            for (int e = 0; e < ES; ++e) {
                const float4 x = *(const float4 *)(sX + (e0 + e) * ldx + k);
                acc[e][0] = fmaf(x.x, w0.x, acc[e][0]); acc[e][1] = fmaf(x.x, w0.y, acc[e][1]);
                acc[e][2] = fmaf(x.x, w0.z, acc[e][2]); acc[e][3] = fmaf(x.x, w0.w, acc[e][3]);
                acc[e][0] = fmaf(x.y, w1.x, acc[e][0]); acc[e][1] = fmaf(x.y, w1.y, acc[e][1]);
                acc[e][2] = fmaf(x.y, w1.z, acc[e][2]); acc[e][3] = fmaf(x.y, w1.w, acc[e][3]);
                acc[e][0] = fmaf(x.z, w2.x, acc[e][0]); acc[e][1] = fmaf(x.z, w2.y, acc[e][1]);
                acc[e][2] = fmaf(x.z, w2.z, acc[e][2]); acc[e][3] = fmaf(x.z, w2.w, acc[e][3]);
                acc[e][0] = fmaf(x.w, w3.x, acc[e][0]); acc[e][1] = fmaf(x.w, w3.y, acc[e][1]);
                acc[e][2] = fmaf(x.w, w3.z, acc[e][2]); acc[e][3] = fmaf(x.w, w3.w, acc[e][3]);
            }
        }
#pragma unroll
        for (int e = 0; e < ES; ++e) {
#pragma unroll
            for (int c = 0; c < 4; ++c) epi(e0 + e, 4 * q + c, ks, acc[e][c]);
        }
    }
}

// narrow residual layers (N = D): split K four ways so the whole CTA works, partial sums meet in a
// scratch tile [KS][kTE][D] and are folded into the residual stream in a fixed order (deterministic).
template <class Fin>
__device__ __forceinline__ void dense_residual(const float *__restrict__ Wt, int K, int D,
                                               const float *__restrict__ sX, int ldx,
                                               float *__restrict__ scratch, float *__restrict__ sx, Fin bias) {
    dense_tile<4, kSplitK>(Wt, K, D, sX, ldx, [&](int e, int n, int ks, float acc) {
        scratch[(ks * kTE + e) * D + n] = acc;
    });
    __syncthreads();
    for (int i = threadIdx.x; i < kTE * D; i += kThreads) {
        float v = sx[i] + bias(i % D);
#pragma unroll
        for (int ks = 0; ks < kSplitK; ++ks) v += scratch[ks * kTE * D + i];
        sx[i] = v;
    }
}

// x-transformers bias-less LayerNorm (eps 1e-5) * gamma; a warp normalises kTE/8 envs.
__device__ __forceinline__ void layer_norm_tile(const float *__restrict__ sIn, float *__restrict__ sOut, int ldo,
                                                const float *__restrict__ gamma, int D) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int e = warp; e < kTE; e += kThreads / 32) {
        const float a = lane < D ? sIn[e * D + lane] : 0.f;
        const float b = lane + 32 < D ? sIn[e * D + lane + 32] : 0.f;
        const float mean = warp_sum(a + b) / (float)D;
        const float da = lane < D ? a - mean : 0.f, db = lane + 32 < D ? b - mean : 0.f;
        const float rstd = rsqrtf(warp_sum(da * da + db * db) / (float)D + 1e-5f);
        if (lane < D) sOut[e * ldo + lane] = da * rstd * __ldg(gamma + lane);
        if (lane + 32 < D) sOut[e * ldo + lane + 32] = db * rstd * __ldg(gamma + lane + 32);
    }
}

// same normalisation, written as the bf16 B-operand image of the next GEMM (kTE rows, K = D)
__device__ __forceinline__ void layer_norm_img(const float *__restrict__ sIn, unsigned char *__restrict__ img,
                                               const float *__restrict__ gamma, int D) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int e = warp; e < kTE; e += kThreads / 32) {
        const float a = lane < D ? sIn[e * D + lane] : 0.f;
        const float b = lane + 32 < D ? sIn[e * D + lane + 32] : 0.f;
        const float mean = warp_sum(a + b) / (float)D;
        const float da = lane < D ? a - mean : 0.f, db = lane + 32 < D ? b - mean : 0.f;
        const float rstd = rsqrtf(warp_sum(da * da + db * db) / (float)D + 1e-5f);
        if (lane < D) *(__nv_bfloat16 *)(img + umma::img_off_bf16(kTE, e, lane)) = __float2bfloat16_rn(da * rstd * __ldg(gamma + lane));
        if (lane + 32 < D) *(__nv_bfloat16 *)(img + umma::img_off_bf16(kTE, e, lane + 32)) = __float2bfloat16_rn(db * rstd * __ldg(gamma + lane + 32));
    }
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float silu(float x) { return x / (1.f + expf(-x)); }
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// KV-cache element traits.  A row (one key or value of one head) is dim_head elements = PPR 16-byte
// pieces of EP elements.  The stream of one (env, head) is read as consecutive 16-byte pieces by
// consecutive lanes -- every LDG.128 of a warp covers 512 contiguous bytes, each 32-byte sector is
// fetched from L2 exactly once -- so lane = (row group, piece-of-row): grp = lane / PPR, sub = lane % PPR.
template <typename KV> struct KvTraits;
template <> struct KvTraits<float> {
    static constexpr int EP = 4, PPR = 4, LOG_PPR = 2;
    static __device__ __forceinline__ void unpack(const uint4 &r, float (&o)[4]) {
        o[0] = __uint_as_float(r.x); o[1] = __uint_as_float(r.y); o[2] = __uint_as_float(r.z); o[3] = __uint_as_float(r.w);
    }
    static __device__ __forceinline__ void store1(float *p, float v) { *p = v; }
};
template <> struct KvTraits<__nv_bfloat16> {
    static constexpr int EP = 8, PPR = 2, LOG_PPR = 1;
    static __device__ __forceinline__ void unpack(const uint4 &r, float (&o)[8]) {
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) { o[2 * j] = __uint_as_float(w[j] << 16); o[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u); }
    }
    static __device__ __forceinline__ void store1(__nv_bfloat16 *p, float v) { *p = __float2bfloat16_rn(v); }
};

template <typename KV, bool TC>
__global__ void __launch_bounds__(kThreads, kMinBlocks)
wm_step_kernel(const StepArgs a) {
    extern __shared__ __align__(128) float sm[];
    const StepDims &d = a.d;
    const Smem so = smem_layout(d);
    float *sx = sm + so.x, *sxn = sm + so.xn, *sqkv = sm + so.qkv, *sv1 = sm + so.v1, *sao = sm + so.ao;
    float *sff = sm + so.ff, *sst = sm + so.st, *shin = sm + so.hin, *sgene = sm + so.gene, *slog = sm + so.logit;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // a CTA has kTE env slots but owns only `epc` envs (the host sizes epc so the grid fills the SMs evenly);
    // slots e >= epc are padding: env index B, i.e. invalid everywhere below
    const int epc = a.epc, env0 = blockIdx.x * epc;
    auto env_of = [&](int e) { return e < epc ? env0 + e : a.B; };
    const int D = d.D, I = d.I, H = d.H, NQ = d.NQ, S = d.S, Spad = d.Spad;
    const float *W = a.weights;
    const int lst = Spad + 4;
    using TR = KvTraits<KV>;
    constexpr int EP = TR::EP, PPR = TR::PPR, RPI = 32 / TR::PPR;     // elems per piece, pieces per row, rows per LDG

    // ---- tensor-core plumbing (TC): one weight stage fed by TMA bulk copies, accumulators in TMEM ----
    unsigned char *wbuf = (unsigned char *)(sm + so.wbuf), *xn_img = (unsigned char *)(sm + so.xn_img);
    unsigned char *ao_img = (unsigned char *)(sm + so.ao_img), *ff_img = (unsigned char *)(sm + so.ff_img);
    unsigned char *hin_img = (unsigned char *)(sm + so.hin_img);
    uint64_t *bars = (uint64_t *)(sm + so.bars);                        // [0] weights landed, [1] MMAs complete
    uint32_t *tslot = (uint32_t *)(bars + 2);
    uint32_t tbase = 0, pw = 0, pm = 0;                                 // TMEM base, barrier phase parities
    constexpr uint32_t kTmemCols = 64;                                  // up to 4 tiles x 16 fp32 columns
    auto tc_load = [&](int64_t off, uint32_t bytes) {                   // thread 0: next weights -> wbuf
        if (tid == 0) { umma::mbar_expect_tx(&bars[0], bytes); umma::bulk_g2s(wbuf, a.tcw + off, bytes, &bars[0]); }
    };
    // D[tile t] = W_tile[128 x K] * X[kTE x K]^T for nt tiles; returns when the accumulators are complete
    auto tc_mma = [&](int nt, int K, const unsigned char *bimg) {
        if (tid == 0) {
            umma::mbar_wait(&bars[0], pw);
            umma::tc_fence_after();
            const uint32_t idesc = umma::idesc_bf16(128, kTE);
            for (int t = 0; t < nt; ++t)
                for (int s = 0; s < K / 16; ++s)
                    umma::mma_bf16_ss(tbase + t * kTE, umma::smem_desc(wbuf + (size_t)t * 128 * K * 2 + s * 2 * (128 * 16), 128 * 16, 128),
                                      umma::smem_desc(bimg + s * 2 * (kTE * 16), kTE * 16, 128), idesc, s > 0);
            umma::mma_commit(&bars[1]);
        }
        pw ^= 1;
        umma::mbar_wait(&bars[1], pm);
        pm ^= 1;
        umma::tc_fence_after();
    };
    // accumulators -> registers: warp (q = warp%4, half = warp/4) reads TMEM lanes 32q.. and 8 of the 16 env
    // columns of every tile; thread <-> output feature.  epi(e, f, acc).  Ends with a CTA barrier.
    auto tc_epilogue = [&](int nt, int tile0, auto epi) {
        const int q = warp & 3, half = warp >> 2;
        for (int t = 0; t < nt; ++t) {
            float v[8];
            umma::tmem_ld8(tbase + t * kTE + half * 8 + ((uint32_t)(q * 32) << 16), v);
            const int f = (tile0 + t) * 128 + q * 32 + lane;
#pragma unroll
            for (int j = 0; j < 8; ++j) epi(half * 8 + j, f, v[j]);
        }
        umma::tc_fence_before();
        umma::fence_proxy_async();
        __syncthreads();
    };
    if constexpr (TC) {
        static_assert(kTE == 16 && kThreads == 256, "the UMMA path is built for 16 envs (N=16) and 8 warps per CTA");
        if (tid == 0) { umma::mbar_init(&bars[0], 1); umma::mbar_init(&bars[1], 1); umma::fence_mbar_init(); }
        if (warp == 0) umma::tmem_alloc(tslot, kTmemCols);
        umma::tc_fence_before();
        __syncthreads();
        umma::tc_fence_after();
        tbase = *tslot;
        const int g0 = min(a.tc.tiles_q, min(kWbufBytes / (128 * D * 2), (int)kTmemCols / kTE));
        tc_load(a.tc.qkv[0], (uint32_t)g0 * 128 * D * 2);               // layer 0 QKV weights fly during the prologue
    }

    // ---- RSNorm.eval on (state || prev reward) (:588-594); gene rows l2-normalised (evolution.py:73-74)
    for (int i = tid; i < kTE * lst; i += kThreads) {
        const int e = i / lst, j = i - e * lst, env = env_of(e);
        float v = 0.f;
        if (env < a.B && j <= Spad && (j < S || j == Spad)) {
            const int f = j < S ? j : S;                       // slot Spad holds the reward
            const float raw = j < S ? __ldg(a.io.state + (int64_t)env * S + j) : __ldg(a.io.prev_reward + env);
            v = (raw - __ldg(a.io.rs_mean + f)) / fmaxf(sqrtf(__ldg(a.io.rs_var + f)), 1e-5f);
        }
        sst[i] = v;
    }
    if (d.G > 0) {
        for (int e = warp; e < kTE; e += kThreads / 32) {
            const int env = env_of(e);
            const float *row = env < a.B ? a.io.genes + (int64_t)__ldg(a.io.gene_ids + env) * d.G : nullptr;
            float ss = 0.f;
            for (int g = lane; g < d.G; g += 32) { const float v = row ? __ldg(row + g) : 0.f; sgene[e * d.G + g] = v; ss = fmaf(v, v, ss); }
            ss = warp_sum(ss);
            const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
            for (int g = lane; g < d.G; g += 32) sgene[e * d.G + g] *= inv;
        }
    }
    __syncthreads();

    // ---- token embedding (:492-503) and the two state-only head inputs (:492, :544)
    dense_tile<kTE / 8, 1>(W + a.w.in_proj, Spad, 2 * D, sst, lst, [&](int e, int n, int, float acc) {
        const int env = env_of(e);
        if (n < D) {
            float v = acc;
            if (env < a.B) {
                if (d.cont && a.io.prev_action_f == nullptr) {
                    // deploy call without a previous action (:1056-1061): no action embedding at all
                } else if (d.cont) {                                                     // nn.Linear(A, D) on the previous action (:309)
                    v += __ldg(W + a.w.action_embed_bias + n);
                    for (int i = 0; i < d.A; ++i)
                        v = fmaf(__ldg(a.io.prev_action_f + (int64_t)env * d.A + i), __ldg(W + a.w.action_embed + (int64_t)i * D + n), v);
                } else {
                    const int pa = __ldg(a.io.prev_action + env);
                    if (pa >= 0) v += __ldg(W + a.w.action_embed + (int64_t)pa * D + n); // SafeEmbedding :181-195
                }
                v += sst[e * lst + Spad] * __ldg(W + a.w.reward_embed + n) * a.io.reward_flag;
            }
            sx[e * D + n] = v;
        } else {
            const float v = acc + __ldg(W + a.w.state_embed_bias + n - D);                  // to_state_embed
            if constexpr (TC) *(__nv_bfloat16 *)(hin_img + umma::img_off_bf16(kTE, e, n)) = __float2bfloat16_rn(v);
            else shin[e * d.HIN + n] = v;
        }
    });
    if (d.G > 0) {
        dense_tile<kTE / 8, 1>(W + a.w.latent_w, d.G, D, sgene, d.G, [&](int e, int n, int, float acc) {
            const float v = acc + __ldg(W + a.w.latent_b + n);                              // latent_to_embed
            if constexpr (TC) *(__nv_bfloat16 *)(hin_img + umma::img_off_bf16(kTE, e, 2 * D + n)) = __float2bfloat16_rn(v);
            else shin[e * d.HIN + 2 * D + n] = v;
        });
    }
    __syncthreads();

    const int64_t kv_head_stride = (int64_t)d.Tmax * kDH;
    const int64_t kv_half = (int64_t)d.L * a.B * H * kv_head_stride;
    KV *kvbase = (KV *)a.kv;

    for (int l = 0; l < d.L; ++l) {
        const auto &wl = a.w.layer[l];
        // 128-row tiles of a K = D GEMM per pass: what fits the weight stage AND the TMEM columns (16 per tile)
        const int wgroup = min(kWbufBytes / (128 * D * 2), (int)kTmemCols / kTE);
        auto qkv_epi = [&](int e, int n, float acc) {
            if (n < NQ) {
                float v = acc + __ldg(W + wl.qkv_b + n);
                if (n >= 3 * I) v = sigmoidf_(v);              // to_v_gate / to_value_residual_mix
                sqkv[e * NQ + n] = v;
            }
        };
        if constexpr (TC) {
            layer_norm_img(sx, xn_img, W + wl.ln1_gamma, D);
            umma::fence_proxy_async();
            __syncthreads();
            for (int t0 = 0; t0 < a.tc.tiles_q; t0 += wgroup) {
                const int nt = min(wgroup, a.tc.tiles_q - t0), t1 = t0 + nt;
                tc_mma(nt, D, xn_img);
                if (t1 < a.tc.tiles_q) tc_load(a.tc.qkv[l] + (int64_t)t1 * 128 * D * 2, (uint32_t)min(wgroup, a.tc.tiles_q - t1) * 128 * D * 2);
                else tc_load(a.tc.out[l], 128u * I * 2);
                tc_epilogue(nt, t0, qkv_epi);
            }
        } else {
            layer_norm_tile(sx, sxn, D, W + wl.ln1_gamma, D);
            __syncthreads();
            // fused q | k | v | gate | mix projection
            dense_tile<kTE / 2, 1>(W + wl.qkv_w, D, NQ, sxn, D, [&](int e, int n, int, float acc) { qkv_epi(e, n, acc); });
            __syncthreads();
        }
        // value residual (App. A.1) and, in true-position mode, rotary on q and the new k
        if (d.vres) {
            const int mix_off = 3 * I + (d.gate ? I : 0);
            for (int i = tid; i < kTE * I; i += kThreads) {
                const int e = i / I, n = i - e * I;
                float *vp = sqkv + e * NQ + 2 * I + n;
                if (l == 0) sv1[i] = *vp;
                else {
                    const float v1 = sv1[i], v = *vp;
                    if (d.mix) { const float w = sqkv[e * NQ + mix_off + n / kDH]; const float df = v - v1; *vp = w < 0.5f ? fmaf(w, df, v1) : fmaf(-df, 1.f - w, v); }
                    else *vp = v1 * 0.5f + v * 0.5f;
                }
            }
        }
        if (d.rot) {
            for (int i = tid; i < kTE * H * 4 * 2; i += kThreads) {
                const int which = i & 1, pr = (i >> 1) & 3, eh = i >> 3, e = eh / H, h = eh - e * H;
                const int env = env_of(e);
                const float p = env < a.B ? (float)(a.io.pos ? __ldg(a.io.pos + env) : a.io.pos_uniform) : 0.f;
                const float ang = p * powf(10000.f, -(float)(2 * pr) / 8.f);
                float sn, cs;
                sincosf(ang, &sn, &cs);
                float *t = sqkv + e * NQ + which * I + h * kDH + 2 * pr;
                const float x0 = t[0], x1 = t[1];
                t[0] = x0 * cs - x1 * sn;
                t[1] = x1 * cs + x0 * sn;
            }
        }
        if (d.vres || d.rot) __syncthreads();

        // ---- attention over the cache: one warp per (env, head)
        for (int eh = warp; eh < epc * H; eh += kThreads / 32) {
            const int e = eh / H, h = eh - e * H, env = env0 + e;
            if (env >= a.B) continue;
            const int len = min(a.io.pos ? __ldg(a.io.pos + env) : a.io.pos_uniform, d.Tmax - 1);   // keys already cached
            KV *kc = kvbase + (((int64_t)l * a.B + env) * H + h) * kv_head_stride;
            KV *vc = kc + kv_half;
            const float *qs = sqkv + e * NQ + h * kDH;
            // append the new key / value (lanes 0-15: k, 16-31: v)
            {
                const float nv = qs[(lane < kDH ? I : 2 * I) + (lane & (kDH - 1))];
                TR::store1((lane < kDH ? kc : vc) + (int64_t)len * kDH + (lane & (kDH - 1)), nv);
            }
            const int sub = lane & (PPR - 1), grp = lane >> TR::LOG_PPR;
            // this lane's piece of every cached row: 16-byte pieces, PPR per row, 32-bit row offsets
            const uint4 *kpl = (const uint4 *)kc + sub, *vpl = (const uint4 *)vc + sub;
            float q[EP], acc[EP];
#pragma unroll
            for (int i = 0; i < EP; ++i) { q[i] = qs[sub * EP + i]; acc[i] = 0.f; }
            float m = -INFINITY, lsum = 0.f;
            // Cached keys 0..len-1.  Loads are unconditional: row indices past the end are clamped to the last
            // cached row (finite data) and masked out of the softmax, so the stream carries no per-load branches.
            const int last = len - 1;
            auto fetch = [&](int c, uint4 (&kr)[PPR], uint4 (&vr)[PPR]) {
#pragma unroll
                for (int i = 0; i < PPR; ++i) {
                    const int j = min(c * 32 + i * RPI + grp, last);
                    kr[i] = ld_stream_u4(kpl + j * PPR);
                    vr[i] = ld_stream_u4(vpl + j * PPR);
                }
            };
            auto process = [&](int c, const uint4 (&kr)[PPR], const uint4 (&vr)[PPR]) {
                float sc[PPR], vf[PPR][EP];
                float mn = m;
#pragma unroll
                for (int i = 0; i < PPR; ++i) {
                    float kf[EP];
                    TR::unpack(kr[i], kf);
                    TR::unpack(vr[i], vf[i]);
                    float part = 0.f;
#pragma unroll
                    for (int t = 0; t < EP; ++t) part = fmaf(q[t], kf[t], part);
#pragma unroll
                    for (int o = 1; o < PPR; o <<= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                    sc[i] = (c * 32 + i * RPI + grp) < len ? part * 0.25f : -INFINITY;     // dim_head^-0.5
                    mn = fmaxf(mn, sc[i]);
                }
                if (mn > -INFINITY) {                                      // a row group with no valid key yet skips
                    const float c0 = expf(m - mn);
                    lsum *= c0;
#pragma unroll
                    for (int t = 0; t < EP; ++t) acc[t] *= c0;
#pragma unroll
                    for (int i = 0; i < PPR; ++i) {
                        const float pi = expf(sc[i] - mn);
                        lsum += pi;
#pragma unroll
                        for (int t = 0; t < EP; ++t) acc[t] = fmaf(pi, vf[i][t], acc[t]);
                    }
                    m = mn;
                }
            };
            // software pipeline over 32-key chunks (two register buffers, ping-pong): the loads of chunk
            // c+1 are in flight while chunk c is folded into the online softmax
            const int nchunks = (len + 31) >> 5;
            if (nchunks > 0) {
                if constexpr (kKvBuf >= 2 && PPR == 2) {
                    // bf16 cache: a chunk is only 2 KB, so four register buffers keep the same 8 KB per warp in flight
                    uint4 ka[PPR], va[PPR], kb[PPR], vb[PPR], kc4[PPR], vc4[PPR], kd[PPR], vd[PPR];
                    fetch(0, ka, va);
                    if (1 < nchunks) fetch(1, kb, vb);
                    if (2 < nchunks) fetch(2, kc4, vc4);
                    for (int c = 0; c < nchunks; c += 4) {
                        if (c + 3 < nchunks) fetch(c + 3, kd, vd);
                        process(c, ka, va);
                        if (c + 4 < nchunks) fetch(c + 4, ka, va);
                        if (c + 1 < nchunks) process(c + 1, kb, vb);
                        if (c + 5 < nchunks) fetch(c + 5, kb, vb);
                        if (c + 2 < nchunks) process(c + 2, kc4, vc4);
                        if (c + 6 < nchunks) fetch(c + 6, kc4, vc4);
                        if (c + 3 < nchunks) process(c + 3, kd, vd);
                    }
                } else if constexpr (kKvBuf >= 2) {
                    uint4 ka[PPR], va[PPR], kb[PPR], vb[PPR];
                    fetch(0, ka, va);
                    for (int c = 0; c < nchunks; c += 2) {
                        if (c + 1 < nchunks) fetch(c + 1, kb, vb);
                        process(c, ka, va);
                        if (c + 2 < nchunks) fetch(c + 2, ka, va);
                        if (c + 1 < nchunks) process(c + 1, kb, vb);
                    }
                } else {                                               // one chunk in flight per warp; more warps per SM
                    uint4 ka[PPR], va[PPR];
                    for (int c = 0; c < nchunks; ++c) { fetch(c, ka, va); process(c, ka, va); }
                }
            }
            // the token itself (key `len`, still in shared memory): every lane scores its piece, row group 0 folds it
            {
                float part = 0.f, vself[EP];
#pragma unroll
                for (int t = 0; t < EP; ++t) { part = fmaf(q[t], qs[I + sub * EP + t], part); vself[t] = qs[2 * I + sub * EP + t]; }
#pragma unroll
                for (int o = 1; o < PPR; o <<= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                if (grp == 0) {
                    const float sc = part * 0.25f, mn = fmaxf(m, sc), c0 = expf(m - mn), pi = expf(sc - mn);
                    lsum = fmaf(lsum, c0, pi);
#pragma unroll
                    for (int t = 0; t < EP; ++t) acc[t] = fmaf(acc[t], c0, pi * vself[t]);
                    m = mn;
                }
            }
            // merge the row groups: rescale to the common max, then butterfly sums across groups
            float M = m;
#pragma unroll
            for (int o = PPR; o < 32; o <<= 1) M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, o));
            const float f = m == -INFINITY ? 0.f : expf(m - M);
            lsum *= f;
#pragma unroll
            for (int t = 0; t < EP; ++t) acc[t] *= f;
#pragma unroll
            for (int o = PPR; o < 32; o <<= 1) {
                lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
#pragma unroll
                for (int t = 0; t < EP; ++t) acc[t] += __shfl_xor_sync(0xffffffffu, acc[t], o);
            }
            if (grp == 0) {
                const float inv = 1.f / lsum;
#pragma unroll
                for (int t = 0; t < EP; ++t) {
                    const int dd = sub * EP + t;
                    float o = acc[t] * inv;
                    if (d.gate) o *= sqkv[e * NQ + 3 * I + h * kDH + dd];
                    if constexpr (TC) *(__nv_bfloat16 *)(ao_img + umma::img_off_bf16(kTE, e, h * kDH + dd)) = __float2bfloat16_rn(o);
                    else sao[e * I + h * kDH + dd] = o;
                }
            }
        }
        if constexpr (TC) {
            umma::fence_proxy_async();
            __syncthreads();
            // to_out + residual
            tc_mma(1, I, ao_img);
            tc_load(a.tc.ff1[l], (uint32_t)min(wgroup, a.tc.tiles_f) * 128 * D * 2);
            tc_epilogue(1, 0, [&](int e, int n, float acc) { if (n < D) sx[e * D + n] += acc; });
            layer_norm_img(sx, xn_img, W + wl.ln2_gamma, D);
            umma::fence_proxy_async();
            __syncthreads();
            for (int t0 = 0; t0 < a.tc.tiles_f; t0 += wgroup) {
                const int nt = min(wgroup, a.tc.tiles_f - t0), t1 = t0 + nt;
                tc_mma(nt, D, xn_img);
                if (t1 < a.tc.tiles_f) tc_load(a.tc.ff1[l] + (int64_t)t1 * 128 * D * 2, (uint32_t)min(wgroup, a.tc.tiles_f - t1) * 128 * D * 2);
                else tc_load(a.tc.ff2[l], 128u * 4 * D * 2);
                tc_epilogue(nt, t0, [&](int e, int n, float acc) {
                    if (n < 4 * D) *(__nv_bfloat16 *)(ff_img + umma::img_off_bf16(kTE, e, n)) = __float2bfloat16_rn(gelu_erf(acc + __ldg(W + wl.ff1_b + n)));
                });
            }
            tc_mma(1, 4 * D, ff_img);
            if (l + 1 < d.L) tc_load(a.tc.qkv[l + 1], (uint32_t)min(wgroup, a.tc.tiles_q) * 128 * D * 2);
            else tc_load(a.tc.head_a, 128u * d.HIN * 2);
            tc_epilogue(1, 0, [&](int e, int n, float acc) { if (n < D) sx[e * D + n] += acc + __ldg(W + wl.ff2_b + n); });
        } else {
        __syncthreads();
        // to_out + residual
        dense_residual(W + wl.out_w, I, D, sao, I, sqkv, sx, [](int) { return 0.f; });   // sqkv is dead here: scratch
        __syncthreads();
        layer_norm_tile(sx, sxn, D, W + wl.ln2_gamma, D);
        __syncthreads();
        dense_tile<kTE / 4, 1>(W + wl.ff1_w, D, 4 * D, sxn, D, [&](int e, int n, int, float acc) {
            sff[e * 4 * D + n] = gelu_erf(acc + __ldg(W + wl.ff1_b + n));
        });
        __syncthreads();
        dense_residual(W + wl.ff2_w, 4 * D, D, sff, 4 * D, sqkv, sx, [&](int n) { return __ldg(W + wl.ff2_b + n); });
        __syncthreads();
        }
    }

    // ---- final norm -> heads (:533-557)
    if constexpr (TC) {
        layer_norm_img(sx, hin_img, W + a.w.final_gamma, D);               // embed -> columns [0, D) of the head input
        umma::fence_proxy_async();
        __syncthreads();
        tc_mma(1, d.HIN, hin_img);                                         // action_head.0
        tc_load(a.tc.head_c, 128u * d.HIN * 2);
        tc_epilogue(1, 0, [&](int e, int n, float acc) { if (n < 2 * D) sff[e * 4 * D + n] = silu(acc + __ldg(W + a.w.head1_b + n)); });
        tc_mma(1, d.HIN, hin_img);                                         // critic_head.0
        tc_epilogue(1, 0, [&](int e, int n, float acc) { if (n < 2 * D) sff[e * 4 * D + 2 * D + n] = silu(acc + __ldg(W + a.w.head1_b + 2 * D + n)); });
        if (warp == 0) umma::tmem_dealloc(tbase, kTmemCols);               // no tcgen05 op follows
    } else {
    layer_norm_tile(sx, shin, d.HIN, W + a.w.final_gamma, D);
    __syncthreads();
    dense_tile<kTE / 4, 1>(W + a.w.head1_w, d.HIN, 4 * D, shin, d.HIN, [&](int e, int n, int, float acc) {
        sff[e * 4 * D + n] = silu(acc + __ldg(W + a.w.head1_b + n));
    });
    __syncthreads();
    }
    const int ll = d.Apad + d.Kpad;
    dense_tile<1, 1>(W + a.w.actor2_w, 2 * D, d.Apad, sff, 4 * D, [&](int e, int n, int, float acc) {
        slog[e * ll + n] = acc + __ldg(W + a.w.actor2_b + n);
    });
    dense_tile<kTE / 8, 1>(W + a.w.critic2_w, 2 * D, d.Kpad, sff + 2 * D, 4 * D, [&](int e, int n, int, float acc) {
        slog[e * ll + d.Apad + n] = acc + __ldg(W + a.w.critic2_b + n);
    });
    __syncthreads();

    // ---- Discrete (:197-222) + HL-Gauss expectation (:843): a warp finishes kTE/8 envs
    for (int e = warp; e < kTE; e += kThreads / 32) {
        const int env = env_of(e);
        if (env >= a.B) continue;
        const float *la = slog + e * ll, *lv = la + d.Apad;
        // critic: softmax over K bins . bin centres
        float mx = -INFINITY;
        for (int k = lane; k < d.K; k += 32) mx = fmaxf(mx, lv[k]);
        mx = warp_max(mx);
        float se = 0.f, sc = 0.f;
        const int half = (d.K + 1) / 2;
        for (int k = lane; k < d.K; k += 32) {
            const float ex = expf(lv[k] - mx);
            const float e0 = k < half ? a.reward_lo + a.hl_step * (float)k : a.reward_hi - a.hl_step * (float)(d.K - k);
            const float e1 = k + 1 < half ? a.reward_lo + a.hl_step * (float)(k + 1) : a.reward_hi - a.hl_step * (float)(d.K - k - 1);
            se += ex; sc = fmaf(ex, (e0 + e1) * 0.5f, sc);
            if (a.io.value_logits) a.io.value_logits[(int64_t)env * d.K + k] = lv[k];
        }
        se = warp_sum(se); sc = warp_sum(sc);
        if (lane == 0 && a.io.value) a.io.value[env] = sc / se;
        // actor
        if (d.cont) {                                          // Continuous (:224-277): lane i <-> action dim i
            if (a.io.raw_actions)
                for (int i = lane; i < 2 * d.A; i += 32) a.io.raw_actions[(int64_t)env * 2 * d.A + i] = la[i];
            if (lane < d.A) {
                const float mean = la[2 * lane], lv = la[2 * lane + 1];
                const float var = expf(tanhf(lv / 3.f) * 3.f);                 // softclamp :124,:237
                const float sd = sqrtf(fmaxf(var, 1e-5f));
                float x;
                bool have = true;
                if (a.io.forced_action_f) x = __ldg(a.io.forced_action_f + (int64_t)env * d.A + lane);
                else if (a.io.normal) {
                    x = fmaf(sd, __ldg(a.io.normal + (int64_t)env * d.A + lane), mean);
                    if (d.squash) x = tanhf(x);
                } else { x = 0.f; have = false; }
                if (have) {
                    const float df = x - mean;
                    float lp = -(df * df) / (2.f * sd * sd) - logf(sd) - 0.91893853320467274178f;   // log sqrt(2 pi)
                    if (d.squash) lp -= logf(fmaxf(1.f - x * x, 1e-20f));    // :271 (evaluated at the squashed value)
                    if (a.io.action_f) a.io.action_f[(int64_t)env * d.A + lane] = x;
                    if (a.io.log_prob) a.io.log_prob[(int64_t)env * d.A + lane] = lp;
                }
            }
            continue;
        }
        const float lg = lane < d.A ? la[lane] : -INFINITY;
        if (lane < d.A && a.io.raw_actions) a.io.raw_actions[(int64_t)env * d.A + lane] = lg;
        const float amx = warp_max(lg);
        const float qx = lane < d.A ? expf(lg - amx) : 0.f;
        const float smx = qx / warp_sum(qx);
        const float p = smx / warp_sum(smx);                   // Categorical(probs) renormalises
        int act = -1;
        if (a.io.forced_action) act = __ldg(a.io.forced_action + env);
        else if (a.io.uniform) {
            const float u = __ldg(a.io.uniform + env);
            float run = 0.f;
            act = d.A - 1;
            for (int i = 0; i < d.A; ++i) {                    // sequential fp32 running sum
                run += __shfl_sync(0xffffffffu, p, i);
                if (run > u) { act = i; break; }
            }
        }
        if (act >= 0) {
            const float feps = 1.1920928955078125e-07f;
            const float pa = __shfl_sync(0xffffffffu, p, act & 31);
            if (lane == 0) {
                if (a.io.action) a.io.action[env] = act;
                if (a.io.log_prob) a.io.log_prob[env] = logf(fminf(fmaxf(pa, feps), 1.f - feps));
            }
        }
    }
}

}  // namespace xtrl

extern "C" int xtrl_wm_step(const xtrl_model_desc *desc, const float *weights, void *kv_cache,
                            const xtrl_step_io *io, int32_t num_envs, void *stream) {
    using namespace xtrl;
    int rc = check_desc(desc);
    if (rc) return rc;
    XTRL_REQUIRE(num_envs >= 0, XTRL_E_BADARG, "xtrl_wm_step: num_envs=%d", num_envs);
    if (num_envs == 0) return XTRL_OK;                      // nothing to do (and the caller's buffers may be empty)
    XTRL_REQUIRE(io && weights && kv_cache, XTRL_E_BADARG, "xtrl_wm_step: null pointer");
    XTRL_REQUIRE(io->state && io->prev_reward && io->rs_mean && io->rs_var, XTRL_E_BADARG, "xtrl_wm_step: null required input");
    XTRL_REQUIRE(desc->continuous_actions || io->prev_action != nullptr, XTRL_E_BADARG, "xtrl_wm_step: prev_action missing");
    XTRL_REQUIRE(io->pos || (io->pos_uniform >= 0 && io->pos_uniform < desc->max_timesteps), XTRL_E_BADARG,
                 "xtrl_wm_step: pos_uniform=%d outside [0, max_timesteps=%d)", io->pos_uniform, desc->max_timesteps);
    XTRL_REQUIRE(desc->dim_gene == 0 || (io->genes && io->gene_ids), XTRL_E_BADARG,
                 "xtrl_wm_step: evolutionary model needs genes and gene_ids");
    StepArgs a;
    make_dims(desc, &a.d);
    rc = xtrl_wm_weight_layout(desc, &a.w);
    if (rc) return rc;
    a.io = *io; a.weights = weights; a.kv = kv_cache; a.B = num_envs;
    a.reward_lo = desc->reward_min; a.reward_hi = desc->reward_max;
    a.hl_step = (desc->reward_max - desc->reward_min) / (float)desc->num_bins;
    const Smem so = smem_layout(a.d);
    const size_t smem = (size_t)so.total * sizeof(float);
    XTRL_REQUIRE(smem <= 200 * 1024, XTRL_E_UNSUPPORTED, "xtrl_wm_step: configuration needs %zu B shared memory per CTA (max 200 KiB)", smem);
    a.tcw = (const unsigned char *)io->tc_weights;
    if (a.d.tc) {
        XTRL_REQUIRE(io->tc_weights, XTRL_E_BADARG, "xtrl_wm_step: dense_bf16 needs io->tc_weights (xtrl_wm_pack_tc)");
        XTRL_REQUIRE(((uintptr_t)io->tc_weights & 15) == 0, XTRL_E_BADARG, "xtrl_wm_step: tc_weights must be 16-byte aligned");
        XTRL_REQUIRE(128 * 4 * a.d.D * 2 <= kWbufBytes && 128 * a.d.I * 2 <= kWbufBytes && 128 * a.d.HIN * 2 <= kWbufBytes,
                     XTRL_E_UNSUPPORTED, "xtrl_wm_step: dense_bf16 weight tile exceeds the %d-byte stage (dim=%d heads=%d)", kWbufBytes, a.d.D, a.d.H);
        a.tc = tc_layout(a.d);
    }
    // spread the envs so that every CTA slot of the chip (2 per SM while shared memory allows) gets the same
    // share: 4096 envs -> 14 per CTA, 293 CTAs on 296 slots instead of 256 CTAs (108 SMs with two, 40 with one)
    const int slots = (smem <= 113 * 1024 ? 2 : 1) * num_sms();
    int epc = (num_envs + slots - 1) / slots;
    if (io->envs_per_cta > 0) epc = io->envs_per_cta;                 // caller overlaps several calls on different streams
    epc = epc < 1 ? 1 : (epc > kTE ? kTE : epc);
    a.epc = epc;
    const int blocks = (num_envs + epc - 1) / epc;
    cudaStream_t s = (cudaStream_t)stream;
    auto launch = [&](auto kern) -> int {
        XTRL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<blocks, kThreads, smem, s>>>(a);
        XTRL_LAUNCH_CHECK("wm_step_kernel");
        return XTRL_OK;
    };
    if (a.d.tc) {
#if XTRL_TE == 16
        return desc->kv_bf16 ? launch(wm_step_kernel<__nv_bfloat16, true>) : launch(wm_step_kernel<float, true>);
#else
        return fail(XTRL_E_UNSUPPORTED, "xtrl_wm_step: dense_bf16 needs the 16-env CTA tile this library was not built with");
#endif
    }
    return desc->kv_bf16 ? launch(wm_step_kernel<__nv_bfloat16, false>) : launch(wm_step_kernel<float, false>);
}

extern "C" int64_t xtrl_wm_tc_weights_bytes(const xtrl_model_desc *desc) {
    using namespace xtrl;
    if (check_desc(desc)) return -1;
    StepDims d;
    make_dims(desc, &d);
    return tc_layout(d).total;
}

extern "C" int xtrl_wm_pack_tc(const xtrl_model_desc *desc, const float *weights, void *tc_weights, void *stream) {
    using namespace xtrl;
    int rc = check_desc(desc);
    if (rc) return rc;
    XTRL_REQUIRE(weights && tc_weights, XTRL_E_BADARG, "xtrl_wm_pack_tc: null pointer");
    StepDims d;
    make_dims(desc, &d);
    xtrl_weight_layout w;
    rc = xtrl_wm_weight_layout(desc, &w);
    if (rc) return rc;
    const TcLayout t = tc_layout(d);
    cudaStream_t s = (cudaStream_t)stream;
    unsigned char *dst = (unsigned char *)tc_weights;
    auto pack = [&](int64_t src_off, int ld, int coff, int nvalid, int K, int tiles, int64_t dst_off) -> int {
        const int64_t total = (int64_t)tiles * 128 * K;
        pack_tc_kernel<<<(int)((total + 255) / 256), 256, 0, s>>>(weights + src_off, ld, coff, nvalid, K, tiles,
                                                                 (__nv_bfloat16 *)(dst + dst_off));
        XTRL_LAUNCH_CHECK("pack_tc_kernel");
        return XTRL_OK;
    };
    for (int l = 0; l < d.L; ++l) {
        if ((rc = pack(w.layer[l].qkv_w, d.NQ, 0, d.NQ, d.D, t.tiles_q, t.qkv[l]))) return rc;
        if ((rc = pack(w.layer[l].out_w, d.D, 0, d.D, d.I, 1, t.out[l]))) return rc;
        if ((rc = pack(w.layer[l].ff1_w, 4 * d.D, 0, 4 * d.D, d.D, t.tiles_f, t.ff1[l]))) return rc;
        if ((rc = pack(w.layer[l].ff2_w, d.D, 0, d.D, 4 * d.D, 1, t.ff2[l]))) return rc;
    }
    if ((rc = pack(w.head1_w, 4 * d.D, 0, 2 * d.D, d.HIN, 1, t.head_a))) return rc;
    if ((rc = pack(w.head1_w, 4 * d.D, 2 * d.D, 2 * d.D, d.HIN, 1, t.head_c))) return rc;
    return XTRL_OK;
}
