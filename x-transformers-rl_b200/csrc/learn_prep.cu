// learn_prep.cu -- the prologue of one PPO minibatch and its backward, one launch each.
//
// Replaces, in Agent.learn (x_transformers_rl/x_transformers_rl.py):
//   the DataLoader's gather of the minibatch episodes (:876-901), the prev-action / prev-reward shifts (:911-918), RSNorm.eval on
//   the packed (state || prev reward) rows (:920-926, :588-594), the masked batch mean the RSNorm copy is updated with (:1005,
//   :598-601), and the input side of WorldModelActorCritic.forward (:492-503): project_in(state) + action_embeds(prev action)
//   + prev reward * reward_embed * flag, to_state_embed(state), and action_embeds(action) for the next-state head (:516-529)
// -- ~40 small ATen launches per minibatch (9 index_select, pads, cat, normalise, masks, three skinny GEMMs, embedding gathers)
// and, in the backward, three skinny weight-gradient passes.
//
// forward : gathers the rows of the b chosen episodes from the episode-major trajectory tensors and writes, per token,
//           packed[S+1] (normalised), x0[48] (the first layer's residual input), se[48] = to_state_embed(state), ae_next[48] =
//           action_embeds(action), the gathered per-row fields the loss kernel reads, and the masked column sums of packed.
// backward: one pass over (dx0, d_se, d_ae_next, packed, actions) accumulating every skinny parameter gradient
//           (project_in, to_state_embed weight + bias, action embedding table or Linear, reward_embed) in registers; one
//           atomicAdd per value per CTA at the end.
// HBM-bound: forward writes 3 x 192 + 4(S+1) + ~21 B per token, backward reads 3 x 192 + 4(S+1) + 4 B per token.
// Thread (feature j of 48, token lane of 4): the 48 threads of a token read / write one contiguous 192-byte row.
#include "common.cuh"

namespace xtrl {
namespace prep {

constexpr int D = 48, TL = 4, NT = D * TL, TOK = 64;          // features, token lanes, threads, tokens per tile
constexpr int SMAX = 16, AMAX = 16;

struct Args {
    xtrl_prep_io io;
    int b, T, S, A, cont;
};

// SM / AM: compile-time bounds of state_dim / num_actions (the loops over them are fully unrolled and predicated)
template <int SM, int AM>
__global__ void __launch_bounds__(NT) prep_fwd_kernel(const Args a) {
    __shared__ float s_packed[TOK][SM + 1];
    __shared__ float s_act[TOK][AM];                          // continuous: prev action / action vectors
    __shared__ float s_nact[TOK][AM];
    __shared__ int s_pa[TOK], s_na[TOK];                      // discrete: prev action (-1 none), action
    __shared__ float s_valid[TOK];                            // 1 where the row lies inside its episode's length
    __shared__ float s_sum[SM + 2];
    const xtrl_prep_io &io = a.io;
    const int tid = threadIdx.x, j = tid % D, tl = tid / D;
    const int S = a.S, S1 = S + 1, A = a.A, T = a.T;
    const int N = a.b * T;                                    // < 2^31 (checked by the launcher): 32-bit row arithmetic throughout
    // this thread's rows of the skinny weights
    float w_in[SM], w_se[SM], w_ac[AM];
#pragma unroll
    for (int s = 0; s < SM; ++s) { w_in[s] = s < S ? __ldg(io.w_in + j * S + s) : 0.f; w_se[s] = s < S ? __ldg(io.w_se + j * S + s) : 0.f; }
#pragma unroll
    for (int i = 0; i < AM; ++i) w_ac[i] = (a.cont && i < A) ? __ldg(io.w_act + j * A + i) : 0.f;
    const float b_se = __ldg(io.b_se + j), b_ac = a.cont ? __ldg(io.b_act + j) : 0.f;
    const float re = __ldg(io.reward_embed + j) * __ldg(io.reward_flag);
    if (tid < SM + 2) s_sum[tid] = 0.f;
    float my_sum = 0.f;                                       // thread tid < S1: column tid of packed; tid == S1: the row count

    const int ntiles = (N + TOK - 1) / TOK;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int r0 = tile * TOK;
        __syncthreads();
        // phase 1: gather + normalise the tile's rows (one (token, column) item per thread and pass)
        for (int item = tid; item < TOK * (S1 + 1); item += NT) {
            const int tk = (unsigned)item / (unsigned)(S1 + 1), c = item - tk * (S1 + 1);
            const int r = r0 + tk;
            if (r >= N) { if (c == S1) s_valid[tk] = 0.f; continue; }
            const int bi = (unsigned)r / (unsigned)T, t = r - bi * T;
            const int64_t e = __ldg(io.idx + bi), src = e * T + t;
            if (c < S1) {
                const float raw = c < S ? __ldg(io.states + src * S + c) : (t > 0 ? __ldg(io.rewards + src - 1) : 0.f);
                const float v = (raw - __ldg(io.rs_mean + c)) / fmaxf(sqrtf(__ldg(io.rs_var + c)), 1e-5f);
                s_packed[tk][c] = v;
                io.packed[(int64_t)r * S1 + c] = v;
            } else {                                          // the row's scalar fields
                io.mb_returns[r] = __ldg(io.returns + src);
                io.mb_old_values[r] = __ldg(io.old_values + src);
                io.mb_dones[r] = __ldg(io.dones + src);
                if (!a.cont) {
                    const int32_t *acts = static_cast<const int32_t *>(io.actions);
                    const int na = __ldg(acts + src);
                    s_na[tk] = na; s_pa[tk] = t > 0 ? __ldg(acts + src - 1) : -1;
                    static_cast<int32_t *>(io.mb_actions)[r] = na;
                    io.mb_old_log_probs[r] = __ldg(io.old_log_probs + src);
                } else {
                    const float *acts = static_cast<const float *>(io.actions);
                    for (int i = 0; i < A; ++i) {
                        const float na = __ldg(acts + src * A + i);
                        s_nact[tk][i] = na; s_act[tk][i] = t > 0 ? __ldg(acts + (src - 1) * A + i) : 0.f;
                        static_cast<float *>(io.mb_actions)[(int64_t)r * A + i] = na;
                        io.mb_old_log_probs[(int64_t)r * A + i] = __ldg(io.old_log_probs + src * A + i);
                    }
                }
                const int len = __ldg(io.lens + e);
                s_valid[tk] = t < len ? 1.f : 0.f;
                if (t == 0) { io.mb_lens[bi] = len; if (io.gene_ids) io.mb_gene_ids[bi] = __ldg(io.gene_ids + e); }
            }
        }
        __syncthreads();
        // masked column sums of packed for the RSNorm batch mean (:1005, :601): thread c sums column c over the tile
        if (tid <= S1) {
            const int lim = N - r0 < TOK ? N - r0 : TOK;
            for (int tk = 0; tk < lim; ++tk) my_sum += s_valid[tk] * (tid < S1 ? s_packed[tk][tid] : 1.f);
        }
        // phase 2: the three 48-wide rows of every token
#pragma unroll 4
        for (int tk = tl; tk < TOK; tk += TL) {
            const int r = r0 + tk;
            if (r >= N) continue;
            float x = 0.f, se = b_se, ae_p = 0.f, ae_n = 0.f;
#pragma unroll
            for (int s = 0; s < SM; ++s) if (s < S) { const float v = s_packed[tk][s]; x = fmaf(w_in[s], v, x); se = fmaf(w_se[s], v, se); }
            if (a.cont) {
                ae_p = b_ac; ae_n = b_ac;
#pragma unroll
                for (int i = 0; i < AM; ++i) if (i < A) { ae_p = fmaf(w_ac[i], s_act[tk][i], ae_p); ae_n = fmaf(w_ac[i], s_nact[tk][i], ae_n); }
            } else {
                const int pa = s_pa[tk], na = s_na[tk];
                ae_p = pa >= 0 ? __ldg(io.w_act + pa * D + j) : 0.f;           // SafeEmbedding (:181-195): id -1 -> zeros
                ae_n = na >= 0 ? __ldg(io.w_act + na * D + j) : 0.f;
            }
            io.x0[(int64_t)r * D + j] = x + ae_p + s_packed[tk][S] * re;
            io.se[(int64_t)r * D + j] = se;
            io.ae_next[(int64_t)r * D + j] = ae_n;
        }
    }
    if (tid <= S1 && my_sum != 0.f) atomicAdd(io.rs_sum + tid, my_sum);
}

struct BwdArgs {
    xtrl_prep_bwd_io io;
    int64_t N;
    int S, A, cont;
};

// backward: 48 features x TLB token lanes per CTA, one CTA per SM (few CTAs = few same-address atomics at the end)
constexpr int TLB = 16, NTB = D * TLB;

template <int SM, int AM>
__global__ void __launch_bounds__(NTB) prep_bwd_kernel(const BwdArgs a) {
    const xtrl_prep_bwd_io &io = a.io;
    const int tid = threadIdx.x, j = tid % D, tl = tid / D;
    const int S = a.S, S1 = S + 1, A = a.A;
    const float flag = __ldg(io.reward_flag);
    float g_in[SM + 1], g_se[SM + 1], g_ac[AM], g_bac = 0.f;      // g_in[S]: reward_embed; g_se[S]: to_state_embed bias
#pragma unroll
    for (int s = 0; s <= SM; ++s) { g_in[s] = 0.f; g_se[s] = 0.f; }
#pragma unroll
    for (int i = 0; i < AM; ++i) g_ac[i] = 0.f;
    const int Nr = (int)a.N;                                  // < 2^31 (checked by the launcher)
    int per = (Nr + gridDim.x - 1) / gridDim.x;
    per = (per + TLB - 1) / TLB * TLB;
    const int lo = blockIdx.x * per, hi = lo + per < Nr ? lo + per : Nr;
    // Rows are taken U at a time: all of their loads are issued before the first accumulate, so U x 14 independent loads are in
    // flight per thread instead of one row's (the loop is latency-bound otherwise).
    constexpr int U = 2;
    struct Row { float dx, dse, dae, pk[SM + 1]; float av[AM]; int na, pa; bool live; };
    auto load_row = [&](int64_t r, int t, Row &w) {
        w.live = r < hi;
        const int64_t rr = w.live ? r : lo;
        w.dx = w.live ? __ldg(io.dx0 + rr * D + j) : 0.f;
        w.dse = (w.live && io.d_se) ? __ldg(io.d_se + rr * D + j) : 0.f;
        w.dae = (w.live && io.d_ae_next) ? __ldg(io.d_ae_next + rr * D + j) : 0.f;
        const float *pk = io.packed + rr * S1;
#pragma unroll
        for (int s = 0; s < SM; ++s) w.pk[s] = s < S ? __ldg(pk + s) : 0.f;
        w.pk[SM] = __ldg(pk + S) * flag;
        if (a.cont) {
            const float *an = static_cast<const float *>(io.mb_actions) + rr * A;
#pragma unroll
            for (int i = 0; i < AM; ++i) w.av[i] = i < A ? __ldg(an + i) : 0.f;
            w.na = t;                                                          // continuous: na carries t (first row of an episode has no prev action)
            w.pa = 0;
        } else {
            const int32_t *acts = static_cast<const int32_t *>(io.mb_actions);
            w.na = __ldg(acts + rr);
            w.pa = t == 0 ? -1 : __ldg(acts + rr - 1);
        }
    };
    auto accumulate = [&](const Row &w, int64_t r) {
        if (!w.live) return;
#pragma unroll
        for (int s = 0; s < SM; ++s) { g_in[s] = fmaf(w.dx, w.pk[s], g_in[s]); g_se[s] = fmaf(w.dse, w.pk[s], g_se[s]); }
        g_in[SM] = fmaf(w.dx, w.pk[SM], g_in[SM]);
        g_se[SM] += w.dse;
        if (a.cont) {
            const float *an = static_cast<const float *>(io.mb_actions) + r * A;
#pragma unroll
            for (int i = 0; i < AM; ++i) if (i < A) g_ac[i] += w.dae * w.av[i] + (w.na == 0 ? 0.f : w.dx * __ldg(an + i - A));
            g_bac += w.dae + w.dx;
        } else {
#pragma unroll
            for (int i = 0; i < AM; ++i) g_ac[i] += (w.na == i ? w.dae : 0.f) + (w.pa == i ? w.dx : 0.f);
        }
    };
    int t = lo + tl < Nr ? (int)((unsigned)(lo + tl) % (unsigned)io.T) : 0;      // timestep of row r, advanced without divisions
    for (int64_t r = lo + tl; r < hi; r += U * TLB) {
        Row w[U];
        int tt = t;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            while (tt >= io.T) tt -= io.T;
            load_row(r + u * TLB, tt, w[u]);
            tt += TLB;
        }
        t = tt;
#pragma unroll
        for (int u = 0; u < U; ++u) accumulate(w[u], r + u * TLB);
    }
    // the TLB token lanes of a feature meet in shared memory (atomics), then one global atomicAdd per value and CTA
    constexpr int NV = 2 * (SM + 1) + AM + 1;
    __shared__ float red[D][NV];
    for (int i = tid; i < D * NV; i += NTB) (&red[0][0])[i] = 0.f;
    __syncthreads();
    float *mine = red[j];
#pragma unroll
    for (int s = 0; s <= SM; ++s) { atomicAdd(mine + s, g_in[s]); atomicAdd(mine + SM + 1 + s, g_se[s]); }
#pragma unroll
    for (int i = 0; i < AM; ++i) atomicAdd(mine + 2 * (SM + 1) + i, g_ac[i]);
    atomicAdd(mine + 2 * (SM + 1) + AM, g_bac);
    __syncthreads();
    if (tl == 0) {
        auto tot = [&](int k) { return red[j][k]; };
        for (int s = 0; s < S; ++s) {
            atomicAdd(io.d_w_in + j * S + s, tot(s));
            if (io.d_w_se) atomicAdd(io.d_w_se + j * S + s, tot(SM + 1 + s));
        }
        atomicAdd(io.d_reward_embed + j, tot(SM));
        if (io.d_b_se) atomicAdd(io.d_b_se + j, tot(2 * SM + 1));
        for (int i = 0; i < A; ++i) {
            const float v = tot(2 * (SM + 1) + i);
            if (a.cont) atomicAdd(io.d_w_act + j * A + i, v); else atomicAdd(io.d_w_act + i * D + j, v);
        }
        if (a.cont) atomicAdd(io.d_b_act + j, tot(2 * (SM + 1) + AM));
    }
}

}  // namespace prep
}  // namespace xtrl

extern "C" int xtrl_learn_prep_fwd(const xtrl_prep_io *io, int32_t b, int32_t T, int32_t state_dim, int32_t num_actions,
                                   int32_t continuous, void *stream) {
    using namespace xtrl; using namespace xtrl::prep;
    XTRL_REQUIRE(io && b >= 0 && T >= 1 && (int64_t)b * T < (1LL << 31), XTRL_E_BADARG, "xtrl_learn_prep_fwd: bad arguments");
    XTRL_REQUIRE(state_dim >= 1 && state_dim <= SMAX && num_actions >= 1 && num_actions <= AMAX, XTRL_E_UNSUPPORTED,
                 "xtrl_learn_prep_fwd: built for state_dim <= %d, num_actions <= %d (got %d, %d)", SMAX, AMAX, state_dim, num_actions);
    if (b == 0) return XTRL_OK;
    XTRL_REQUIRE(io->states && io->actions && io->rewards && io->old_log_probs && io->returns && io->old_values && io->dones && io->lens &&
                 io->idx && io->rs_mean && io->rs_var && io->w_in && io->w_se && io->b_se && io->w_act && io->reward_embed && io->reward_flag &&
                 io->x0 && io->se && io->ae_next && io->packed && io->mb_actions && io->mb_old_log_probs && io->mb_returns &&
                 io->mb_old_values && io->mb_dones && io->mb_lens && io->rs_sum && (!continuous || io->b_act) && (!io->gene_ids || io->mb_gene_ids),
                 XTRL_E_BADARG, "xtrl_learn_prep_fwd: null pointer");
    Args a;
    a.io = *io; a.b = b; a.T = T; a.S = state_dim; a.A = num_actions; a.cont = continuous != 0;
    const int64_t ntiles = ((int64_t)b * T + TOK - 1) / TOK, cap = 6LL * num_sms();
    const int grid = (int)(ntiles < cap ? ntiles : cap);
    if (state_dim <= 8 && num_actions <= 4) prep_fwd_kernel<8, 4><<<grid, NT, 0, (cudaStream_t)stream>>>(a);
    else prep_fwd_kernel<SMAX, AMAX><<<grid, NT, 0, (cudaStream_t)stream>>>(a);
    XTRL_LAUNCH_CHECK("prep_fwd_kernel");
    return XTRL_OK;
}

extern "C" int xtrl_learn_prep_bwd(const xtrl_prep_bwd_io *io, int32_t b, int32_t T, int32_t state_dim, int32_t num_actions,
                                   int32_t continuous, void *stream) {
    using namespace xtrl; using namespace xtrl::prep;
    XTRL_REQUIRE(io && b >= 0 && T >= 1 && (int64_t)b * T < (1LL << 31), XTRL_E_BADARG, "xtrl_learn_prep_bwd: bad arguments");
    XTRL_REQUIRE(state_dim >= 1 && state_dim <= SMAX && num_actions >= 1 && num_actions <= AMAX, XTRL_E_UNSUPPORTED,
                 "xtrl_learn_prep_bwd: built for state_dim <= %d, num_actions <= %d (got %d, %d)", SMAX, AMAX, state_dim, num_actions);
    if (b == 0) return XTRL_OK;
    XTRL_REQUIRE(io->dx0 && io->packed && io->mb_actions && io->reward_flag && io->d_w_in && io->d_w_act && io->d_reward_embed &&
                 (!continuous || io->d_b_act) && io->T == T, XTRL_E_BADARG, "xtrl_learn_prep_bwd: null pointer / inconsistent T");
    BwdArgs a;
    a.io = *io; a.N = (int64_t)b * T; a.S = state_dim; a.A = num_actions; a.cont = continuous != 0;
    const int64_t want = (a.N + 2 * TLB - 1) / (2 * TLB), cap = num_sms();
    const int grid = (int)(want < cap ? want : cap);
    if (state_dim <= 8 && num_actions <= 4) prep_bwd_kernel<8, 4><<<grid, NTB, 0, (cudaStream_t)stream>>>(a);
    else prep_bwd_kernel<SMAX, AMAX><<<grid, NTB, 0, (cudaStream_t)stream>>>(a);
    XTRL_LAUNCH_CHECK("prep_bwd_kernel");
    return XTRL_OK;
}
