/* xtrl.h -- C ABI of libxtrl.so: the B200 (sm_100a) rollout-and-learn hot path that stands in for
 * the arithmetic of tinycrops/x-transformers-rl (citations: paths relative to the reference tree).
 *
 * Contract
 *   - extern "C", plain pointers and sizes only; no torch / C++ types cross this boundary.
 *   - every pointer is a DEVICE pointer owned by the caller unless the name says `host`; the
 *     library never allocates, frees or retains a pointer past the call.
 *   - every entry point is asynchronous on the caller's stream (`stream` is a cudaStream_t passed
 *     as void*), never synchronises, and keeps no mutable global state beyond one-time
 *     cudaFuncSetAttribute calls.
 *   - return value: 0 = ok, negative = error (XTRL_E_*); text via xtrl_last_error() (thread-local).
 *   - an unsupported configuration is an ERROR, never a fallback: there is no CPU path in here.
 */
#ifndef XTRL_H
#define XTRL_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XTRL_ABI_VERSION 9

#define XTRL_OK            0
#define XTRL_E_BADARG     -1   /* null pointer, negative size, inconsistent shapes            */
#define XTRL_E_UNSUPPORTED -2  /* configuration outside what the kernels are built for        */
#define XTRL_E_CUDA       -3   /* a CUDA runtime call failed (message holds cudaGetErrorString) */

int         xtrl_abi_version(void);
const char *xtrl_last_error(void);
/* compute capability of the current device as major*10+minor, or a negative error.           */
int         xtrl_device_sm(void);
/* number of SMs of the current device (grids are sized from this, never from a constant).      */
int         xtrl_device_num_sms(void);

/* ------------------------------------------------------------------------------------------
 * (1) world-model / actor-critic rollout step
 *     replaces Learner.forward's state_to_pred_action_and_value closure + dist.sample/log_prob
 *     (x_transformers_rl/x_transformers_rl.py:1250-1289): RSNorm.eval (:588-594) ->
 *     WorldModelActorCritic.forward (:479-559) with the x-transformers Decoder and its KV cache
 *     (call site :721-734, :505-512) -> Discrete (:197-222) -> HL-Gauss decode (:843).
 *     One call advances EVERY env of the batch by one timestep.
 * ---------------------------------------------------------------------------------------- */

typedef struct xtrl_model_desc {
    int32_t state_dim;        /* S                                                            */
    int32_t num_actions;      /* A (discrete)                                                 */
    int32_t dim;              /* D, Agent hidden_dim (:666); kernels are built for 48         */
    int32_t heads;            /* H                                                            */
    int32_t dim_head;         /* kernels are built for 16                                     */
    int32_t depth;            /* L                                                            */
    int32_t num_bins;         /* K, critic_pred_num_bins (:665)                               */
    int32_t dim_gene;         /* G, 0 = not evolutionary                                      */
    int32_t gate_values;      /* world_model['attn_gate_values']                              */
    int32_t value_residual;   /* world_model['add_value_residual']                            */
    int32_t learned_mix;      /* world_model['learned_value_residual_mix']                    */
    int32_t max_timesteps;    /* Tmax: KV-cache capacity per env                              */
    int32_t rotary_true_pos;  /* 0 = reference behaviour (all rollout positions 0) */
    int32_t kv_bf16;          /* 0 = fp32 KV cache, 1 = bf16 KV cache                         */
    float   reward_min, reward_max;   /* HL-Gauss support (reward_range)                      */
    int32_t dense_bf16;       /* 0 = dense layers in fp32 FFMA (1e-5 parity with the fp32 reference);
                                 1 = QKV/out/FFN/head GEMMs on tcgen05 tensor cores, bf16 operands, fp32
                                 accumulation in TMEM (the 1e-2 "bf16" tolerance class)              */
    int32_t continuous_actions; /* 0 = Discrete (:197-222); 1 = Continuous (:224-277): the actor head emits
                                 (mean, log-variance) pairs, actions are float vectors of num_actions    */
    int32_t squash_continuous;  /* tanh-squash sampled actions (Learner default :1077)                     */
} xtrl_model_desc;

/* offsets (in floats) of every tensor inside the packed fp32 weight blob the step kernel reads.
 * The host packs a reference-style state_dict into this layout (weights are stored K-major,
 * i.e. transposed nn.Linear.weight, so a warp reads consecutive output features).             */
#define XTRL_MAX_DEPTH 16
typedef struct xtrl_weight_layout {
    int64_t total;                 /* blob length in floats                                   */
    int32_t s_pad;                 /* S rounded up to a multiple of 4                         */
    int32_t n_qkv;                 /* fused q|k|v|gate|mix output width (padded to even)      */
    int32_t head_in;               /* 2D or 3D (evolutionary)                                 */
    int64_t in_proj;               /* [s_pad][2D]: project_in (no bias) | to_state_embed      */
    int64_t state_embed_bias;      /* [D]                                                     */
    int64_t action_embed;          /* [A][D]                                                  */
    int64_t reward_embed;          /* [D]                                                     */
    int64_t latent_w;              /* [G][D] (evolutionary)                                   */
    int64_t latent_b;              /* [D]                                                     */
    int64_t final_gamma;           /* [D]                                                     */
    int64_t head1_w;               /* [head_in][4D]: action_head.0 | critic_head.0            */
    int64_t head1_b;               /* [4D]                                                    */
    int64_t actor2_w;              /* [2D][A]                                                 */
    int64_t actor2_b;              /* [A]                                                     */
    int64_t critic2_w;             /* [2D][K]                                                 */
    int64_t critic2_b;             /* [K]                                                     */
    struct {
        int64_t ln1_gamma;         /* [D]                                                     */
        int64_t qkv_w;             /* [D][n_qkv]: to_q | to_k | to_v | to_v_gate | mix        */
        int64_t qkv_b;             /* [n_qkv] (zeros for q,k,v)                               */
        int64_t out_w;             /* [I][D]                                                  */
        int64_t ln2_gamma;         /* [D]                                                     */
        int64_t ff1_w;             /* [D][4D]                                                 */
        int64_t ff1_b;             /* [4D]                                                    */
        int64_t ff2_w;             /* [4D][D]                                                 */
        int64_t ff2_b;             /* [D]                                                     */
    } layer[XTRL_MAX_DEPTH];
    int64_t action_embed_bias;     /* [D] continuous only: action_embeds is nn.Linear(A, D) (:309)           */
} xtrl_weight_layout;

/* fills *out; HOST pointers.  Pure host arithmetic (usable without a GPU).                    */
int xtrl_wm_weight_layout(const xtrl_model_desc *desc, xtrl_weight_layout *out);

/* tensor-core path (desc->dense_bf16): size of, and packing into, the bf16 operand images the UMMA
 * descriptors read (128-row tiles, canonical K-major no-swizzle core-matrix layout).  xtrl_wm_pack_tc
 * converts the fp32 blob above on the device; pass the result as xtrl_step_io.tc_weights.            */
int64_t xtrl_wm_tc_weights_bytes(const xtrl_model_desc *desc);
int     xtrl_wm_pack_tc(const xtrl_model_desc *desc, const float *weights, void *tc_weights, void *stream);

/* bytes of KV cache the caller must allocate for `num_envs` envs:
 *   [2 (k,v)][L][num_envs][H][Tmax][dim_head] in fp32 or bf16.                                */
int64_t xtrl_wm_kv_cache_bytes(const xtrl_model_desc *desc, int64_t num_envs);

typedef struct xtrl_step_io {
    /* inputs */
    const float   *state;         /* [B][S] raw (un-normalised) observation                   */
    const int32_t *prev_action;   /* [B], -1 = none (:1244)                                   */
    const float   *prev_reward;   /* [B] raw previous reward (0 at t=0, :1246)                */
    const int32_t *pos;           /* [B] number of tokens already cached for this env (its t);
                                     NULL = every env is at `pos_uniform` (lock-step rollout)  */
    const float   *rs_mean;       /* [S+1] RSNorm running_mean                                */
    const float   *rs_var;        /* [S+1] RSNorm running_variance                            */
    const float   *genes;         /* [num_genes][G] un-normalised gene matrix, or NULL        */
    const int32_t *gene_ids;      /* [B] or NULL                                              */
    const float   *uniform;       /* [B] u in [0,1): inverse-CDF sampling; NULL =
                                     do not sample (action_out untouched unless forced)      */
    const int32_t *forced_action; /* [B] or NULL: evaluate log-prob of this action instead    */
    float          reward_flag;   /* the per-forward reward-conditioning flag (:501-503)      */
    int32_t        pos_uniform;   /* used when pos == NULL                                    */
    const void    *tc_weights;    /* xtrl_wm_pack_tc output; required iff desc->dense_bf16     */
    /* outputs (any may be NULL) */
    int32_t *action;              /* [B]                                                      */
    float   *log_prob;            /* [B] log(clamp(p[a], eps, 1-eps))                         */
    float   *value;               /* [B] HL-Gauss expectation of the critic logits            */
    float   *raw_actions;         /* [B][A] actor logits                                      */
    float   *value_logits;        /* [B][K] critic logits                                     */
    /* continuous actions (desc->continuous_actions): float action vectors replace the int ids above;
       log_prob is then [B][A] (per action dim) and raw_actions [B][2A] (mean, log-variance interleaved) */
    const float   *prev_action_f;   /* [B][A] previous action, zeros at t=0 (:1242)             */
    const float   *normal;          /* [B][A] standard-normal draws for Normal.sample */
    const float   *forced_action_f; /* [B][A] or NULL: evaluate log-prob of these actions instead */
    float         *action_f;        /* [B][A]                                                    */
    int32_t        envs_per_cta;    /* 0 = size the grid to fill the chip with THIS call's envs; >0 = envs per CTA
                                       (<= 16), for callers that run several calls concurrently on different streams */
} xtrl_step_io;

int xtrl_wm_step(const xtrl_model_desc *desc, const float *weights, void *kv_cache,
                 const xtrl_step_io *io, int32_t num_envs, void *stream);

/* ------------------------------------------------------------------------------------------
 * (2) GAE / returns -- replaces calc_gae (x_transformers_rl.py:616-640), i.e. the seam the
 *     reference itself exposes as AssocScan(reverse=True, use_accelerated=...) (:634-636).
 *     returns[e,t] = gae[e,t] + values[e,t],  gae_t = delta_t + gamma*lam*m_t*gae_{t+1},
 *     delta_t = r_t + gamma*v_{t+1}*m_t - v_t,  v_T = 0,  m = !is_boundary.
 *     Element (e,t) of every array lives at  e*stride_e + t*stride_t  (in elements); either
 *     episode-major (stride_t == 1, the reference's (E,T)) or time-major (stride_e == 1, the
 *     rollout buffer's (T,B)).
 * ---------------------------------------------------------------------------------------- */
int xtrl_gae(const float *rewards, const float *values, const uint8_t *is_boundary,
             float *returns, int32_t num_episodes, int32_t num_timesteps,
             int64_t stride_e, int64_t stride_t, float gamma, float lam, void *stream);

/* ------------------------------------------------------------------------------------------
 * (3) fused PPO loss, forward + backward w.r.t. the head outputs -- replaces
 *     compute_actor_loss (:413-444), compute_critic_loss (:446-477),
 *     compute_autoregressive_loss (:398-404, with Continuous.mean_variance :232-241),
 *     compute_done_loss (:406-411) and the assembly of Agent.learn (:939-978).
 *     Rows are the b*T timesteps of one minibatch, row r = (episode r / T, step r % T).
 * ---------------------------------------------------------------------------------------- */
typedef struct xtrl_loss_cfg {
    int32_t num_actions, num_bins, state_dim;
    float   reward_min, reward_max;
    float   eps_clip, value_clip, entropy_weight;
    float   actor_weight, critic_weight, autoreg_weight;
    int32_t normalize_advantages;   /* 1 (reference default, :298)                            */
    int32_t bf16_io;                /* 0: head outputs / grads fp32; 1: bf16 (math stays fp32) */
    int32_t continuous;             /* Continuous actor (:224-277): raw_actions [N][2A], actions_f [N][A],
                                       old_log_probs [N][A]                                                   */
    int32_t squash;                 /* squash_continuous: entropy term = -log_prob (:425)                     */
} xtrl_loss_cfg;

typedef struct xtrl_loss_io {
    /* head outputs of the model, row-major (fp32, or bf16 when cfg.bf16_io) */
    const void *raw_actions;        /* [N][A] actor logits                                    */
    const void *value_logits;       /* [N][K] critic logits                                   */
    const void *state_pred_raw;     /* [N][2(S+1)] to_pred output, (mean,logvar) interleaved; NULL = skip */
    const void *done_logit;         /* [N] pre-sigmoid to_pred_done; NULL = skip              */
    /* episode buffer (always fp32 / int) */
    const int32_t *actions;         /* [N]                                                    */
    const float   *old_log_probs;   /* [N]                                                    */
    const float   *returns;         /* [N]                                                    */
    const float   *old_values;      /* [N] scalar (HL-Gauss decode of the stored logits, :427) */
    const float   *target;          /* [N][S+1] normalised (state||prev reward), :920-926     */
    const uint8_t *dones;           /* [N] is_boundary                                        */
    const int32_t *episode_lens;    /* [b]                                                    */
    /* outputs */
    void  *g_raw_actions;           /* [N][A]   d loss / d raw_actions                        */
    void  *g_value_logits;          /* [N][K]                                                 */
    void  *g_state_pred_raw;        /* [N][2(S+1)]                                            */
    void  *g_done_logit;            /* [N]                                                    */
    float *per_step;                /* [N][2] (actor, critic) per-timestep losses or NULL     */
    float *scalars;                 /* [8]: total, actor_mean, critic_mean, autoreg_mean, done_mean,
                                        n_masked, adv_mean, adv_var                           */
    double *workspace;              /* xtrl_ppo_loss_workspace_bytes() bytes, any contents     */
    const float *actions_f;         /* [N][A] continuous actions (cfg.continuous), else NULL   */
} xtrl_loss_io;

int64_t xtrl_ppo_loss_workspace_bytes(void);
int xtrl_ppo_loss_fwd_bwd(const xtrl_loss_cfg *cfg, const xtrl_loss_io *io,
                          int32_t batch, int32_t num_timesteps, void *stream);

/* HL-Gauss decode (hl_gauss(logits), :843/:427): out[n] = sum_k softmax(logits[n])_k * centre_k */
int xtrl_hl_decode(const float *logits, float *out, int64_t n, int32_t num_bins,
                   float reward_min, float reward_max, void *stream);

/* ------------------------------------------------------------------------------------------
 * (4) latent gene pool -- replaces LatentGenePool.evolve_ (evolution.py:76-184).  The three
 *     random tensors the reference draws (evolution.py:113,133,173) are INPUTS.
 * ---------------------------------------------------------------------------------------- */
typedef struct xtrl_pool_desc {
    int32_t dim;                    /* G                                                      */
    int32_t num_islands;            /* I                                                      */
    int32_t num_genes_per_island;   /* g                                                      */
    int32_t num_selected;           /* s                                                      */
    int32_t tournament_size;        /* k                                                      */
    int32_t num_elites;
    int32_t migrate;                /* 1 = this call is a migration step (evolution.py:137-141) */
    int32_t num_migrate;            /* rows moved when migrate (host validates the reference's
                                       int/float quirk)                           */
    float   mutation_std_dev;
    float   temperature;            /* 1.5 (evolution.py:79)                                  */
} xtrl_pool_desc;

/* genes [I*g][G] updated IN PLACE; fitness [I*g]; noise_tourn [I][c][s]; noise_cross [I][c][G];
 * noise_mut [I][g - (elites?1:0)][G]; selected_ids [I][s] (int64, as the reference returns);
 * scratch: unused since ABI 8 (the step is in place through shared memory); may be NULL.          */
int xtrl_evolve(float *genes, const float *fitness, const float *noise_tourn,
                const float *noise_cross, const float *noise_mut, int64_t *selected_ids,
                float *scratch, const xtrl_pool_desc *desc, void *stream);

/* ------------------------------------------------------------------------------------------
 * (5) optimiser step over the flat fp32 parameter buffer -- replaces the python loops of
 *     clip_grad_norm_ (:987) + AdoptAtan2.step (:749,:991) + the EMA post-step hook (:753).
 * ---------------------------------------------------------------------------------------- */
typedef struct xtrl_opt_cfg {
    float lr, init_lr, beta1, beta2, regen_rate, cautious_factor, a, b, max_grad_norm;
    int32_t first_step;             /* 1: initialise v=g*g, param_init=p, no update (ADOPT)   */
    float ema_mode;                 /* 0 none, 1 copy online->ema, 2 lerp with ema_decay      */
    float ema_decay;
    int32_t copy_ema_to_online;     /* update_model_with_ema_every hit                        */
    float grad_scale;               /* gradients are read as grads * grad_scale: 1 / world size after a SUM
                                       all-reduce (DDP's mean, :885/:981), 1 on one GPU; must be > 0        */
} xtrl_opt_cfg;

/* params, grads, m, v, param_init, ema: [n] fp32; seg_offsets [num_tensors+1] (int64, device):
 * the cautious mask is normalised per parameter tensor; workspace: 2 + num_tensors doubles.
 * ONE cooperative launch (two grid-wide barriers): the grid is sized to be co-resident.       */
int64_t xtrl_opt_workspace_bytes(int32_t num_tensors);
int xtrl_adopt_atan2_step(float *params, const float *grads, float *m, float *v, float *param_init,
                          float *ema, const int64_t *seg_offsets, int32_t num_tensors, int64_t n,
                          const xtrl_opt_cfg *cfg, double *workspace, float *grad_norm_out,
                          void *stream);

/* The same step with the data-parallel gradient all-reduce fused in as its first phase -- replaces DDP's gradient
 * all-reduce inside accelerator.backward (x_transformers_rl.py:885, :981) together with the optimiser (:987-991).
 * Every rank of the node keeps [gradients | extra sums] (n_reduce floats) followed by 2*world uint32 flags in a block from
 * xtrl_peer_alloc; the other ranks map it with xtrl_peer_open (CUDA IPC, NVLink / NVSwitch peer access).  The kernel
 * waits until every rank's block is complete, sums the `world` buffers in rank order into `reduced` (identical bits on
 * every rank), runs clip + AdoptAtan2 + EMA on reduced[0:n] * cfg->grad_scale, and does not return before every peer has
 * finished reading this rank's buffer (the caller may overwrite it afterwards).  All ranks must call it the same number
 * of times with epoch = 1, 2, 3, ...; a rank that waits longer than timeout_s (<= 0: 10 s) sets *status = 1.           */
typedef struct xtrl_peer_ctx {
    const float *const *grads;      /* HOST array [world]: every rank's gradient block as mapped on this device;
                                       grads[rank] == this rank's own `grads` argument                             */
    uint32_t *const *flags;         /* HOST array [world]: every rank's flag words (ready[world] | done[world])  */
    float *reduced;                 /* [n_reduce] fp32 on this device: receives the sum over ranks                */
    int32_t *status;                /* device int, 0 = ok                                                         */
    int64_t n_reduce;               /* >= n: elements summed (the optimiser consumes the first n)                 */
    int32_t rank, world;            /* world <= 8 (one NVSwitch node)                                             */
    uint32_t epoch;
    float timeout_s;
} xtrl_peer_ctx;

int xtrl_adopt_atan2_step_peers(float *params, const float *grads, float *m, float *v, float *param_init,
                                float *ema, const int64_t *seg_offsets, int32_t num_tensors, int64_t n,
                                const xtrl_opt_cfg *cfg, double *workspace, float *grad_norm_out,
                                const xtrl_peer_ctx *peers, void *stream);
/* peer memory: a zero-filled cudaMalloc block + its 64-byte IPC handle; map / unmap a peer's block; free one's own */
int xtrl_peer_alloc(int64_t bytes, void **ptr, void *handle64);
int xtrl_peer_open(const void *handle64, void **ptr);
int xtrl_peer_close(void *ptr);
int xtrl_peer_free(void *ptr);

/* ------------------------------------------------------------------------------------------
 * (6) learn-side sequence model, fused sub-layers on tcgen05 -- replace the
 *     teacher-forced Decoder pass inside Agent.learn (x_transformers_rl.py:928-935; x-transformers
 *     Decoder, App. A.1) piece by piece.  bf16 operands, fp32 accumulate, fp32 residual stream.
 *
 *     feed-forward sub-layer:  y = x + W2 . drop(gelu(W1 . (LayerNorm(x) * gamma) + b1)) + b2
 *     x, y, dy, dx: [N][dim] fp32 (N tokens, any order).  The forward saves nothing: the backward
 *     recomputes from x and regenerates the dropout mask from (*rng_step, rng_key).  Gradients of the
 *     parameters are ACCUMULATED (+=) into d_* (fp32, the layouts of the torch parameters).
 * ---------------------------------------------------------------------------------------- */
typedef struct xtrl_ff_io {
    const float *x;                 /* [N][dim]                                               */
    const float *dy;                /* [N][dim]            (backward)                         */
    const float *ln_gamma;          /* [dim]                                                  */
    const float *b1;                /* [hidden]                                               */
    const float *b2;                /* [dim]                                                  */
    const void  *w_img;             /* xtrl_ff_pack output (bf16 operand images of W1, W2)    */
    float *y;                       /* [N][dim]            (forward)                          */
    float *dx;                      /* [N][dim]            (backward)                         */
    float *d_ln_gamma, *d_w1, *d_b1, *d_w2, *d_b2;        /* += (backward)                    */
    const int64_t *rng_step;        /* device counter the caller bumps once per fwd/bwd pair; may
                                       be NULL when dropout == 0                              */
    uint32_t rng_key;               /* distinguishes call sites (layer index)                 */
    float dropout;                  /* p in [0,1); 0 = none                                   */
} xtrl_ff_io;

int64_t xtrl_ff_img_bytes(int32_t dim, int32_t hidden);          /* -1 if the shape is not built */
int xtrl_ff_pack(const float *w1 /*[hidden][dim]*/, const float *w2 /*[dim][hidden]*/, void *img,
                 int32_t dim, int32_t hidden, void *stream);
/* the same for nlayers (<= 8) layers in one launch: HOST arrays of device pointers */
int xtrl_ff_pack_layers(const float *const *w1, const float *const *w2, void *const *img, int32_t nlayers, int32_t dim,
                        int32_t hidden, void *stream);
int xtrl_ff_fwd(const xtrl_ff_io *io, int64_t N, int32_t dim, int32_t hidden, void *stream);
int xtrl_ff_bwd(const xtrl_ff_io *io, int64_t N, int32_t dim, int32_t hidden, void *stream);

/*     attention sub-layer around the attention core (plain block: no value gate / value residual):
 *       pre :  [q|k|v] = (LayerNorm(x) * gamma) . Wqkv^T, partial rotary on q, k; bf16 [b][H][n][16]
 *       post:  y = x + mask . (o . Wo^T), o bf16 [b][H][n][16], mask = position < lens[episode]
 *     x, y, dy, dx: [b*n][dim] fp32, token t = episode * n + position.  cos_tab / sin_tab: [n][dim_head/4]
 *     fp32 (angle = position * inv_freq, as the caller's rotary embedding defines it).  Parameter gradients
 *     are ACCUMULATED (+=).  pre_bwd returns only the LayerNorm path's dx (the residual path is the caller's).  */
typedef struct xtrl_attn_io {
    const float *x;                 /* [b*n][dim]              pre fwd/bwd, post fwd           */
    const float *ln_gamma;          /* [dim]                   pre                             */
    const float *cos_tab, *sin_tab; /* [n][4]                  pre                             */
    const void  *w_img;             /* xtrl_attn_pack output                                   */
    void *q, *k, *v;                /* bf16 [b][H][n][16]      pre fwd out                     */
    const void *dq, *dk, *dv;       /* bf16 [b][H][n][16]      pre bwd in                      */
    float *dx;                      /* [b*n][dim]              pre bwd out                     */
    float *d_ln_gamma, *d_wq, *d_wk, *d_wv;               /* += pre bwd                        */
    const void *o;                  /* bf16 [b][H][n][16]      post fwd/bwd in                 */
    const int32_t *lens;            /* [b]                     post                            */
    float *y;                       /* [b*n][dim]              post fwd out                    */
    const float *dy;                /* [b*n][dim]              post bwd in (dx of the residual path = dy);
                                       pre bwd: optional -- when non-NULL, dx = dy + the LayerNorm path's gradient, i.e.
                                       the whole block's dx (y = x + attn(LN(x))) without a separate add          */
    void *d_o;                      /* bf16 [b][H][n][16]      post bwd out                    */
    float *d_wo;                    /* += post bwd                                             */
    /* optional extra projection blocks of `pre` (value gate: to_v_gate; learned value-residual mix:
     * to_value_residual_mix) -- present when the output pointer (fwd) / gradient pointer (bwd) is non-NULL.
     * pre emits the PRE-activations (bias added); sigmoid / multiply / lerp are the caller's.            */
    const float *gate_bias;         /* [H*16]                                                  */
    const float *mix_bias;          /* [H]                                                     */
    void *gate;                     /* bf16 [b][H][n][16]      pre fwd out                     */
    void *mix;                      /* bf16 [b*n][16], first H columns valid   pre fwd out     */
    const void *d_gate, *d_mix;     /* same layouts            pre bwd in                      */
    float *d_wg, *d_bg, *d_wm, *d_bm;                     /* += pre bwd                        */
} xtrl_attn_io;

/* built: heads 4 or 8, gate 0/1, mix 0/1 except (8 heads, gate, mix); -1 otherwise */
int64_t xtrl_attn_img_bytes(int32_t dim, int32_t heads, int32_t dim_head, int32_t gate, int32_t mix);
int xtrl_attn_pack(const float *wq, const float *wk, const float *wv /*[H*16][dim]*/, const float *wo /*[dim][H*16]*/,
                   const float *wg /*[H*16][dim] or NULL*/, const float *wm /*[H][dim] or NULL*/,
                   void *img, int32_t dim, int32_t heads, int32_t dim_head, void *stream);
/* nlayers (<= 8) layers of ONE variant (wg / wm all given or NULL) in one launch: HOST arrays of device pointers */
int xtrl_attn_pack_layers(const float *const *wq, const float *const *wk, const float *const *wv, const float *const *wo,
                          const float *const *wg, const float *const *wm, void *const *img, int32_t nlayers,
                          int32_t dim, int32_t heads, int32_t dim_head, void *stream);
int xtrl_attn_pre_fwd(const xtrl_attn_io *io, int32_t b, int32_t n, int32_t heads, void *stream);
int xtrl_attn_pre_bwd(const xtrl_attn_io *io, int32_t b, int32_t n, int32_t heads, void *stream);
int xtrl_attn_post_fwd(const xtrl_attn_io *io, int32_t b, int32_t n, int32_t heads, void *stream);
int xtrl_attn_post_bwd(const xtrl_attn_io *io, int32_t b, int32_t n, int32_t heads, void *stream);

/*     attention core for dim_head = 16, n <= 512:  o = dropout(softmax(scale . q k^T + causal)) v  per (episode, head).
 *     q, k, v, o, d_o, dq, dk, dv: bf16 [bh][n][16] (bh = episodes * heads); lse: fp32 [bh][n], natural-log
 *     sum-exp of the scaled scores (written by fwd, read by bwd).  Deterministic (no atomics).               */
typedef struct xtrl_attn_core_io {
    const void *q, *k, *v;
    const void *o, *d_o;            /* bwd in                                                 */
    void *o_out;                    /* fwd out                                                */
    void *dq, *dk, *dv;             /* bwd out                                                */
    float *lse;
    const int64_t *rng_step;        /* as in xtrl_ff_io                                       */
    uint32_t rng_key;
    float dropout;
    float scale;                    /* dim_head^-1/2                                          */
} xtrl_attn_core_io;

int xtrl_attn_core_supported(int32_t n, int32_t dim_head);
int xtrl_attn_core_fwd(const xtrl_attn_core_io *io, int32_t bh, int32_t n, int32_t dim_head, void *stream);
int xtrl_attn_core_bwd(const xtrl_attn_core_io *io, int32_t bh, int32_t n, int32_t dim_head, void *stream);

/*     two-layer head MLPs:  y = W2 . act(W1 . [src0 | src1 | (src2)] + b1) + b2,  act = SiLU on hidden units < n_silu,
 *     identity on the rest (the done logit rides through the world-model head as one linear hidden unit).
 *     Sources are 48-wide fp32 rows ([N][48]; src2 optionally one row per episode); weights are ZERO-PADDED dense
 *     matrices w1 [hidden][in_dim], w2 [out][hidden] (built: (in_dim, hidden, out) in {(96|144, 96, 112), (96|144, 96, 16),
 *     (96, 64, 32)}); y / gy: bf16 [N][out_cols], out_cols <= out.  The backward recomputes the hidden layer, returns
 *     d_src* ([N][48] fp32 each, per token also for a per-episode src2) and ACCUMULATES (+=) the padded parameter
 *     gradients d_w1 [hidden][in_dim], d_b1 [hidden], d_w2 [out][hidden], d_b2 [out].                              */
typedef struct xtrl_mlp_io {
    const float *src0, *src1, *src2;
    int32_t src2_per_episode;
    const float *b1, *b2;           /* [hidden], [out] (padded)                               */
    const void *w_img;              /* xtrl_mlp_pack output                                   */
    void *y;                        /* fwd out                                                */
    const void *gy;                 /* bwd in                                                 */
    float *d_src0, *d_src1, *d_src2;
    float *d_w1, *d_b1, *d_w2, *d_b2;
    int32_t n_silu;
    int32_t out_cols;
    /* backward: d_src[s] = (bit s of d_src_accumulate ? d_src[s] : 0) + d_src_scale[s] * dX_s  (scale 0 is read as 1).  Lets
     * several heads that share an input add their gradients into ONE buffer (with frac_gradient's factor, :114-116, :533)
     * instead of three tensors and the adds that sum them.                                                          */
    float d_src_scale[3];
    int32_t d_src_accumulate;
} xtrl_mlp_io;

/*     rectangular blocks between the live parameters (several nn.Linear weights / biases) and the zero-padded dense
 *     buffers the head kernels use: dst[r][c] = src[r][c] (accumulate = 0: refresh before the forward) or += (1: add the
 *     padded gradient scratch to the parameters' gradients after the backward); up to 8 blocks per launch.            */
typedef struct xtrl_seg2d {
    const float *src;
    float *dst;
    int32_t rows, cols, src_pitch, dst_pitch;   /* pitches in floats */
} xtrl_seg2d;
int xtrl_seg2d_copy(const xtrl_seg2d *segs, int32_t nseg, int32_t accumulate, void *stream);

int64_t xtrl_mlp_img_bytes(int32_t in_dim, int32_t hidden, int32_t out);      /* -1 if not built */
int xtrl_mlp_pack(const float *w1, const float *w2, void *img, int32_t in_dim, int32_t hidden, int32_t out, void *stream);
int xtrl_mlp_fwd(const xtrl_mlp_io *io, int64_t N, int32_t n, int32_t in_dim, int32_t hidden, int32_t out, void *stream);
int xtrl_mlp_bwd(const xtrl_mlp_io *io, int64_t N, int32_t n, int32_t in_dim, int32_t hidden, int32_t out, void *stream);

/*     the Decoder's final LayerNorm (F.layer_norm(x) * gamma over dim = 48 rows), forward and backward;
 *     d_gamma is ACCUMULATED (+=).                                                                      */
int xtrl_ln_fwd(const float *x, const float *gamma, float *y, int64_t N, int32_t dim, void *stream);
int xtrl_ln_bwd(const float *x, const float *dy, const float *gamma, float *dx, float *d_gamma, int64_t N, int32_t dim,
                void *stream);

/*     weight gradient of a skinny input layer into 48 outputs (+=): d_w[k][j] (out_kj = 1, embedding layout [K][48]) or
 *     d_w[j][k] (out_kj = 0, Linear layout [48][K]) = sum_t f[t][k] dy[t][j]; f = feats [N][K] fp32, or the one-hot of
 *     ids[t] in [-1, K) (int64, -1 = no row); d_b[j] += sum_t dy[t][j] when d_b is given.  K <= 16.                   */
int xtrl_small_dw(const float *feats, const int64_t *ids, const float *dy, float *d_w, float *d_b, int64_t N, int32_t K,
                  int32_t dim, int32_t out_kj, void *stream);

/* ------------------------------------------------------------------------------------------
 * (7) minibatch prologue of Agent.learn and its backward -- replace the DataLoader gather (:876-901), the prev-action /
 *     prev-reward shifts (:911-918), RSNorm.eval on the packed rows (:920-926), the masked batch mean of the RSNorm copy
 *     update (:1005, :601) and the input side of WorldModelActorCritic.forward (:492-503, :516-529).
 *     Sources are EPISODE-major [E][T][..]; outputs are minibatch-major [b][T][..], hidden size 48.
 * ---------------------------------------------------------------------------------------- */
typedef struct xtrl_prep_io {
    const float   *states;          /* [E][T][S]                                              */
    const void    *actions;         /* [E][T] int32, or [E][T][A] float when continuous       */
    const float   *rewards;         /* [E][T]                                                 */
    const float   *old_log_probs;   /* [E][T], or [E][T][A] when continuous                   */
    const float   *returns, *old_values;                    /* [E][T]                         */
    const uint8_t *dones;           /* [E][T]                                                 */
    const int32_t *lens;            /* [E]                                                    */
    const int32_t *gene_ids;        /* [E] or NULL                                            */
    const int64_t *idx;             /* [b] the minibatch's episodes                           */
    const float   *rs_mean, *rs_var;                        /* [S+1] RSNorm statistics         */
    const float   *w_in;            /* transformer.project_in.weight [48][S]                  */
    const float   *w_se, *b_se;     /* to_state_embed weight [48][S], bias [48]               */
    const float   *w_act;           /* action_embeds: table [A][48], or Linear weight [48][A] when continuous */
    const float   *b_act;           /* Linear bias [48] (continuous), else NULL               */
    const float   *reward_embed;    /* [48]                                                   */
    const float   *reward_flag;     /* device scalar: the reward-conditioning Bernoulli (:501) */
    float *x0;                      /* [b][T][48] project_in(state) + action_embeds(prev action) + prev reward * reward_embed * flag */
    float *se;                      /* [b][T][48] to_state_embed(state)                       */
    float *ae_next;                 /* [b][T][48] action_embeds(action)                       */
    float *packed;                  /* [b][T][S+1] normalised (state || prev reward)          */
    void  *mb_actions;              /* gathered copies of the per-row fields, layouts as the sources */
    float *mb_old_log_probs, *mb_returns, *mb_old_values;
    uint8_t *mb_dones;
    int32_t *mb_lens, *mb_gene_ids; /* [b]                                                    */
    float *rs_sum;                  /* [S+2] += masked column sums of packed, then the masked row count */
} xtrl_prep_io;

typedef struct xtrl_prep_bwd_io {
    const float *dx0;               /* [b][T][48]                                             */
    const float *d_se;              /* [b][T][48] or NULL                                     */
    const float *d_ae_next;         /* [b][T][48] or NULL                                     */
    const float *packed;            /* [b][T][S+1]                                            */
    const void  *mb_actions;        /* [b][T] int32 / [b][T][A] float                         */
    const float *reward_flag;
    int32_t T;                      /* timesteps per episode (row r starts an episode when r % T == 0) */
    float *d_w_in, *d_w_se, *d_b_se, *d_w_act, *d_b_act, *d_reward_embed;      /* += ; layouts of the parameters */
} xtrl_prep_bwd_io;

int xtrl_learn_prep_fwd(const xtrl_prep_io *io, int32_t b, int32_t T, int32_t state_dim, int32_t num_actions,
                        int32_t continuous, void *stream);
int xtrl_learn_prep_bwd(const xtrl_prep_bwd_io *io, int32_t b, int32_t T, int32_t state_dim, int32_t num_actions,
                        int32_t continuous, void *stream);

/* sizeof of the ABI structures as this library was compiled (binding self-check): 0 xtrl_model_desc, 1 xtrl_weight_layout,
 * 2 xtrl_step_io, 3 xtrl_loss_cfg, 4 xtrl_loss_io, 5 xtrl_pool_desc, 6 xtrl_opt_cfg, 7 xtrl_ff_io, 8 xtrl_attn_io,
 * 9 xtrl_attn_core_io, 10 xtrl_mlp_io, 11 xtrl_prep_io, 12 xtrl_prep_bwd_io, 13 xtrl_peer_ctx, 14 xtrl_seg2d; -1 otherwise. */
int64_t xtrl_struct_size(int32_t which);

#ifdef __cplusplus
}
#endif
#endif /* XTRL_H */
