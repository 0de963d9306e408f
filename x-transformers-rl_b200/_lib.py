"""ctypes binding of libxtrl.so (include/xtrl.h).  The library is the ONLY compute backend: if it is
missing or does not load, importing the ops fails loudly -- there is no CPU / eager fallback.

Structures below mirror include/xtrl.h field for field; `tests/test_abi.py` checks every
declared symbol is exported and that sizes agree with what the C side reports.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

XTRL_ABI_VERSION = 9
XTRL_MAX_DEPTH = 16

c_i32, c_i64, c_f32, c_vp = C.c_int32, C.c_int64, C.c_float, C.c_void_p


class ModelDesc(C.Structure):
    _fields_ = [(n, c_i32) for n in (
        'state_dim', 'num_actions', 'dim', 'heads', 'dim_head', 'depth', 'num_bins', 'dim_gene',
        'gate_values', 'value_residual', 'learned_mix', 'max_timesteps', 'rotary_true_pos', 'kv_bf16')] + [
        ('reward_min', c_f32), ('reward_max', c_f32), ('dense_bf16', c_i32), ('continuous_actions', c_i32),
        ('squash_continuous', c_i32)]


class _LayerLayout(C.Structure):
    _fields_ = [(n, c_i64) for n in ('ln1_gamma', 'qkv_w', 'qkv_b', 'out_w', 'ln2_gamma',
                                     'ff1_w', 'ff1_b', 'ff2_w', 'ff2_b')]


class WeightLayout(C.Structure):
    _fields_ = [('total', c_i64), ('s_pad', c_i32), ('n_qkv', c_i32), ('head_in', c_i32)] + [
        (n, c_i64) for n in ('in_proj', 'state_embed_bias', 'action_embed', 'reward_embed', 'latent_w',
                             'latent_b', 'final_gamma', 'head1_w', 'head1_b', 'actor2_w', 'actor2_b',
                             'critic2_w', 'critic2_b')] + [('layer', _LayerLayout * XTRL_MAX_DEPTH), ('action_embed_bias', c_i64)]


class StepIO(C.Structure):
    _fields_ = [(n, c_vp) for n in ('state', 'prev_action', 'prev_reward', 'pos', 'rs_mean', 'rs_var',
                                    'genes', 'gene_ids', 'uniform', 'forced_action')] + [
        ('reward_flag', c_f32), ('pos_uniform', c_i32), ('tc_weights', c_vp)] + [(n, c_vp) for n in ('action', 'log_prob', 'value', 'raw_actions', 'value_logits',
                                           'prev_action_f', 'normal', 'forced_action_f', 'action_f')] + [('envs_per_cta', c_i32)]


class LossCfg(C.Structure):
    _fields_ = [('num_actions', c_i32), ('num_bins', c_i32), ('state_dim', c_i32),
                ('reward_min', c_f32), ('reward_max', c_f32),
                ('eps_clip', c_f32), ('value_clip', c_f32), ('entropy_weight', c_f32),
                ('actor_weight', c_f32), ('critic_weight', c_f32), ('autoreg_weight', c_f32),
                ('normalize_advantages', c_i32), ('bf16_io', c_i32), ('continuous', c_i32), ('squash', c_i32)]


class LossIO(C.Structure):
    _fields_ = [(n, c_vp) for n in (
        'raw_actions', 'value_logits', 'state_pred_raw', 'done_logit',
        'actions', 'old_log_probs', 'returns', 'old_values', 'target', 'dones', 'episode_lens',
        'g_raw_actions', 'g_value_logits', 'g_state_pred_raw', 'g_done_logit', 'per_step', 'scalars', 'workspace',
        'actions_f')]


class PoolDesc(C.Structure):
    _fields_ = [(n, c_i32) for n in ('dim', 'num_islands', 'num_genes_per_island', 'num_selected',
                                     'tournament_size', 'num_elites', 'migrate', 'num_migrate')] + [
        ('mutation_std_dev', c_f32), ('temperature', c_f32)]


class OptCfg(C.Structure):
    _fields_ = [(n, c_f32) for n in ('lr', 'init_lr', 'beta1', 'beta2', 'regen_rate', 'cautious_factor',
                                     'a', 'b', 'max_grad_norm')] + [
        ('first_step', c_i32), ('ema_mode', c_f32), ('ema_decay', c_f32), ('copy_ema_to_online', c_i32), ('grad_scale', c_f32)]


class PeerCtx(C.Structure):
    _fields_ = [('grads', c_vp), ('flags', c_vp), ('reduced', c_vp), ('status', c_vp), ('n_reduce', c_i64),
                ('rank', c_i32), ('world', c_i32), ('epoch', C.c_uint32), ('timeout_s', c_f32)]


class Seg2d(C.Structure):
    _fields_ = [('src', c_vp), ('dst', c_vp), ('rows', c_i32), ('cols', c_i32), ('src_pitch', c_i32), ('dst_pitch', c_i32)]


class FfIO(C.Structure):
    _fields_ = [(n, c_vp) for n in ('x', 'dy', 'ln_gamma', 'b1', 'b2', 'w_img', 'y', 'dx',
                                    'd_ln_gamma', 'd_w1', 'd_b1', 'd_w2', 'd_b2', 'rng_step')] + [
        ('rng_key', C.c_uint32), ('dropout', c_f32)]


class AttnIO(C.Structure):
    _fields_ = [(n, c_vp) for n in ('x', 'ln_gamma', 'cos_tab', 'sin_tab', 'w_img', 'q', 'k', 'v', 'dq', 'dk', 'dv', 'dx',
                                    'd_ln_gamma', 'd_wq', 'd_wk', 'd_wv', 'o', 'lens', 'y', 'dy', 'd_o', 'd_wo',
                                    'gate_bias', 'mix_bias', 'gate', 'mix', 'd_gate', 'd_mix', 'd_wg', 'd_bg', 'd_wm', 'd_bm')]


class AttnCoreIO(C.Structure):
    _fields_ = [(n, c_vp) for n in ('q', 'k', 'v', 'o', 'd_o', 'o_out', 'dq', 'dk', 'dv', 'lse', 'rng_step')] + [
        ('rng_key', C.c_uint32), ('dropout', c_f32), ('scale', c_f32)]


class MlpIO(C.Structure):
    _fields_ = [('src0', c_vp), ('src1', c_vp), ('src2', c_vp), ('src2_per_episode', c_i32), ('b1', c_vp), ('b2', c_vp), ('w_img', c_vp),
                ('y', c_vp), ('gy', c_vp), ('d_src0', c_vp), ('d_src1', c_vp), ('d_src2', c_vp), ('d_w1', c_vp), ('d_b1', c_vp),
                ('d_w2', c_vp), ('d_b2', c_vp), ('n_silu', c_i32), ('out_cols', c_i32), ('d_src_scale', c_f32 * 3), ('d_src_accumulate', c_i32)]


class PrepIO(C.Structure):
    _fields_ = [(n, c_vp) for n in ('states', 'actions', 'rewards', 'old_log_probs', 'returns', 'old_values', 'dones', 'lens', 'gene_ids', 'idx',
                                    'rs_mean', 'rs_var', 'w_in', 'w_se', 'b_se', 'w_act', 'b_act', 'reward_embed', 'reward_flag',
                                    'x0', 'se', 'ae_next', 'packed', 'mb_actions', 'mb_old_log_probs', 'mb_returns', 'mb_old_values',
                                    'mb_dones', 'mb_lens', 'mb_gene_ids', 'rs_sum')]


class PrepBwdIO(C.Structure):
    _fields_ = [(n, c_vp) for n in ('dx0', 'd_se', 'd_ae_next', 'packed', 'mb_actions', 'reward_flag')] + [('T', c_i32)] + [
        (n, c_vp) for n in ('d_w_in', 'd_w_se', 'd_b_se', 'd_w_act', 'd_b_act', 'd_reward_embed')]


# every symbol include/xtrl.h declares: name -> (restype, argtypes)
SYMBOLS = {
    'xtrl_abi_version': (c_i32, []),
    'xtrl_last_error': (C.c_char_p, []),
    'xtrl_device_sm': (c_i32, []),
    'xtrl_device_num_sms': (c_i32, []),
    'xtrl_wm_weight_layout': (c_i32, [C.POINTER(ModelDesc), C.POINTER(WeightLayout)]),
    'xtrl_wm_kv_cache_bytes': (c_i64, [C.POINTER(ModelDesc), c_i64]),
    'xtrl_wm_tc_weights_bytes': (c_i64, [C.POINTER(ModelDesc)]),
    'xtrl_wm_pack_tc': (c_i32, [C.POINTER(ModelDesc), c_vp, c_vp, c_vp]),
    'xtrl_wm_step': (c_i32, [C.POINTER(ModelDesc), c_vp, c_vp, C.POINTER(StepIO), c_i32, c_vp]),
    'xtrl_gae': (c_i32, [c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i64, c_i64, c_f32, c_f32, c_vp]),
    'xtrl_ppo_loss_workspace_bytes': (c_i64, []),
    'xtrl_ppo_loss_fwd_bwd': (c_i32, [C.POINTER(LossCfg), C.POINTER(LossIO), c_i32, c_i32, c_vp]),
    'xtrl_hl_decode': (c_i32, [c_vp, c_vp, c_i64, c_i32, c_f32, c_f32, c_vp]),
    'xtrl_evolve': (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, C.POINTER(PoolDesc), c_vp]),
    'xtrl_opt_workspace_bytes': (c_i64, [c_i32]),
    'xtrl_adopt_atan2_step': (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i64,
                                      C.POINTER(OptCfg), c_vp, c_vp, c_vp]),
    'xtrl_adopt_atan2_step_peers': (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i64,
                                            C.POINTER(OptCfg), c_vp, c_vp, C.POINTER(PeerCtx), c_vp]),
    'xtrl_peer_alloc': (c_i32, [c_i64, C.POINTER(c_vp), c_vp]),
    'xtrl_peer_open': (c_i32, [c_vp, C.POINTER(c_vp)]),
    'xtrl_peer_close': (c_i32, [c_vp]),
    'xtrl_peer_free': (c_i32, [c_vp]),
    'xtrl_seg2d_copy': (c_i32, [c_vp, c_i32, c_i32, c_vp]),
    'xtrl_ff_img_bytes': (c_i64, [c_i32, c_i32]),
    'xtrl_ff_pack': (c_i32, [c_vp, c_vp, c_vp, c_i32, c_i32, c_vp]),
    'xtrl_ff_pack_layers': (c_i32, [c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_vp]),
    'xtrl_ff_fwd': (c_i32, [C.POINTER(FfIO), c_i64, c_i32, c_i32, c_vp]),
    'xtrl_ff_bwd': (c_i32, [C.POINTER(FfIO), c_i64, c_i32, c_i32, c_vp]),
    'xtrl_attn_img_bytes': (c_i64, [c_i32, c_i32, c_i32, c_i32, c_i32]),
    'xtrl_attn_pack': (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_vp]),
    'xtrl_attn_pack_layers': (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_vp]),
    'xtrl_attn_pre_fwd': (c_i32, [C.POINTER(AttnIO), c_i32, c_i32, c_i32, c_vp]),
    'xtrl_attn_pre_bwd': (c_i32, [C.POINTER(AttnIO), c_i32, c_i32, c_i32, c_vp]),
    'xtrl_attn_post_fwd': (c_i32, [C.POINTER(AttnIO), c_i32, c_i32, c_i32, c_vp]),
    'xtrl_attn_post_bwd': (c_i32, [C.POINTER(AttnIO), c_i32, c_i32, c_i32, c_vp]),
    'xtrl_attn_core_supported': (c_i32, [c_i32, c_i32]),
    'xtrl_attn_core_fwd': (c_i32, [C.POINTER(AttnCoreIO), c_i32, c_i32, c_i32, c_vp]),
    'xtrl_attn_core_bwd': (c_i32, [C.POINTER(AttnCoreIO), c_i32, c_i32, c_i32, c_vp]),
    'xtrl_mlp_img_bytes': (c_i64, [c_i32, c_i32, c_i32]),
    'xtrl_mlp_pack': (c_i32, [c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_vp]),
    'xtrl_mlp_fwd': (c_i32, [C.POINTER(MlpIO), c_i64, c_i32, c_i32, c_i32, c_i32, c_vp]),
    'xtrl_mlp_bwd': (c_i32, [C.POINTER(MlpIO), c_i64, c_i32, c_i32, c_i32, c_i32, c_vp]),
    'xtrl_struct_size': (c_i64, [c_i32]),
    'xtrl_small_dw': (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_i32, c_i32, c_vp]),
    'xtrl_ln_fwd': (c_i32, [c_vp, c_vp, c_vp, c_i64, c_i32, c_vp]),
    'xtrl_ln_bwd': (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_vp]),
    'xtrl_learn_prep_fwd': (c_i32, [C.POINTER(PrepIO), c_i32, c_i32, c_i32, c_i32, c_i32, c_vp]),
    'xtrl_learn_prep_bwd': (c_i32, [C.POINTER(PrepBwdIO), c_i32, c_i32, c_i32, c_i32, c_i32, c_vp]),
}


# order of xtrl_struct_size(which)
ABI_STRUCTS = (ModelDesc, WeightLayout, StepIO, LossCfg, LossIO, PoolDesc, OptCfg, FfIO, AttnIO, AttnCoreIO, MlpIO, PrepIO, PrepBwdIO, PeerCtx, Seg2d)


class XtrlError(RuntimeError):
    pass


_lib = None


def lib_path() -> str:
    return _build.LIB_PATH


def load():
    """dlopen libxtrl.so (building it first if the in-tree copy is absent or stale)."""
    global _lib
    if _lib is not None:
        return _lib
    import torch  # noqa: F401  (loads libcudart.so.12 into the process before our dlopen)
    path = _build.LIB_PATH
    if not _build.is_fresh():
        try:
            path = _build.build()
        except Exception as exc:
            if not os.path.exists(path):
                raise XtrlError(f'libxtrl.so is missing and could not be built ({exc}); '
                                'this package has no CPU fallback') from exc
    try:
        lib = C.CDLL(path)
    except OSError as exc:
        raise XtrlError(f'cannot load {path}: {exc}; this package has no CPU fallback') from exc
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError = ABI mismatch, let it surface
        fn.restype, fn.argtypes = res, args
    got = lib.xtrl_abi_version()
    if got != XTRL_ABI_VERSION:
        raise XtrlError(f'libxtrl.so ABI {got} != binding {XTRL_ABI_VERSION}')
    for which, st in enumerate(ABI_STRUCTS):
        if lib.xtrl_struct_size(which) != C.sizeof(st):
            raise XtrlError(f'{st.__name__}: ctypes size {C.sizeof(st)} != library {lib.xtrl_struct_size(which)}')
    _lib = lib
    return lib


_cudart = None


def cudart():
    """the CUDA runtime already mapped into this process (torch's), for raw cudaMemcpyAsync on hot host loops."""
    global _cudart
    if _cudart is None:
        import torch  # noqa: F401
        rt = C.CDLL('libcudart.so.12')
        rt.cudaMemcpyAsync.restype = c_i32
        rt.cudaMemcpyAsync.argtypes = [c_vp, c_vp, C.c_size_t, c_i32, c_vp]
        _cudart = rt
    return _cudart


def check(rc: int, what: str):
    if rc != 0:
        msg = load().xtrl_last_error()
        raise XtrlError(f'{what} failed ({rc}): {msg.decode() if msg else "?"}')
