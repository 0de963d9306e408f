"""CPU: the C-ABI library builds, loads and exports every symbol include/xtrl.h declares; host-only
entry points (weight layout, sizes, argument validation) behave.  No kernel is launched here."""
import ctypes as C
import os
import re

import pytest

from helpers import xtrl

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, 'include', 'xtrl.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(xtrl_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    pkg = xtrl()
    lib = pkg.load_library()
    names = header_symbols()
    assert len(names) >= 12
    from importlib import import_module
    L = import_module('x-transformers-rl_b200._lib')
    assert sorted(L.SYMBOLS) == names, (sorted(L.SYMBOLS), names)
    for n in names:
        assert hasattr(lib, n), n
    assert lib.xtrl_abi_version() == L.XTRL_ABI_VERSION == 9
    for which, st in enumerate(L.ABI_STRUCTS):                       # every ctypes mirror has the size the C side compiled
        assert lib.xtrl_struct_size(which) == C.sizeof(st), st.__name__
    assert lib.xtrl_struct_size(len(L.ABI_STRUCTS)) == -1
    assert os.path.basename(pkg.lib_path()) == 'libxtrl.so' and ROOT in pkg.lib_path()   # in-tree, not site-packages


def test_weight_layout_and_sizes():
    ops = xtrl().ops
    cfg = ops.ModelConfig(state_dim=8, num_actions=4, reward_range=(-5., 5.), max_timesteps=500, heads=8, depth=4)
    lay = ops.weight_layout(cfg)
    D, I = 48, 128
    assert lay.n_qkv == 3 * I and lay.head_in == 2 * D and lay.s_pad == 8
    n_params = (8 * 2 * D + D + 4 * D + D + D + 2 * D * 4 * D + 4 * D + 2 * D * 4 + 4 + 2 * D * 100 + 100
                + 4 * (D + D * 3 * I + 3 * I + I * D + D + D * 4 * D + 4 * D + 4 * D * D + D))
    assert n_params <= lay.total <= n_params + 4 * 64          # only alignment padding on top
    assert lay.layer[3].ff2_b > 0 and lay.layer[4].ff2_b == -1
    lib = xtrl().load_library()
    d = cfg.c_struct()
    assert lib.xtrl_wm_kv_cache_bytes(C.byref(d), 4096) == 2 * 4 * 4096 * 8 * 500 * 16 * 4   # 8.4 GB fp32, SURVEY 8d c3
    cfg.kv_bf16 = True
    d = cfg.c_struct()
    assert lib.xtrl_wm_kv_cache_bytes(C.byref(d), 4096) == 2 * 4 * 4096 * 8 * 500 * 16 * 2


def test_unsupported_configs_are_errors_not_fallbacks():
    pkg = xtrl()
    ops = pkg.ops
    with pytest.raises(pkg.XtrlError, match='dim_head'):
        ops.weight_layout(ops.ModelConfig(state_dim=8, num_actions=4, reward_range=(-1, 1), max_timesteps=10, dim_head=64))
    with pytest.raises(pkg.XtrlError, match='not supported'):
        ops.ModelConfig.from_learner_kwargs(5, 2, (-1, 1), 10, dict(depth=1, heads=4, attn_dim_head=16, attn_flash=True))
    import torch
    with pytest.raises(pkg.XtrlError, match='CUDA tensor'):
        ops.gae(torch.zeros(2, 3), torch.zeros(2, 3), torch.zeros(2, 3, dtype=torch.bool))


def test_pack_weights_roundtrip_on_cpu():
    """host packing logic: every state_dict entry lands K-major at its layout offset."""
    import torch
    from oracle import restated as R
    from helpers import random_state_dict, model_config
    cfg = R.Config(state_dim=5, num_actions=3, heads=4, depth=2, gate_values=True, value_residual=True,
                   learned_mix=True, evolutionary=True, dim_gene=32)
    sd = random_state_dict(cfg, seed=3)
    ops = xtrl().ops
    mc = model_config(cfg, 10)
    blob = ops.pack_weights(sd, mc, device='cpu')
    lay = ops.weight_layout(mc)
    D, I = 48, 64
    w = blob[lay.layer[1].qkv_w: lay.layer[1].qkv_w + D * lay.n_qkv].view(D, lay.n_qkv)
    pa, _ = R._L(1)
    assert torch.equal(w[:, :I], sd[f'{pa}.1.to_q.weight'].T)
    assert torch.equal(w[:, 3 * I:4 * I], sd[f'{pa}.1.to_v_gate.weight'].T)
    assert torch.equal(w[:, 4 * I:4 * I + 4], sd[f'{pa}.1.to_value_residual_mix.0.weight'].T)
    a2 = blob[lay.actor2_w: lay.actor2_w + 2 * D * 4].view(2 * D, 4)
    assert torch.equal(a2[:, :3], sd['action_head.2.weight'].T) and torch.all(a2[:, 3] == 0)
    inp = blob[lay.in_proj: lay.in_proj + 8 * 2 * D].view(8, 2 * D)
    assert torch.equal(inp[:5, :D], sd['transformer.project_in.weight'].T) and torch.all(inp[5:] == 0)


def test_learn_side_shape_queries_are_host_only():
    """which learn-side shapes are built is answered without a GPU (the caller keeps its torch ops for the rest)."""
    lib = xtrl().load_library()
    assert lib.xtrl_ff_img_bytes(48, 192) == 2 * 2 * 48 * 192 and lib.xtrl_ff_img_bytes(64, 256) == -1
    assert lib.xtrl_attn_img_bytes(48, 8, 16, 0, 0) == 2 * (48 * 128 + 384 * 48)
    assert lib.xtrl_attn_img_bytes(48, 4, 16, 1, 1) == 2 * (48 * 64 + (4 * 64 + 16) * 48)
    assert lib.xtrl_attn_img_bytes(48, 8, 16, 1, 1) == -1 and lib.xtrl_attn_img_bytes(48, 8, 32, 0, 0) == -1
    assert lib.xtrl_attn_core_supported(500, 16) == 1 and lib.xtrl_attn_core_supported(513, 16) == 0
    assert lib.xtrl_mlp_img_bytes(96, 96, 112) == 2 * (96 * 96 + 112 * 96) and lib.xtrl_mlp_img_bytes(144, 64, 32) == -1
