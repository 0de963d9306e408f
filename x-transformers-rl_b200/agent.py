"""Agent with the reference's interface (x_transformers_rl/x_transformers_rl.py:644-1065).

Owns the online model (flat fp32 parameters), the EMA copy (a second flat buffer the rollout kernel
reads), AdoptAtan2 state (three more flat buffers), RSNorm statistics and the gene pool.
`learn` = GAE kernel -> epochs x minibatches of [torch sequence model fwd -> fused loss kernel
(fwd+bwd) -> autograd through the model -> one flat-gradient all-reduce -> optimiser/EMA kernel].
"""
from __future__ import annotations

from collections import namedtuple
from dataclasses import dataclass
from pathlib import Path

import torch
import torch.nn.functional as F
from torch import nn, tensor

from . import ops
from .distributed import Accelerator, PeerGroup, all_reduce_sum_, broadcast_from_rank0_, world_size
from .evolution import LatentGenePool
from .model import FusedPPOLoss, WorldModelActorCritic
from .schedule import EmaSchedule

Memory = namedtuple('Memory', ['state', 'action', 'action_log_prob', 'reward', 'is_boundary', 'value'])


@dataclass
class RolloutBuffer:
    """device-resident trajectories of one learning update, TIME-major as the step kernel writes them:
    states (T,B,S) f32, actions (T,B) i32, log_probs (T,B), rewards (T,B), is_boundary (T,B) bool,
    values (T,B) scalar HL-Gauss expectations; episode_lens (B,) i32, gene_ids (B,) i32."""
    states: torch.Tensor
    actions: torch.Tensor
    log_probs: torch.Tensor
    rewards: torch.Tensor
    is_boundary: torch.Tensor
    values: torch.Tensor
    episode_lens: torch.Tensor
    gene_ids: torch.Tensor

    @classmethod
    def allocate(cls, T, B, S, device, continuous_dim=0):
        """continuous_dim = A for continuous actions: actions (T,B,A) f32 and per-dim log_probs (T,B,A)."""
        z = lambda *s, dt=torch.float32: torch.zeros(*s, device=device, dtype=dt)
        acts = z(T, B, continuous_dim) if continuous_dim else z(T, B, dt=torch.int32)
        logp = z(T, B, continuous_dim) if continuous_dim else z(T, B)
        return cls(z(T, B, S), acts, logp, z(T, B), z(T, B, dt=torch.bool), z(T, B), z(B, dt=torch.int32), z(B, dt=torch.int32))

    @classmethod
    def cat(cls, bufs):
        if len(bufs) == 1:
            return bufs[0]
        c = lambda name, dim: torch.cat([getattr(b, name) for b in bufs], dim=dim)
        return cls(c('states', 1), c('actions', 1), c('log_probs', 1), c('rewards', 1), c('is_boundary', 1),
                   c('values', 1), c('episode_lens', 0), c('gene_ids', 0))


class RSNorm(nn.Module):
    """running mean / variance of (state || prev reward), x_transformers_rl.py:565-612.  The eval
    normalisation itself is fused into the kernels; this module only holds and updates the statistics."""

    def __init__(self, dim, eps=1e-5):
        super().__init__()
        self.dim, self.eps = dim, eps
        self.register_buffer('step', tensor(1))
        self.register_buffer('running_mean', torch.zeros(dim))
        self.register_buffer('running_variance', torch.ones(dim))
        self.time = 1                    # host mirror of `step`: no .item() sync on the hot path

    def normalize(self, x):
        return (x - self.running_mean) / self.running_variance.sqrt().clamp(min=self.eps)

    @torch.no_grad()
    def update_(self, batch_mean):
        """the train-branch update (:598-610) given the (already all-reduced) batch mean."""
        t = float(self.time)
        delta = batch_mean - self.running_mean
        self.running_variance.copy_((t - 1) / t * (self.running_variance + delta ** 2 / t))
        self.running_mean.add_(delta / t)
        self.step.add_(1)
        self.time += 1

    def copy_from(self, other):
        self.running_mean.copy_(other.running_mean); self.running_variance.copy_(other.running_variance)
        self.step.copy_(other.step); self.time = other.time


class Hiddens:
    """opaque recurrent state returned by Agent.forward (the KV cache of one env)."""

    def __init__(self, stepper, weights_version):
        self.stepper, self.weights_version = stepper, weights_version


class _LearnGraph:
    """static minibatch buffers + the CUDA graph of Agent._fwd_bwd over them.  Captured lazily on the first real
    minibatch (after two eager warm-up passes on a side stream, as graph capture requires)."""

    def __init__(self, agent, b, T, E):
        self.agent, dev, cfg = agent, agent.device, agent.cfg
        A = cfg.num_actions
        z = lambda *s, dt=torch.float32: torch.zeros(*s, device=dev, dtype=dt)
        cont = agent.continuous_actions
        if agent.fused_learn:          # bf16 class: the prologue kernel gathers straight out of the persistent episode-major sources
            self.src, self.idx, self.mb = agent._learn_sources(E, T), z(b, dt=torch.int64), None
        else:
            self.mb = dict(states=z(b, T, cfg.state_dim), actions=z(b, T, A) if cont else z(b, T, dt=torch.int32), rewards=z(b, T),
                           old_log_probs=z(b, T, A) if cont else z(b, T), returns=z(b, T), old_values=z(b, T),
                           dones=z(b, T, dt=torch.bool), lens=z(b, dt=torch.int32), gene_ids=z(b, dt=torch.int32))
        self.flag = torch.ones((), device=dev)
        self.graph, self.scalars = None, None

    def _pass(self, collective=True):
        agent = self.agent
        if agent.fused_learn:
            return agent._fwd_bwd_prepared(self.src, self.idx, self.flag, collective=collective)
        return agent._fwd_bwd(self.mb, self.flag, collective=collective)

    def capture(self):
        """two eager warm-up passes on a side stream, then the capture.  A failure here is an error (there is no eager
        fallback): the gradient buffer is left zeroed either way, so nothing half-accumulated survives."""
        agent = self.agent
        try:
            cur = torch.cuda.current_stream(agent.device)
            side = torch.cuda.Stream(agent.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                for _ in range(2):
                    self._pass(collective=False)
            cur.wait_stream(side)
            agent._comm.zero_()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self.scalars = self._pass()
            self.graph = graph
        finally:
            torch.cuda.current_stream(agent.device).synchronize()
            agent._comm.zero_()                     # capture does not execute; start the replay from clean gradients

    def run(self):
        if self.graph is None:
            self.capture()
        self.graph.replay()
        return self.scalars


class Agent(nn.Module):
    def __init__(self, state_dim, num_actions, reward_range, epochs, max_timesteps, batch_size, lr, betas, lam, gamma,
                 beta_s, regen_reg_rate, cautious_factor, eps_clip, value_clip, ema_decay, continuous_actions=False,
                 squash_continuous=True, critic_pred_num_bins=100, hidden_dim=48, evolutionary=False, evolve_every=1,
                 evolve_after_step=20, latent_gene_pool=None, world_model=None, dropout=0.25, max_grad_norm=0.5,
                 frac_actor_critic_head_gradient=0.5, ema_kwargs=None, save_path='./ppo.pt', accelerator=None,
                 actor_loss_weight=1., critic_loss_weight=1., autoregressive_loss_weight=1.,
                 rollout_positions='reference', kv_cache_dtype='fp32', rollout_dense_dtype='fp32', peer_allreduce=True):
        super().__init__()
        latent_gene_pool = latent_gene_pool if latent_gene_pool is not None else dict(
            dim=128, num_genes_per_island=3, num_selected=2, tournament_size=2)
        world_model = world_model if world_model is not None else dict(
            attn_dim_head=16, heads=4, depth=4, attn_gate_values=True, add_value_residual=True,
            learned_value_residual_mix=True)
        ema_kwargs = ema_kwargs if ema_kwargs is not None else dict(update_model_with_ema_every=1250)
        self.accelerator = accelerator if accelerator is not None else Accelerator()
        device = self.accelerator.device
        if device.type != 'cuda':
            raise ops.L.XtrlError('Agent needs a CUDA device: the hot path has no CPU fallback')

        self.model_dim = hidden_dim
        self.evolutionary, self.evolve_every, self.evolve_after_step = evolutionary, evolve_every, evolve_after_step
        self.gene_pool = LatentGenePool(**latent_gene_pool) if evolutionary else None
        self.cfg = ops.ModelConfig.from_learner_kwargs(
            state_dim, num_actions, reward_range, max_timesteps, world_model, hidden_dim=hidden_dim,
            num_bins=critic_pred_num_bins, dim_gene=self.gene_pool.dim_gene if evolutionary else 0,
            rotary_true_pos=(rollout_positions == 'true'), kv_bf16=(kv_cache_dtype == 'bf16'),
            dense_bf16=(rollout_dense_dtype == 'bf16'), continuous=bool(continuous_actions),
            squash=bool(squash_continuous))
        self.model = WorldModelActorCritic(self.cfg, dropout=dropout,
                                           frac_actor_critic_head_gradient=frac_actor_critic_head_gradient)
        self.loss_cfg = ops.LossConfig(num_actions=num_actions, num_bins=critic_pred_num_bins, state_dim=state_dim,
                                       reward_range=tuple(reward_range), eps_clip=eps_clip, value_clip=value_clip,
                                       entropy_weight=beta_s, actor_weight=actor_loss_weight,
                                       critic_weight=critic_loss_weight, autoreg_weight=autoregressive_loss_weight,
                                       continuous=bool(continuous_actions), squash=bool(squash_continuous))
        self.rsnorm = RSNorm(state_dim + 1)
        self.continuous_actions = continuous_actions
        self.max_grad_norm = max_grad_norm
        self.batch_size, self.epochs, self.lam, self.gamma, self.beta_s = batch_size, epochs, lam, gamma, beta_s
        self.eps_clip, self.value_clip = eps_clip, value_clip
        self.lr, self.betas, self.regen_reg_rate, self.cautious_factor = lr, tuple(betas), regen_reg_rate, cautious_factor
        self.save_path = Path(save_path)
        self.actor_loss_weight, self.critic_loss_weight = actor_loss_weight, critic_loss_weight
        self.autoregressive_loss_weight = autoregressive_loss_weight
        self.register_buffer('step', tensor(0))
        self._step_host = 0
        self.to(device)

        # optimiser / EMA state: flat fp32 buffers parallel to model.flat_params
        n = self.model.flat_params.numel()
        z = lambda: torch.zeros(n, device=device, dtype=torch.float32)
        self.opt_m, self.opt_v, self.opt_param_init = z(), z(), z()
        self.ema_params = self.model.flat_params.clone()      # EMA(model) starts as a deep copy (:747)
        self.ema_schedule = EmaSchedule(beta=ema_decay, **ema_kwargs)
        self.opt_steps = 0
        self._opt_ws = torch.empty(ops.L.load().xtrl_opt_workspace_bytes(len(self.model._offsets) - 1) // 8,
                                   device=device, dtype=torch.float64)
        # gradient all-reduce buffer: [flat grads | RSNorm batch mean] so one collective carries both
        # On several GPUs of one node the buffer lives in a block the other ranks map (distributed.PeerGroup) and the
        # optimiser kernel sums the ranks' buffers itself; where that cannot be set up: torch.distributed.all_reduce.
        self._peers = PeerGroup.create(n + state_dim + 1, torch.device(device)) if peer_allreduce else None
        self._comm = self._peers.local if self._peers is not None else torch.zeros(n + state_dim + 1, device=device, dtype=torch.float32)
        self.model._bind(self.model.flat_params, self._comm[:n])
        self.last_logs = None
        self.keep_last_grads = False
        self.last_grads = None
        self.event_log = None                # bench.py: a list -> (start, graph done, optimiser done) CUDA events per minibatch
        # The bf16 class (the default, see distributed.Accelerator) runs the sequence model as fused tcgen05 kernels
        # (csrc/learn_*.cu); a shape those kernels are not built for is an ERROR there (model.fused_state), never a silent
        # switch of back end.  mixed_precision='no' is the explicit fp32 parity class (the torch expression).
        self.fused_learn = self.accelerator.mixed_precision == 'bf16'
        self.use_cuda_graph_learn = True     # fwd + losses + bwd of a minibatch as one graph replay
        # the flat-gradient all-reduce as the last node of that graph (True) or as one eager NCCL call right after the replay
        # (False, the default: the captured collective hung on the 2-GPU box of this image -- torch 2.11 / NCCL 2.28 -- while
        # the eager call is what round 1 measured 7.9x on 8 GPUs with)
        self.capture_allreduce = False
        self._learn_graphs = {}
        self._learn_src = {}
        # replicas start from rank 0's values, as DDP's wrap does in the reference (:885), whatever each rank's RNG drew
        if self.evolutionary:
            broadcast_from_rank0_(self.model.flat_params, self.gene_pool.genes.data)
        else:
            broadcast_from_rank0_(self.model.flat_params)
        self.ema_params.copy_(self.model.flat_params)

    @property
    def device(self):
        return self.step.device

    # ---- weights for the rollout kernel ------------------------------------------------------
    def ema_state_dict(self):
        return self.model.state_dict_from_flat(self.ema_params)

    def save(self):
        if not self.accelerator.is_main_process:
            return
        torch.save({'model': self.model.state_dict()}, str(self.save_path))

    def load(self):
        if not self.save_path.exists():
            return
        data = torch.load(str(self.save_path), weights_only=True)
        self.model.load_state_dict(data['model'])

    # ---- learn ---------------------------------------------------------------------------------
    def _buffer_from_memories(self, memories, episode_lens, gene_ids):
        """reference format (list of episodes, each a list of Memory with value LOGITS) -> episode-major
        tensors, padded like pad_sequence (:822-837)."""
        dev = self.device
        eps = []
        for episode in memories:
            cols = list(zip(*episode))
            eps.append([torch.stack([torch.as_tensor(v) for v in col]).to(dev) for col in cols])
        pad = lambda i, value=0.: nn.utils.rnn.pad_sequence([e[i] for e in eps], batch_first=True, padding_value=value)
        states, actions, logp, rewards, bd = (pad(i) for i in range(5))
        if eps[0][5].ndim == 2:                    # value LOGITS, zero-padded as in the reference, decoded after (:837, :843)
            values = ops.hl_decode(pad(5).float().contiguous(), self.cfg.reward_range)
        else:                                      # scalar values: pad with what zero logits decode to (the support's midpoint)
            values = pad(5, float(sum(self.cfg.reward_range)) / 2.)
        return dict(states=states.float(), actions=actions.float() if self.continuous_actions else actions.int(),
                    old_log_probs=logp.float(), rewards=rewards.float(),
                    dones=bd.bool(), old_values=values.float(), lens=episode_lens.to(dev).int(),
                    gene_ids=gene_ids.to(dev).int(), time_major=False)

    def learn(self, memories, episode_lens=None, gene_ids=None, fitnesses=None, permutations=None, reward_flags=None):
        """reference signature (:808-814).  `memories` is either the reference's list[list[Memory]] or a
        RolloutBuffer (vectorised rollouts).  permutations / reward_flags: optional replay of the reference's random
        draws for parity tests -- one episode order per epoch (DataLoader shuffle :878) and one reward-conditioning
        Bernoulli per minibatch (:501)."""
        cfg, dev, model = self.cfg, self.device, self.model
        if isinstance(memories, RolloutBuffer):
            buf = memories
            returns_tm = ops.gae(buf.rewards, buf.values, buf.is_boundary, self.gamma, self.lam, time_major=True)
            em = lambda t: t.transpose(0, 1).contiguous()
            data = dict(states=em(buf.states), actions=em(buf.actions), old_log_probs=em(buf.log_probs),
                        rewards=em(buf.rewards), dones=em(buf.is_boundary), old_values=em(buf.values),
                        returns=em(returns_tm), lens=buf.episode_lens, gene_ids=buf.gene_ids)
        else:
            data = self._buffer_from_memories(memories, episode_lens, gene_ids)
            data.pop('time_major')
            data['returns'] = ops.gae(data['rewards'].contiguous(), data['old_values'].contiguous(),
                                      data['dones'].contiguous(), self.gamma, self.lam)
        E, T = data['actions'].shape[:2]
        S = cfg.state_dim
        n = model.flat_params.numel()
        rs_copy = RSNorm(S + 1).to(dev)
        rs_copy.copy_from(self.rsnorm)
        model.train()
        self.last_grads = None
        fields = ('states', 'actions', 'rewards', 'old_log_probs', 'returns', 'old_values', 'dones', 'lens', 'gene_ids')
        reward_flags = iter(reward_flags) if reward_flags is not None else None
        if self.fused_learn:                 # persistent episode-major sources: captured graphs keep reading the same addresses
            src = self._learn_sources(E, T)
            for k in fields:
                src[k].copy_(data[k])
        for epoch in range(self.epochs):
            perm = permutations[epoch].to(dev) if permutations is not None else torch.randperm(E, device=dev)
            for lo in range(0, E, self.batch_size):
                idx = perm[lo:lo + self.batch_size]
                flag = float(torch.rand(()) >= 0.5) if model.training else 1.                        # reward dropout :501
                if reward_flags is not None:
                    flag = float(next(reward_flags))
                lg = self._learn_graph(idx.numel(), T, E) if self.use_cuda_graph_learn else None
                if self.event_log is not None:
                    evs = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                if lg is not None:                       # static minibatch buffers -> one graph replay (fwd + losses + bwd)
                    if self.fused_learn:
                        lg.idx.copy_(idx)                # the prologue kernel gathers inside the graph
                    else:
                        for k in fields:
                            torch.index_select(data[k], 0, idx, out=lg.mb[k])
                    lg.flag.fill_(flag)
                    if self.event_log is not None:
                        evs[0].record()
                    scalars = lg.run()
                elif self.fused_learn:
                    scalars = self._fwd_bwd_prepared(src, idx.long(), torch.full((), flag, device=dev))
                else:
                    scalars = self._fwd_bwd({k: data[k][idx] for k in fields}, flag)
                inv_world = 1. / world_size()
                if self.event_log is not None and self._peers is not None:
                    evs[1].record()
                if self._peers is None:
                    if not self.capture_allreduce:
                        all_reduce_sum_(self._comm)      # ONE collective: gradients (+ RSNorm batch mean in the tail)
                    summed = self._comm
                else:
                    summed = self._peers.reduced         # filled by the optimiser kernel's first phase (csrc/optim.cu)
                # `summed` holds the SUM over ranks of [gradients | RSNorm batch mean]; the mean's 1/world is folded into the
                # optimiser kernel (grad_scale) and the 9-float tail
                if self._peers is None:
                    if self.keep_last_grads:
                        self.last_grads = summed[:n] * inv_world
                    rs_mean = summed[n:] * inv_world
                    if self.event_log is not None:
                        evs[1].record()
                self._optimizer_step(inv_world)
                if self._peers is not None:
                    if self.keep_last_grads:
                        self.last_grads = summed[:n] * inv_world
                    rs_mean = summed[n:] * inv_world
                if self.event_log is not None:
                    evs[2].record()
                    if lg is not None:
                        self.event_log.append(tuple(evs))
                rs_copy.update_(rs_mean)
                self.last_logs = scalars
                if (self.evolutionary and fitnesses is not None and self._step_host > self.evolve_after_step
                        and self._step_host % self.evolve_every == 0):                                # :1009-1015
                    self.gene_pool.evolve_(fitnesses)
                self.accelerator.log(dict(losses=scalars))
        self.rsnorm.copy_from(rs_copy)
        if self._peers is not None:
            self._peers.check()
        self.step.add_(1)
        self._step_host += 1
        return self.epochs * int(data['lens'].sum().item())        # learn-timesteps processed

    def _fwd_bwd(self, mb, flag, collective=True):
        """one minibatch: prep (:903-926) -> sequence model -> fused loss kernel (fwd + grads w.r.t. head outputs) ->
        autograd backward into the flat gradient buffer; the RSNorm batch mean lands in the buffer's tail; then the one
        SUM all-reduce of that buffer across ranks (NCCL, captured with the rest).  No host synchronisation and only
        static shapes, so the whole thing is one CUDA graph (_LearnGraph)."""
        model, cfg = self.model, self.cfg
        S, n = cfg.state_dim, model.flat_params.numel()
        states, actions, rewards, lens = mb['states'], mb['actions'], mb['rewards'], mb['lens']
        T = actions.shape[1]
        mask = torch.arange(T, device=states.device)[None, :] < lens[:, None]                 # :908-909
        if self.continuous_actions:                                                          # :911-916
            prev_actions = F.pad(actions, (0, 0, 1, -1), value=0.)
        else:
            prev_actions = F.pad(actions, (1, -1), value=-1)
        prev_rewards = F.pad(rewards, (1, -1), value=0.)                                     # :918
        packed = self.rsnorm.normalize(torch.cat((states, prev_rewards[..., None]), dim=-1))  # :920-926
        gene = self.gene_pool[mb['gene_ids'].long()] if self.evolutionary else None          # :906
        cast = (lambda t: t) if self.continuous_actions else (lambda t: t.long())
        with torch.autocast('cuda', dtype=torch.bfloat16, enabled=self.accelerator.mixed_precision == 'bf16'):
            ra, vl, spr, dl = model(packed[..., :S], packed[..., S], cast(prev_actions), cast(actions), mask,
                                    latent_gene=gene, reward_flag=flag, fused=self.fused_learn)
        loss, scalars = FusedPPOLoss.apply(self.loss_cfg, ra, vl, spr, dl, actions.contiguous(), mb['old_log_probs'].contiguous(),
                                           mb['returns'].contiguous(), mb['old_values'].contiguous(), packed.contiguous(),
                                           mb['dones'].contiguous(), lens.contiguous(), True)
        loss.backward()
        # RSNorm batch mean over the masked rows rides in the tail of the gradient buffer (:1005, :601)
        mf = mask[..., None].float()
        self._comm[n:] = (packed * mf).sum(dim=(0, 1)) / mf.sum().clamp(min=1.)
        if collective and self.capture_allreduce and self._peers is None:
            all_reduce_sum_(self._comm)              # gradients (+ RSNorm batch mean in the tail), one collective
        return scalars

    def _learn_graph(self, b, T, E):
        """the captured fwd + bwd (+ all-reduce) of a (b, T) minibatch drawn from E episodes"""
        key = (b, T, E)
        if key not in self._learn_graphs:
            self._learn_graphs[key] = _LearnGraph(self, b, T, E)
        return self._learn_graphs[key]

    def _learn_sources(self, E, T):
        """episode-major trajectory tensors of one learning update at fixed addresses (what xtrl_learn_prep_fwd gathers from)"""
        key = (E, T)
        if key not in self._learn_src:
            cfg, dev, A, cont = self.cfg, self.device, self.cfg.num_actions, self.continuous_actions
            z = lambda *s, dt=torch.float32: torch.zeros(*s, device=dev, dtype=dt)
            self._learn_src = {key: dict(                  # one set alive at a time
                states=z(E, T, cfg.state_dim), actions=z(E, T, A) if cont else z(E, T, dt=torch.int32), rewards=z(E, T),
                old_log_probs=z(E, T, A) if cont else z(E, T), returns=z(E, T), old_values=z(E, T), dones=z(E, T, dt=torch.bool),
                lens=z(E, dt=torch.int32), gene_ids=z(E, dt=torch.int32))}
            self._learn_graphs = {k: v for k, v in self._learn_graphs.items() if (k[2], k[1]) == key}
        return self._learn_src[key]

    def _fwd_bwd_prepared(self, src, idx, flag, collective=True):
        """one minibatch of the bf16 class: prologue kernel (gather, shifts, RSNorm, input embeddings :903-926, :492-503) ->
        tcgen05 sequence model -> fused loss kernel -> backward into the flat gradient buffer; the RSNorm batch mean lands in the
        buffer's tail; then the SUM all-reduce.  `flag` is a device scalar."""
        model, cfg = self.model, self.cfg
        S, n = cfg.state_dim, model.flat_params.numel()
        gene_fn = (lambda ids: self.gene_pool[ids]) if self.evolutionary else None                  # :906
        ra, vl, spr, dl, box = model.forward_prepared(src, idx, self.rsnorm.running_mean, self.rsnorm.running_variance, flag, gene_fn)
        loss, scalars = FusedPPOLoss.apply(self.loss_cfg, ra, vl, spr, dl, box['actions'], box['old_log_probs'], box['returns'],
                                           box['old_values'], box['packed'], box['dones'], box['lens'], True)
        loss.backward()
        rs = box['rs_sum']                                                                           # :1005, :601
        self._comm[n:] = rs[:S + 1] / rs[S + 1].clamp(min=1.)
        if collective and self.capture_allreduce and self._peers is None:
            all_reduce_sum_(self._comm)
        return scalars

    def _optimizer_step(self, grad_scale=1.):
        model = self.model
        mode, decay, back = self.ema_schedule.next()
        ops.adopt_atan2_step(model.flat_params, model.flat_grads, self.opt_m, self.opt_v, self.opt_param_init,
                             self.ema_params, model.seg_offsets, lr=self.lr, init_lr=self.lr, betas=self.betas,
                             regen_rate=self.regen_reg_rate, cautious_factor=self.cautious_factor,
                             max_grad_norm=self.max_grad_norm, first_step=(self.opt_steps == 0), ema_mode=mode,
                             ema_decay=decay, copy_ema_to_online=back, workspace=self._opt_ws, grad_scale=grad_scale,
                             peers=self._peers)
        self.opt_steps += 1
        self._comm.zero_()                   # (with peers: the kernel returned only after every rank had read this buffer)

    # ---- deploy --------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, state, reward=None, hiddens=None, latent_gene_id=0):
        """(:1025-1065) one env, one step, ONLINE weights; returns (raw_actions (A,), hiddens)."""
        dev, cfg = self.device, self.cfg
        if hiddens is None:
            st = ops.WorldModelStepper(cfg, 1, dev)
            st.load_state_dict(self.model.state_dict())
            hiddens = Hiddens(st, self.opt_steps)
        st = hiddens.stepper
        if hiddens.weights_version != self.opt_steps:          # the online weights moved on since this cache was created
            st.load_state_dict(self.model.state_dict())
            hiddens.weights_version = self.opt_steps
        state = torch.as_tensor(state, dtype=torch.float32).to(dev).reshape(1, cfg.state_dim).contiguous()
        has_reward = reward is not None
        prev_r = torch.as_tensor(float(reward) if has_reward else 0., dtype=torch.float32, device=dev).reshape(1)
        no_action = None if self.continuous_actions else torch.full((1,), -1, dtype=torch.int32, device=dev)
        out = st.step(state, no_action, prev_r, self.rsnorm.running_mean,
                      self.rsnorm.running_variance, genes=self.gene_pool.genes.data if self.evolutionary else None,
                      gene_ids=torch.full((1,), int(latent_gene_id), dtype=torch.int32, device=dev) if self.evolutionary else None,
                      reward_flag=1. if has_reward else 0., want_logits=True)
        return out['raw_actions'][0].clone(), hiddens
