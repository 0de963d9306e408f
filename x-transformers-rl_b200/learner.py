"""Learner with the reference's constructor and call signature
(x_transformers_rl/x_transformers_rl.py:1069-1380):

    learner = Learner(state_dim, num_actions, reward_range, max_timesteps=..., world_model=dict(...))
    learner(env, num_learning_updates)

The rollout loop is re-architected: instead of batch-1 episodes with ~10 host syncs per step, all
episodes of an update advance in lock step, one xtrl_wm_step launch per timestep, trajectories stay
on the device.  A vectorised env (`num_envs` attribute, see envs.py) is stepped B-wide; a plain env
keeps the reference's one-episode-at-a-time semantics through the same kernel with B = 1.
"""
from __future__ import annotations

import weakref

import numpy as np
import torch
from torch import nn

from . import ops
from .agent import Agent, Memory, RolloutBuffer  # noqa: F401
from .distributed import Accelerator, reduce_fitness, shard_episode_genes, sync_seed


def _to_device_state(x, device, pinned=None):
    if torch.is_tensor(x):
        return x.to(device=device, dtype=torch.float32, non_blocking=True)
    x = np.asarray(x, dtype=np.float32)
    if pinned is not None:
        pinned.copy_(torch.from_numpy(x).reshape(pinned.shape))
        return pinned.to(device, non_blocking=True)
    return torch.from_numpy(x).to(device)


class Learner(nn.Module):
    def __init__(self, state_dim, num_actions, reward_range, world_model: dict, continuous_actions=False,
                 squash_continuous=True, continuous_actions_clamp=None, evolutionary=False, evolve_every=10,
                 evolve_after_step=20, latent_gene_pool=None, max_timesteps=500, batch_size=8,
                 num_episodes_per_update=64, lr=0.0008, betas=(0.9, 0.99), lam=0.95, gamma=0.99, eps_clip=0.2,
                 value_clip=0.4, beta_s=.01, regen_reg_rate=1e-4, cautious_factor=0.1, epochs=4, ema_decay=0.9,
                 save_every=100, frac_actor_critic_head_gradient=0.5, accelerate_kwargs: dict = dict(),
                 agent_kwargs: dict = dict()):
        super().__init__()
        assert num_episodes_per_update % batch_size == 0
        self.accelerator = Accelerator(**accelerate_kwargs)
        self.agent = Agent(state_dim=state_dim, num_actions=num_actions, continuous_actions=continuous_actions,
                           squash_continuous=squash_continuous, reward_range=reward_range, world_model=world_model,
                           evolutionary=evolutionary, evolve_every=evolve_every, evolve_after_step=evolve_after_step,
                           latent_gene_pool=latent_gene_pool, epochs=epochs, max_timesteps=max_timesteps,
                           batch_size=batch_size, lr=lr, betas=betas, lam=lam, gamma=gamma, beta_s=beta_s,
                           regen_reg_rate=regen_reg_rate, cautious_factor=cautious_factor, eps_clip=eps_clip,
                           value_clip=value_clip, ema_decay=ema_decay,
                           frac_actor_critic_head_gradient=frac_actor_critic_head_gradient,
                           accelerator=self.accelerator, **agent_kwargs)
        self.num_episodes_per_update, self.max_timesteps = num_episodes_per_update, max_timesteps
        num_genes = self.agent.gene_pool.num_genes if self.agent.gene_pool is not None else 1
        self.episode_genes_for_process = shard_episode_genes(num_episodes_per_update, num_genes,
                                                             self.accelerator.num_processes,
                                                             self.accelerator.process_index)
        self.num_actions, self.continuous_actions = num_actions, continuous_actions
        self.continuous_actions_clamp = continuous_actions_clamp
        self.save_every = save_every
        self._steppers = {}
        self.use_cuda_graph = True       # device envs that declare cuda_graph_safe: the T-step rollout is one graph replay
        self._graphs = {}
        # host-memory envs that can split(): groups of envs stepped on separate streams through the raw-pointer loop
        self.rollout_groups = 2
        self.rollout_group_epc = 'share'     # 'share' the CTA slots between the groups' concurrent kernels, or 'fill' per call
        self._group_res = {}
        self.timings = dict(rollout_s=0., learn_s=0., env_steps=0, learn_timesteps=0)

    @property
    def device(self):
        return self.accelerator.device

    def _stepper(self, num_envs, max_timesteps, group=None):
        key = (num_envs, max_timesteps, group)
        if group is not None:                       # pipelined groups: one stepper (KV cache) per group, kept together
            if key not in self._steppers:
                if any(k[2] is None for k in self._steppers):
                    self._steppers = {}
                self._steppers[key] = ops.WorldModelStepper(self.agent.cfg, num_envs, self.device)
            return self._steppers[key]
        if key not in self._steppers:
            cfg = self.agent.cfg
            if max_timesteps > cfg.max_timesteps:
                raise ops.L.XtrlError(f'max_timesteps={max_timesteps} exceeds the KV-cache capacity {cfg.max_timesteps}')
            self._steppers = {key: ops.WorldModelStepper(cfg, num_envs, self.device)}     # one live KV cache at a time
        return self._steppers[key]

    # ---- vectorised rollout: B envs in lock step, one kernel launch per timestep -------------------
    @torch.no_grad()
    def rollout_vectorized(self, env, gene_ids, max_timesteps, seed=None, uniforms=None, allow_graph=True):
        """returns (RolloutBuffer (T,B,..), cumulative_rewards (B,)).  Sampling consumes `uniforms`
        (T,B) in [0,1) -- drawn here from torch's CUDA generator unless given (parity tests)."""
        agent, dev, cfg = self.agent, self.device, self.agent.cfg
        B, T, S = env.num_envs, max_timesteps, cfg.state_dim
        cont = cfg.continuous
        if (not cont and self.rollout_groups >= 1 and hasattr(env, 'split') and getattr(env, 'never_terminates', False)
                and B % self.rollout_groups == 0 and B // self.rollout_groups >= 256):
            return self._rollout_pipelined(env, gene_ids, max_timesteps, seed, uniforms)
        if (not cont and self.use_cuda_graph and allow_graph and uniforms is None and getattr(env, 'cuda_graph_safe', False)
                and getattr(env, 'never_terminates', False)):
            return self._rollout_graphed(env, gene_ids, max_timesteps, seed)
        st = self._stepper(B, max_timesteps)
        st.load_state_dict(agent.ema_state_dict())               # rollouts read the EMA weights (:1194)
        A = cfg.num_actions
        buf = RolloutBuffer.allocate(T, B, S, dev, continuous_dim=A if cont else 0)
        buf.gene_ids.copy_(gene_ids)
        # the random inputs of the sampler: uniforms (inverse CDF) or, for continuous actions, standard normals
        if uniforms is not None:
            u = uniforms
        else:
            u = torch.randn(T, B, A, device=dev) if cont else torch.rand(T, B, device=dev)
        first_a = torch.zeros(B, A, device=dev) if cont else torch.full((B,), -1, dtype=torch.int32, device=dev)   # :1241-1244
        first_r = torch.zeros(B, device=dev)                              # :1246
        pin_s = pin_r = None
        state = env.reset(**({} if seed is None else dict(seed=seed)))
        if isinstance(state, tuple):
            state = state[0]
        if not torch.is_tensor(state):
            pin_s = torch.empty(B, S, dtype=torch.float32).pin_memory()
            pin_r = torch.empty(B, dtype=torch.float32).pin_memory()
        never_done = getattr(env, 'never_terminates', False)
        alive = torch.ones(B, dtype=torch.bool, device=dev)
        lens = torch.zeros(B, dtype=torch.int32, device=dev)
        genes = agent.gene_pool.genes.data if agent.evolutionary else None
        rs_mean, rs_var = agent.rsnorm.running_mean, agent.rsnorm.running_variance
        pad_value = float(sum(cfg.reward_range)) / 2.     # decode of all-zero logits = mean of the bin centres
        for t in range(T):
            buf.states[t].copy_(_to_device_state(state, dev, pin_s), non_blocking=True)
            out = dict(action=buf.actions[t], log_prob=buf.log_probs[t], value=buf.values[t])
            st.step(buf.states[t], first_a if t == 0 else buf.actions[t - 1], first_r if t == 0 else buf.rewards[t - 1],
                    rs_mean, rs_var, uniform=None if cont else u[t], normal=u[t] if cont else None, genes=genes,
                    gene_ids=buf.gene_ids if genes is not None else None, out=out, pos_uniform=t)
            action = buf.actions[t]
            if cont and self.continuous_actions_clamp is not None:      # environment clamping (:1291-1295)
                action.clamp_(*self.continuous_actions_clamp)
            step_out = env.step(action if pin_s is None else action.cpu().numpy())     # host env: D2H sync
            if len(step_out) >= 4:
                state, reward, terminated, truncated = step_out[:4]
            elif len(step_out) == 3:
                state, reward, terminated = step_out
                truncated = None
            else:
                raise RuntimeError('invalid number of returns from environment .step')
            if torch.is_tensor(reward):
                buf.rewards[t].copy_(reward, non_blocking=True)
            else:
                pin_r.copy_(torch.from_numpy(np.asarray(reward, dtype=np.float32).reshape(B)))
                buf.rewards[t].copy_(pin_r, non_blocking=True)
            if not never_done:
                term = torch.as_tensor(terminated, device=dev).bool()
                done = term if truncated is None else term | torch.as_tensor(truncated, device=dev).bool()
                buf.is_boundary[t] = term & alive
                # finished envs keep stepping (lock step) but their rows become pad_sequence padding (:837)
                dead = ~alive
                dead_a = dead[:, None] if cont else dead
                buf.states[t].masked_fill_(dead[:, None], 0.); buf.actions[t].masked_fill_(dead_a, 0)
                buf.log_probs[t].masked_fill_(dead_a, 0.); buf.rewards[t].masked_fill_(dead, 0.)
                buf.values[t].masked_fill_(dead, pad_value)
                lens += alive.int()
                alive = alive & ~done
        if never_done:
            lens.fill_(T)
        buf.episode_lens.copy_(lens)
        return buf, buf.rewards.sum(dim=0)

    @torch.no_grad()
    def _rollout_graphed(self, env, gene_ids, max_timesteps, seed=None):
        """device-resident, never-terminating env whose step() is pure views: the whole rollout -- T x (state copy,
        xtrl_wm_step, reward copy) -- is captured ONCE into a CUDA graph over static buffers (trajectory buffer,
        KV cache, packed weights, RSNorm stats, uniforms) and replayed per update: one launch from the host
        instead of 3T.  Everything that changes between updates is refreshed in place before the replay."""
        agent, dev, cfg = self.agent, self.device, self.agent.cfg
        B, T, S = env.num_envs, max_timesteps, cfg.state_dim
        key = (id(env), B, T)
        st = self._stepper(B, T)
        st.load_state_dict(agent.ema_state_dict())               # in place: the captured pointers stay valid
        state0 = env.reset(**({} if seed is None else dict(seed=seed)))
        ent = self._graphs.get(key)
        if ent is None or ent['st'] is not st or ent['env']() is not env:       # id() can be recycled: check identity
            buf = RolloutBuffer.allocate(T, B, S, dev)
            buf.episode_lens.fill_(T)
            u = torch.empty(T, B, device=dev)
            first_a = torch.full((B,), -1, dtype=torch.int32, device=dev)
            first_r = torch.zeros(B, device=dev)
            genes = agent.gene_pool.genes.data if agent.evolutionary else None
            rs_mean, rs_var = agent.rsnorm.running_mean, agent.rsnorm.running_variance

            def body():
                state = state0
                for t in range(T):
                    buf.states[t].copy_(state)
                    st.step(buf.states[t], first_a if t == 0 else buf.actions[t - 1], first_r if t == 0 else buf.rewards[t - 1],
                            rs_mean, rs_var, uniform=u[t], genes=genes, gene_ids=buf.gene_ids if genes is not None else None,
                            out=dict(action=buf.actions[t], log_prob=buf.log_probs[t], value=buf.values[t]), pos_uniform=t)
                    state, reward = env.step(buf.actions[t])[:2]
                    buf.rewards[t].copy_(reward)
                env._t = 0

            buf.gene_ids.copy_(gene_ids)
            u.uniform_()
            side = torch.cuda.Stream(dev)                        # one eager pass first: lazy module loads cannot be captured
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                body()
            torch.cuda.current_stream(dev).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                body()
            # every tensor whose address the graph baked in must outlive the graph
            ent = self._graphs[key] = dict(graph=graph, buf=buf, u=u, st=st, env=weakref.ref(env), keep=(first_a, first_r, genes, rs_mean, rs_var, state0))
            if len(self._graphs) > 2:
                self._graphs.pop(next(iter(self._graphs)))
        buf = ent['buf']
        buf.gene_ids.copy_(gene_ids)
        ent['u'].uniform_()
        ent['graph'].replay()
        return buf, buf.rewards.sum(dim=0)

    @torch.no_grad()
    def _rollout_pipelined(self, env, gene_ids, max_timesteps, seed=None, uniforms=None):
        """host-memory vector env split into groups: while the host steps group A's simulators and moves its
        states over PCIe, the GPU runs group B's timestep (one stream, one KV cache, one pinned staging set per
        group).  Same buffers and results as rollout_vectorized; never-terminating envs only."""
        agent, dev, cfg = self.agent, self.device, self.agent.cfg
        B, T, S, G = env.num_envs, max_timesteps, cfg.state_dim, self.rollout_groups
        Bg = B // G
        env.reset(**({} if seed is None else dict(seed=seed)))
        subs = env.split(G)
        buf = RolloutBuffer.allocate(T, B, S, dev)
        buf.gene_ids.copy_(gene_ids)
        u = torch.rand(T, B, device=dev) if uniforms is None else uniforms
        first_a = torch.full((Bg,), -1, dtype=torch.int32, device=dev)
        first_r = torch.zeros(Bg, device=dev)
        genes = agent.gene_pool.genes.data if agent.evolutionary else None
        rs_mean, rs_var = agent.rsnorm.running_mean, agent.rsnorm.running_variance
        sd = agent.ema_state_dict()
        main = torch.cuda.current_stream(dev)
        grp = []
        for g, sub in enumerate(subs):
            st = self._stepper(Bg, T, group=g)
            st.load_state_dict(sd)
            # streams, events and pinned staging live on the Learner: raw cudaMemcpyAsync calls are invisible to torch's
            # caching host allocator, so these blocks must never go back to it while a copy may be in flight
            res = self._group_res.get((g, Bg, S))
            if res is None:
                res = self._group_res[(g, Bg, S)] = dict(
                    stream=torch.cuda.Stream(dev), ev=torch.cuda.Event(), pin_s=torch.empty(Bg, S).pin_memory(),
                    pin_r=torch.empty(Bg).pin_memory(), pin_a=torch.empty(Bg, dtype=torch.int32).pin_memory())
            grp.append(dict(env=sub, st=st, lo=g * Bg, hi=(g + 1) * Bg, state=sub.reset(), **res))
        for q in grp:
            q['stream'].wait_stream(main)
        # Hot loop on raw pointers: the per-timestep host work (2 x G times per timestep) is a numpy copy into pinned
        # memory, three cudaMemcpyAsync, one xtrl_wm_step and one event record -- no tensor views, no marshalling.
        import ctypes as C
        rt = ops.L.cudart()
        lib = ops.L.load()
        H2D, D2H = 1, 2
        f4 = 4
        for q in grp:
            lo, st = q['lo'], q['st']
            q['sid'] = C.c_void_p(q['stream'].cuda_stream)
            q['io'] = ops.L.StepIO()
            io = q['io']
            io.rs_mean, io.rs_var = rs_mean.data_ptr(), rs_var.data_ptr()
            io.genes = genes.data_ptr() if genes is not None else None
            io.gene_ids = buf.gene_ids.data_ptr() + lo * 4 if genes is not None else None
            io.reward_flag, io.tc_weights = 1., (st.tc_weights.data_ptr() if st.tc_weights is not None else None)
            # the groups' kernels run concurrently: size each grid as if the whole batch were one call, so that a group
            # takes its share of the CTA slots instead of paying the per-CTA fixed cost twice
            io.envs_per_cta = min(16, max(1, -(-B // (2 * lib.xtrl_device_num_sms())))) if self.rollout_group_epc == 'share' else 0
            q['p'] = dict(states=buf.states.data_ptr() + lo * S * f4, actions=buf.actions.data_ptr() + lo * 4,
                          logp=buf.log_probs.data_ptr() + lo * f4, values=buf.values.data_ptr() + lo * f4,
                          rewards=buf.rewards.data_ptr() + lo * f4, u=u.data_ptr() + lo * f4,
                          pin_s=q['pin_s'].data_ptr(), pin_r=q['pin_r'].data_ptr(), pin_a=q['pin_a'].data_ptr(),
                          first_a=first_a.data_ptr(), first_r=first_r.data_ptr())
            q['np_s'], q['np_r'], q['np_a'] = q['pin_s'].numpy(), q['pin_r'].numpy(), q['pin_a'].numpy()
            q['desc'], q['w'], q['kv'] = C.byref(st.desc), C.c_void_p(st.weights.data_ptr()), C.c_void_p(st.kv_cache.data_ptr())
        row_s, row = B * S * f4, B * f4                               # byte strides of one timestep in the (T, B, .) buffers
        for t in range(T + 1):
            for q in grp:
                p, sid = q['p'], q['sid']
                if t > 0:
                    q['ev'].synchronize()                              # actions of step t-1 are on the host
                    step_out = q['env'].step(q['np_a'])
                    q['state'] = step_out[0]
                    np.copyto(q['np_r'], step_out[1], casting='same_kind')
                    rt.cudaMemcpyAsync(C.c_void_p(p['rewards'] + (t - 1) * row), C.c_void_p(p['pin_r']), Bg * f4, H2D, sid)
                if t == T:
                    continue
                np.copyto(q['np_s'], q['state'], casting='same_kind')
                rt.cudaMemcpyAsync(C.c_void_p(p['states'] + t * row_s), C.c_void_p(p['pin_s']), Bg * S * f4, H2D, sid)
                io = q['io']
                io.state = p['states'] + t * row_s
                io.prev_action = p['first_a'] if t == 0 else p['actions'] + (t - 1) * row
                io.prev_reward = p['first_r'] if t == 0 else p['rewards'] + (t - 1) * row
                io.uniform, io.pos_uniform = p['u'] + t * row, t
                io.action, io.log_prob, io.value = p['actions'] + t * row, p['logp'] + t * row, p['values'] + t * row
                ops.L.check(lib.xtrl_wm_step(q['desc'], q['w'], q['kv'], C.byref(io), Bg, sid), 'xtrl_wm_step')
                rt.cudaMemcpyAsync(C.c_void_p(p['pin_a']), C.c_void_p(p['actions'] + t * row), Bg * 4, D2H, sid)
                q['ev'].record(q['stream'])
        for q in grp:
            main.wait_stream(q['stream'])
            q['stream'].synchronize()             # the last H2D out of the pinned blocks is done before the host may touch them again
        buf.episode_lens.fill_(T)
        return buf, buf.rewards.sum(dim=0)

    # ---- plain env: reference semantics, one episode at a time (:1220-1355) -------------------------
    @torch.no_grad()
    def rollout_episode(self, env, gene_id, max_timesteps, reset_kwargs):
        agent, dev, cfg = self.agent, self.device, self.agent.cfg
        st = self._stepper(1, max_timesteps)
        st.reset()
        reset_out = env.reset(**reset_kwargs)
        state = reset_out[0] if isinstance(reset_out, tuple) else reset_out
        genes = agent.gene_pool.genes.data if agent.evolutionary else None
        gid = torch.full((1,), int(gene_id), dtype=torch.int32, device=dev)
        cont, A = cfg.continuous, cfg.num_actions
        prev_a = torch.zeros(1, A, device=dev) if cont else torch.full((1,), -1, dtype=torch.int32, device=dev)
        prev_r = torch.zeros(1, device=dev)
        memories, cumulative = [], 0.
        for timestep in range(max_timesteps):
            s = torch.as_tensor(np.asarray(state, dtype=np.float32) if not torch.is_tensor(state) else state,
                                dtype=torch.float32).to(dev).reshape(1, cfg.state_dim).contiguous()
            out = st.step(s, prev_a, prev_r, agent.rsnorm.running_mean, agent.rsnorm.running_variance,
                          uniform=None if cont else torch.rand(1, device=dev), normal=torch.randn(1, A, device=dev) if cont else None,
                          genes=genes, gene_ids=gid if genes is not None else None)
            action = out['action'].clone()
            if cont and self.continuous_actions_clamp is not None:
                action.clamp_(*self.continuous_actions_clamp)
            step_out = env.step(action.tolist()[0])
            if len(step_out) >= 4:
                state, reward, terminated, truncated = step_out[:4]
            elif len(step_out) == 3:
                state, reward, terminated = step_out
                truncated = False
            else:
                raise RuntimeError('invalid number of returns from environment .step')
            reward = float(np.asarray(reward, dtype=np.float64).reshape(-1)[0])      # accepts shape-(1,) arrays (README.md:21)
            cumulative += reward
            memories.append(Memory(s[0], action[0], out['log_prob'][0].clone(), torch.tensor(reward, device=dev),
                                   torch.tensor(bool(terminated), device=dev), out['value'][0].clone()))
            prev_a, prev_r = action, torch.full((1,), reward, device=dev)
            if terminated or truncated:
                break
        return memories, timestep + 1, cumulative

    def forward(self, env, num_learning_updates: int, seed=None, max_timesteps=None):
        import time
        max_timesteps = max_timesteps if max_timesteps is not None else self.max_timesteps
        agent, dev = self.agent, self.device
        if seed is not None:
            torch.manual_seed(seed)
            np.random.seed(seed)
        agent.eval()
        vectorized = hasattr(env, 'num_envs')
        pairs = self.episode_genes_for_process
        for learning_update in range(num_learning_updates):
            fitnesses, episode_seeds = None, None
            if agent.evolutionary:
                fitnesses = torch.zeros(agent.gene_pool.num_genes, device=dev)
                torch.manual_seed(sync_seed(dev))                                              # :1211-1212
                episode_seeds = torch.randint(0, int(1e7), (self.num_episodes_per_update,))
            torch.cuda.synchronize(dev); t0 = time.perf_counter()
            if vectorized:
                B = env.num_envs
                if len(pairs) % B != 0:
                    raise ValueError(f'{len(pairs)} (episode, gene) pairs on this rank are not a multiple of env.num_envs={B}')
                bufs = []
                for w in range(0, len(pairs), B):
                    wave = pairs[w:w + B]
                    gids = torch.tensor([p[1] for p in wave], dtype=torch.int32, device=dev)
                    wave_seed = int(episode_seeds[wave[0][0]]) if episode_seeds is not None else None
                    buf, cum = self.rollout_vectorized(env, gids, max_timesteps, seed=wave_seed, allow_graph=(len(pairs) == B))
                    if fitnesses is not None:
                        fitnesses.index_add_(0, gids.long(), cum)                              # :1345-1346
                    bufs.append(buf)
                memories = RolloutBuffer.cat(bufs)
                lens_t = gids_t = None
                steps = len(pairs) * max_timesteps
            else:
                memories, lens, gids = [], [], []
                self._stepper(1, max_timesteps).load_state_dict(agent.ema_state_dict())   # EMA weights (:1194)
                for episode, gene_id in pairs:
                    kw = dict(seed=int(episode_seeds[episode])) if episode_seeds is not None else {}
                    mem, n, cum = self.rollout_episode(env, gene_id, max_timesteps, kw)
                    if fitnesses is not None:
                        fitnesses[gene_id] += cum
                    memories.append(mem); lens.append(n); gids.append(gene_id)
                lens_t, gids_t = torch.tensor(lens, device=dev), torch.tensor(gids, device=dev)
                steps = sum(lens)
            self.accelerator.wait_for_everyone()                                               # :1359
            if fitnesses is not None:
                fitnesses = reduce_fitness(fitnesses)                                          # :1362
            torch.cuda.synchronize(dev); t1 = time.perf_counter()
            learned = agent.learn(memories, lens_t, gids_t, fitnesses)                         # :1364
            torch.cuda.synchronize(dev); t2 = time.perf_counter()
            self.timings['rollout_s'] += t1 - t0; self.timings['learn_s'] += t2 - t1
            self.timings['env_steps'] += steps; self.timings['learn_timesteps'] += learned
            if learning_update % self.save_every == 0:
                agent.save()
        agent.save()
        if self.accelerator.is_main_process:
            print('training complete')
