"""Batched rollout / training driver: train_fortattack.train + Learner for E environments at once.

The reference steps ONE env per process (train_fortattack.py:24-25,51-104) and keeps one RolloutStorage
per agent (rlagent.py:13).  This driver keeps the same protocol and the same objects -- one
RolloutStorage per agent with num_processes = E, one MPNN + JointPPO per team -- but
  * all buffers live in shared device blocks [T+1, A, E, ...]; each agent's RolloutStorage is a view,
    so the step kernel writes observations and rewards of step t straight into slot t+1 / t of every
    agent's storage (no per-agent copies: learner.py:239-243 does A host->device copies per step);
  * episode ends are per env: the kernel auto-resets finished envs, the returned row is the new episode's
    first observation (what initialize_new_episode stores, rlagent.py:28-31) and `ends[t+1, e]` records
    the end point (train_fortattack.py:98);
  * GAE runs once over all envs with those per-env end points (RolloutStorage.compute_returns_batched ==
    Learner.wrap_horizon per env, learner.py:191-211);
  * multi-GPU: every rank owns the env shard [rank*E, (rank+1)*E); rollouts need no communication, the
    PPO step all-reduces one flat gradient buffer per optimizer step (JointPPO(process_group=...)).
Checkpoints use the reference's file format ({'models': [state_dict per agent], 'ob_rms': (None, None)},
train_fortattack.py:121-128) and its loading rule (models[0] -> guards, models[-1] -> attackers,
learner.py:245-249).
"""

import torch

from . import _capi
from .batched_env import FortAttackBatch
from .mpnn import MPNN
from .policy_kernel import FusedPolicy, MODE_ARGMAX, MODE_SAMPLE, forward_ensemble
from .rlcore.algo import JointPPO
from .rlcore.storage import RolloutStorage


class _Shape(object):
    def __init__(self, *shape):
        self.shape = shape


class SharedRollouts(object):
    """A RolloutStorage per agent, all views of shared [T(+1), A, E, ...] blocks."""

    def __init__(self, num_steps, n_agents, n_envs, device):
        T, A, E = num_steps, n_agents, n_envs
        z = lambda *s, **k: torch.zeros(*s, device=device, **k)
        self.obs = z(T + 1, A, E, 6)
        self.rewards = z(T, A, E)
        self.value_preds = z(T + 1, A, E, 1)
        self.returns = z(T + 1, A, E, 1)
        self.action_log_probs = z(T, A, E, 1)
        self.actions = z(T, A, E, 1, dtype=torch.long)
        self.actions_i32 = z(T, A, E, dtype=torch.int32)            # what the step kernel reads
        self.masks = torch.ones(T + 1, A, E, 1, device=device)
        self.hidden = z(T + 1, A, E, 1)
        self.ends = z(T + 1, E, dtype=torch.bool)                   # per-env end points (train_fortattack.py:98)
        self.done = z(T, E, dtype=torch.uint8)
        self.result = z(T, E, dtype=torch.uint8)
        self.agents = []
        for i in range(A):
            r = RolloutStorage.__new__(RolloutStorage)
            r.obs, r.rewards = self.obs[:, i], self.rewards[:, i].unsqueeze(-1)
            r.value_preds, r.returns = self.value_preds[:, i], self.returns[:, i]
            r.action_log_probs, r.actions = self.action_log_probs[:, i], self.actions[:, i]
            r.masks, r.recurrent_hidden_states = self.masks[:, i], self.hidden[:, i]
            r.num_steps, r.step = T, 0
            self.agents.append(r)

    def bookkeeping(self, step, episode_rewards):
        """masks[step+1], ends[step+1] and episode_rewards from obs[step], obs[step+1], done[step], rewards[step]
        (rl_rollout_bookkeeping, include/fortattack_rollout.h)."""
        L = _capi.lib()
        T, A, E = self.rewards.shape
        _capi.check(L.rl_rollout_bookkeeping(self.obs[step].data_ptr(), self.obs[step + 1].data_ptr(), self.done[step].data_ptr(),
                                             self.rewards[step].data_ptr(), self.masks[step + 1].data_ptr(),
                                             self.ends[step + 1].data_ptr(), episode_rewards.data_ptr(), A, E,
                                             torch.cuda.current_stream(self.rewards.device).cuda_stream))

    def compute_returns(self, next_value, gamma, tau):
        """Learner.wrap_horizon for every agent and env in ONE launch (rl_gae, include/fortattack_rollout.h):
        next_value float [A, E]; per-env episode boundaries from `ends`."""
        L = _capi.lib()
        T, A, E = self.rewards.shape
        nv = next_value.to(torch.float32).contiguous()
        assert nv.shape == (A, E) and self.ends.dtype == torch.bool
        _capi.check(L.rl_gae(self.rewards.data_ptr(), self.value_preds.data_ptr(), nv.data_ptr(), self.masks.data_ptr(),
                             self.ends.data_ptr(), self.returns.data_ptr(), T, A, E, float(gamma), float(tau),
                             torch.cuda.current_stream(self.rewards.device).cuda_stream))

    def after_update(self):
        # RolloutStorage.after_update for every agent at once (storage.py:51-56)
        self.obs[0].copy_(self.obs[-1])
        self.obs[1:] = 0
        self.masks[0].copy_(self.masks[-1])
        self.ends.zero_()


class BatchedTrainer(object):
    def __init__(self, n_envs, n_guards=3, n_attackers=3, num_steps=128, max_episode_steps=100, device="cuda:0",
                 seed=0, env_id0=0, hidden_dim=128, gamma=0.99, tau=0.95, clip_param=0.2, ppo_epoch=4,
                 num_mini_batch=32, value_loss_coef=0.5, entropy_coef=0.01, lr=1e-4, max_grad_norm=0.5,
                 use_clipped_value_loss=True, process_group=None, fused_policy="auto", allow_tf32=False,
                 graph_rollouts=True, attacker_ensemble=None, fused_update=True, graph_update=None, exact_old="auto",
                 overlap_teams=True):
        self.device = torch.device(device)
        # graph_update: replay the optimizer steps from CUDA graphs (JointPPO(graph_update=True)); an eager step is bound by the
        # host's ~140 launches.  Default: on for a single-rank CUDA trainer.  With a process group it stays opt-in, because a
        # captured step then holds the group's NCCL kernels and dist.destroy_process_group() waits for ever unless
        # release_graphs() ran first (seen as a teardown hang, profiles/r4g_dist_train_2gpu.log) -- a caller who asks for it
        # (as bench.py does) also gets the joint two-team step, and owes the trainer a release_graphs() before teardown.
        if graph_update is None:
            graph_update = self.device.type == "cuda" and bool(fused_update) and process_group is None
        self.overlap_teams = overlap_teams if overlap_teams == "joint" else bool(overlap_teams)
        self.E, self.ng, self.na, self.T = n_envs, n_guards, n_attackers, num_steps
        self.A = n_guards + n_attackers
        self.gamma, self.tau = gamma, tau
        self.env = FortAttackBatch(n_envs, n_guards, n_attackers, max_steps=max_episode_steps, seed=seed,
                                   env_id0=env_id0, device=self.device)
        mk = lambda n, m: MPNN(action_space=_Shape(8), num_agents=n, num_opp_agents=m, num_entities=0,
                               input_size=6, hidden_dim=hidden_dim, pos_index=2).to(self.device)
        self.policies = [mk(n_guards, n_attackers), mk(n_attackers, n_guards)]     # guards first (learner.py:57-69)
        self.trainers = [JointPPO(p, clip_param, ppo_epoch, num_mini_batch, value_loss_coef, entropy_coef, lr=lr,
                                  max_grad_norm=max_grad_norm, use_clipped_value_loss=use_clipped_value_loss,
                                  process_group=process_group, allow_tf32=allow_tf32, graph_update=graph_update)
                         for p in self.policies]
        self.process_group = process_group
        for pol in self.policies:                     # training forward: attention kernels instead of bmm chains (mpnn.py)
            pol.fused_attention = bool(fused_update) and self.device.type == "cuda"
        self.fused_update = bool(fused_update)      # minibatch gather + clipped-PPO loss kernels (rlcore/fused.py)
        if process_group is not None:                    # replicas start from rank 0's weights
            import torch.distributed as dist
            for p in self.policies:
                for t in p.parameters():
                    dist.broadcast(t.data, src=dist.get_global_rank(process_group, 0), group=process_group)
        # rollout-time forwards: the fused tcgen05 kernel (csrc/mp_policy.cu) when the network has the reference's
        # shape (hidden 128); the torch module otherwise (small test networks).  Training always uses the module.
        if fused_policy == "auto":
            fused_policy = hidden_dim == 128
        self.fused = [FusedPolicy(p, seed=seed * 2 + t, env_id0=env_id0) for t, p in enumerate(self.policies)] \
            if fused_policy else None
        # exact_old: the rollout kernel computes with fp16 tensor-core operands; with trained weights its stored
        # log-probs / values differ from the update's fp32-grade forward by up to ~1e-2, so the PPO ratio would start at
        # 1 +- 1e-2 instead of exactly 1 (rlcore/algo/ppo.py:163-173).  When on (default whenever the fused rollout
        # kernel is used) train_once() re-evaluates the stored (obs, action) rows with the update's own forward before
        # GAE, the way RL stacks with a separate inference engine recompute the behaviour log-probs in the trainer.
        self.exact_old = (self.fused is not None) if exact_old == "auto" else bool(exact_old)
        # Ensemble play (train_fortattack_v2.py:34-35,110-111; Learner.sample_attacker / select_attacker,
        # learner.py:119-140): K frozen attacker checkpoints; every env draws one uniformly at each of its episode
        # starts.  One forward per checkpoint over all envs, each writing only the envs assigned to it.
        self.ensemble = None
        if attacker_ensemble is not None:
            if not fused_policy:
                raise ValueError("attacker_ensemble needs the fused policy kernel (hidden_dim 128)")
            self.ensemble = []
            for k, sd in enumerate(attacker_ensemble):
                pol = mk(n_attackers, n_guards)
                pol.load_state_dict(sd)
                pol.requires_grad_(False)
                self.ensemble.append(FusedPolicy(pol, seed=seed * 2 + 1, env_id0=env_id0))
            for f in self.ensemble[1:]:                    # one sampling stream for the attacker team
                f.counter = self.ensemble[0].counter
            K = len(self.ensemble)
            self.att_id = torch.randint(0, K, (n_envs,), device=self.device, dtype=torch.int32)
            # per checkpoint: episodes ended by [attackers all dead, time limit, fort reached] (world.gameResult),
            # and the sums needed for the reference's test_fortattack_v2.py table (:94-124)
            self.ensemble_results = torch.zeros(K, 4, device=self.device, dtype=torch.int64)
            # sums behind the reference's evaluation table (test_fortattack_v2.py:94-124), per checkpoint:
            # episodes, alive guards / attackers when the episode ended, mean per-agent return of guards / attackers
            self.ensemble_sums = torch.zeros(K, 5, device=self.device, dtype=torch.float64)
            self.alive_end = torch.zeros(n_envs, dtype=torch.uint8, device=self.device)
            self.env.set_alive_end_buffer(self.alive_end)
            self.ep_return = torch.zeros(self.A, n_envs, device=self.device)
        # the T-step collection loop is captured into ONE CUDA graph on its second use and replayed afterwards
        # (3 kernels of ours + ~8 small bookkeeping kernels per step; the sampling counter lives on the device)
        self.graph_rollouts = bool(graph_rollouts) and self.fused is not None
        self._collect_calls, self._graph = 0, None
        self.roll = SharedRollouts(num_steps, self.A, n_envs, self.device)
        self.teams = [list(range(0, n_guards)), list(range(n_guards, self.A))]
        self.roll.obs[0].copy_(self.env.reset())
        self.episode_rewards = torch.zeros(self.A, n_envs, device=self.device)

    # -- Learner.act (learner.py:143-172): one forward per team over agent-major rows ----------------
    @torch.no_grad()
    def act(self, step):
        R = self.roll
        for t, (team, opp, policy) in enumerate(((self.teams[0], self.teams[1], self.policies[0]),
                                                 (self.teams[1], self.teams[0], self.policies[1]))):
            lo, hi, olo, ohi = team[0], team[-1] + 1, opp[0], opp[-1] + 1
            if t == 1 and self.ensemble is not None:
                out = {"value": R.value_preds[step, lo:hi], "action": R.actions[step, lo:hi],
                       "action_i32": R.actions_i32[step, lo:hi], "logp": R.action_log_probs[step, lo:hi]}
                order, offsets = self._ensemble_lists()
                # one launch: CTA c keeps checkpoint c % K resident and serves the envs that currently play it
                forward_ensemble(self.ensemble, R.obs[step, lo:hi], R.obs[step, olo:ohi], order, offsets, MODE_SAMPLE, out=out)
                continue
            if self.fused is not None:
                # one launch: value, sampled action (int64 for the storage, int32 for the step kernel) and its
                # log-probability are written straight into slot `step` of the shared rollout blocks
                self.fused[t].forward(R.obs[step, lo:hi], R.obs[step, olo:ohi], MODE_SAMPLE,
                                      out={"value": R.value_preds[step, lo:hi], "action": R.actions[step, lo:hi],
                                           "action_i32": R.actions_i32[step, lo:hi],
                                           "logp": R.action_log_probs[step, lo:hi]})
                continue
            own = R.obs[step, lo:hi].reshape(-1, 6)
            oth = R.obs[step, olo:ohi].reshape(-1, 6)
            value, action, logp, _ = policy.act(own, None, oth, None, deterministic=False)
            n = len(team)
            R.value_preds[step, lo:hi] = value.view(n, self.E, 1)
            R.actions[step, lo:hi] = action.view(n, self.E, 1)
            R.action_log_probs[step, lo:hi] = logp.view(n, self.E, 1)
            R.actions_i32[step, lo:hi].copy_(R.actions[step, lo:hi, :, 0])
        return R.actions_i32[step]

    # -- the data-collection loop of train_fortattack.train (:51-104) for E envs ---------------------
    @torch.no_grad()
    def collect(self):
        self._collect_calls += 1
        if not self.graph_rollouts or self._collect_calls == 1:       # first pass eager: lazy initialisation happens here
            return self._collect()
        if self._graph is None:
            torch.cuda.synchronize(self.device)
            self._graph = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream(self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                with torch.cuda.graph(self._graph, stream=side):
                    self._collect()
            torch.cuda.current_stream(self.device).wait_stream(side)
        self._graph.replay()
        return self.episode_rewards

    def _ensemble_lists(self):
        """Envs grouped by the checkpoint they currently play: (order int32 [E], offsets int32 [K+1]), device-side."""
        K = len(self.ensemble)
        ids = self.att_id.long()
        order = torch.argsort(ids, stable=True).to(torch.int32)
        offsets = torch.zeros(K + 1, dtype=torch.int32, device=self.device)
        counts = (ids[:, None] == torch.arange(K, device=self.device)[None, :]).sum(0)     # (bincount would synchronise)
        offsets[1:] = torch.cumsum(counts, 0).to(torch.int32)
        return order, offsets

    def ensemble_table(self):
        """The reference's evaluation table (test_fortattack_v2.py:50,94-124), one row per attacker checkpoint, averaged
        over the episodes finished so far: [P(all attackers dead), P(time limit), their sum = guards win, P(fort reached),
        alive guards, alive attackers, mean guard return, mean attacker return]."""
        n = self.ensemble_sums[:, 0].clamp_min(1.0)
        r = self.ensemble_results.double()
        cols = [r[:, 1] / n, r[:, 2] / n, (r[:, 1] + r[:, 2]) / n, r[:, 3] / n] + [self.ensemble_sums[:, k] / n for k in (1, 2, 3, 4)]
        return torch.stack(cols, dim=1).cpu().numpy()

    def invalidate_graph(self):
        """Drop the captured rollout graph (it bakes in kernel parameters such as the episode cap and the tensors'
        addresses): call after changing anything collect() depends on; the next collect() captures again."""
        self._graph = None

    @torch.no_grad()
    def _collect(self):
        R = self.roll
        self.episode_rewards.zero_()
        for step in range(self.T):
            masks = R.obs[step, :, :, 0]                                   # alive flags before the step (:53)
            actions = self.act(step)
            # obs of step+1, rewards of step, done/result are written in place by the kernel
            # ... and so are RolloutStorage.insert's masks (storage.py:41), the end point (train_fortattack.py:98),
            # initialize_new_episode's masks for finished envs (:100-104) and the episode reward sums: same launch
            self.env.step(actions, auto_reset=True, out=(R.obs[step + 1], R.rewards[step], R.done[step], R.result[step]),
                          bookkeeping=(R.masks[step + 1], R.ends[step + 1], self.episode_rewards))
            if self.ensemble is not None:
                K = len(self.ensemble)
                finished = R.done[step] != 0
                code = R.result[step].long()                                  # 0 running, 1 all dead, 2 time limit, 3 reached
                self.ensemble_results.view(-1).index_add_(0, self.att_id.long() * 4 + code, finished.long())
                self.ep_return += R.rewards[step] * masks
                fin = finished.double()
                ae = self.alive_end.long()
                rows = torch.stack((fin, fin * (ae & 15), fin * (ae >> 4), fin * self.ep_return[:self.ng].mean(0),
                                    fin * self.ep_return[self.ng:].mean(0)), dim=1)           # [E, 5]
                self.ensemble_sums.index_add_(0, self.att_id.long(), rows)
                self.ep_return *= (~finished).float()
                self.att_id.copy_(torch.where(finished, torch.randint(0, K, (self.E,), device=self.device, dtype=torch.int32),
                                              self.att_id))
        R.ends[self.T] = True                                              # (:108-109)
        return self.episode_rewards

    # -- old log-probs / values in the update's arithmetic ----------------------------------------------
    @torch.no_grad()
    def recompute_old(self, chunk_rows=1 << 18):
        """Overwrite action_log_probs[0:T] and value_preds[0:T] of the TRAINED teams with MPNN.evaluate_actions of the
        stored observations and actions, computed by the same forward the update uses (mpnn.py:194-200 through
        rlcore/fused.py), so that ratio = exp(new - old) is 1 at the first minibatch of the update (ppo.py:163).
        Rows are agent-major per chunk of time steps, exactly the minibatch layout (ppo.py:222-234)."""
        R, T, E = self.roll, self.T, self.E
        two = self.overlap_teams and self.device.type == "cuda" and self.ensemble is None
        main = torch.cuda.current_stream(self.device) if two else None
        if two and getattr(self, "_team_streams", None) is None:
            self._team_streams = [torch.cuda.Stream(self.device) for _ in self.policies]
        for t, policy in enumerate(self.policies):
            if t == 1 and self.ensemble is not None:       # frozen attackers are not trained
                continue
            if two:                                        # the teams write disjoint slices: one stream each
                self._team_streams[t].wait_stream(main)
            if two:
                with torch.cuda.stream(self._team_streams[t]):
                    self._recompute_team(t, policy, chunk_rows)
            else:
                self._recompute_team(t, policy, chunk_rows)
        if two:
            for st in self._team_streams:
                main.wait_stream(st)

    def _recompute_team(self, t, policy, chunk_rows):
        R, T, E = self.roll, self.T, self.E
        team, opp = self.teams[t], self.teams[1 - t]
        lo, hi, olo, ohi = team[0], team[-1] + 1, opp[0], opp[-1] + 1
        n = hi - lo
        steps = max(1, chunk_rows // (max(n, ohi - olo) * E))
        policy.fused_no_grad = True
        try:
            for s0 in range(0, T, steps):
                s1 = min(T, s0 + steps)
                rows = lambda x, a, b: x[s0:s1, a:b].transpose(0, 1).reshape(-1, x.shape[-1])
                v, lp, _, _ = policy.evaluate_actions(rows(R.obs, lo, hi), None, rows(R.obs, olo, ohi), None,
                                                      rows(R.actions, lo, hi))
                R.value_preds[s0:s1, lo:hi] = v.view(n, s1 - s0, E, 1).transpose(0, 1)
                R.action_log_probs[s0:s1, lo:hi] = lp.view(n, s1 - s0, E, 1).transpose(0, 1)
        finally:
            policy.fused_no_grad = False

    # -- Learner.wrap_horizon (learner.py:191-211) ---------------------------------------------------
    @torch.no_grad()
    def wrap_horizon(self):
        R, T = self.roll, self.T
        nv = torch.empty(self.A, self.E, device=self.device)
        for t, (team, opp, policy) in enumerate(((self.teams[0], self.teams[1], self.policies[0]),
                                                 (self.teams[1], self.teams[0], self.policies[1]))):
            lo, hi, olo, ohi = team[0], team[-1] + 1, opp[0], opp[-1] + 1
            if t == 1 and self.ensemble is not None:
                order, offsets = self._ensemble_lists()
                forward_ensemble(self.ensemble, R.obs[T, lo:hi], R.obs[T, olo:ohi], order, offsets, MODE_ARGMAX,
                                 out={"value": nv[lo:hi]})
            elif self.fused is not None and self.exact_old:
                # bootstrap values in the same arithmetic as the recomputed value_preds (recompute_old)
                policy.fused_no_grad = True
                try:
                    nv[lo:hi] = policy.get_value(R.obs[T, lo:hi].reshape(-1, 6), None, R.obs[T, olo:ohi].reshape(-1, 6),
                                                 None).view(len(team), self.E)
                finally:
                    policy.fused_no_grad = False
            elif self.fused is not None:
                self.fused[t].forward(R.obs[T, lo:hi], R.obs[T, olo:ohi], MODE_ARGMAX, out={"value": nv[lo:hi]})
            else:
                nv[lo:hi] = policy.get_value(R.obs[T, lo:hi].reshape(-1, 6), None, R.obs[T, olo:ohi].reshape(-1, 6),
                                             None).view(len(team), self.E)
        R.compute_returns(nv, self.gamma, self.tau)            # segment GAE of all agents and envs: one launch
        if self.fused is not None:
            for f in self.fused + (self.ensemble or []):
                f.check_status()

    # -- Learner.update (learner.py:175-188) ---------------------------------------------------------
    def update(self, train_guards_only=None):
        if train_guards_only is None:                      # an attacker ensemble is frozen by construction
            train_guards_only = self.ensemble is not None
        vals = []
        trainers = self.trainers[:1] if train_guards_only else self.trainers
        jobs = []
        for t, trainer in enumerate(trainers):
            own = [self.roll.agents[i] for i in self.teams[t]]
            opp = [self.roll.agents[i] for i in self.teams[1 - t]]
            lo, olo = self.teams[t][0], self.teams[1 - t][0]
            shared = (self.roll, lo, len(own), olo, len(opp)) if self.fused_update else None
            jobs.append((trainer, own, opp, shared))
        if self._joint_ok(jobs):
            vals = self._update_joint(jobs)
        elif self._overlap_ok(jobs):
            vals = self._update_overlapped(jobs)
        else:
            vals = [trainer.update(own, opp, shared=shared) for trainer, own, opp, shared in jobs]
        if self.fused is not None:                         # the optimizer moved the weights: re-pack the kernel's blob
            for f in self.fused:
                f.refresh()
        return vals

    def _overlap_ok(self, jobs):
        """The two teams' updates share nothing but read-only rollout blocks (each team has its own network, optimizer and
        scratch), so on one rank they are enqueued on two streams and run at the same time.  Several ranks keep them one after
        the other: two captured collectives of one communicator must not be in flight together."""
        return (self.overlap_teams and len(jobs) == 2 and self.device.type == "cuda" and self.process_group is None
                and all(sh is not None and tr.use_clipped_value_loss for tr, _o, _p, sh in jobs))

    def _update_overlapped(self, jobs):
        main = torch.cuda.current_stream(self.device)
        if getattr(self, "_team_streams", None) is None:
            self._team_streams = [torch.cuda.Stream(self.device) for _ in jobs]
        handles = []
        for (trainer, own, opp, shared), st in zip(jobs, self._team_streams):
            st.wait_stream(main)
            with torch.cuda.stream(st):
                prev_tf32 = torch.backends.cuda.matmul.allow_tf32
                torch.backends.cuda.matmul.allow_tf32 = bool(trainer.allow_tf32)
                try:
                    handles.append(trainer.update_begin(own, opp, shared))
                finally:
                    torch.backends.cuda.matmul.allow_tf32 = prev_tf32
        for st in self._team_streams:
            main.wait_stream(st)
        return [trainer.update_end(h) for (trainer, _o, _p, _s), h in zip(jobs, handles)]

    def _joint_ok(self, jobs):
        """Several ranks (or overlap_teams="joint"): the two teams' optimizer steps run as the two branches of ONE captured
        graph around ONE all-reduce of a flat buffer that holds both teams' gradients -- the teams overlap as on one rank, and
        there is never more than one collective of the communicator in flight."""
        want = self.overlap_teams == "joint" or (self.overlap_teams and self.process_group is not None)
        return (want and len(jobs) == 2 and self.device.type == "cuda"
                and all(sh is not None and tr.use_clipped_value_loss and tr.graph_update and tr._tg_adam for tr, _o, _p, sh in jobs))

    def _update_joint(self, jobs):
        dev = self.device
        main = torch.cuda.current_stream(dev)
        trs = [j[0] for j in jobs]
        if getattr(self, "_team_streams", None) is None:
            self._team_streams = [torch.cuda.Stream(dev) for _ in jobs]
        sizes = [sum(p.numel() for p in tr.actor_critic.parameters()) + 5 for tr in trs]
        J = getattr(trs[0], "_joint", None)
        if J is None or J["flat"].numel() != sum(sizes):
            J = trs[0]._joint = {"flat": torch.zeros(sum(sizes), device=dev), "key": None}
        off = 0
        for tr, sz in zip(trs, sizes):                   # each trainer's flat gradient buffer is its segment of the shared one
            tr.use_flat_buffer(J["flat"][off:off + sz])
            off += sz
        sts = []
        for tr, own, _opp, shared in jobs:
            tr._shared = shared
            prev_tf32 = torch.backends.cuda.matmul.allow_tf32
            torch.backends.cuda.matmul.allow_tf32 = bool(tr.allow_tf32)
            try:
                sts.append(tr._prepare_fused(own))
            finally:
                torch.backends.cuda.matmul.allow_tf32 = prev_tf32
        nb, mbs = len(sts[0]["batches"]), sts[0]["mini_batch_size"]
        assert len(sts[1]["batches"]) == nb and sts[1]["mini_batch_size"] == mbs
        key = (id(self.roll), tuple(st["team"] for st in sts), mbs, tuple(tuple(st["advantages"].shape) for st in sts))
        if J["key"] != key:
            J.update(key=key, eager=0, graph=None, idx=[torch.empty_like(st["batches"][0]) for st in sts],
                     adv=[torch.empty_like(st["advantages"]) for st in sts], totals=[torch.zeros_like(st["totals"]) for st in sts])
        adv_fresh = True
        for k in range(nb):
            idxs = [st["batches"][k] for st in sts]
            full = all(i.numel() == mbs for i in idxs)
            if not full or (J["graph"] is None and J["eager"] < 3):
                # a ragged last minibatch, or the warm-up steps a capture needs: one team after the other, each with its own
                # all-reduce (of its segment of the shared buffer)
                J["eager"] += int(full)
                for tr, st, idx in zip(trs, sts, idxs):
                    tr._minibatch_step(st["fused"], st["R"], st["team"], idx, st["advantages"], st["totals"], st["params"], st["world"])
                continue
            if J["graph"] is None:
                self._capture_joint(J, trs, sts, idxs, main)
            for t, st in enumerate(sts):
                J["idx"][t].copy_(idxs[t])
                if adv_fresh:
                    J["adv"][t].copy_(st["advantages"])
                J["totals"][t].zero_()
            adv_fresh = False
            J["graph"].replay()
            for t, st in enumerate(sts):
                st["totals"] += J["totals"][t]
        return [tr.update_end(st["totals"]) for tr, st in zip(trs, sts)]

    def _capture_joint(self, J, trs, sts, idxs, main):
        dev = self.device
        for t, st in enumerate(sts):
            J["idx"][t].copy_(idxs[t])
            J["adv"][t].copy_(st["advantages"])
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(dev)
        side.wait_stream(main)
        for tr in trs:
            tr.optimizer.zero_grad(set_to_none=True)
        # (a process group's watchdog thread polls its events while we capture -> thread-local mode, as in JointPPO._graphed_step)
        with torch.cuda.graph(graph, stream=side, capture_error_mode="thread_local" if self.process_group is not None else "global"):
            for t, (tr, st) in enumerate(zip(trs, sts)):
                ts = self._team_streams[t]
                ts.wait_stream(side)                     # fork: each team's forward / backward on its own branch
                with torch.cuda.stream(ts):
                    tr._minibatch_step(st["fused"], st["R"], st["team"], J["idx"][t], J["adv"][t], J["totals"][t], st["params"],
                                       st["world"], fill_only=True)
            for ts in self._team_streams:
                side.wait_stream(ts)                     # join
            trs[0]._allreduce(J["flat"])                 # ONE collective: both teams' gradients, normalisers and loss sums
            for t, (tr, st) in enumerate(zip(trs, sts)):
                tr._ranks_apply_flat(J["totals"][t], st["params"])
        main.wait_stream(side)
        J["graph"] = graph

    def release_graphs(self):
        """Drop every captured graph (rollout collection, optimizer steps, the joint two-team step); the next collect() /
        update() captures again.  Required before dist.destroy_process_group() when the trainer has a process group and
        graph_update=True, and after changing anything a captured step bakes in as a kernel argument (learning rate, clip,
        loss coefficients, max_grad_norm, the episode cap) -- the reference's scripts keep all of these constant."""
        for trn in self.trainers:
            trn.release_graphs()
        self._graph = None

    def after_update(self):
        self.roll.after_update()

    def train_once(self, train_guards_only=None):
        rewards = self.collect()
        if self.exact_old:
            self.recompute_old()
        self.wrap_horizon()
        vals = self.update(train_guards_only)
        self.after_update()
        return rewards, vals

    # -- checkpoints in the reference's format --------------------------------------------------------
    def state(self):
        sd = [self.policies[0].state_dict()] * self.ng + [self.policies[1].state_dict()] * self.na
        return {"models": sd, "ob_rms": (None, None)}

    def save(self, path):
        torch.save(self.state(), path)

    def load_models(self, models):
        self.policies[0].load_state_dict(models[0])        # learner.py:245-249
        self.policies[1].load_state_dict(models[-1])
        if self.fused is not None:
            for f in self.fused:
                f.refresh()
