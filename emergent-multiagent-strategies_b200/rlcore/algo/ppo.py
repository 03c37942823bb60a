"""Team-shared clipped PPO of the reference (rlcore/algo/ppo.py:89-246).

`JointPPO(actor_critic, clip_param, ppo_epoch, num_mini_batch, value_loss_coef, entropy_coef, lr, eps,
max_grad_norm, use_clipped_value_loss).update(rollouts_list, opp_rollouts_list) -> (value_loss,
action_loss, dist_entropy)` with the same arithmetic: per-agent advantage normalisation, one shared
random permutation of the T*P sample indices per epoch so the teammates' rows stay time-aligned,
agent-major minibatches, alive-mask weighting of every loss term divided by mask.mean(), Adam,
gradient-norm clipping.  Random draws (torch.randperm per epoch, as SubsetRandomSampler does) match the
reference, so equal seeds give equal updates.

Multi-GPU (new): pass `process_group=`; each rank holds its own env shard, gradients are
summed with ONE all-reduce over a flat buffer per optimizer step and the loss normalisers
(mask.mean(), advantage mean/std) are made global, so N ranks x P/N envs reproduce 1 rank x P envs.
The three `.item()` syncs per minibatch of the reference (ppo.py:194-196) are replaced by on-device
accumulation with a single sync per update.
"""
import torch
import torch.nn as nn
import torch.optim as optim


def magent_feed_forward_generator(rollouts_list, opp_rollouts_list, advantages_list, num_mini_batch, perm=None,
                                  index_batches=None):
    """Time-aligned multi-agent minibatches (ppo.py:207-246): for every chunk of one random permutation of
    the T*P indices, the rows of all teammates for those indices, concatenated agent-major."""
    num_steps, num_processes = rollouts_list[0].rewards.size()[0:2]
    batch_size = num_processes * num_steps
    mini_batch_size = int(batch_size / num_mini_batch)
    if perm is None and index_batches is None:
        perm = torch.randperm(batch_size)                     # == SubsetRandomSampler(range(batch_size))
    # rows are addressed as (t, env) = divmod(index, P): no flattening, so the buffers may be strided
    # views of shared [T, A, E, ...] blocks (rollout.SharedRollouts) as well as the reference's own layout
    own_obs = [r.obs[:-1] for r in rollouts_list]
    opp_obs = [r.obs[:-1] for r in opp_rollouts_list]
    hid = [r.recurrent_hidden_states[:-1] for r in rollouts_list]
    act = [r.actions for r in rollouts_list]
    val = [r.value_preds[:-1] for r in rollouts_list]
    ret = [r.returns[:-1] for r in rollouts_list]
    msk = [r.masks[:-1] for r in rollouts_list]
    olp = [r.action_log_probs for r in rollouts_list]
    adv = list(advantages_list)
    dev = own_obs[0].device
    if index_batches is None:                                  # BatchSampler(..., drop_last=False)
        index_batches = [perm[i:i + mini_batch_size] for i in range(0, batch_size, mini_batch_size)]
    for idx in index_batches:
        idx = idx.to(dev)
        ti, ei = torch.div(idx, num_processes, rounding_mode="floor"), idx % num_processes
        take = lambda lst: torch.cat([t[ti, ei] for t in lst], 0)
        obs_batch = take(own_obs)
        mask = obs_batch[:, 0].clone().view(-1, 1)            # alive flag = observation feature 0 (ppo.py:224)
        yield (obs_batch, mask, take(opp_obs), take(hid), take(act), take(val), take(ret), take(msk),
               take(olp), take(adv))


class JointPPO(object):
    def __init__(self, actor_critic, clip_param, ppo_epoch, num_mini_batch, value_loss_coef, entropy_coef,
                 lr=None, eps=None, max_grad_norm=None, use_clipped_value_loss=False, process_group=None,
                 allow_tf32=False, graph_update=False, tg_optimizer=True):
        self.actor_critic = actor_critic
        # graph_update (new, fused path): after three eager minibatch steps the whole optimizer step
        # (gather -> forward -> loss -> backward -> clip -> Adam) is captured in a CUDA graph and replayed; with TF32
        # GEMMs the eager step is CPU-launch-bound (~9 ms of Python/autograd dispatch vs ~7 ms of GPU work at config 3)
        self.graph_update = bool(graph_update)
        self._g = None
        # allow_tf32 (new, off by default = the reference's fp32 arithmetic): run the update's cuBLAS GEMMs on the
        # tensor cores in TF32.  On B200 the fp32 path is SIMT sgemm and takes ~70% of the update
        # (profiles/r1c_ppo_update_torch_profile.txt).
        self.allow_tf32 = allow_tf32
        self.clip_param, self.ppo_epoch, self.num_mini_batch = clip_param, ppo_epoch, num_mini_batch
        self.value_loss_coef, self.entropy_coef = value_loss_coef, entropy_coef
        self.max_grad_norm, self.use_clipped_value_loss = max_grad_norm, use_clipped_value_loss
        on_cuda = next(actor_critic.parameters()).is_cuda
        # eps is ignored by the reference too (:114); capturable keeps Adam's step counters on the device (graph capture)
        # on CUDA the whole Adam step is ONE fused kernel (24 parameter tensors; the foreach form cost 240 us of GPU time per
        # optimizer step, 7 % of it)
        # on CUDA: gradient-norm clip + Adam are ONE pair of launches of this repo's own kernel (rlcore/fused.TgAdam ->
        # tg_adam_step), step counter on the device; on the CPU (tests, the reference's --no-cuda runs) torch's Adam
        self._tg_adam = False
        if on_cuda and tg_optimizer:
            try:
                from .. import fused as _fused
            except ImportError:
                from rlcore import fused as _fused
            self.optimizer = _fused.TgAdam(actor_critic.parameters(), lr=lr)
            self._tg_adam = True
        else:
            self.optimizer = optim.Adam(actor_critic.parameters(), lr=lr, capturable=bool(graph_update) and on_cuda,
                                        fused=True if on_cuda else None)
        self.process_group = process_group

    def _clip_and_step(self):
        if self._tg_adam:
            self.optimizer.step(self.max_grad_norm)
        else:
            if self.max_grad_norm:
                nn.utils.clip_grad_norm_(self.actor_critic.parameters(), self.max_grad_norm)
            self.optimizer.step()

    # -- distributed helpers (identity on one rank) ------------------------------------------------
    def _world(self):
        import torch.distributed as dist
        return dist.get_world_size(self.process_group) if self.process_group is not None else 1

    def _allreduce(self, t):
        import torch.distributed as dist
        if self.process_group is not None:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.process_group)
        return t

    def _perm(self, rollout):
        """One permutation of the T*P sample indices per epoch; rank 0's draw is used by every rank."""
        T, P = rollout.rewards.size()[0:2]
        perm = torch.randperm(T * P)
        if self.process_group is not None:
            import torch.distributed as dist
            perm = perm.to(rollout.rewards.device)
            dist.broadcast(perm, src=dist.get_global_rank(self.process_group, 0), group=self.process_group)
        return perm

    def _advantages(self, rollout):
        adv = rollout.returns[:-1] - rollout.value_preds[:-1]
        if self.process_group is None:
            return (adv - adv.mean()) / (adv.std() + 1e-5)               # ppo.py:121-123 (unbiased std)
        s = self._allreduce(torch.stack([adv.sum(), (adv * adv).sum(), adv.new_tensor(float(adv.numel()))]).double())
        n, mean = s[2], s[0] / s[2]
        var = (s[1] - n * mean * mean) / (n - 1)
        return ((adv - mean.to(adv.dtype)) / (var.clamp_min(0).sqrt().to(adv.dtype) + 1e-5))

    def update(self, rollouts_list, opp_rollouts_list, index_batches=None, shared=None):
        """index_batches (optional, testing): per epoch, a list of index tensors to use as the minibatches
        instead of chunks of a fresh random permutation.
        shared (optional): (SharedRollouts, a0, n, o0, m) when the per-agent storages are views of the shared rollout
        blocks -- minibatches are then gathered by ONE kernel and the clipped-PPO loss and its gradient are evaluated
        by ONE kernel (rlcore/fused.py), instead of ~60 index/cat and ~30 elementwise launches per minibatch."""
        self._shared = shared
        prev_tf32 = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = bool(self.allow_tf32)
        try:
            if shared is not None and self.use_clipped_value_loss and rollouts_list[0].rewards.is_cuda:
                return self._update_fused(rollouts_list, index_batches)
            return self._update(rollouts_list, opp_rollouts_list, index_batches)
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev_tf32

    def _update_fused(self, rollouts_list, index_batches=None):
        return self._update_fused_end(self._update_fused_begin(rollouts_list, index_batches))

    def update_begin(self, rollouts_list, opp_rollouts_list, shared):
        """Enqueue a whole fused update on the CURRENT stream without waiting for it; update_end(handle) returns the three
        averaged losses (the update's only host synchronisation).  BatchedTrainer.update uses the pair to run the two teams'
        updates on two streams at once.  Returns None when this update cannot take the fused path (the caller then uses
        update())."""
        if not (shared is not None and self.use_clipped_value_loss and rollouts_list[0].rewards.is_cuda):
            return None
        self._shared = shared
        return self._update_fused_begin(rollouts_list, None)

    def update_end(self, handle):
        return self._update_fused_end(handle)

    def _prepare_fused(self, rollouts_list, index_batches=None):
        """Everything a fused update needs before its first optimizer step: normalised advantages [n, T, E], the minibatch index
        tensors of all epochs in order (on the device), the loss accumulator."""
        try:
            from .. import fused
        except ImportError:
            from rlcore import fused
        R, a0, n, o0, m = self._shared
        advantages = torch.stack([self._advantages(r)[..., 0] for r in rollouts_list]).contiguous()      # [n, T, E]
        T, P = rollouts_list[0].rewards.size()[0:2]
        dev = advantages.device
        batch_size = T * P
        mini_batch_size = int(batch_size / self.num_mini_batch)
        perms = None
        if index_batches is None:
            # every epoch's permutation is drawn up front (same draws, same order as one per epoch) and goes to the device
            # in one asynchronous copy from pinned memory: a copy from pageable memory makes the host wait for the stream,
            # and the host has to stay ahead of the device for the whole update
            perms = [self._perm(rollouts_list[0]) for _ in range(self.ppo_epoch)]
            perms = [q if q.is_cuda else q.pin_memory().to(dev, non_blocking=True) for q in perms]
        batches = []
        for epoch in range(self.ppo_epoch):
            if index_batches is not None:
                batches += [idx.to(dev).contiguous() for idx in index_batches[epoch]]
            else:
                perm = perms[epoch]
                batches += [perm[i:i + mini_batch_size].contiguous() for i in range(0, batch_size, mini_batch_size)]
        self._adv_fresh = True               # the captured step copies the advantages into its own buffer once per update
        return {"fused": fused, "R": R, "team": (a0, n, o0, m), "advantages": advantages, "batches": batches,
                "mini_batch_size": mini_batch_size, "totals": torch.zeros(3, device=dev),
                "params": [p for p in self.actor_critic.parameters()], "world": self._world()}

    def _update_fused_begin(self, rollouts_list, index_batches=None):
        st = self._prepare_fused(rollouts_list, index_batches)
        fused, R, team, advantages, totals = st["fused"], st["R"], st["team"], st["advantages"], st["totals"]
        for idx in st["batches"]:
            if self._graphed_step(fused, R, team, idx, advantages, totals, st["mini_batch_size"], st["world"]):
                continue
            self._minibatch_step(fused, R, team, idx, advantages, totals, st["params"], st["world"])
        return totals

    def _update_fused_end(self, totals):
        # (several ranks: the per-step sums were made global inside _minibatch_step, nothing left to reduce)
        # the reference divides the summed losses by ppo_epoch * num_mini_batch whatever the number of minibatches the
        # sampler produced (ppo.py:198-202: a ragged tail adds one more term to the sums)
        v, a, e = (totals / (self.ppo_epoch * self.num_mini_batch)).tolist()
        return v, a, e

    def _minibatch_step(self, fused, R, team, idx, advantages, totals, params, world, fill_only=False):
        """fill_only: stop after the flat buffer holds this rank's gradients and loss sums (the several-ranks form of the
        step, whatever the world size); the caller runs the collective and _ranks_apply_flat itself (BatchedTrainer's joint
        step: both teams' halves around ONE all-reduce)."""
        fused.pack_cache(True)               # weights are constant from here to the optimizer kernel: packs are reused
        prev_scope = fused.scratch_scope(id(self))      # this trainer's own partial-sum / counter / status storage
        try:
            self._minibatch_step_body(fused, R, team, idx, advantages, totals, params, world, fill_only)
        finally:
            fused.scratch_scope(prev_scope)
            fused.pack_cache(False)

    def _minibatch_step_body(self, fused, R, team, idx, advantages, totals, params, world, fill_only=False):
        a0, n, o0, m = team
        (obs_batch, mask, obs_opp_batch, actions_batch, value_preds_batch, return_batch, masks_batch,
         old_log_probs_batch, adv_targ, alive_sum) = fused.gather_minibatch(R, idx, a0, n, o0, m, advantages)
        # the Categorical head (log-prob, entropy) and its backward live inside the loss kernel when the policy offers its
        # logits (MPNN.evaluate_logits on the fused path); otherwise the distribution is evaluated by torch as before
        head = getattr(self.actor_critic, "evaluate_logits", None)
        vl = head(obs_batch, obs_opp_batch) if head is not None and torch.is_grad_enabled() else None
        if vl is not None:
            values, logits = vl

            def loss_of(norm_):
                return fused.ppo_loss_logits(values, logits, actions_batch, value_preds_batch, return_batch, old_log_probs_batch,
                                             adv_targ, mask, norm_, self.clip_param, self.value_loss_coef, self.entropy_coef)
        else:
            values, action_log_probs, dist_entropy, _ = self.actor_critic.evaluate_actions(
                obs_batch, None, obs_opp_batch, masks_batch, actions_batch)

            def loss_of(norm_):
                return fused.ppo_loss(values, action_log_probs, dist_entropy, value_preds_batch, return_batch, old_log_probs_batch,
                                      adv_targ, mask, norm_, self.clip_param, self.value_loss_coef, self.entropy_coef)
        count = mask.new_full((1,), float(mask.numel()))
        if fill_only:
            self._ranks_fill_flat(loss_of, mask, alive_sum, count, params)
            return
        if world == 1:
            norm = torch.where(alive_sum != 0, alive_sum, count)          # mask.mean() != 0 else 1 (ppo.py:150-187)
            loss, stats = loss_of(norm)
            self.optimizer.zero_grad()
            loss.backward()
            self._clip_and_step()
            totals += stats[:3]
            return
        # Several ranks: ONE collective per optimizer step.  Every loss term is sum_r(local sum) / sum_r(local alive count)
        # (ppo.py:150-187 with global sums), so each rank differentiates its UN-normalised local sums (norm = 1), the
        # gradients travel together with the two normaliser terms and the three loss sums in one persistent flat buffer,
        # and the division by the global normaliser happens after the all-reduce (inside the optimizer kernel).
        self._ranks_fill_flat(loss_of, mask, alive_sum, count, params)
        self._allreduce(self._flat_grads(params)[0])
        self._ranks_apply_flat(totals, params)

    def _ranks_fill_flat(self, loss_of, mask, alive_sum, count, params):
        """Backward of the un-normalised local loss sums; gradients + [alive sum, count, three loss sums] into the flat buffer."""
        if getattr(self, "_one", None) is None or self._one.device != mask.device:
            self._one = torch.ones(1, device=mask.device)
        loss, stats = loss_of(self._one)
        self.optimizer.zero_grad()
        loss.backward()
        flat, _views = self._flat_grads(params)
        total = flat.numel() - 5
        torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params], out=flat[:total])
        torch.cat([alive_sum, count, stats[:3]], out=flat[total:])

    def _ranks_apply_flat(self, totals, params):
        """After the all-reduce of the flat buffer: divide by the global normaliser, clip, Adam, loss sums."""
        flat, views = self._flat_grads(params)
        total = flat.numel() - 5
        inv = 1.0 / torch.where(flat[total:total + 1] != 0, flat[total:total + 1], flat[total + 1:total + 2])
        if self._tg_adam:
            self.optimizer.step(self.max_grad_norm, grad_scale=inv, grads=views)
        else:
            for p, v in zip(params, views):
                p.grad = v * inv
            self._clip_and_step()
        totals += flat[total + 2:] * inv

    def use_flat_buffer(self, flat, params=None):
        """Install a caller-owned flat gradient buffer ([sum(numel) + 5] floats, e.g. one segment of a buffer that several
        trainers share so that ONE collective serves all of them: BatchedTrainer's joint step)."""
        params = [p for p in self.actor_critic.parameters()] if params is None else params
        total = sum(p.numel() for p in params)
        if flat.numel() != total + 5 or flat.dtype != torch.float32 or not flat.is_contiguous():
            raise ValueError("flat buffer must be contiguous float32 with %d elements" % (total + 5))
        views, off = [], 0
        for p in params:
            views.append(flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        self._fg = (tuple(id(p) for p in params), flat, views)

    def _flat_grads(self, params):
        """Persistent flat buffer [sum(numel) + 5] and its per-parameter views (allocated once per parameter set)."""
        key = tuple(id(p) for p in params)
        fg = getattr(self, "_fg", None)
        if fg is None or fg[0] != key:
            total = sum(p.numel() for p in params)
            flat = torch.zeros(total + 5, device=params[0].device)
            views, off = [], 0
            for p in params:
                views.append(flat[off:off + p.numel()].view_as(p))
                off += p.numel()
            fg = self._fg = (key, flat, views)
        return fg[1], fg[2]

    def release_graphs(self):
        """Drop the captured optimizer-step graph (it holds this process group's NCCL kernels): call before
        dist.destroy_process_group(), or teardown waits on the graph's communicator references."""
        self._g = None
        self._joint = None                   # (BatchedTrainer keeps its joint two-team step graph on the first trainer)

    def _graphed_step(self, fused, R, team, idx, advantages, totals, mini_batch_size, world):
        """Replay (or, on its fourth call, capture) the optimizer step as one CUDA graph.  Returns False when the step
        has to run eagerly (option off, a ragged last minibatch, or still warming up).  With several ranks the two
        all-reduces of the step (loss normaliser, flat gradient) are captured with it -- NCCL collectives are graph nodes --
        so the replicas replay in lockstep and the multi-rank step costs no host work either; every rank takes the same
        eager / capture / replay decisions because they depend only on call counts and sizes."""
        if not self.graph_update or idx.numel() != mini_batch_size or not idx.is_cuda:
            return False
        g = self._g
        if g is None or g["key"] != (id(R), team, mini_batch_size, tuple(advantages.shape)):
            g = self._g = {"key": (id(R), team, mini_batch_size, tuple(advantages.shape)), "eager": 0, "graph": None,
                           "idx": torch.empty_like(idx), "adv": torch.empty_like(advantages), "totals": torch.zeros_like(totals)}
        if g["graph"] is None:
            if g["eager"] < 3:                      # real training steps double as the warm-up torch asks for before a capture
                g["eager"] += 1
                return False
            params = [p for p in self.actor_critic.parameters()]
            g["idx"].copy_(idx)
            g["adv"].copy_(advantages)
            torch.cuda.synchronize(idx.device)
            graph = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream(idx.device)
            side.wait_stream(torch.cuda.current_stream(idx.device))
            self.optimizer.zero_grad(set_to_none=True)
            with torch.cuda.stream(side):
                # several ranks: the process group's watchdog thread polls its events while we capture -> thread-local mode
                with torch.cuda.graph(graph, stream=side, capture_error_mode="global" if world == 1 else "thread_local"):
                    self._minibatch_step(fused, R, team, g["idx"], g["adv"], g["totals"], params, world)
            torch.cuda.current_stream(idx.device).wait_stream(side)
            g["graph"] = graph
        g["idx"].copy_(idx)
        if getattr(self, "_adv_fresh", True):   # the advantages are one tensor for a whole update: copied at its first replay
            g["adv"].copy_(advantages)
            self._adv_fresh = False
        g["totals"].zero_()
        g["graph"].replay()
        totals += g["totals"]
        return True

    def _update(self, rollouts_list, opp_rollouts_list, index_batches=None):
        advantages_list = [self._advantages(r) for r in rollouts_list]
        dev = rollouts_list[0].rewards.device
        totals = torch.zeros(3, device=dev)
        params = [p for p in self.actor_critic.parameters()]
        world = self._world()
        n_updates = 0
        for epoch in range(self.ppo_epoch):
            if self.actor_critic.is_recurrent:
                raise NotImplementedError("sampler not implemented for recurrent policies")
            if index_batches is not None:
                gen = magent_feed_forward_generator(rollouts_list, opp_rollouts_list, advantages_list,
                                                    self.num_mini_batch, index_batches=index_batches[epoch])
            else:
                gen = magent_feed_forward_generator(rollouts_list, opp_rollouts_list, advantages_list,
                                                    self.num_mini_batch, perm=self._perm(rollouts_list[0]))
            for sample in gen:
                (obs_batch, mask, obs_opp_batch, hid_batch, actions_batch, value_preds_batch, return_batch,
                 masks_batch, old_log_probs_batch, adv_targ) = sample
                values, action_log_probs, dist_entropy, _ = self.actor_critic.evaluate_actions(
                    obs_batch, hid_batch, obs_opp_batch, masks_batch, actions_batch)

                # Every loss term is `(x * mask).mean()`, divided by mask.mean() unless that is 0
                # (ppo.py:150-187).  One rank: literally that.  G ranks: the same quotient with GLOBAL sums,
                # sum_r(local sum) / sum_r(mask sum); each rank scales its local sum by G / global mask sum so
                # that the average of the ranks' gradients is the gradient of the global loss.
                mmean = mask.mean()
                if world == 1:
                    denom = torch.where(mmean != 0, mmean, torch.ones_like(mmean))
                    mmean_of = lambda x: x.mean() / denom
                else:
                    g = self._allreduce(torch.stack([mask.sum(), mask.new_tensor(float(mask.numel()))]))
                    denom = torch.where(g[0] != 0, g[0], g[1]) / world
                    mmean_of = lambda x: x.sum() / denom
                    mmean = g[0] / g[1]

                entropy = mmean_of(dist_entropy * mask[:, 0])
                ratio = mask * torch.exp(action_log_probs - old_log_probs_batch)
                surr1 = ratio * adv_targ
                surr2 = torch.clamp(ratio, 1.0 - self.clip_param, 1.0 + self.clip_param) * adv_targ
                action_loss = mmean_of(mask * -torch.min(surr1, surr2))
                if self.use_clipped_value_loss:
                    clipped = value_preds_batch + (values - value_preds_batch).clamp(-self.clip_param, self.clip_param)
                    v = 0.5 * torch.max((values - return_batch).pow(2), (clipped - return_batch).pow(2))
                    value_loss = mmean_of(v * mask)
                elif world == 1:
                    # scalar mse times the mask, averaged, over mask.mean(): the mask cancels (ppo.py:182-187)
                    value_loss = mmean_of(0.5 * (return_batch - values).pow(2).mean() * mask)
                else:
                    value_loss = 0.5 * (return_batch - values).pow(2).mean() * (mmean != 0)

                self.optimizer.zero_grad()
                (value_loss * self.value_loss_coef + action_loss - entropy * self.entropy_coef).backward()
                if world > 1:
                    grads = [p.grad for p in params if p.grad is not None]
                    flat = torch.cat([g_.reshape(-1) for g_ in grads])          # one 158 153-float buffer
                    self._allreduce(flat).div_(world)
                    off = 0
                    for g_ in grads:
                        g_.copy_(flat[off:off + g_.numel()].view_as(g_))
                        off += g_.numel()
                self._clip_and_step()
                totals += torch.stack([value_loss.detach(), action_loss.detach(), entropy.detach()])
                n_updates += 1
        if world > 1:
            totals = self._allreduce(totals) / world
        v, a, e = (totals / (self.ppo_epoch * self.num_mini_batch)).tolist()          # the only host sync of the update (ppo.py:198-202)
        return v, a, e


class PPO(JointPPO):
    """rlagent.Neo builds one of these per agent and never uses it (rlagent.py:16-17, ppo.py:8): kept so
    that the reference's rlagent.py imports and constructs unchanged."""

    def __init__(self, actor_critic, clip_param, ppo_epoch, num_mini_batch, value_loss_coef, entropy_coef,
                 lr=None, eps=None, max_grad_norm=None, use_clipped_value_loss=True):
        super().__init__(actor_critic, clip_param, ppo_epoch, num_mini_batch, value_loss_coef, entropy_coef,
                         lr=lr, eps=eps, max_grad_norm=max_grad_norm, use_clipped_value_loss=use_clipped_value_loss)

    def update(self, rollouts):
        raise NotImplementedError("single-agent PPO is unused by the FortAttack scripts (rlcore/algo/ppo.py:8)")
