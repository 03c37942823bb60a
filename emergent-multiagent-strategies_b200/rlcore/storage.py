"""Per-agent rollout buffers of the reference (rlcore/storage.py:9-149), device-resident.

Same constructor, attributes ([T+1, P, ...] tensors named obs, recurrent_hidden_states, rewards,
value_preds, returns, action_log_probs, actions (int64), masks, plus num_steps / step) and the same
methods with the same meaning, so Neo / Learner / magent_feed_forward_generator (rlagent.py:23-42,
learner.py:110-111,150-152,196-199, rlcore/algo/ppo.py:207-246) work on it unchanged.  Additions for the
batched engine: the buffers can be created directly on a device (`device=`), and `compute_returns_batched`
runs the reference's segment-wise GAE for P envs whose episodes end at different steps.
"""
import torch


def _flatten_helper(T, N, _tensor):
    return _tensor.view(T * N, *_tensor.size()[2:])


class RolloutStorage(object):
    def __init__(self, num_steps, num_processes, obs_shape, action_space, recurrent_hidden_state_size, device=None):
        z = lambda *s, **k: torch.zeros(*s, device=device, **k)
        self.obs = z(num_steps + 1, num_processes, *obs_shape)
        self.recurrent_hidden_states = z(num_steps + 1, num_processes, recurrent_hidden_state_size)
        self.rewards = z(num_steps, num_processes, 1)
        self.value_preds = z(num_steps + 1, num_processes, 1)
        self.returns = z(num_steps + 1, num_processes, 1)
        self.action_log_probs = z(num_steps, num_processes, 1)
        self.actions = z(num_steps, num_processes, 1, dtype=torch.long)
        self.masks = torch.ones(num_steps + 1, num_processes, 1, device=device)
        self.num_steps = num_steps
        self.step = 0

    _FIELDS = ("obs", "recurrent_hidden_states", "rewards", "value_preds", "returns", "action_log_probs",
               "actions", "masks")

    def to(self, device):
        for name in self._FIELDS:
            setattr(self, name, getattr(self, name).to(device))

    def insert(self, obs, recurrent_hidden_states, actions, action_log_probs, value_preds, rewards, masks):
        s = self.step
        self.obs[s + 1].copy_(obs)
        self.recurrent_hidden_states[s + 1].copy_(recurrent_hidden_states)
        self.actions[s].copy_(actions)
        self.action_log_probs[s].copy_(action_log_probs)
        self.value_preds[s].copy_(value_preds)
        self.rewards[s].copy_(rewards)
        self.masks[s + 1].copy_(masks)
        self.step = (s + 1) % self.num_steps

    def reset(self):
        self.step = 0

    def after_update(self):
        # last slot becomes the first; everything after it is cleared (storage.py:51-56)
        self.obs[0].copy_(self.obs[-1])
        self.obs[1:] = 0
        self.recurrent_hidden_states[0].copy_(self.recurrent_hidden_states[-1])
        self.masks[0].copy_(self.masks[-1])
        self.step = 0

    def compute_returns(self, next_value, use_gae, gamma, tau, start_pt, end_pt):
        """Returns for the episode segment [start_pt, end_pt) (storage.py:59-70): GAE with
        delta_t = r_t + gamma V_{t+1} m_{t+1} - V_t, A_t = delta_t + gamma tau m_{t+1} A_{t+1}."""
        if use_gae:
            self.value_preds[end_pt] = next_value
            gae = 0
            for step in reversed(range(start_pt, end_pt)):
                delta = self.rewards[step] + gamma * self.value_preds[step + 1] * self.masks[step + 1] \
                    - self.value_preds[step]
                gae = delta + gamma * tau * self.masks[step + 1] * gae
                self.returns[step] = gae + self.value_preds[step]
        else:
            self.returns[end_pt] = next_value
            for step in reversed(range(start_pt, end_pt)):
                self.returns[step] = self.returns[step + 1] * gamma * self.masks[step + 1] + self.rewards[step]

    def compute_returns_batched(self, next_value, ends, gamma, tau):
        """The same GAE for P envs with per-env episode boundaries, in one reverse sweep.

        ends: bool/uint8 [T+1, P]; ends[t, p] != 0 marks t as an end point of env p exactly as the
        reference's `end_pts` list does (train_fortattack.py:97-109: the step count right after a done,
        and always T).  Per env this reproduces Learner.wrap_horizon (learner.py:191-211): every segment
        [start, end) is swept with the accumulator cleared, `start = end + 1`, so index `end` itself is
        skipped and keeps whatever `returns` held.  value_preds[t] must already hold V(obs[t]) for t < T
        (it does: it is written when acting at step t on the same observation the reference would
        re-evaluate), next_value is V(obs[T])."""
        T = self.num_steps
        ends = ends.to(self.rewards.device).bool()
        self.value_preds[T] = next_value
        gae = torch.zeros_like(self.rewards[0])
        for step in reversed(range(T)):
            is_end = ends[step].view(-1, 1)                      # step is an end point: skipped
            nxt_end = ends[step + 1].view(-1, 1)                 # a segment ends right after this step
            gae = torch.where(nxt_end, torch.zeros_like(gae), gae)
            delta = self.rewards[step] + gamma * self.value_preds[step + 1] * self.masks[step + 1] \
                - self.value_preds[step]
            gae = delta + gamma * tau * self.masks[step + 1] * gae
            self.returns[step] = torch.where(is_end, self.returns[step], gae + self.value_preds[step])
            gae = torch.where(is_end, torch.zeros_like(gae), gae)

    def feed_forward_generator(self, advantages, num_mini_batch, sampler=None):
        num_steps, num_processes = self.rewards.size()[0:2]
        batch_size = num_processes * num_steps
        if batch_size < num_mini_batch:
            raise AssertionError("PPO requires num_processes (%d) * num_steps (%d) >= num_mini_batch (%d)"
                                 % (num_processes, num_steps, num_mini_batch))
        mini_batch_size = batch_size // num_mini_batch
        if sampler is None:
            perm = torch.randperm(batch_size)
            sampler = [perm[i:i + mini_batch_size].tolist() for i in range(0, batch_size, mini_batch_size)]
        flat = lambda t: t.view(-1, t.size(-1))
        for indices in sampler:
            yield (flat(self.obs[:-1])[indices], flat(self.recurrent_hidden_states[:-1])[indices],
                   flat(self.actions)[indices], flat(self.value_preds[:-1])[indices], flat(self.returns[:-1])[indices],
                   flat(self.masks[:-1])[indices], flat(self.action_log_probs)[indices],
                   advantages.view(-1, 1)[indices])
