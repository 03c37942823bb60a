"""Gym-style single-environment facade over the CUDA engine -- the reference's env seam.

Mirrors gym_fortattack/fortattack.py:17-27,31-302 as its callers see it:
    env = make_fortattack_env(num_steps)
    env.n, env.action_space[i].n, env.observation_space[i].shape, env.action_spaces[i].shape, env.ob_rms
    env.world.policy_agents[i].attacker, env.world.numGuards / numAttackers / numAliveGuards /
        numAliveAttackers / gameResult / max_time_steps / time_step
    obs = env.reset()                        -> float64 ndarray [A, 6]
    obs, reward, done, info = env.step(a)    -> ndarray [A, 6], list of A floats, bool, {'n': [{}]*A}
so learner.setup_master / Learner / train_fortattack.train run on it unchanged (learner.py:21-70,137;
train_fortattack.py:25-29,60,100,124).  One env = a batch of E=1 on the GPU: every step is one
fa_step_host call (kernel reads the actions from, and writes obs/reward/done into, pinned host memory).
The reference hard-codes 5 guards v 5 attackers (fortattack_env_v1.py:18-19); here they are arguments.
"""
import os

import numpy as np
import torch

try:
    from ..batched_env import FortAttackBatch
    from .. import render as _render
except ImportError:      # drop-in mode: this package was imported top-level as `gym_fortattack`
    import importlib
    import sys
    _root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    if _root not in sys.path:
        sys.path.insert(0, _root)
    FortAttackBatch = importlib.import_module("emergent-multiagent-strategies_b200").FortAttackBatch
    _render = importlib.import_module("emergent-multiagent-strategies_b200.render")


class Discrete(object):
    """gym.spaces.Discrete as used by the callers (learner.py:34 -> rlagent.py:13)."""

    def __init__(self, n):
        self.n, self.shape, self.dtype = n, (), np.int64

    def __repr__(self):
        return "Discrete(%d)" % self.n


class Box(object):
    """gym.spaces.Box / malib.spaces.Box: only .shape/.low/.high are read (learner.py:49-67, mpnn.py:73)."""

    def __init__(self, low, high, shape, dtype=np.float32):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype

    def __repr__(self):
        return "Box%r" % (self.shape,)


class _Agent(object):
    def __init__(self, world, index, attacker):
        self._world, self.index, self.attacker = world, index, attacker
        self.name = "agent %d" % index
        self.movable, self.silent, self.collide = True, True, True

    @property
    def alive(self):
        return bool(self._world._env._last_obs[self.index, 0] > 0)


class _World(object):
    """The attributes of gym_fortattack.core.World that code above the env reads (core.py:109-128)."""
    dim_p, dim_c, dt = 3, 0, 0.1           # core.py:115-121 (dim_p = 3: x, y, rotation -> 2*3+2 = 8 actions)

    def __init__(self, env, n_guards, n_attackers):
        self._env = env
        self.numGuards, self.numAttackers = n_guards, n_attackers
        self.numAgents = n_guards + n_attackers
        self.agents = [_Agent(self, i, i >= n_guards) for i in range(self.numAgents)]
        self.gameResult = np.array([0, 0, 0])        # [all attackers dead, time up, attacker reached] fortattack.py:205-222
        self.fortDim, self.doorLoc = 0.15, np.array([0, 0.8])
        self.wall_pos = [-1, 1, -0.8, 0.8]
        self.vizDead, self.vizAttn = False, True     # fortattack_env_v1.py:42-43 (read by render)

    @property
    def policy_agents(self):
        return self.agents

    @property
    def max_time_steps(self):
        return self._env._batch.max_steps

    @max_time_steps.setter
    def max_time_steps(self, v):
        self._env._batch.set_max_steps(v)

    # world.time_step / numAliveGuards / numAliveAttackers (core.py:123-126, fortattack.py:171): host-side mirrors of what
    # the last step returned -- no launch, no device->host copy per read (test_fortattack_v2.py:94-99 reads them per episode)
    @property
    def time_step(self):
        return self._env._time_step

    @property
    def numAliveGuards(self):
        return int((self._env._last_obs[:self.numGuards, 0] > 0).sum())

    @property
    def numAliveAttackers(self):
        return int((self._env._last_obs[self.numGuards:, 0] > 0).sum())


class FortAttackGlobalEnv(object):
    metadata = {"render.modes": ["human", "rgb_array"]}

    def __init__(self, num_steps, n_guards=5, n_attackers=5, seed=0, device="cuda:0", dtype=torch.float32):
        self.ob_rms = None                                         # fortattack.py:42
        self._cfg = dict(n_guards=n_guards, n_attackers=n_attackers, device=device, dtype=dtype)
        self._num_steps, self._seed = num_steps, seed
        self._make_batch()
        self.world = _World(self, n_guards, n_attackers)
        self.agents = self.world.policy_agents
        self.n = self.agent_num = len(self.agents)
        self.discrete_action_space = self.discrete_action_input = True
        self.shared_reward = False                                 # fortattack.py:61
        n_act = self.world.dim_p * 2 + 2
        self.action_space = [Discrete(n_act) for _ in range(self.n)]                            # fortattack.py:73,94
        self.observation_space = [Box(-np.inf, np.inf, (6,)) for _ in range(self.n)]           # fortattack.py:96-98
        self.action_spaces = tuple(Box(0., 1., (n_act,)) for _ in range(self.n))               # MASpace, fortattack.py:107
        self.observation_spaces = tuple(Box(-np.inf, np.inf, (6,)) for _ in range(self.n))
        self.action_range = [0., 1.]
        self._last_obs = np.zeros((self.n, 6))
        self._time_step = 0
        self._last_actions = np.zeros(self.n, np.int32)            # agent.action.shoot of the last step (for render)
        self.reset()                                               # the reference's constructor resets once (v1:45)

    def _make_batch(self):
        c = self._cfg
        self._batch = FortAttackBatch(1, c["n_guards"], c["n_attackers"], max_steps=self._num_steps, seed=self._seed,
                                      device=c["device"], dtype=c["dtype"])
        self._bufs = self._batch.make_host_buffers()

    def seed(self, seed=None):
        """The reference's env ignores seeds (gym.Env.seed default; resets use np.random).  Here the reset
        streams are keyed by it."""
        if seed is not None and seed != self._seed:
            cap = self._batch.max_steps
            self._seed = int(seed)
            self._make_batch()
            self._batch.set_max_steps(cap)
        return [self._seed]

    def reset(self):
        obs = self._batch.reset()                                  # [A,1,6] on device
        self.world.gameResult[:] = 0                               # v1:57
        self._last_obs = obs[:, 0].double().cpu().numpy()
        self._last_actions = np.zeros(self.n, np.int32)
        self._time_step = 0                                        # v1:51
        return self._last_obs.copy()

    def step(self, action_n):
        act = np.asarray(action_n).reshape(-1)
        if act.shape[0] != self.n:
            raise ValueError("expected %d actions, got %d" % (self.n, act.shape[0]))
        h_act, h_obs, h_rew, h_done, h_res = self._bufs
        h_act[:, 0] = torch.from_numpy(act.astype(np.int32))
        self._last_actions = act.astype(np.int32)
        self._batch.step_host(h_act, h_obs, h_rew, h_done, h_res, auto_reset=False)
        self._last_obs = h_obs[:, 0].double().numpy()
        self._time_step += 1                                       # fortattack.py:171
        done = bool(h_done[0])
        res = int(h_res[0])
        if res:                                                    # fortattack.py:205-222
            self.world.gameResult[{1: 0, 2: 1, 3: 2}[res]] = 1
        return self._last_obs.copy(), [float(r) for r in h_rew[:, 0].tolist()], done, {"n": [{} for _ in range(self.n)]}

    def render(self, attn_list=None, mode="human", close=False):
        """The scene of the reference's render (fortattack.py:368-596) rasterised on the device from the current state
        (render.render_batch / fr_render): no window is opened.  mode='rgb_array' returns [uint8 array 700 x 700 x 3] (the
        reference returns one entry per viewer, :583-591); any other mode returns [True] after drawing into
        `self.last_frame`.  attn_list = [[team_attn, opp_attn], ...] as the reference's callers pass it
        (train_fortattack_v2.py:69): the guards' matrices become the yellow attention halos (:441-466)."""
        attention_halos, render_batch = _render.attention_halos, _render.render_batch
        obs = torch.from_numpy(self._last_obs).to(device=self._batch.device, dtype=torch.float32).view(self.n, 1, 6).contiguous()
        act = torch.as_tensor(self._last_actions, dtype=torch.int32, device=obs.device).view(self.n, 1).contiguous()
        halo = None
        ng = self.world.numGuards
        if attn_list is not None and self.world.vizAttn:
            team = torch.as_tensor(np.asarray(attn_list[0][0]), dtype=torch.float32, device=obs.device).reshape(1, ng, ng)
            opp = torch.as_tensor(np.asarray(attn_list[0][1]), dtype=torch.float32, device=obs.device).reshape(1, ng, self.n - ng)
            halo = attention_halos(obs, ng, team, opp)
        self.last_frame = render_batch(obs, ng, actions=act, halo=halo, draw_dead=bool(self.world.vizDead))[0].cpu().numpy()
        return [self.last_frame] if mode == "rgb_array" else [True]

    def terminate(self):
        pass

    def close(self):
        self._batch.close()


def make_fortattack_env(num_steps, benchmark=False, n_guards=None, n_attackers=None, seed=0, device="cuda:0"):
    """make_fortattack_env(num_steps) of the reference (fortattack.py:17-27): world.max_time_steps = num_steps.
    Team sizes default to the reference's hard-coded 5v5 (fortattack_env_v1.py:18-19); callers that cannot pass
    arguments (the reference's own scripts call make_fortattack_env(num_steps), utils.py:24) set FORTATTACK_TEAMS=3v3."""
    if n_guards is None or n_attackers is None:
        g, a = (int(x) for x in os.environ.get("FORTATTACK_TEAMS", "5v5").lower().split("v"))
        n_guards = g if n_guards is None else n_guards
        n_attackers = a if n_attackers is None else n_attackers
    return FortAttackGlobalEnv(num_steps, n_guards=n_guards, n_attackers=n_attackers, seed=seed, device=device)
