"""Batched generic particle world: World.step of the reference's multi-agent particle environment
(multiagent/core.py:118-225) for E worlds that share one entity configuration, through mw_step (include/mape_world.h).

    w = MapeWorldBatch(4096, agents=[dict(size=0.15)] * 3, landmarks=[dict(size=0.05, collide=False)] * 3)
    w.pos[:] = ...; w.vel.zero_()          # [n_entities, E, 2] device tensors, agents first (scenario reset_world)
    w.step(u)                              # u [n_agents, E, 2] = agent.action.u (environment.py:_set_action)
    reward = -(w.pos[:3, :, None] - w.pos[None, 3:]).norm(dim=-1).min(0)...   # scenario callbacks are tensor code

Entity defaults are the reference's (core.py:29-75): size 0.05, mass 1, max_speed None, collide True; agents movable,
landmarks not.  No CPU path."""
import ctypes

import torch

from . import _capi

MW_MAX_ENTITIES = 12


class MwConfig(ctypes.Structure):
    _fields_ = [("n_envs", ctypes.c_int32), ("n_agents", ctypes.c_int32), ("n_entities", ctypes.c_int32),
                ("scalar", ctypes.c_int32), ("dt", ctypes.c_double), ("damping", ctypes.c_double),
                ("contact_force", ctypes.c_double), ("contact_margin", ctypes.c_double), ("wall", ctypes.c_double * 4),
                ("size", ctypes.c_double * MW_MAX_ENTITIES), ("mass", ctypes.c_double * MW_MAX_ENTITIES),
                ("max_speed", ctypes.c_double * MW_MAX_ENTITIES), ("collide", ctypes.c_uint8 * MW_MAX_ENTITIES),
                ("movable", ctypes.c_uint8 * MW_MAX_ENTITIES)]


class MapeWorldBatch(object):
    def __init__(self, n_envs, agents, landmarks=(), device="cuda:0", dtype=torch.float32, dt=0.1, damping=0.25,
                 contact_force=1e2, contact_margin=1e-10, wall_pos=(-1.0, 1.0, -1.0, 1.0)):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise _capi.FaError("MapeWorldBatch needs a CUDA device; there is no CPU path")
        ents = [dict(dict(size=0.05, mass=1.0, max_speed=None, collide=True, movable=True), **a) for a in agents] + \
               [dict(dict(size=0.05, mass=1.0, max_speed=None, collide=True, movable=False), **l) for l in landmarks]
        if not 1 <= len(agents) <= len(ents) <= MW_MAX_ENTITIES:
            raise ValueError("need 1 <= n_agents <= n_entities <= %d" % MW_MAX_ENTITIES)
        self.device, self.dtype, self.E = dev, dtype, int(n_envs)
        self.n_agents, self.n_entities = len(agents), len(ents)
        c = MwConfig(self.E, self.n_agents, self.n_entities, 1 if dtype == torch.float64 else 0, dt, damping,
                     contact_force, contact_margin)
        c.wall[:] = list(wall_pos)
        for i, e in enumerate(ents):
            c.size[i], c.mass[i] = e["size"], e["mass"]
            c.max_speed[i] = -1.0 if e["max_speed"] is None else e["max_speed"]
            c.collide[i], c.movable[i] = int(bool(e["collide"])), int(bool(e["movable"]))
        self.cfg = c
        self.pos = torch.zeros(self.n_entities, self.E, 2, device=dev, dtype=dtype)
        self.vel = torch.zeros(self.n_entities, self.E, 2, device=dev, dtype=dtype)
        self._lib = _capi.lib()
        self.launches = 0

    def step(self, u):
        """u [n_agents, E, 2]: action forces.  Updates self.pos / self.vel in place."""
        if u.shape != (self.n_agents, self.E, 2) or u.dtype != self.dtype or not u.is_contiguous() or u.device != self.pos.device:
            raise ValueError("u must be a contiguous %s tensor [%d, %d, 2] on %s" % (self.dtype, self.n_agents, self.E, self.device))
        _capi.check(self._lib.mw_step(ctypes.byref(self.cfg), self.pos.data_ptr(), self.vel.data_ptr(), u.data_ptr(),
                                      torch.cuda.current_stream(self.device).cuda_stream))
        self.launches += 1


# ---- scenarios: the reference's MultiAgentEnv surface (multiagent/environment.py:9-140) for E worlds ---------------------
SCENARIOS = {
    # name: (scenario id, agents, landmarks, n_adversaries, agent accel (None -> sensitivity 5.0), reset boxes, shared reward)
    # multiagent/scenarios/simple_spread.py:8-30
    "simple_spread": dict(sid=0, agents=[dict(size=0.15)] * 3, landmarks=[dict(collide=False)] * 3, n_adv=0, accel=[None] * 3,
                          boxes=(-1.0, 1.0, -1.0, 1.0), shared=True),
    # multiagent/scenarios/simple_tag.py:8-40 (3 adversaries of size 0.075 chase 1 agent of size 0.05; accel 4, max_speed 1.3)
    "simple_tag": dict(sid=1, agents=[dict(size=0.075, max_speed=1.3)] * 3 + [dict(size=0.05, max_speed=1.3)],
                       landmarks=[dict(size=0.2)] * 2, n_adv=3, accel=[4.0] * 4, boxes=(-1.0, 1.0, -0.9, 0.9), shared=False),
}


class MultiAgentEnvBatch(object):
    """MultiAgentEnv(world, scenario.reset_world, scenario.reward, scenario.observation) of the reference for E worlds on the
    GPU: `reset()` -> obs_n, `step(action_n)` -> (obs_n, reward_n, done_n, info_n), with obs_n[i] a device tensor [E, obs_dim_i]
    and reward_n[i] [E].  action_n[i]: float [E, 5] action vectors (the reference's default, discrete_action_input False:
    u = (a[1] - a[2], a[3] - a[4]) * sensitivity, environment.py:171-177) or int [E] when discrete_action_input is set
    (1/2 -> -x/+x, 3/4 -> -y/+y, :163-168).  World.step is mw_step, the callbacks one launch of mw_scenario_callbacks."""

    def __init__(self, scenario, n_envs, device="cuda:0", dtype=torch.float32, seed=0, discrete_action_input=False):
        sc = SCENARIOS[scenario]
        self.sc, self.scenario = sc, scenario
        self.world = MapeWorldBatch(n_envs, sc["agents"], sc["landmarks"], device=device, dtype=dtype)
        w = self.world
        self.n, self.E, self.seed, self.episode = w.n_agents, w.E, int(seed), 0
        self.discrete_action_input = bool(discrete_action_input)
        self.shared_reward = sc["shared"]
        L = w._lib
        self.obs_stride = L.mw_scenario_obs_dim(sc["sid"], w.n_agents, w.n_entities, sc["n_adv"])
        n_good = w.n_agents - sc["n_adv"]
        base = 4 + 2 * (w.n_entities - w.n_agents) + 2 * (w.n_agents - 1)
        # per-agent observation length: simple_tag's good agents do not see their own velocity again (simple_tag.py:176-178)
        self.obs_dims = [self.obs_stride] * w.n_agents if sc["sid"] == 0 else \
            [base + 2 * (n_good - (0 if i < sc["n_adv"] else 1)) for i in range(w.n_agents)]
        self.sens = torch.tensor([5.0 if a is None else a for a in sc["accel"]], device=w.device, dtype=dtype).view(-1, 1, 1)
        self._obs = torch.zeros(w.n_agents, w.E, self.obs_stride, device=w.device, dtype=dtype)
        self._rew = torch.zeros(w.n_agents, w.E, device=w.device, dtype=dtype)

    def _callbacks(self):
        w = self.world
        _capi.check(w._lib.mw_scenario_callbacks(ctypes.byref(w.cfg), self.sc["sid"], self.sc["n_adv"], w.pos.data_ptr(), w.vel.data_ptr(),
                                                 self._obs.data_ptr(), self.obs_stride, self._rew.data_ptr(),
                                                 torch.cuda.current_stream(w.device).cuda_stream))
        return [self._obs[i, :, :d] for i, d in enumerate(self.obs_dims)], [self._rew[i] for i in range(self.n)]

    def reset(self, mask=None):
        """scenario.reset_world for every world (or those with mask[e] != 0) -> obs_n."""
        w, b = self.world, self.sc["boxes"]
        if mask is not None and (mask.dtype != torch.uint8 or mask.numel() != w.E or not mask.is_contiguous() or mask.device != w.pos.device):
            raise ValueError("mask must be a contiguous uint8 [E] tensor on %s" % w.device)
        _capi.check(w._lib.mw_scenario_reset(ctypes.byref(w.cfg), b[0], b[1], b[2], b[3], self.seed, self.episode,
                                             None if mask is None else mask.data_ptr(), w.pos.data_ptr(), w.vel.data_ptr(),
                                             torch.cuda.current_stream(w.device).cuda_stream))
        self.episode += 1
        return self._callbacks()[0]

    def set_action(self, action_n):
        """environment.py:_set_action for every agent -> u [n, E, 2]."""
        w = self.world
        if self.discrete_action_input:
            a = torch.stack([torch.as_tensor(x, device=w.device).long().view(w.E) for x in action_n])            # [n, E]
            u = torch.stack(((a == 2).to(w.dtype) - (a == 1).to(w.dtype), (a == 4).to(w.dtype) - (a == 3).to(w.dtype)), dim=-1)
        else:
            a = torch.stack([torch.as_tensor(x, device=w.device, dtype=w.dtype).view(w.E, 5) for x in action_n])  # [n, E, 5]
            u = torch.stack((a[..., 1] - a[..., 2], a[..., 3] - a[..., 4]), dim=-1)
        return (u * self.sens).contiguous()

    def step(self, action_n):
        self.world.step(self.set_action(action_n))
        obs_n, reward_n = self._callbacks()
        done_n = [torch.zeros(self.E, dtype=torch.bool, device=self.world.device)] * self.n      # no done_callback (environment.py:131-134)
        return obs_n, reward_n, done_n, {"n": [{}] * self.n}
