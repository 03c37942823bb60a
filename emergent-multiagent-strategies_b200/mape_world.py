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
