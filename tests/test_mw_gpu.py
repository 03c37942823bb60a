"""GPU: the batched generic particle world (mw_step) against the reference's recorded transitions and the float64 oracle."""
import os
from importlib import import_module

import numpy as np
import pytest
import torch

import mw_oracle

pytestmark = pytest.mark.gpu
PKG = "emergent-multiagent-strategies_b200"
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mape_world.npz")


def _world(cfg, na, world, E, dtype):
    mw = import_module(PKG + ".mape_world")
    ents = [dict(size=float(s), mass=float(m), max_speed=None if ms < 0 else float(ms), collide=bool(c), movable=bool(mv))
            for s, m, ms, c, mv in cfg]
    return mw.MapeWorldBatch(E, ents[:na], ents[na:], dtype=dtype, dt=world[0], damping=world[1], contact_force=world[2],
                             contact_margin=world[3], wall_pos=tuple(world[4:]))


def _dev(a, dtype):          # [E, N, 2] -> entity-major device tensor
    return torch.from_numpy(np.ascontiguousarray(a.transpose(1, 0, 2))).to("cuda", dtype)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("name", ["spread", "tag", "push"])
def test_golden_world_steps(name, dtype):
    """Teacher-forced on the reference's own states: every recorded World.step is reproduced."""
    g = np.load(GOLD)
    cfg, na, world = g[name + "/cfg"], int(g[name + "/na"]), g[name + "/world"]
    E = g[name + "/pos0"].shape[0]
    w = _world(cfg, na, world, E, dtype)
    pos, vel = g[name + "/pos0"], g[name + "/vel0"]
    tol = dict(rtol=0, atol=1e-10) if dtype == torch.float64 else dict(rtol=1e-5, atol=1e-5)
    for t in range(g[name + "/u"].shape[0]):
        w.pos.copy_(_dev(pos, dtype)); w.vel.copy_(_dev(vel, dtype))
        w.step(_dev(g[name + "/u"][t], dtype))
        p, v = w.pos.cpu().numpy().transpose(1, 0, 2), w.vel.cpu().numpy().transpose(1, 0, 2)
        # contact forces divide by the centre distance: float32 rounding of the state is amplified by 100 * dt / dist
        slack = 0.0
        if dtype == torch.float32:
            d = np.linalg.norm(pos[:, :, None] - pos[:, None], axis=-1) + np.eye(pos.shape[1])[None]
            slack = 2e-6 / d.min(axis=(1, 2))[:, None, None]
        assert (np.abs(p - g[name + "/pos"][t]) <= tol["atol"] + tol["rtol"] * np.abs(g[name + "/pos"][t]) + 0.1 * slack).all(), (name, t)
        assert (np.abs(v - g[name + "/vel"][t]) <= tol["atol"] + tol["rtol"] * np.abs(g[name + "/vel"][t]) + slack).all(), (name, t)
        pos, vel = g[name + "/pos"][t], g[name + "/vel"][t]


def test_large_batch_matches_oracle_and_is_loud():
    """4096 simple_tag-like worlds x 50 free-running steps in double track the oracle; bad arguments raise."""
    g = np.load(GOLD)
    cfg, na, world = g["tag/cfg"], int(g["tag/na"]), g["tag/world"]
    E, T = 4096, 50
    rng = np.random.RandomState(0)
    pos = rng.uniform(-1, 1, (E, cfg.shape[0], 2)); vel = np.zeros_like(pos)
    w = _world(cfg, na, world, E, torch.float64)
    w.pos.copy_(_dev(pos, torch.float64)); w.vel.zero_()
    for t in range(T):
        u = rng.uniform(-3, 3, (E, na, 2))
        pos, vel = mw_oracle.step(pos, vel, u, cfg, na, world)
        w.step(_dev(u, torch.float64))
    p, v = w.pos.cpu().numpy().transpose(1, 0, 2), w.vel.cpu().numpy().transpose(1, 0, 2)
    ok = np.isfinite(pos).all(axis=(1, 2))
    assert ok.mean() > 0.99 and np.allclose(p[ok], pos[ok], atol=1e-6) and np.allclose(v[ok], vel[ok], atol=1e-5)
    assert w.launches == T
    mw = import_module(PKG + ".mape_world")
    with pytest.raises(Exception):
        mw.MapeWorldBatch(4, [dict()] * 13)
    with pytest.raises(Exception):
        w.step(torch.zeros(1, 2, 2, device="cuda"))
    with pytest.raises(Exception):
        mw.MapeWorldBatch(4, [dict()], device="cpu")
