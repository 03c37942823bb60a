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


@pytest.mark.parametrize("name", ["simple_spread", "simple_tag"])
@pytest.mark.parametrize("dtype,atol", [(torch.float64, 1e-10), (torch.float32, 2e-4)])
def test_scenario_env_matches_reference(name, dtype, atol):
    """MultiAgentEnvBatch.step (set_action -> mw_step -> mw_scenario_callbacks) against the transitions recorded from the
    reference's MultiAgentEnv with its own simple_spread / simple_tag scenarios (tests/golden/mape_scenarios.npz)."""
    from importlib import import_module
    mwm = import_module("emergent-multiagent-strategies_b200.mape_world")
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mape_scenarios.npz"))
    T, E = g[name + "/act"].shape[:2]
    env = mwm.MultiAgentEnvBatch(name, E, dtype=dtype)
    assert env.n == int(g[name + "/na"]) and env.obs_dims == [int(d) for d in g[name + "/obs_dims"]]
    assert env.shared_reward == bool(g[name + "/shared"])
    flips = 0
    for t in range(T):
        env.world.pos.copy_(torch.from_numpy(g[name + "/pos_before"][t]).transpose(0, 1).to(dtype))      # teacher-forced
        env.world.vel.copy_(torch.from_numpy(g[name + "/vel_before"][t]).transpose(0, 1).to(dtype))
        act = g[name + "/act"][t]
        obs_n, rew_n, done_n, info = env.step([torch.from_numpy(act[:, i]) for i in range(env.n)])
        assert np.allclose(env.world.pos.transpose(0, 1).cpu().numpy(), g[name + "/pos_after"][t], rtol=0, atol=atol)
        for i in range(env.n):
            assert tuple(obs_n[i].shape) == (E, env.obs_dims[i])
            assert np.allclose(obs_n[i].cpu().numpy(), g["%s/obs%d" % (name, i)][t], rtol=0, atol=atol * 10), (t, i)
            d = np.abs(rew_n[i].cpu().numpy() - g[name + "/rew"][t][:, i])
            # a contact / catch decided within float rounding of the threshold moves a reward by a whole unit: count, do not hide
            flips += int((d > 0.5).sum())
            assert (d[d <= 0.5] < atol * 50).all(), (t, i, d.max())
        assert not any(bool(x.any()) for x in done_n)
    assert flips == 0 if dtype == torch.float64 else flips <= 2


def test_scenario_reset_and_discrete_actions():
    from importlib import import_module
    mwm = import_module("emergent-multiagent-strategies_b200.mape_world")
    env = mwm.MultiAgentEnvBatch("simple_tag", 4096, discrete_action_input=True, seed=3)
    obs = env.reset()
    p = env.world.pos
    assert float(p[:4].abs().max()) < 1.0 and float(p[4:].abs().max()) < 0.9 and float(p[4:].abs().max()) > 0.85
    assert abs(float(p[:4].mean())) < 0.02 and abs(float(p[:4].var()) - 1 / 3) < 0.02 and float(env.world.vel.abs().max()) == 0.0
    assert torch.equal(obs[0][:, 2:4], p[0])
    before = p.clone()
    mask = torch.zeros(4096, dtype=torch.uint8, device="cuda"); mask[::2] = 1
    env.reset(mask)
    assert torch.equal(p[:, 1::2], before[:, 1::2]) and not torch.equal(p[:, ::2], before[:, ::2])
    # discrete input: 1 -> -x, 2 -> +x, 3 -> -y, 4 -> +y (environment.py:163-168), times accel 4
    u = env.set_action([torch.full((4096,), k, device="cuda") for k in (1, 2, 3, 4)])
    assert torch.equal(u[:, 0].cpu(), torch.tensor([[-4.0, 0.0], [4.0, 0.0], [0.0, -4.0], [0.0, 4.0]]))
