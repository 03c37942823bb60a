"""CPU: the generic particle-world oracle (oracle/mw_oracle.py) against transitions recorded from the unchanged
reference's multiagent/core.py World.step (tests/golden/mape_world.npz)."""
import os

import numpy as np
import pytest

import mw_oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mape_world.npz")


@pytest.mark.parametrize("name", ["spread", "tag", "push"])
def test_oracle_reproduces_reference_world_step(name):
    g = np.load(GOLD)
    cfg, na, world = g[name + "/cfg"], int(g[name + "/na"]), g[name + "/world"]
    pos, vel = g[name + "/pos0"], g[name + "/vel0"]
    T = g[name + "/u"].shape[0]
    contacts = 0
    for t in range(T):
        # teacher-forced: every step starts from the reference's own previous state
        p, v = mw_oracle.step(pos, vel, g[name + "/u"][t], cfg, na, world)
        assert np.allclose(p, g[name + "/pos"][t], rtol=0, atol=1e-12), (name, t)
        assert np.allclose(v, g[name + "/vel"][t], rtol=0, atol=1e-11), (name, t)
        contacts += int((np.abs(v - vel * 0.75).max(axis=(1, 2)) > 0.5).sum())
        pos, vel = g[name + "/pos"][t], g[name + "/vel"][t]
    assert contacts > 0                                    # the fixture exercises contact / wall forces
    # free-running from the initial state also tracks the reference (same float64 arithmetic, same order)
    pos, vel = g[name + "/pos0"], g[name + "/vel0"]
    for t in range(T):
        pos, vel = mw_oracle.step(pos, vel, g[name + "/u"][t], cfg, na, world)
    assert np.allclose(pos, g[name + "/pos"][-1], atol=1e-7) and np.allclose(vel, g[name + "/vel"][-1], atol=1e-6)


SCN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mape_scenarios.npz")
# entity configurations of the two scenarios (multiagent/scenarios/simple_spread.py:8-30, simple_tag.py:8-40) as mw_oracle.step
# takes them: size, mass, max_speed (< 0: None), collide, movable; world = dt, damping, contact force / margin, walls
SCN_CFG = {
    "simple_spread": (np.array([[0.15, 1, -1, 1, 1]] * 3 + [[0.05, 1, -1, 0, 0]] * 3, dtype=float), 0, [5.0] * 3),
    "simple_tag": (np.array([[0.075, 1, 1.3, 1, 1]] * 3 + [[0.05, 1, 1.3, 1, 1]] + [[0.2, 1, -1, 1, 0]] * 2, dtype=float), 3, [4.0] * 4),
}
SCN_WORLD = np.array([0.1, 0.25, 1e2, 1e-10, -1.0, 1.0, -1.0, 1.0])


def scenario_step(name, g, t):
    """One MultiAgentEnv.step of golden record t through the oracle: (pos, vel, obs list, reward [E, na])."""
    cfg, n_adv, sens = SCN_CFG[name]
    na = int(g[name + "/na"])
    act = g[name + "/act"][t]                                                       # [E, na, 5]
    u = np.stack((act[..., 1] - act[..., 2], act[..., 3] - act[..., 4]), axis=-1) * np.array(sens)[None, :, None]
    pos, vel = mw_oracle.step(g[name + "/pos_before"][t], g[name + "/vel_before"][t], u, cfg, na, SCN_WORLD)
    if name == "simple_spread":
        obs, rew = mw_oracle.spread_callbacks(pos, vel, cfg[:, 0], na)
        obs = [obs[:, i] for i in range(na)]
    else:
        obs, rew = mw_oracle.tag_callbacks(pos, vel, cfg[:, 0], na, n_adv)
    return pos, vel, obs, rew


@pytest.mark.parametrize("name", ["simple_spread", "simple_tag"])
def test_oracle_reproduces_reference_scenarios(name):
    """_set_action + World.step + scenario.observation / reward (+ the shared sum) of the reference's MultiAgentEnv."""
    g = np.load(SCN)
    assert np.array_equal(g[name + "/size"], SCN_CFG[name][0][:, 0])
    T, events = g[name + "/act"].shape[0], 0
    for t in range(T):
        pos, vel, obs, rew = scenario_step(name, g, t)
        assert np.allclose(pos, g[name + "/pos_after"][t], rtol=0, atol=1e-12) and np.allclose(vel, g[name + "/vel_after"][t], rtol=0, atol=1e-11)
        for i, o in enumerate(obs):
            assert o.shape[1] == int(g[name + "/obs_dims"][i])
            assert np.allclose(o, g["%s/obs%d" % (name, i)][t], rtol=0, atol=1e-11), (name, t, i)
        assert np.allclose(rew, g[name + "/rew"][t], rtol=0, atol=1e-10), (name, t)
        events += int((np.abs(g[name + "/rew"][t]) >= 9.9).sum()) if name == "simple_tag" else int((g[name + "/rew"][t][:, 0] < -9).sum())
    assert events > 0                                   # catches / collisions are exercised
