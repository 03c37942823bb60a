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
