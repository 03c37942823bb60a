"""Golden transitions of the reference's GENERIC particle world (multiagent/core.py:118-225 World.step: action force,
pairwise contact force, wall force, integration) for three entity configurations.  Imports the unchanged reference
from /root/reference (numpy only); run in the build container:  python tests/golden/make_mw_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install_stubs()
sys.path.insert(0, ref_shim.REF)
from multiagent.core import Agent, Landmark, World  # noqa: E402

CONFIGS = {
    # name: (agents [(size, mass, max_speed, collide, movable)], landmarks [...])
    "spread": ([(0.15, 1.0, None, True, True)] * 3, [(0.05, 1.0, None, False, False)] * 3),
    "tag": ([(0.075, 1.0, 1.0, True, True)] * 3 + [(0.05, 1.0, 1.3, True, True)], [(0.2, 1.0, None, True, False)] * 2),
    "push": ([(0.1, 1.0, None, True, True)] * 2, [(0.12, 2.0, 0.8, True, True), (0.08, 1.0, None, True, False)]),
}


def make_world(agents, landmarks):
    w = World()
    w.agents = [Agent() for _ in agents]
    w.landmarks = [Landmark() for _ in landmarks]
    for ent, (size, mass, ms, col, mov) in zip(w.agents + w.landmarks, list(agents) + list(landmarks)):
        ent.size, ent.initial_mass, ent.max_speed, ent.collide, ent.movable = size, mass, ms, col, mov
    for a in w.agents:
        a.silent = True
        a.action.c = np.zeros(0)
    return w


def main():
    out = {}
    E, T = 24, 30
    for name, (agents, landmarks) in CONFIGS.items():
        rng = np.random.RandomState({"spread": 11, "tag": 12, "push": 13}[name])
        na, ne = len(agents), len(agents) + len(landmarks)
        pos0 = rng.uniform(-1, 1, (E, ne, 2))
        pos0[: E // 4, :na] *= 1.05                      # some agents start in contact with a wall
        vel0 = rng.uniform(-0.5, 0.5, (E, ne, 2)) * np.array([[[1.0 if m else 0.0] for (_, _, _, _, m) in list(agents) + list(landmarks)]])
        u = rng.uniform(-1, 1, (T, E, na, 2)) * rng.choice([0.0, 1.0, 3.0, 4.0], size=(T, E, na, 1))
        pos, vel = np.zeros((T, E, ne, 2)), np.zeros((T, E, ne, 2))
        for e in range(E):
            w = make_world(agents, landmarks)
            for k, ent in enumerate(w.entities):
                ent.state.p_pos, ent.state.p_vel = pos0[e, k].copy(), vel0[e, k].copy()
            for t in range(T):
                for i, a in enumerate(w.agents):
                    a.action.u = u[t, e, i].copy()
                w.step()
                for k, ent in enumerate(w.entities):
                    pos[t, e, k], vel[t, e, k] = ent.state.p_pos, ent.state.p_vel
        cfg = np.array([[s, m, -1.0 if ms is None else ms, float(c), float(mv)] for (s, m, ms, c, mv) in list(agents) + list(landmarks)])
        for k, v in (("cfg", cfg), ("na", np.array(na)), ("pos0", pos0), ("vel0", vel0), ("u", u), ("pos", pos), ("vel", vel),
                     ("world", np.array([w.dt, w.damping, w.contact_force, w.contact_margin] + list(w.wall_pos)))):
            out["%s/%s" % (name, k)] = v
    np.savez_compressed(os.path.join(HERE, "mape_world.npz"), **out)
    print("wrote", os.path.join(HERE, "mape_world.npz"), {k: v.shape for k, v in out.items() if k.startswith("spread")})


if __name__ == "__main__":
    main()
