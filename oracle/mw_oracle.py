"""CPU restatement of the reference's generic particle-world step (multiagent/core.py:118-225), vectorised over
environments in numpy float64.  TEST INFRASTRUCTURE, NOT PRODUCT: only tests/ may import it.  Pinned to the unchanged
reference by tests/golden/mape_world.npz (tests/golden/make_mw_golden.py).

cfg [NE, 5] = size, mass, max_speed (< 0: None), collide, movable; agents first (the first `na` entities).
world = dt, damping, contact_force, contact_margin, xmin, xmax, ymin, ymax."""
import numpy as np


def _softplus_pen(x, k):
    # np.logaddexp(0, -x / k) * k                                   core.py:205,221
    return np.logaddexp(0.0, -x / k) * k


def step(pos, vel, u, cfg, na, world):
    """pos, vel [E, NE, 2] (updated copies are returned), u [E, na, 2]."""
    dt, damping, cf, k, xmin, xmax, ymin, ymax = [float(x) for x in world]
    size, mass, max_speed, collide, movable = cfg[:, 0], cfg[:, 1], cfg[:, 2], cfg[:, 3] != 0, cfg[:, 4] != 0
    E, NE, _ = pos.shape
    force = np.zeros((E, NE, 2))
    has = np.zeros(NE, bool)
    for i in range(na):                                             # apply_action_force          core.py:139-145
        if movable[i]:
            force[:, i] = u[:, i]
            has[i] = True
    for a in range(NE):                                             # apply_environment_force      core.py:148-160
        for b in range(a + 1, NE):
            if not (collide[a] and collide[b]):
                continue
            d = pos[:, a] - pos[:, b]
            dist = np.sqrt(np.sum(np.square(d), axis=1))
            pen = _softplus_pen(dist - (size[a] + size[b]), k)
            with np.errstate(divide="ignore", invalid="ignore"):
                f = cf * d / dist[:, None] * pen[:, None]
            if movable[a]:
                force[:, a] = f + force[:, a]
                has[a] = True
            if movable[b]:
                force[:, b] = -f + force[:, b]
                has[b] = True
    for a in range(na):                                             # apply_wall_collision_force   core.py:163-169,212-225
        if not collide[a]:
            continue
        x, y = pos[:, a, 0], pos[:, a, 1]
        dists = np.stack([x - size[a] - xmin, xmax - x - size[a], y - size[a] - ymin, ymax - y - size[a]], axis=1)
        p = cf * _softplus_pen(dists, k)
        force[:, a] = np.stack([p[:, 0] - p[:, 1], p[:, 2] - p[:, 3]], axis=1) + force[:, a]
        has[a] = True
    pos, vel = pos.copy(), vel.copy()
    for i in range(NE):                                             # integrate_state              core.py:172-184
        if not movable[i]:
            continue
        v = vel[:, i] * (1 - damping)
        if has[i]:
            v = v + force[:, i] / mass[i] * dt
        if max_speed[i] >= 0:
            speed = np.sqrt(np.square(v[:, 0]) + np.square(v[:, 1]))
            over = speed > max_speed[i]
            with np.errstate(divide="ignore", invalid="ignore"):
                v = np.where(over[:, None], v / speed[:, None] * max_speed[i], v)
        vel[:, i] = v
        pos[:, i] = pos[:, i] + v * dt
    return pos, vel


# ---- scenario callbacks (multiagent/scenarios/simple_spread.py:72-101, simple_tag.py:83-179; MultiAgentEnv.step,
# multiagent/environment.py:97-108), vectorised over environments.  Pinned by tests/golden/mape_scenarios.npz.
def _dist(a, b):
    return np.sqrt(np.sum(np.square(a - b), axis=-1))


def spread_callbacks(pos, vel, size, na):
    """pos, vel [E, NE, 2] -> (obs [E, na, 4 + 2L + 4(na-1)], reward [E, na]) with the shared (summed) reward."""
    E, NE, _ = pos.shape
    obs, rew = [], np.zeros((E, na))
    cover = np.zeros(E)
    for l in range(na, NE):
        cover -= np.min(np.stack([_dist(pos[:, a], pos[:, l]) for a in range(na)], axis=1), axis=1)
    for i in range(na):
        r = cover.copy()
        for a in range(na):
            r -= (_dist(pos[:, a], pos[:, i]) < size[a] + size[i]).astype(float)
        rew[:, i] = r
        others = [j for j in range(na) if j != i]
        obs.append(np.concatenate([vel[:, i], pos[:, i]] + [pos[:, l] - pos[:, i] for l in range(na, NE)] +
                                  [pos[:, j] - pos[:, i] for j in others] + [np.zeros((E, 2)) for _ in others], axis=1))
    total = rew.sum(axis=1, keepdims=True)
    return np.stack(obs, axis=1), np.repeat(total, na, axis=1)


def _bound(x):
    return np.where(x < 0.9, 0.0, np.where(x < 1.0, (x - 0.9) * 10, np.minimum(np.exp(2 * x - 2), 10)))


def tag_callbacks(pos, vel, size, na, n_adv):
    """-> (list of per-agent obs [E, d_i], reward [E, na])."""
    E, NE, _ = pos.shape
    obs, rew = [], np.zeros((E, na))
    catches = np.zeros(E)
    for g in range(n_adv, na):
        for a in range(n_adv):
            catches += 10.0 * (_dist(pos[:, g], pos[:, a]) < size[g] + size[a])
    for i in range(na):
        if i < n_adv:
            rew[:, i] = catches
        else:
            r = np.zeros(E)
            for a in range(n_adv):
                r -= 10.0 * (_dist(pos[:, a], pos[:, i]) < size[a] + size[i])
            for d in range(2):
                r -= _bound(np.abs(pos[:, i, d]))
            rew[:, i] = r
        others = [j for j in range(na) if j != i]
        obs.append(np.concatenate([vel[:, i], pos[:, i]] + [pos[:, l] - pos[:, i] for l in range(na, NE)] +
                                  [pos[:, j] - pos[:, i] for j in others] + [vel[:, j] for j in others if j >= n_adv], axis=1))
    return obs, rew
