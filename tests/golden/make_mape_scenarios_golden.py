"""Golden MultiAgentEnv transitions of the reference's simple_spread and simple_tag scenarios: the unchanged
multiagent/scenarios/*.py + multiagent/environment.py + multiagent/core.py from /root/reference (numpy only; gym replaced by
the stub of ref_shim.py, which does no arithmetic).  For E worlds x T steps: state before the step, action vectors, and the
env.step() outputs (obs_n, reward_n) plus the state after.  Run in the build container:
    python tests/golden/make_mape_scenarios_golden.py"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install_stubs()
sys.path.insert(0, ref_shim.REF)
from multiagent.environment import MultiAgentEnv  # noqa: E402


def load_scenario(name):
    # multiagent/scenarios/__init__.py uses the `imp` module (gone in python 3.12): load the scenario file itself
    spec = importlib.util.spec_from_file_location("ref_scn_" + name, os.path.join(ref_shim.REF, "multiagent", "scenarios", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.Scenario()


def main():
    out = {}
    E, T = 16, 25
    for name in ("simple_spread", "simple_tag"):
        np.random.seed({"simple_spread": 21, "simple_tag": 22}[name])
        rng = np.random.RandomState(5)
        pos0 = pos = vel0 = None
        recs = {k: [] for k in ("pos_before", "vel_before", "act", "pos_after", "vel_after", "rew")}
        obs_rec = None
        for e in range(E):
            sc = load_scenario(name)
            world = sc.make_world()
            env = MultiAgentEnv(world, sc.reset_world, sc.reward, sc.observation)
            obs0 = env.reset()
            if e % 3 == 0:                                   # crowd some worlds so that contacts / catches happen
                for ent in world.entities:
                    ent.state.p_pos = ent.state.p_pos * 0.25
            if e % 4 == 1 and name == "simple_tag":          # push the prey towards the screen edge (bound() penalty branches)
                world.agents[-1].state.p_pos = np.array([0.93, -1.15])
            n = env.n
            if obs_rec is None:
                obs_rec = [[] for _ in range(n)]
                out[name + "/obs_dims"] = np.array([len(o) for o in obs0])
                out[name + "/size"] = np.array([ent.size for ent in world.entities])
                out[name + "/na"] = np.array(n)
            row = {k: [] for k in recs}
            orow = [[] for _ in range(n)]
            for t in range(T):
                row["pos_before"].append(np.stack([ent.state.p_pos.copy() for ent in world.entities]))
                row["vel_before"].append(np.stack([ent.state.p_vel.copy() for ent in world.entities]))
                act = np.zeros((n, 5))
                hot = rng.randint(0, 5, size=n)
                act[np.arange(n), hot] = 1.0
                if t % 5 == 4:
                    act = rng.uniform(0, 1, (n, 5))           # soft action vectors take the same path (environment.py:171-177)
                obs_n, rew_n, done_n, _ = env.step([a.copy() for a in act])
                row["act"].append(act)
                row["pos_after"].append(np.stack([ent.state.p_pos.copy() for ent in world.entities]))
                row["vel_after"].append(np.stack([ent.state.p_vel.copy() for ent in world.entities]))
                row["rew"].append(np.array(rew_n, dtype=np.float64))
                for i in range(n):
                    orow[i].append(np.asarray(obs_n[i], dtype=np.float64))
            for k in recs:
                recs[k].append(np.stack(row[k]))
            for i in range(n):
                obs_rec[i].append(np.stack(orow[i]))
        for k, v in recs.items():
            out["%s/%s" % (name, k)] = np.stack(v, axis=1)                 # [T, E, ...]
        for i, o in enumerate(obs_rec):
            out["%s/obs%d" % (name, i)] = np.stack(o, axis=1)               # [T, E, d_i]
        out[name + "/shared"] = np.array(bool(getattr(world, "collaborative", False)))
    np.savez_compressed(os.path.join(HERE, "mape_scenarios.npz"), **out)
    print("wrote mape_scenarios.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
