"""Generate golden env transitions by RUNNING THE UNCHANGED REFERENCE (this container only).

    python tests/golden/make_env_golden.py

Writes tests/golden/env_{3v3,5v5,2v1,1v1,4v2,1v5,5v1}.npz and copies the reference's own
recorded trajectory out_files/1.npy (written by test_fortattack.py:129-133) to
tests/golden/ref_traj_5v5.npy.

Every transition is one call of FortAttackGlobalEnv.step (gym_fortattack/fortattack.py:127-173)
on the reference's World (gym_fortattack/core.py:191-218) with the scenario callbacks of
gym_fortattack/envs/fortattack_env_v1.py.  The full pre-state is stored so the transition can be
replayed teacher-forced by the C oracle (CPU tests) and by the CUDA kernel (GPU tests).

Arrays (N transitions, A agents, guards first):
  pre_f   [N,A,6] f64  x, y, vx, vy, ang, prevDist (NaN = None)
  pre_i   [N,A,6] u8   alive, justDied, hit, wasHit, numHit, numWasHit
  t_pre   [N]     i32  world.time_step before the step
  cap     [N]     i32  world.max_time_steps in force (episode cap, fortattack.py:21)
  act     [N,A]   i32  discrete action 0..7
  obs     [N,A,6] f64  returned observation [alive,x,y,ang,vx,vy]
  rew     [N,A]   f64  returned reward
  done    [N]     u8
  result  [N]     u8   0 none, 1 all attackers dead, 2 time limit, 3 attacker reached (gameResult)
  post_i  [N,A,6] u8   alive, justDied, hit, wasHit, numHit, numWasHit after the step
  post_pd [N,A]   f64  prevDist after the step
  margin  [N]     f64  min |lambda| over every laser_hit test evaluated in the step (inf if none):
                       distance of the hit predicate from its decision boundary (core.py:384-390)
  reset_obs [R,A,6] f64 observations returned by env.reset() (distribution checks only)
"""
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402


def snapshot(w):
    A = len(w.agents)
    f = np.zeros((A, 6))
    i = np.zeros((A, 6), np.uint8)
    for k, a in enumerate(w.agents):
        f[k] = [a.state.p_pos[0], a.state.p_pos[1], a.state.p_vel[0], a.state.p_vel[1],
                a.state.p_ang, np.nan if a.prevDist is None else float(a.prevDist)]
        i[k] = [a.alive, a.justDied, a.hit, a.wasHit, a.numHit, a.numWasHit]
    return f, i


def laser_margin(w, act):
    """min |lambda| over the laser tests World.apply_laser_effect (core.py:254-285) is about to do."""
    m = np.inf
    alive = [a for a in w.agents if a.alive]
    idx = {id(a): k for k, a in enumerate(w.agents)}
    for a in alive:
        if act[idx[id(a)]] != 7:
            continue
        Amat = w.get_tri_pts_arr(a)
        for b in alive:
            if a.attacker == b.attacker:
                continue
            bb = np.array([[b.state.p_pos[0]], [b.state.p_pos[1]], [1]])
            lam = w.svd_sol(Amat, bb)
            m = min(m, float(np.min(np.abs(lam))))
    return m


def run(n_guards, n_attackers, n_trans, caps, seed):
    """Roll the reference env; `caps` cycles world.max_time_steps per episode (recorded as `cap`)."""
    rng = np.random.RandomState(seed)
    np.random.seed(seed)
    env, _ = ref_shim.make_ref_env(n_guards, n_attackers, caps[0])
    w = env.world
    A = env.n
    keys = "pre_f pre_i t_pre act obs rew done result post_i post_pd margin cap".split()
    rec = {k: [] for k in keys}
    reset_obs = []
    ep = 0
    with ref_shim.quiet():
        reset_obs.append(env.reset().copy())
    while len(rec["done"]) < n_trans:
        # action streams cycle per episode: uniform, shoot-heavy, and "rush"
        # (attackers head for the door so that result 3 / the guards' -10 term are exercised)
        p = np.full(8, 1 / 8.0) if ep % 3 == 0 else np.array([.1] * 7 + [.3])
        act = rng.choice(8, size=A, p=p).astype(np.int32)
        if ep % 3 == 2:
            for k, a in enumerate(w.agents):
                if a.attacker and rng.rand() < 0.8:
                    x = a.state.p_pos[0]
                    act[k] = 3 if (abs(x) < 0.1 or rng.rand() < 0.6) else (2 if x > 0 else 1)
        f, i = snapshot(w)
        rec["pre_f"].append(f)
        rec["pre_i"].append(i)
        rec["t_pre"].append(w.time_step)
        rec["cap"].append(w.max_time_steps)
        rec["act"].append(act)
        rec["margin"].append(laser_margin(w, act))
        with ref_shim.quiet():
            obs, rew, done, _ = env.step(act)
        rec["obs"].append(np.asarray(obs, np.float64))
        rec["rew"].append(np.asarray(rew, np.float64))
        rec["done"].append(done)
        res = 0
        if done:
            gr = w.gameResult
            res = 3 if gr[2] else (1 if gr[0] else 2)
        rec["result"].append(res)
        f2, i2 = snapshot(w)
        rec["post_i"].append(i2)
        rec["post_pd"].append(f2[:, 5])
        if done:
            ep += 1
            w.max_time_steps = caps[ep % len(caps)]
            with ref_shim.quiet():
                reset_obs.append(env.reset().copy())
    return dict(
        pre_f=np.array(rec["pre_f"]), pre_i=np.array(rec["pre_i"], np.uint8),
        t_pre=np.array(rec["t_pre"], np.int32), cap=np.array(rec["cap"], np.int32),
        act=np.array(rec["act"], np.int32), obs=np.array(rec["obs"]), rew=np.array(rec["rew"]),
        done=np.array(rec["done"], np.uint8), result=np.array(rec["result"], np.uint8),
        post_i=np.array(rec["post_i"], np.uint8), post_pd=np.array(rec["post_pd"]),
        margin=np.array(rec["margin"]), reset_obs=np.array(reset_obs),
        n_guards=np.int32(n_guards), n_attackers=np.int32(n_attackers))


def main():
    specs = [("env_3v3.npz", 3, 3, 1500, [100, 25, 60], 11),
             ("env_5v5.npz", 5, 5, 500, [100, 30], 12),
             ("env_2v1.npz", 2, 1, 200, [40, 15], 13),
             # the smallest and the most lopsided team shapes the kernels are instantiated for (oracle pinning on the CPU)
             ("env_1v1.npz", 1, 1, 300, [30, 12], 14),
             ("env_4v2.npz", 4, 2, 300, [50, 20], 15),
             ("env_1v5.npz", 1, 5, 300, [60, 25], 16),
             ("env_5v1.npz", 5, 1, 300, [40, 25], 17)]
    for name, g, a, n, caps, seed in specs:
        out = run(g, a, n, caps, seed)
        path = os.path.join(HERE, name)
        np.savez_compressed(path, **out)
        kills = int((out["pre_i"][:, :, 0].astype(int) - out["post_i"][:, :, 0].astype(int)).sum())
        print("%s: %d transitions, %d episode ends (results %s), %d kills, min margin %.3g, %d bytes"
              % (name, n, int(out["done"].sum()), np.bincount(out["result"], minlength=4).tolist(),
                 kills, float(out["margin"].min()), os.path.getsize(path)))
    shutil.copyfile(os.path.join(ref_shim.REF, "out_files", "1.npy"),
                    os.path.join(HERE, "ref_traj_5v5.npy"))
    print("copied out_files/1.npy -> ref_traj_5v5.npy")


if __name__ == "__main__":
    main()
