"""GPU parity of the fused CUDA step (through the C ABI) against the oracle and the golden transitions
produced by the unchanged reference.  Run on the B200 box: pytest -m gpu."""
import numpy as np
import pytest
import torch

import fa_oracle
import golden_util
from gpu_util import MAPPINGS, MARGIN_F32, acts_dev, close, contact_slack, make, pull, push, to_env_major

pytestmark = pytest.mark.gpu
DTYPES = [torch.float64, torch.float32]


def _compare_transition(env, dtype, pre_f, pre_i, t, ep, act, ref, margin, auto_reset=False, what=""):
    """Teacher-forced: inject the pre-state, take ONE CUDA step, compare everything with `ref`
    (dict obs,rew,done,result,post_i,post_pd,t_post).  Returns #envs excluded from the exact mask
    comparison (fp32 only: oracle hit predicate closer than MARGIN_F32 to its boundary)."""
    push(env, pre_f, pre_i, t, ep)
    obs, rew, done, result = env.step(acts_dev(act), auto_reset=auto_reset)
    st_f, st_i, t_post, _ = pull(env)
    obs, rew = to_env_major(obs), to_env_major(rew)
    done, result = done.cpu().numpy(), result.cpu().numpy()
    keep = np.ones(len(act), bool)
    if dtype == torch.float32:
        keep = ~(margin < MARGIN_F32)
    k = keep
    assert np.array_equal(st_i[k], ref["post_i"][k]), what + " alive/justDied/hit/wasHit/counters"
    assert np.array_equal(done[k], ref["done"][k]) and np.array_equal(result[k], ref["result"][k]), what + " done"
    # envs that were auto-reset report the new episode's first obs: no contact conditioning there
    sl = contact_slack(pre_f, ref["post_i"][:, :, 0], dtype)
    sl_obs = np.where(ref["done"][:, None] & bool(auto_reset), 0.0, sl)[k]
    obs_slack = sl_obs[:, :, None] * np.array([0, .1, .1, 0, 1, 1])       # x,y move by dt * dv
    close(obs[k], ref["obs"][k], dtype, what + " obs", obs_slack)
    close(rew[k], ref["rew"][k], dtype, what + " reward", 0.2 * sl[k])     # 2 * (prevDist - d)
    pd, rpd = st_f[k][:, :, 5], ref["post_pd"][k]
    assert np.array_equal(np.isnan(pd), np.isnan(rpd)), what + " prevDist None-ness"
    close(np.nan_to_num(pd), np.nan_to_num(rpd), dtype, what + " prevDist", 0.1 * sl[k])
    if "t_post" in ref:
        assert np.array_equal(t_post[k], ref["t_post"][k]), what + " time_step"
    return int((~keep).sum())


@pytest.mark.parametrize("mapping", MAPPINGS)
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("name", ["env_3v3.npz", "env_5v5.npz", "env_2v1.npz"])
def test_golden_reference_transitions(name, dtype, mapping):
    """Every transition recorded from the unchanged reference (tests/golden/make_env_golden.py)."""
    g = golden_util.load(name)
    N = g["act"].shape[0]
    env = make(N, g["n_guards"], g["n_attackers"], dtype, max_steps=golden_util.CAP, mapping=mapping)
    assert env.kernel_info()["mapping"] == mapping
    ref = dict(obs=g["obs"], rew=g["rew"], done=g["done"], result=g["result"], post_i=g["post_i"],
               post_pd=g["post_pd"], t_post=g["t_shift"] + 1)
    excluded = _compare_transition(env, dtype, g["pre_f"], g["pre_i"], g["t_shift"], np.zeros(N, np.uint32),
                                   g["act"], ref, g["margin"], what=name)
    assert excluded <= 2
    assert ref["done"].sum() > 0 and (g["pre_i"][:, :, 0] != g["post_i"][:, :, 0]).sum() > 0


@pytest.mark.parametrize("mapping", MAPPINGS)
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("ng,na,E,steps", [(3, 3, 4096, 130), (5, 5, 1024, 110), (1, 1, 257, 60), (4, 2, 100, 60)])
def test_teacher_forced_vs_oracle(ng, na, E, steps, dtype, mapping):
    """SURVEY 8d parity gate: oracle state -> fa_set_state -> fa_step -> compare, every step, with the
    shoot-heavy action stream and in-kernel auto-reset (Philox streams must match the oracle's)."""
    rng = np.random.RandomState(5)
    ora = fa_oracle.OracleEnv(E, ng, na, max_steps=40, seed=11, env_id0=1000, n_threads=8)
    env = make(E, ng, na, dtype, max_steps=40, seed=11, env_id0=1000, mapping=mapping)
    o0 = ora.reset()
    close(to_env_major(env.reset()), o0, dtype, "reset obs")
    excluded = kills = dones = 0
    for s in range(steps):
        act = rng.choice(8, size=(E, ng + na), p=[.1] * 7 + [.3]).astype(np.int32)
        pre = (ora.st_f.copy(), ora.st_i.copy(), ora.time_step.copy(), ora.episode.copy())
        obs, rew, done, result, margin = ora.step(act, auto_reset=True, want_margin=True)
        ref = dict(obs=obs, rew=rew, done=done, result=result, post_i=ora.st_i, post_pd=ora.st_f[:, :, 5],
                   t_post=ora.time_step)
        excluded += _compare_transition(env, dtype, *pre, act, ref, margin, auto_reset=True, what="step %d" % s)
        kills += int((pre[1][:, :, 0] > ora.st_i[:, :, 0]).sum())
        dones += int(done.sum())
        _, _, _, ep = pull(env)
        assert np.array_equal(ep[~(margin < MARGIN_F32)], ora.episode[~(margin < MARGIN_F32)])
    assert dones >= E and kills > 0
    if dtype == torch.float32:
        _quantised_teacher_forcing(ng, na, min(E, 1024), 60, mapping)
    assert excluded <= max(2, int(2e-5 * E * steps * (ng + na)))   # ~1e-5 * 0.23 per laser test


def _quantise(st_f):
    """Round an oracle state to the nearest state the fp32 engine can hold (heading: reduced + turns)."""
    q = st_f.astype(np.float32).astype(np.float64)
    turns = np.floor(st_f[:, :, 4] / (2 * np.pi))
    q[:, :, 4] = (st_f[:, :, 4] - turns * 2 * np.pi).astype(np.float32).astype(np.float64) + turns * 2 * np.pi
    return q


def _quantised_teacher_forcing(ng, na, E, steps, mapping):
    """Same gate with the oracle started from the fp32-representable state each step: isolates the
    kernel's arithmetic from state rounding, so the plain 1e-5 bound must hold with NO contact slack."""
    rng = np.random.RandomState(6)
    ora = fa_oracle.OracleEnv(E, ng, na, max_steps=40, seed=12, n_threads=8)
    env = make(E, ng, na, torch.float32, max_steps=40, seed=12, mapping=mapping)
    ora.reset(); env.reset()
    for s in range(steps):
        ora.st_f[:] = _quantise(ora.st_f)
        act = rng.choice(8, size=(E, ng + na), p=[.1] * 7 + [.3]).astype(np.int32)
        push(env, ora.st_f, ora.st_i, ora.time_step, ora.episode)
        obs, rew, done, result, margin = ora.step(act, auto_reset=True, want_margin=True)
        o, r, d, rs = env.step(acts_dev(act), auto_reset=True)
        k = ~(margin < MARGIN_F32)
        close(to_env_major(o)[k], obs[k], torch.float32, "quantised step %d obs" % s)
        close(to_env_major(r)[k], rew[k], torch.float32, "quantised step %d reward" % s)
        assert np.array_equal(d.cpu().numpy()[k], done[k]) and np.array_equal(rs.cpu().numpy()[k], result[k])


@pytest.mark.parametrize("mapping", MAPPINGS)
@pytest.mark.parametrize("ng,na", [(3, 3), (5, 5)])
def test_free_run_double_tracks_oracle(ng, na, mapping):
    """Double mode, no re-sync for 300 steps (several episodes with resets): same trajectories."""
    E, T = 512, 300
    rng = np.random.RandomState(9)
    ora = fa_oracle.OracleEnv(E, ng, na, max_steps=50, seed=3, n_threads=8)
    env = make(E, ng, na, torch.float64, max_steps=50, seed=3, mapping=mapping)
    ora.reset(); env.reset()
    acts = rng.randint(0, 8, size=(T, E, ng + na)).astype(np.int32)
    o_obs, o_rew, o_done, o_res = ora.step_many(acts)
    a = torch.from_numpy(np.ascontiguousarray(np.swapaxes(acts, 1, 2))).cuda()
    obs, rew, done, res = env.step_many(a)
    obs = np.swapaxes(obs.cpu().numpy(), 1, 2); rew = np.swapaxes(rew.cpu().numpy(), 1, 2)
    assert np.array_equal(done.cpu().numpy(), o_done) and np.array_equal(res.cpu().numpy(), o_res)
    assert np.abs(obs - o_obs).max() < 1e-7 and np.abs(rew - o_rew).max() < 1e-7
    assert o_done.sum() >= 5 * E


@pytest.mark.parametrize("mapping", MAPPINGS)
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("ng,na,E", [(3, 3, 4096), (5, 5, 333), (2, 1, 31), (3, 3, 1)])
def test_step_many_equals_repeated_step(ng, na, E, dtype, mapping):
    """The persistent T-step launch and T single-step launches are the same arithmetic: bit-equal."""
    T = 70
    g = torch.Generator(device="cuda").manual_seed(1)
    acts = torch.randint(0, 8, (T, ng + na, E), generator=g, device="cuda", dtype=torch.int32)
    e1 = make(E, ng, na, dtype, max_steps=30, seed=2, mapping=mapping)
    e2 = make(E, ng, na, dtype, max_steps=30, seed=2, mapping=mapping)
    assert torch.equal(e1.reset(), e2.reset())
    obs, rew, done, res = e2.step_many(acts)
    for t in range(T):
        o, r, d, rs = e1.step(acts[t], auto_reset=True)
        assert torch.equal(o, obs[t]) and torch.equal(r, rew[t]), "step %d" % t
        assert torch.equal(d, done[t]) and torch.equal(rs, res[t])
    for x, y in zip(e1.get_state(), e2.get_state()):
        assert torch.equal(torch.nan_to_num(x.double()), torch.nan_to_num(y.double()))
    assert int(done.sum()) >= 2 * E


@pytest.mark.parametrize("dtype", DTYPES)
def test_state_roundtrip_and_alive_counts(dtype):
    g = golden_util.load("env_5v5.npz")
    N = g["act"].shape[0]
    env = make(N, 5, 5, dtype)
    ep = np.arange(N).astype(np.uint32)
    push(env, g["pre_f"], g["pre_i"], g["t_pre"], ep)
    st_f, st_i, t, ep2 = pull(env)
    assert np.array_equal(st_i, g["pre_i"]) and np.array_equal(t, g["t_pre"]) and np.array_equal(ep2, ep)
    tol = 0 if dtype == torch.float64 else 1e-5
    assert np.nanmax(np.abs(st_f - g["pre_f"]) / (1 + np.abs(g["pre_f"]))) <= tol
    assert np.array_equal(np.isnan(st_f), np.isnan(g["pre_f"]))
    ng_alive, na_alive = env.alive_counts()
    assert np.array_equal(ng_alive.cpu().numpy(), g["pre_i"][:, :5, 0].sum(1))
    assert np.array_equal(na_alive.cpu().numpy(), g["pre_i"][:, 5:, 0].sum(1))


def test_shard_invariance_and_seed():
    """Envs are keyed by global id: a shard [1000,1100) reproduces rows of the full batch."""
    T = 120
    g = torch.Generator(device="cuda").manual_seed(4)
    acts = torch.randint(0, 8, (T, 6, 2048), generator=g, device="cuda", dtype=torch.int32)
    full = make(2048, 3, 3, torch.float32, max_steps=25, seed=77)
    part = make(100, 3, 3, torch.float32, max_steps=25, seed=77, env_id0=1000)
    other = make(100, 3, 3, torch.float32, max_steps=25, seed=78, env_id0=1000)
    o_full, o_part, o_other = full.reset(), part.reset(), other.reset()
    assert torch.equal(o_full[:, 1000:1100], o_part) and not torch.equal(o_part, o_other)
    of, rf, df, _ = full.step_many(acts)
    op, rp, dp, _ = part.step_many(acts[:, :, 1000:1100].contiguous())
    assert torch.equal(of[:, :, 1000:1100], op) and torch.equal(rf[:, :, 1000:1100], rp)
    assert torch.equal(df[:, 1000:1100], dp)


@pytest.mark.parametrize("path,pinned", [("auto", True), ("staged", True), ("auto", False)])
def test_step_host_matches_device_step(path, pinned, monkeypatch):
    """fa_step_host: mapped (zero-copy) path, packed staged path, and pageable buffers."""
    E = 1000
    monkeypatch.setenv("FA_HOST_PATH", path)
    e1, e2 = make(E, 3, 3, torch.float32, seed=5), make(E, 3, 3, torch.float32, seed=5)
    e1.reset(); e2.reset()
    bufs = e2.make_host_buffers()
    if not pinned:
        bufs = tuple(torch.empty_like(b, pin_memory=False) for b in bufs)
    rng = np.random.RandomState(0)
    for _ in range(20):
        a = torch.from_numpy(rng.randint(0, 8, size=(6, E)).astype(np.int32))
        o, r, d, rs = e1.step(a.cuda())
        bufs[0].copy_(a)
        ho, hr, hd, hrs = e2.step_host(*bufs)
        assert torch.equal(o.cpu(), ho) and torch.equal(r.cpu(), hr) and torch.equal(d.cpu(), hd) and torch.equal(rs.cpu(), hrs)


@pytest.mark.parametrize("chunk,pinned", [(None, True), (7, True), (1, True), (0, True), (5, False)])
@pytest.mark.parametrize("ng,na,E,dtype", [(3, 3, 1000, torch.float32), (5, 5, 333, torch.float64)])
def test_step_many_host_matches_device_steps(ng, na, E, dtype, chunk, pinned):
    """fa_step_many_host: chunk pipeline (default chunk, ragged last chunk, one-step chunks, pageable buffers) and the
    mapped single-launch form (chunk 0) are bit-equal to T device-side steps, alive-end output included; the
    state left behind is the same (a following step agrees)."""
    T, A = 45, ng + na
    e1, e2 = make(E, ng, na, dtype, seed=5, max_steps=30), make(E, ng, na, dtype, seed=5, max_steps=30)
    e1.reset(); e2.reset()
    hs = e2.make_host_streams(T)
    if not pinned:
        hs = tuple(torch.empty_like(b, pin_memory=False) for b in hs)
    ae = torch.zeros(T, E, dtype=torch.uint8, device="cuda")
    e2.set_alive_end_buffer(ae)
    rng = np.random.RandomState(1)
    acts = torch.from_numpy(rng.choice(8, size=(T + 1, A, E), p=[.1] * 7 + [.3]).astype(np.int32))
    hs[0].copy_(acts[:T])
    ho, hr, hd, hrs = e2.step_many_host(*hs, chunk_steps=chunk)
    ae1 = torch.zeros(E, dtype=torch.uint8, device="cuda")
    e1.set_alive_end_buffer(ae1)
    for t in range(T):
        o, r, d, rs = e1.step(acts[t].cuda())
        assert torch.equal(o.cpu(), ho[t]) and torch.equal(r.cpu(), hr[t]), t
        assert torch.equal(d.cpu(), hd[t]) and torch.equal(rs.cpu(), hrs[t]) and torch.equal(ae1, ae[t]), t
    assert hd.sum() > 0
    e1.set_alive_end_buffer(None); e2.set_alive_end_buffer(None)
    for x, y in zip(e1.step(acts[T].cuda()), e2.step(acts[T].cuda())):
        assert torch.equal(x, y)
    # outputs may be skipped
    hs2 = e2.make_host_streams(3, store_obs=False)
    e2.step_many_host(*hs2, chunk_steps=chunk if pinned else 2)


def test_full_size_properties():
    """BASELINE config 2 size (3v3 x 4096 envs x 1000 steps, cap 100) and a >L2 batch: size-independent
    properties of the reference semantics."""
    for E, T in ((4096, 1000), (1 << 20, 12)):
        env = make(E, 3, 3, torch.float32, max_steps=100, seed=0)
        prev = env.reset()
        g = torch.Generator(device="cuda").manual_seed(0)
        n_done = 0
        chunk = 50 if E == 4096 else T
        for c in range(T // chunk):
            acts = torch.randint(0, 8, (chunk, 6, E), generator=g, device="cuda", dtype=torch.int32)
            obs, rew, done, res = env.step_many(acts)
            assert torch.isfinite(obs).all() and torch.isfinite(rew).all()
            allobs = torch.cat([prev[None], obs], 0)
            cont = (done[:, None, :] == 0)                                   # [chunk,1,E]: no reset at t
            was_dead = (allobs[:-1, :, :, 0] == 0) & cont
            # dead agents are frozen and stay dead within an episode (core.py:324 loops over alive only)
            assert torch.equal(allobs[1:][was_dead], allobs[:-1][was_dead])
            # positions stay inside the walls' soft band, |v| <= max_speed
            assert (obs[..., 1].abs() < 1.2).all() and (obs[..., 2].abs() < 1.0).all()
            assert ((obs[..., 4] ** 2 + obs[..., 5] ** 2) <= 9.0001).all()
            # done <=> result != 0 ; after a reset everyone is alive with zero velocity
            assert torch.equal(done != 0, res != 0)
            fresh = obs.permute(0, 2, 1, 3)[done != 0]                       # [n,A,6]
            assert (fresh[:, :, 0] == 1).all() and (fresh[:, :, 4:] == 0).all()
            # rewards are bounded by the scenario's terms
            assert rew.abs().max() <= 10 + 10 + 3 * 3 + 3 + 2
            n_done += int(done.sum())
            prev = obs[-1]
        if E == 4096:
            assert n_done >= 10 * E          # cap 100 over 1000 steps
            counts = torch.bincount(res.flatten().long(), minlength=4)
            assert counts[2] > 0


@pytest.mark.parametrize("ng,na", [(3, 3), (5, 5)])
def test_episode_statistics_match_oracle_at_scale(ng, na):
    """Free-running float32 CUDA envs and the float64 oracle, same reset streams and the same uniform random actions,
    2048 envs x 400 steps (~8 000 episodes): episode count, game-result split (all dead / time limit / fort reached,
    fortattack.py:202-225), kills and mean team rewards agree statistically (the trajectories themselves separate
    chaotically after a near-threshold laser test, so this is a distribution-level check; SURVEY appendix A gives the
    reference's own numbers: ~9 % / 90 % / 1 % at 3v3, ~2.2 kills per episode)."""
    E, T, A = 2048, 400, ng + na
    env = make(E, ng, na, torch.float32, max_steps=100, seed=21)
    ora = fa_oracle.OracleEnv(E, ng, na, max_steps=100, seed=21)
    env.reset(); ora.reset()
    rng = np.random.RandomState(5)
    res_g, res_o = np.zeros(4), np.zeros(4)
    kills_g = kills_o = 0
    rew_g, rew_o = np.zeros(A), np.zeros(A)
    prev_g = prev_o = None
    for c in range(T // 50):
        act = rng.randint(0, 8, size=(50, E, A)).astype(np.int32)
        obs, rew, done, res = env.step_many(torch.from_numpy(np.ascontiguousarray(act.transpose(0, 2, 1))).cuda())
        alive = obs[..., 0].cpu().numpy()                                   # [50, A, E]
        d = done.cpu().numpy() != 0
        res_g += np.bincount(res.cpu().numpy().ravel(), minlength=4)
        rew_g += rew.sum(dim=(0, 2)).cpu().numpy()
        a_prev = np.concatenate([np.ones((1, A, E)) if prev_g is None else prev_g[None], alive[:-1]])
        kills_g += int(((a_prev == 1) & (alive == 0) & ~d[:, None, :]).sum())
        prev_g = alive[-1]
        o, r, dn, rs = ora.step_many(act)                                   # [50,E,A,6], [50,E,A], [50,E], [50,E]
        res_o += np.bincount(rs.ravel(), minlength=4)
        rew_o += r.sum(axis=(0, 1))
        al = o[..., 0].transpose(0, 2, 1)                                   # [50, A, E]
        ap = np.concatenate([np.ones((1, A, E)) if prev_o is None else prev_o[None], al[:-1]])
        kills_o += int(((ap == 1) & (al == 0) & ~(dn != 0)[:, None, :]).sum())
        prev_o = al[-1]
    n_g, n_o = res_g[1:].sum(), res_o[1:].sum()
    assert n_o > 3 * E and abs(n_g - n_o) <= 0.01 * n_o                     # same number of episodes to 1 %
    for k in (1, 2, 3):                                                     # each outcome's share: within 4 sigma + 0.3 %
        p_o, p_g = res_o[k] / n_o, res_g[k] / n_g
        assert abs(p_g - p_o) < 4 * np.sqrt(max(p_o, 1e-3) * (1 - p_o) / n_o) + 3e-3, (k, p_g, p_o)
    assert abs(kills_g - kills_o) < 0.05 * kills_o + 20, (kills_g, kills_o)
    assert np.all(np.abs(rew_g - rew_o) < 0.05 * np.abs(rew_o) + 0.02 * E * T / 100), (rew_g, rew_o)
    if (ng, na) == (3, 3):                                                   # the reference's own split
        assert 0.05 < res_g[1] / n_g < 0.14 and 0.84 < res_g[2] / n_g < 0.94 and res_g[3] / n_g < 0.03


@pytest.mark.parametrize("mapping", MAPPINGS)
def test_alive_counts_at_step_end(mapping):
    """fa_set_alive_end_buffer: numAliveGuards | numAliveAttackers << 4 as every step leaves them (before the auto-reset),
    against the alive flags of the double-precision oracle run on the same actions (core.py:113-114,293-302)."""
    E, ng, na = 512, 3, 2
    env = make(E, ng, na, torch.float64, max_steps=20, seed=3, mapping=mapping)
    ora = fa_oracle.OracleEnv(E, ng, na, max_steps=20, seed=3)
    env.reset(); ora.reset()
    buf = torch.full((E,), 255, dtype=torch.uint8, device="cuda")
    env.set_alive_end_buffer(buf)
    rng = np.random.RandomState(1)
    seen_kill = False
    for t in range(60):
        act = rng.choice(8, size=(E, ng + na), p=[.1] * 7 + [.3]).astype(np.int32)
        env.step(acts_dev(act), auto_reset=True)
        # the oracle without auto-reset shows the terminal state; then reset the finished envs by hand
        obs, rew, done, res = ora.step(act, auto_reset=False)
        alive = ora.st_i[:, :, 0]
        want = alive[:, :ng].sum(1) | (alive[:, ng:].sum(1) << 4)
        assert np.array_equal(buf.cpu().numpy(), want.astype(np.uint8)), t
        seen_kill |= bool((want != (ng | na << 4)).any())
        ora.reset(mask=done.astype(np.uint8))
    assert seen_kill
    many = torch.zeros(7, E, dtype=torch.uint8, device="cuda")
    env.set_alive_end_buffer(many)
    env.step_many(torch.randint(0, 8, (7, ng + na, E), device="cuda", dtype=torch.int32))
    assert int(many.max()) <= (ng | na << 4) and int((many & 15).max()) == ng
    env.set_alive_end_buffer(None)


def test_both_mappings_agree_and_auto_picks_by_batch_size():
    """Thread-per-env and thread-per-agent run the same physics: 200 free-running steps stay within float
    rounding of each other (contact forces are summed in a different order), masks identical."""
    T, E = 60, 2048
    g = torch.Generator(device="cuda").manual_seed(8)
    acts = torch.randint(0, 8, (T, 6, E), generator=g, device="cuda", dtype=torch.int32)
    ea, eb = make(E, 3, 3, torch.float64, max_steps=20, seed=1, mapping="env"), make(E, 3, 3, torch.float64, max_steps=20, seed=1, mapping="agent")
    assert torch.equal(ea.reset(), eb.reset())
    oa, ra, da, sa = ea.step_many(acts)
    ob, rb, db, sb = eb.step_many(acts)
    assert torch.equal(da, db) and torch.equal(sa, sb) and torch.equal(oa[..., 0], ob[..., 0])
    assert (oa - ob).abs().max() < 1e-9 and (ra - rb).abs().max() < 1e-9
    assert make(4096, 3, 3, torch.float32).kernel_info()["mapping"] == "group"
    assert make(1 << 17, 3, 3, torch.float32).kernel_info()["mapping"] == "env"


@pytest.mark.parametrize("ng,na,E", [(3, 3, 2050), (5, 5, 1001), (1, 1, 515), (2, 1, 300), (4, 2, 777), (5, 4, 64)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_group_mapping_is_bit_equal_to_agent_mapping(ng, na, E, dtype):
    """The sub-warp-group kernel (2/4/8/16 lanes per env, ballots instead of block barriers) performs the per-agent
    arithmetic of the thread-per-agent kernel statement for statement: free-running, with resets, every output and the
    final state are identical bit for bit -- for every group width, with ragged tails (E not a multiple of 32 / G)."""
    T, A = 50, ng + na
    g = torch.Generator(device="cuda").manual_seed(ng * 10 + na)
    acts = torch.randint(0, 8, (T, A, E), generator=g, device="cuda", dtype=torch.int32)
    ea, eb = make(E, ng, na, dtype, max_steps=17, seed=4, mapping="agent"), make(E, ng, na, dtype, max_steps=17, seed=4, mapping="group")
    assert eb.kernel_info()["mapping"] == "group"
    assert torch.equal(ea.reset(), eb.reset())
    for x, y in zip(ea.step_many(acts[:30]), eb.step_many(acts[:30])):
        assert torch.equal(x, y)
    for t in range(30, T):                       # single-step launches, auto-reset on
        for x, y in zip(ea.step(acts[t]), eb.step(acts[t])):
            assert torch.equal(x, y)
    for x, y in zip(ea.get_state(), eb.get_state()):
        assert torch.equal(x, y)


@pytest.mark.parametrize("mapping", MAPPINGS)
def test_fused_rollout_bookkeeping_equals_separate_kernel(mapping):
    """fa_set_rollout_outputs (mask of the next slot, end point, episode reward sums written by the step launch itself)
    == rl_rollout_bookkeeping run on the step's outputs afterwards (train_fortattack.py:53,97-104, storage.py:41)."""
    from importlib import import_module
    ro = import_module("emergent-multiagent-strategies_b200.rollout")
    ng, na, E, T = 3, 3, 1500, 40
    A = ng + na
    env = make(E, ng, na, torch.float32, max_steps=11, seed=5, mapping=mapping)
    env2 = make(E, ng, na, torch.float32, max_steps=11, seed=5, mapping=mapping)
    R1, R2 = ro.SharedRollouts(T, A, E, "cuda"), ro.SharedRollouts(T, A, E, "cuda")
    ep1, ep2 = torch.zeros(A, E, device="cuda"), torch.zeros(A, E, device="cuda")
    o0 = env.reset()
    assert torch.equal(o0, env2.reset())
    R1.obs[0].copy_(o0); R2.obs[0].copy_(o0)
    g = torch.Generator(device="cuda").manual_seed(3)
    for t in range(T):
        a = torch.randint(0, 8, (A, E), generator=g, device="cuda", dtype=torch.int32)
        env.step(a, out=(R1.obs[t + 1], R1.rewards[t], R1.done[t], R1.result[t]))
        R1.bookkeeping(t, ep1)
        env2.step(a, out=(R2.obs[t + 1], R2.rewards[t], R2.done[t], R2.result[t]), bookkeeping=(R2.masks[t + 1], R2.ends[t + 1], ep2))
    assert torch.equal(R1.obs, R2.obs) and torch.equal(R1.masks, R2.masks) and torch.equal(R1.ends, R2.ends)
    assert torch.equal(ep1, ep2) and int(R1.ends.sum()) > E and float(R1.masks.min()) == 0.0
    with pytest.raises(ValueError):
        env.step(a, bookkeeping=(R2.masks[0, :, :5], None, None))


def test_step_many_host_errors_are_loud():
    import fortattack_b200 as fab
    env = make(64, 3, 3, torch.float32)
    env.reset()
    hs = env.make_host_streams(4)
    lib, h = fab._capi.lib(), env._h
    small = torch.empty(1024, dtype=torch.uint8, device="cuda")
    assert lib.fa_step_many_host(h, 4, hs[0].data_ptr(), hs[1].data_ptr(), hs[2].data_ptr(), hs[3].data_ptr(),
                                 hs[4].data_ptr(), small.data_ptr(), 1024, None) == -1
    assert b"staging buffer" in lib.fa_last_error()
    pageable = torch.zeros(4, 6, 64, dtype=torch.int32)
    assert lib.fa_step_many_host(h, 4, pageable.data_ptr(), hs[1].data_ptr(), hs[2].data_ptr(), hs[3].data_ptr(),
                                 hs[4].data_ptr(), None, 0, None) == -1
    assert b"page-locked" in lib.fa_last_error()
    assert lib.fa_step_many_host(h, 0, hs[0].data_ptr(), None, None, None, None, None, 0, None) == -1
    with pytest.raises(ValueError):
        env.step_many_host(hs[0][:, :5], *hs[1:])


def test_errors_are_loud():
    import fortattack_b200 as fab
    with pytest.raises(fab.FaError):
        fab.FortAttackBatch(16, 6, 3)                      # unsupported team size
    with pytest.raises(fab.FaError):
        fab.FortAttackBatch(0, 3, 3)
    env = make(8, 3, 3, torch.float32)
    with pytest.raises(ValueError):
        env.step(torch.zeros(6, 9, dtype=torch.int32, device="cuda"))
    info = env.kernel_info()
    assert info["regs"] > 0 and info["block"] in (32, 64, 128, 192) and env.launch_count() >= 1


def test_raw_pointer_arguments_are_validated():
    """Every tensor handed to the library as a raw pointer is checked for shape / dtype / device / contiguity first
    (a short mask, a wrong `out`, or an [E] alive-end buffer under a T-step call would be silent out-of-bounds access)."""
    import fortattack_b200 as fab
    E = 64
    env = fab.FortAttackBatch(E, 3, 3, max_steps=20, seed=1, device="cuda:0")
    env.reset()
    acts = torch.zeros(6, E, dtype=torch.int32, device="cuda")
    with pytest.raises(ValueError):
        env.reset(mask=torch.ones(E - 1, dtype=torch.uint8, device="cuda"))
    with pytest.raises(ValueError):
        env.reset(out=torch.empty(6, E, 5, device="cuda"))
    good = (torch.empty(6, E, 6, device="cuda"), torch.empty(6, E, device="cuda"),
            torch.empty(E, dtype=torch.uint8, device="cuda"), torch.empty(E, dtype=torch.uint8, device="cuda"))
    env.step(acts, out=good)
    for k, bad in ((0, torch.empty(6, E, 6, device="cuda", dtype=torch.float64)), (1, torch.empty(E, 6, device="cuda").t()),
                   (2, torch.empty(E, dtype=torch.int32, device="cuda")), (3, torch.empty(E, dtype=torch.uint8))):
        out = list(good)
        out[k] = bad
        with pytest.raises(ValueError):
            env.step(acts, out=tuple(out))
    with pytest.raises(ValueError):
        env.step_many(acts.expand(4, 6, E).contiguous(), out=(None, torch.empty(3, 6, E, device="cuda"), None, None))
    env.set_alive_end_buffer(torch.zeros(E, dtype=torch.uint8, device="cuda"))
    env.step(acts, out=good)
    with pytest.raises(ValueError):
        env.step_many(acts.expand(4, 6, E).contiguous())
    env.set_alive_end_buffer(torch.zeros(4, E, dtype=torch.uint8, device="cuda"))
    env.step_many(acts.expand(4, 6, E).contiguous())
    env.set_alive_end_buffer(None)
    # a mask given as [E, 1] / bool is accepted (reshaped, converted)
    o = env.reset(mask=torch.ones(E, 1, dtype=torch.bool, device="cuda"))
    assert (o[:, :, 0] == 1).all()
