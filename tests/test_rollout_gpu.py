"""GPU: the gym facade and the batched rollout driver (collect -> GAE -> PPO) on top of the CUDA engine."""
import os
import subprocess
import sys
from importlib import import_module

import numpy as np
import pytest
import torch

import fa_oracle
import fortattack_b200 as fab

pytestmark = pytest.mark.gpu
PKG = "emergent-multiagent-strategies_b200"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_gym_facade_surface_and_semantics():
    gf = import_module(PKG + ".gym_fortattack.fortattack")
    env = gf.make_fortattack_env(30, n_guards=5, n_attackers=5, seed=3)
    # what learner.setup_master / Learner / the scripts touch
    assert env.n == 10 and env.action_space[9].n == 8 and env.observation_space[0].shape == (6,)
    assert env.action_spaces[0].shape == (8,) and env.ob_rms is None and env.world.numGuards == 5
    assert [a.attacker for a in env.world.policy_agents] == [False] * 5 + [True] * 5
    assert env.world.max_time_steps == 30 and env.world.time_step == 0
    obs = env.reset()
    assert isinstance(obs, np.ndarray) and obs.shape == (10, 6) and obs.dtype == np.float64
    # the same trajectory as the oracle, free-running in lockstep with teacher-forced oracle state
    ora = fa_oracle.OracleEnv(1, 5, 5, max_steps=30, seed=3)
    ora.reset(); ora.reset()                       # constructor reset + env.reset() (fortattack_env_v1.py:45)
    assert np.abs(obs - ora.observe()[0]).max() < 1e-6
    rng = np.random.RandomState(0)
    n_done = 0
    for s in range(200):
        act = rng.choice(8, size=10, p=[.1] * 7 + [.3])
        o, r, d, info = env.step(act)
        st = env._batch.get_state()
        assert isinstance(r, list) and len(r) == 10 and isinstance(d, bool) and list(info) == ["n"] and len(info["n"]) == 10
        assert env.world.numAliveAttackers == int(o[5:, 0].sum()) and env.world.numAliveGuards == int(o[:5, 0].sum())
        if d:
            n_done += 1
            assert env.world.gameResult.sum() == 1
            o = env.reset()
            assert env.world.gameResult.sum() == 0 and env.world.time_step == 0 and (o[:, 0] == 1).all()
    assert n_done >= 6
    env.reset()
    env.world.max_time_steps = 5
    for s in range(5):
        o, r, d, _ = env.step(np.zeros(10, int))
    assert d and env.world.gameResult[1] == 1


def test_drop_in_import_paths():
    """With the package directory first on sys.path the reference's own import statements resolve here:
    `from gym_fortattack.fortattack import make_fortattack_env` (utils.py:5), `from rlcore.algo import JointPPO`
    (learner.py:3), `from rlcore.storage import RolloutStorage` (rlagent.py:2), `from mpnn import MPNN` (learner.py:5)."""
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from gym_fortattack.fortattack import make_fortattack_env\n"
            "from rlcore.algo import JointPPO, PPO\nfrom rlcore.storage import RolloutStorage\nfrom mpnn import MPNN\n"
            "env = make_fortattack_env(20, n_guards=2, n_attackers=2)\n"
            "o = env.reset(); o, r, d, i = env.step([0, 1, 2, 7])\nprint('ok', o.shape)\n"
            "f = env.render(mode='rgb_array')\nprint('frame', f[0].shape, f[0].dtype)\n") % os.path.join(ROOT, PKG)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ok (4, 6)" in out.stdout and "frame (700, 700, 3) uint8" in out.stdout, out.stderr[-2000:]


def test_batched_collect_matches_per_env_protocol():
    """collect() + wrap_horizon() for E envs == the reference protocol applied to each env separately."""
    ro = import_module(PKG + ".rollout")
    storage = import_module(PKG + ".rlcore.storage")
    torch.manual_seed(0)
    tr = ro.BatchedTrainer(64, 3, 3, num_steps=48, max_episode_steps=15, hidden_dim=32, seed=9)
    tr.collect()
    tr.wrap_horizon()
    R = tr.roll
    assert int(R.done.sum()) >= 3 * 64
    # stored obs of step t+1 is the reset obs where the env finished at step t: everyone alive, at rest
    fin = R.done != 0
    nxt = R.obs[1:].permute(0, 2, 1, 3)[fin]
    assert (nxt[:, :, 0] == 1).all() and (nxt[:, :, 4:] == 0).all()
    assert torch.equal(R.ends[1:-1], fin[:-1]) and bool(R.ends[-1].all())
    # masks[t+1] = alive flags of obs[t], except after a reset (all alive)
    exp = torch.where(fin[:, None, :], torch.ones_like(R.obs[:-1, :, :, 0]), R.obs[:-1, :, :, 0])
    assert torch.equal(R.masks[1:, :, :, 0], exp)
    for i in (0, 4):
        for e in (0, 17, 63):
            one = storage.RolloutStorage(48, 1, (6,), None, 1)
            one.rewards.copy_(R.agents[i].rewards[:, e:e + 1].cpu())
            one.value_preds.copy_(R.agents[i].value_preds[:, e:e + 1].cpu())
            one.masks.copy_(R.agents[i].masks[:, e:e + 1].cpu())
            start = 0
            for end_pt in torch.nonzero(R.ends[:, e]).flatten().tolist():
                one.compute_returns(one.value_preds[end_pt].clone(), True, tr.gamma, tr.tau, start, end_pt)
                start = end_pt + 1
            assert torch.allclose(one.returns[:, 0], R.agents[i].returns[:, e].cpu(), atol=1e-5)


def test_training_runs_and_checkpoint_format(tmp_path):
    ro = import_module(PKG + ".rollout")
    torch.manual_seed(1)
    tr = ro.BatchedTrainer(256, 3, 3, num_steps=32, max_episode_steps=25, hidden_dim=128, ppo_epoch=2, num_mini_batch=4)
    before = [p.detach().clone() for p in tr.policies[0].parameters()]
    for _ in range(2):
        rewards, vals = tr.train_once()
        assert torch.isfinite(rewards).all() and all(np.isfinite(v).all() for v in vals) and len(vals) == 2
    assert any(not torch.equal(a, b) for a, b in zip(before, tr.policies[0].parameters()))
    path = str(tmp_path / "ep0.pt")
    tr.save(path)
    ck = torch.load(path, map_location="cpu")
    assert set(ck) == {"models", "ob_rms"} and ck["ob_rms"] == (None, None) and len(ck["models"]) == 6
    assert len(ck["models"][0]) == 24 and sum(v.numel() for v in ck["models"][0].values()) == 158153
    tr2 = ro.BatchedTrainer(8, 3, 3, num_steps=4, hidden_dim=128)
    tr2.load_models(ck["models"])
    for a, b in zip(tr.policies[1].parameters(), tr2.policies[1].parameters()):
        assert torch.equal(a, b)


def test_fused_gae_equals_torch_mirror_bitwise():
    """rl_gae (one launch for all agents/envs) == RolloutStorage.compute_returns_batched per agent, bit for bit."""
    ro = import_module(PKG + ".rollout")
    g = torch.Generator().manual_seed(3)
    T, A, E = 37, 5, 1000
    R = ro.SharedRollouts(T, A, E, "cuda")
    R.rewards.copy_(torch.randn(T, A, E, generator=g))
    R.value_preds.copy_(torch.randn(T + 1, A, E, 1, generator=g))
    R.masks.copy_((torch.rand(T + 1, A, E, 1, generator=g) > 0.2).float())
    R.ends.copy_(torch.rand(T + 1, E, generator=g) < 0.08)
    R.ends[T] = True
    R.returns.copy_(torch.randn(T + 1, A, E, 1, generator=g))        # end-point slots must survive untouched
    nv = torch.randn(A, E, generator=g).cuda()
    ref_val, ref_ret = R.value_preds.clone(), R.returns.clone()
    R.compute_returns(nv, 0.99, 0.95)
    torch.cuda.synchronize()
    got_ret, got_val = R.returns.clone(), R.value_preds.clone()
    R.value_preds.copy_(ref_val)
    R.returns.copy_(ref_ret)
    for i in range(A):
        R.agents[i].compute_returns_batched(nv[i].view(E, 1), R.ends, 0.99, 0.95)
    assert torch.equal(R.returns, got_ret) and torch.equal(R.value_preds, got_val)


def test_fused_policy_rollout_statistics():
    """Rollouts driven by the fused policy kernel: stored log-probs / values agree with the torch module evaluated on
    the stored observations and actions (the PPO ratio starts at ~1), and sampled actions follow its distribution."""
    ro = import_module(PKG + ".rollout")
    torch.manual_seed(4)
    tr = ro.BatchedTrainer(512, 3, 3, num_steps=16, max_episode_steps=25, hidden_dim=128, seed=5)
    assert tr.fused is not None
    tr.collect()
    tr.wrap_horizon()
    R = tr.roll
    for t, (lo, hi, olo, ohi) in enumerate(((0, 3, 3, 6), (3, 6, 0, 3))):
        own, opp = R.obs[:-1, lo:hi], R.obs[:-1, olo:ohi]                 # [T, n, E, 6]
        T = own.shape[0]
        for s in (0, T // 2, T - 1):
            with torch.no_grad():
                v, lp, ent, _ = tr.policies[t].evaluate_actions(own[s].reshape(-1, 6), None, opp[s].reshape(-1, 6), None,
                                                                R.actions[s, lo:hi].reshape(-1, 1))
            assert (v.view(3, -1) - R.value_preds[s, lo:hi, :, 0]).abs().max() < 2e-2
            assert (lp.view(3, -1) - R.action_log_probs[s, lo:hi, :, 0]).abs().max() < 2e-2
    assert int(R.actions.min()) >= 0 and int(R.actions.max()) <= 7
    assert len(torch.unique(R.actions)) == 8
    assert torch.isfinite(R.returns).all()


def test_ppo_ratio_starts_at_one_with_trained_weights():
    """Rollout log-probs come from the fp16-operand kernel, the update re-evaluates in fp32: with the shipped trained
    policy the PPO ratio of the first minibatch (ppo.py:163) is 1 +- ~1e-2 as collected, and 1 +- 1e-5 after
    BatchedTrainer.recompute_old() (default in train_once)."""
    import policy_util as pu
    ro = import_module(PKG + ".rollout")
    fused = import_module(PKG + ".rlcore.fused")
    torch.manual_seed(0)
    tr = ro.BatchedTrainer(512, 5, 5, num_steps=16, max_episode_steps=100, seed=4)
    sd = pu.trained_state_dict()[0]
    tr.load_models([sd] * 10)
    assert tr.exact_old
    tr.collect()
    R = tr.roll

    def first_minibatch_ratio():
        trainer, pol = tr.trainers[0], tr.policies[0]
        adv = torch.zeros(5, tr.T, tr.E, device=R.obs.device)
        idx = torch.arange(0, tr.T * tr.E, 3, device=R.obs.device)
        ob, mask, oo, act, vp, ret, msk, olp, ad, alive = fused.gather_minibatch(R, idx, 0, 5, 5, 5, adv)
        with torch.no_grad():
            pol.fused_no_grad = True
            v, lp, ent, _ = pol.evaluate_actions(ob, None, oo, msk, act)
            pol.fused_no_grad = False
        return (torch.exp(lp - olp) - 1).abs(), (v - vp).abs()
    r0, v0 = first_minibatch_ratio()
    tr.recompute_old()
    r1, v1 = first_minibatch_ratio()
    print("|ratio - 1| as collected: max %.2e mean %.2e; after recompute_old: max %.2e;  |value - value_pred|: %.2e -> %.2e"
          % (float(r0.max()), float(r0.mean()), float(r1.max()), float(v0.max()), float(v1.max())))
    assert float(r0.max()) < 4e-2 and float(r0.mean()) < 3e-3
    assert float(r1.max()) < 1e-5 and float(v1.max()) < 1e-4
    tr.wrap_horizon()
    vals = tr.update()
    assert all(np.isfinite(v).all() for v in vals)


def test_graph_replayed_rollouts_equal_eager():
    """collect() replayed from one CUDA graph == the same loop launched eagerly, bit for bit, across updates
    (sampling counter on the device, weight blob refreshed in place)."""
    ro = import_module(PKG + ".rollout")
    trs = []
    for graphed in (True, False):
        torch.manual_seed(11)
        trs.append(ro.BatchedTrainer(128, 3, 3, num_steps=12, max_episode_steps=9, hidden_dim=128, seed=3, ppo_epoch=1,
                                     num_mini_batch=2, graph_rollouts=graphed))
    for it in range(4):
        for tr in trs:
            tr.collect()
            tr.wrap_horizon()
        a, b = trs[0].roll, trs[1].roll
        for name in ("obs", "rewards", "actions", "actions_i32", "value_preds", "action_log_probs", "masks", "returns", "ends"):
            assert torch.equal(getattr(a, name), getattr(b, name)), (it, name)
        assert torch.equal(trs[0].episode_rewards, trs[1].episode_rewards)
        if it == 1:                                   # weights change under the captured graph
            torch.manual_seed(5)
            trs[0].update()
            torch.manual_seed(5)
            trs[1].update()
        for tr in trs:
            tr.after_update()
    assert trs[0]._graph is not None and trs[1]._graph is None


def test_fused_gather_and_loss_match_torch_path():
    """rl_gather_minibatch == magent_feed_forward_generator (bit-equal rows); rl_ppo_loss == the torch expressions of
    ppo.py:150-187 (values and autograd gradients); a fused update tracks the unfused one."""
    ro = import_module(PKG + ".rollout")
    ppo = import_module(PKG + ".rlcore.algo.ppo")
    fused = import_module(PKG + ".rlcore.fused")
    torch.manual_seed(3)
    tr = ro.BatchedTrainer(200, 3, 2, num_steps=24, max_episode_steps=9, seed=4, ppo_epoch=1, num_mini_batch=4)
    tr.collect(); tr.wrap_horizon()
    R = tr.roll
    own, opp = [R.agents[i] for i in (0, 1, 2)], [R.agents[i] for i in (3, 4)]
    trainer = tr.trainers[0]
    adv_list = [trainer._advantages(r) for r in own]
    idx = torch.randperm(24 * 200)[:1111].cuda()
    ref = next(ppo.magent_feed_forward_generator(own, opp, adv_list, 4, index_batches=[idx]))
    (obs_b, mask_b, opp_b, hid_b, act_b, val_b, ret_b, msk_b, olp_b, adv_b) = ref
    adv = torch.stack([a[..., 0] for a in adv_list]).contiguous()
    g = fused.gather_minibatch(R, idx, 0, 3, 3, 2, adv)
    for got, want in zip(g[:9], (obs_b, mask_b, opp_b, act_b, val_b, ret_b, msk_b, olp_b, adv_b)):
        assert got.shape == want.shape and torch.equal(got, want)
    assert float(g[9]) == float(mask_b.sum())
    # loss + gradients
    gen = torch.Generator().manual_seed(0)
    N = mask_b.numel()
    values = (val_b + 0.3 * torch.randn(N, 1, generator=gen).cuda()).requires_grad_()
    logp = (olp_b + 0.3 * torch.randn(N, 1, generator=gen).cuda()).requires_grad_()
    ent = (1.5 + 0.2 * torch.randn(N, generator=gen).cuda()).requires_grad_()
    clip, vc, ec = 0.2, 0.5, 0.01
    norm = mask_b.sum().view(1)
    total, stats = fused.ppo_loss(values, logp, ent, val_b, ret_b, olp_b, adv_b, mask_b, norm, clip, vc, ec)
    total.backward()
    got_grads = [values.grad.clone(), logp.grad.clone(), ent.grad.clone()]
    for t in (values, logp, ent):
        t.grad = None
    mm = lambda x: x.mean() / mask_b.mean()
    entropy = mm(ent * mask_b[:, 0])
    ratio = mask_b * torch.exp(logp - olp_b)
    action_loss = mm(mask_b * -torch.min(ratio * adv_b, torch.clamp(ratio, 1 - clip, 1 + clip) * adv_b))
    clipped = val_b + (values - val_b).clamp(-clip, clip)
    value_loss = mm(0.5 * torch.max((values - ret_b).pow(2), (clipped - ret_b).pow(2)) * mask_b)
    ref_total = value_loss * vc + action_loss - entropy * ec
    ref_total.backward()
    assert torch.allclose(stats[:3], torch.stack([value_loss, action_loss, entropy]).detach(), rtol=2e-5, atol=1e-6)
    assert abs(float(total) - float(ref_total)) < 2e-5 * max(1.0, abs(float(ref_total)))
    for got, t in zip(got_grads, (values, logp, ent)):
        assert torch.allclose(got, t.grad, rtol=1e-4, atol=1e-9)
    # whole update: fused and unfused trainers see the same rollouts and permutations
    trs = []
    for fu in (True, False):
        torch.manual_seed(8)
        t2 = ro.BatchedTrainer(128, 3, 3, num_steps=16, max_episode_steps=9, seed=6, ppo_epoch=2, num_mini_batch=4, fused_update=fu)
        t2.collect(); t2.wrap_horizon()
        torch.manual_seed(9)
        trs.append(t2.update())
    for a, b in zip(trs[0], trs[1]):
        assert np.allclose(a, b, rtol=2e-3, atol=2e-4), (trs[0], trs[1])


@pytest.mark.parametrize("rows,cols,ld", [(1, 8, 8), (1184, 128, 128), (196608, 1, 1), (70001, 8, 8), (5000, 64, 192), (333, 256, 256),
                                          (100000, 32, 32), (7, 3, 3)])
def test_colsum_matches_float64_and_is_reproducible(rows, cols, ld):
    """rl_colsum (bias gradients) == the float64 column sums within fp32 summation error, bit-reproducible; widths the kernel
    does not cover (not a power of two) fall back to torch."""
    fused = import_module(PKG + ".rlcore.fused")
    gen = torch.Generator().manual_seed(rows + cols)
    big = torch.randn(rows, ld, generator=gen).cuda()
    x = big[:, :cols]
    got = fused.colsum(x)
    ref = x.double().sum(0)
    bound = x.double().abs().sum(0)
    assert got.shape == (cols,) and float(((got.double() - ref).abs() / (bound + 1e-30)).max()) < 2e-6
    assert torch.equal(got, fused.colsum(x))


def test_fold_weights_matches_autograd():
    """fused.fold_weights (rl_small_matmul: the [d, d] folds of the attention projections, forward and backward in two launches
    each) == the torch expressions W_query W_key^T, (W_val W_out) U2^T and their autograd gradients; bit-reproducible."""
    fused = import_module(PKG + ".rlcore.fused")
    gen = torch.Generator().manual_seed(9)
    for d, k, h in ((128, 128, 128), (64, 48, 72), (128, 37, 200)):
        mk = lambda *sh: (torch.randn(*sh, generator=gen) / sh[-1] ** 0.5).cuda().requires_grad_()
        Wq, Wk, Wv, Wout, W = mk(d, k), mk(d, k), mk(d, k), mk(k, d), mk(h, 2 * d)
        U2 = W[:, d:]
        g1, g2 = torch.randn(d, d, generator=gen).cuda(), torch.randn(d, h, generator=gen).cuda()
        Mqk, Wz = fused.fold_weights(Wq, Wk, Wv, Wout, U2)
        got = torch.autograd.grad((Mqk * g1).sum() + (Wz * g2).sum(), (Wq, Wk, Wv, Wout, W))
        rq, rz = Wq.double() @ Wk.double().t(), (Wv.double() @ Wout.double()) @ U2.double().t()
        ref = torch.autograd.grad((rq * g1.double()).sum() + (rz * g2.double()).sum(), (Wq, Wk, Wv, Wout, W))
        assert float((Mqk.double() - rq).abs().max()) < 1e-5 and float((Wz.double() - rz).abs().max()) < 1e-5
        for a, b in zip(got, ref):
            assert float((a.double() - b.double()).abs().max()) < 1e-5 * (1.0 + float(b.abs().max())), (d, k, h)
        M2, Z2 = fused.fold_weights(Wq, Wk, Wv, Wout, U2)
        assert torch.equal(M2, Mqk) and torch.equal(Z2, Wz)


def test_loss_from_logits_matches_categorical_and_torch_loss():
    """rl_ppo_loss_logits == FixedCategorical(logits).log_probs / .entropy (rlcore/distributions.py:9-17) fed to the torch
    expressions of ppo.py:150-187: statistics, the gradient with respect to values and LOGITS (torch autograd through
    log_softmax / gather / entropy), bit-reproducible statistics, and the evaluation-only form."""
    fused = import_module(PKG + ".rlcore.fused")
    dist = import_module(PKG + ".rlcore.distributions")
    gen = torch.Generator().manual_seed(5)
    for N in (1, 777, 70001):
        r = lambda *s: torch.randn(*s, generator=gen).cuda()
        logits = (2.0 * r(N, 8)).requires_grad_()
        values = r(N, 1).requires_grad_()
        actions = torch.randint(0, 8, (N, 1), generator=gen).cuda()
        old_v, ret, adv = values.detach() + 0.3 * r(N, 1), r(N, 1), r(N, 1)
        mask = (torch.rand(N, 1, generator=gen) < 0.8).float().cuda()
        mask[0] = 1.0
        clip, vc, ec = 0.2, 0.5, 0.01
        d = dist.FixedCategorical(logits=logits)
        logp, ent = d.log_probs(actions), d.entropy()
        old_lp = logp.detach() + 0.3 * r(N, 1)
        norm = mask.sum().view(1)
        total, stats = fused.ppo_loss_logits(values, logits, actions, old_v, ret, old_lp, adv, mask, norm, clip, vc, ec)
        total.backward()
        got = [values.grad.clone(), logits.grad.clone()]
        values.grad = logits.grad = None
        mm = lambda x: x.mean() / mask.mean()
        entropy = mm(ent * mask[:, 0])
        ratio = mask * torch.exp(logp - old_lp)
        action_loss = mm(mask * -torch.min(ratio * adv, torch.clamp(ratio, 1 - clip, 1 + clip) * adv))
        clipped = old_v + (values - old_v).clamp(-clip, clip)
        value_loss = mm(0.5 * torch.max((values - ret).pow(2), (clipped - ret).pow(2)) * mask)
        ref_total = value_loss * vc + action_loss - entropy * ec
        ref_total.backward()
        assert torch.allclose(stats[:3], torch.stack([value_loss, action_loss, entropy]).detach(), rtol=3e-5, atol=1e-6)
        assert abs(float(total) - float(ref_total)) < 3e-5 * max(1.0, abs(float(ref_total)))
        assert torch.allclose(got[0], values.grad, rtol=1e-4, atol=1e-9)
        scale = float(logits.grad.abs().max())
        assert float((got[1] - logits.grad).abs().max()) < 2e-5 * scale + 1e-10, (N, float((got[1] - logits.grad).abs().max()), scale)
        # same inputs, same bits (block-ordered partial sums, no atomics on the statistics)
        _, again = fused.ppo_loss_logits(values, logits, actions, old_v, ret, old_lp, adv, mask, norm, clip, vc, ec)
        assert torch.equal(again, stats)
        lp2, ent2 = fused.categorical_eval(logits, actions)
        assert float((lp2 - logp.detach()[:, 0]).abs().max()) < 2e-6 and float((ent2 - ent.detach()).abs().max()) < 2e-6


@pytest.mark.parametrize("n,m,hid", [(3, 3, 128), (5, 4, 128), (1, 2, 128), (2, 5, 64)])
def test_fused_training_attention_matches_bmm_path(n, m, hid):
    """MPNN.evaluate_actions with the attention kernels (rl_attn_forward / rl_attn_backward) == the bmm/softmax mirror
    of mpnn.py:249-331,376-443: outputs, attention matrices and every parameter gradient."""
    mp = import_module(PKG + ".mpnn")
    ro = import_module(PKG + ".rollout")
    torch.manual_seed(n * 7 + m)
    net = mp.MPNN(action_space=ro._Shape(8), num_agents=n, num_opp_agents=m, input_size=6, hidden_dim=hid).cuda()
    with torch.no_grad():
        for p in net.parameters():
            if p.dim() == 1:
                p.uniform_(-0.3, 0.3)
    B = 777
    own, opp = torch.randn(n * B, 6, device="cuda"), torch.randn(m * B, 6, device="cuda")
    act = torch.randint(0, 8, (n * B, 1), device="cuda")
    w = torch.randn(n * B, 1, device="cuda")
    fused = import_module(PKG + ".rlcore.fused")

    def run(fused_on, fold, dense, trace=None):
        net.fused_attention, net.fold_projections, fused.DENSE, fused.RELU_TRACE = fused_on, fold, dense, trace
        try:
            net.zero_grad()
            v, lp, ent, _ = net.evaluate_actions(own, None, opp, None, act)
            ((v * w).sum() + (lp * w).sum() * 0.7 + ent.sum() * 0.3).backward()
        finally:
            fused.DENSE, fused.RELU_TRACE = "tcgen05", None
        grads = {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}
        return (v.detach(), lp.detach(), ent.detach(), net.attn_mat, net.opp_attn_mat, grads)

    def same(a, b, what):
        for x, y in zip(a[:3], b[:3]):
            assert torch.allclose(x, y, rtol=1e-4, atol=2e-5), what
        assert np.allclose(a[3], b[3], atol=1e-5) and np.allclose(a[4], b[4], atol=1e-5), what
        assert set(a[5]) == set(b[5])
        for k in a[5]:
            scale = float(a[5][k].abs().max()) + 1e-6
            err = float((a[5][k] - b[5][k]).abs().max())
            assert err < 2e-4 * scale, (what, k, err, scale)

    # (1) the restructuring: attention kernels with explicit projections, then with the projections folded into the weights,
    #     against the bmm/softmax mirror, all three on the library GEMMs (the same products round the same way)
    mirror = run(False, False, "cublas")
    same(mirror, run(True, False, "cublas"), "attention kernels")
    rec = {"mode": "record", "outs": []}
    folded_lib = run(True, True, "cublas", rec)
    same(mirror, folded_lib, "folded projections")
    # (2) the arithmetic: the production path (folded, every dense product on the tcgen05 kernels) against the same path
    #     on the library GEMMs, with the ReLU decisions of the backward pinned to the library run's (fused.RELU_TRACE:
    #     a pre-activation within rounding of zero may legitimately fall either side) -- and those are counted
    rep = {"mode": "replay", "outs": rec["outs"], "pos": 0, "flips": 0}
    folded_tg = run(True, True, "tcgen05", rep)
    fused.tg_check_status("cuda:0")
    assert rep["pos"] == len(rec["outs"]) > 0
    n_act = sum(o.numel() for o in rec["outs"])
    assert rep["flips"] <= max(4, n_act // 200000), (rep["flips"], n_act)      # observed: 0-2 sign flips in ~1.5M activations
    same(folded_lib, folded_tg, "tcgen05 dense kernels (%d ReLU sign flips of %d pinned)" % (rep["flips"], n_act))
    same(mirror, folded_tg, "production path against the mirror")


def test_graph_captured_update_tracks_eager_update():
    """JointPPO(graph_update=True): three eager steps, then the optimizer step replayed from one CUDA graph; losses and
    weights follow the eager fused update (Adam's capturable code path rounds its bias correction on the device)."""
    ro = import_module(PKG + ".rollout")
    out = []
    for gu in (True, False):
        torch.manual_seed(21)
        tr = ro.BatchedTrainer(256, 3, 3, num_steps=16, max_episode_steps=9, seed=2, ppo_epoch=2, num_mini_batch=8, graph_update=gu)
        vals = []
        for it in range(2):
            tr.collect(); tr.wrap_horizon()
            torch.manual_seed(40 + it)
            vals.append(tr.update())
            tr.after_update()
        out.append((vals, [p.detach().clone() for pol in tr.policies for p in pol.parameters()], tr))
    assert out[0][2].trainers[0]._g is not None and out[0][2].trainers[0]._g["graph"] is not None
    for va, vb in zip(out[0][0], out[1][0]):
        assert np.allclose(va, vb, rtol=5e-3, atol=5e-4), (va, vb)
    worst = max(float((a - b).abs().max()) for a, b in zip(out[0][1], out[1][1]))
    assert worst < 2e-3, worst


@pytest.mark.parametrize("graph_update", [False, True])
def test_overlapped_team_updates_equal_sequential(graph_update):
    """BatchedTrainer.update enqueues the two teams' updates on two streams (overlap_teams, the default on one rank).  The
    teams share nothing but read-only rollout blocks, every kernel has a fixed summation order and its own scratch scope,
    so losses and weights must equal those of one update after the other BIT FOR BIT."""
    ro = import_module(PKG + ".rollout")
    out = []
    for ov in (True, False):
        torch.manual_seed(33)
        tr = ro.BatchedTrainer(512, 3, 3, num_steps=16, max_episode_steps=9, seed=5, ppo_epoch=2, num_mini_batch=8,
                               graph_update=graph_update, overlap_teams=ov)
        vals = []
        for it in range(3):
            tr.collect(); tr.wrap_horizon()
            torch.manual_seed(70 + it)
            vals.append(tr.update())
            tr.after_update()
        torch.cuda.synchronize()
        out.append((vals, [p.detach().clone() for pol in tr.policies for p in pol.parameters()]))
        assert (getattr(tr, "_team_streams", None) is not None) == ov
    assert out[0][0] == out[1][0], (out[0][0], out[1][0])
    for a, b in zip(out[0][1], out[1][1]):
        assert torch.equal(a, b), float((a - b).abs().max())
    fused = import_module(PKG + ".rlcore.fused")
    fused.tg_check_status("cuda:0")


def test_joint_two_team_step_tracks_sequential_updates():
    """overlap_teams="joint" (the default form with a process group, forced here on one rank): both teams' forward / backward
    as the two branches of ONE captured graph, the several-ranks form of the step (un-normalised sums, one flat buffer for
    both teams, division by the normaliser inside the optimizer kernel).  Losses and weights follow one team after the other."""
    ro = import_module(PKG + ".rollout")
    out = []
    for ov in ("joint", False):
        torch.manual_seed(33)
        tr = ro.BatchedTrainer(512, 3, 3, num_steps=16, max_episode_steps=9, seed=5, ppo_epoch=2, num_mini_batch=8,
                               graph_update=True, overlap_teams=ov)
        vals = []
        for it in range(2):
            tr.collect(); tr.wrap_horizon()
            torch.manual_seed(70 + it)
            vals.append(tr.update())
            tr.after_update()
        torch.cuda.synchronize()
        out.append((vals, [p.detach().clone() for pol in tr.policies for p in pol.parameters()], tr))
    J = out[0][2].trainers[0]._joint
    assert J is not None and J["graph"] is not None and out[1][2].trainers[0]._g["graph"] is not None
    for va, vb in zip(out[0][0], out[1][0]):
        assert np.allclose(va, vb, rtol=5e-3, atol=5e-4), (va, vb)
    worst = max(float((a - b).abs().max()) for a, b in zip(out[0][1], out[1][1]))
    assert worst < 2e-3, worst
    import_module(PKG + ".rlcore.fused").tg_check_status("cuda:0")


def test_attacker_ensemble_play():
    """K frozen attacker checkpoints, one drawn per env at every episode start (learner.py:119-140,
    train_fortattack_v2.py:34-35,110-111): each env's attacker rows come from the checkpoint it is assigned to."""
    ro = import_module(PKG + ".rollout")
    pk = import_module(PKG + ".policy_kernel")
    mp = import_module(PKG + ".mpnn")
    sds = []
    for k in range(3):
        torch.manual_seed(100 + k)
        sds.append(mp.MPNN(action_space=ro._Shape(8), num_agents=3, num_opp_agents=3, input_size=6, hidden_dim=128).state_dict())
    torch.manual_seed(0)
    tr = ro.BatchedTrainer(300, 3, 3, num_steps=20, max_episode_steps=7, seed=2, attacker_ensemble=sds, ppo_epoch=1,
                           num_mini_batch=2, graph_rollouts=False)
    ids0 = tr.att_id.clone()
    seen = [ids0.clone()]
    R = tr.roll
    tr.collect()
    tr.wrap_horizon()
    # step 0: attacker values of env e equal checkpoint ids0[e]'s own forward on the same observations
    for k in range(3):
        o = tr.ensemble[k].forward(R.obs[0, 3:6].contiguous(), R.obs[0, 0:3].contiguous(), pk.MODE_ARGMAX)
        sel = ids0 == k
        assert sel.any()
        assert torch.equal(o["value"][:, sel], R.value_preds[0, 3:6, :, 0][:, sel])
    # the one-launch ensemble forward equals K compacted single-checkpoint launches and K masked full launches
    order, offsets = tr._ensemble_lists()
    own, opp = R.obs[3, 3:6].contiguous(), R.obs[3, 0:3].contiguous()
    one = pk.forward_ensemble(tr.ensemble, own, opp, order, offsets, pk.MODE_ARGMAX)
    comp = {k: torch.zeros_like(v) for k, v in one.items()}
    mask = {k: torch.zeros_like(v) for k, v in one.items()}
    for k in range(3):
        tr.ensemble[k].forward(own, opp, pk.MODE_ARGMAX, out=comp, sel_value=k, env_order=order, env_offsets=offsets)
        tr.ensemble[k].forward(own, opp, pk.MODE_ARGMAX, out=mask, env_sel=tr.att_id, sel_value=k)
    for key in one:
        assert torch.equal(one[key], comp[key]) and torch.equal(one[key], mask[key]), key
    assert not torch.equal(ids0, tr.att_id)                          # episodes ended (cap 7): new draws
    n_done = int((R.done != 0).sum())
    assert int(tr.ensemble_results.sum()) == n_done and int(tr.ensemble_results[:, 0].sum()) == 0
    assert (tr.ensemble_results.sum(1) > 0).all()
    tab = tr.ensemble_table()                                        # test_fortattack_v2.py's stats, per checkpoint
    assert tab.shape == (3, 8) and np.allclose(tab[:, 0] + tab[:, 1] + tab[:, 3], 1.0) and np.allclose(tab[:, 2], tab[:, 0] + tab[:, 1])
    assert (tab[:, 4] >= 0).all() and (tab[:, 4] <= 3).all() and (tab[:, 5] >= 0).all() and (tab[:, 5] <= 3).all()
    assert (tab[tab[:, 0] == 1.0, 5] == 0).all() and np.isfinite(tab).all()
    before = [p.detach().clone() for p in tr.policies[1].parameters()]
    vals = tr.update()                                               # guards only
    assert len(vals) == 1 and all(torch.equal(a, b) for a, b in zip(before, tr.policies[1].parameters()))
    # graph-replayed ensemble rollouts run and keep resampling
    tr2 = ro.BatchedTrainer(300, 3, 3, num_steps=20, max_episode_steps=7, seed=2, attacker_ensemble=sds)
    for _ in range(3):
        tr2.collect(); tr2.wrap_horizon(); tr2.after_update()
    assert tr2._graph is not None and int(tr2.ensemble_results.sum()) > 2 * n_done


def test_two_gpu_shards_and_nccl_update():
    """World size 2 over NCCL (skipped on a one-GPU box; profiles/r1d_dist_train_2gpu.log holds a recorded run)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "dist_train_gpu.py")],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "-> OK" in out.stdout, (out.stdout[-1500:], out.stderr[-1500:])
