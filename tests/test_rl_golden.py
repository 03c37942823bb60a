"""CPU: MPNN / RolloutStorage / JointPPO mirrors against vectors produced by the unchanged reference
(tests/golden/make_rl_golden.py).  float32 torch ops on both sides; tolerances are a few ulps of the sums."""
import os
from importlib import import_module

import numpy as np
import pytest
import torch

PKG = "emergent-multiagent-strategies_b200"
MPNN = import_module(PKG + ".mpnn").MPNN
storage = import_module(PKG + ".rlcore.storage")
algo = import_module(PKG + ".rlcore.algo")
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class Shape(object):
    def __init__(self, *shape):
        self.shape = shape


@pytest.fixture(scope="module")
def gm():
    return dict(np.load(os.path.join(GOLD, "rl_mpnn.npz")))


@pytest.fixture(scope="module")
def gp():
    return dict(np.load(os.path.join(GOLD, "rl_ppo.npz")))


def _params(g, prefix):
    return {k[len(prefix):]: torch.from_numpy(v) for k, v in g.items() if k.startswith(prefix)}


def _t(g, k):
    return torch.from_numpy(g[k])


@pytest.mark.parametrize("name", ["a", "b", "c", "d"])
def test_mpnn_same_seed_same_init_and_outputs(gm, name):
    n, m, hid, B, seed = [int(x) for x in gm[name + "/meta"]]
    torch.manual_seed(seed)
    net = MPNN(action_space=Shape(8), num_agents=n, num_opp_agents=m, num_entities=0, input_size=6, hidden_dim=hid)
    # identical construction-time random draws => identical parameters
    if hid == 128:
        for k, v in net.state_dict().items():
            s = gm["%s/init_sum/%s" % (name, k)]
            assert abs(v.double().sum().item() - s[0]) <= 1e-9 * max(1, s[1]), k
            assert abs(v.double().abs().sum().item() - s[1]) <= 1e-9 * max(1, s[1]), k
    else:
        ref = _params(gm, name + "/param/")
        assert set(ref) == set(net.state_dict())
        for k, v in net.state_dict().items():
            assert torch.equal(v, ref[k]), k
    own, opp, act = _t(gm, name + "/own"), _t(gm, name + "/opp"), _t(gm, name + "/act")
    with torch.no_grad():
        v, lp, ent, _ = net.evaluate_actions(own, None, opp, None, act)
        attn, opp_attn = net.attn_mat, net.opp_attn_mat
        gv = net.get_value(own, None, opp, None)
        _, a_det, lp_det, _ = net.act(own, None, opp, None, deterministic=True)
        torch.manual_seed(999)
        _, a_smp, lp_smp, _ = net.act(own, None, opp, None, deterministic=False)
    tol = dict(atol=2e-5, rtol=1e-4)
    assert np.allclose(v.numpy(), gm[name + "/value"], **tol) and np.allclose(gv.numpy(), gm[name + "/get_value"], **tol)
    assert np.allclose(lp.numpy(), gm[name + "/logp"], **tol) and np.allclose(ent.numpy(), gm[name + "/entropy"], **tol)
    assert np.allclose(attn, gm[name + "/attn"], atol=1e-5) and np.allclose(opp_attn, gm[name + "/opp_attn"], atol=1e-5)
    assert np.array_equal(a_det.numpy(), gm[name + "/a_det"]) and np.allclose(lp_det.numpy(), gm[name + "/lp_det"], **tol)
    assert np.array_equal(a_smp.numpy(), gm[name + "/a_smp"]), "sampling consumes the RNG like the reference"
    assert np.allclose(lp_smp.numpy(), gm[name + "/lp_smp"], **tol)
    assert v.shape == (n * B, 1) and lp.shape == (n * B, 1) and ent.shape == (n * B,) and a_smp.dtype == torch.int64


def test_shipped_checkpoint_loads_and_matches(gm):
    """marlsave/tmp_1/ep2520.pt guard policy: state_dict compatibility (learner.py:245-249, rlagent.py:20-21)."""
    net = MPNN(action_space=Shape(8), num_agents=5, num_opp_agents=5, num_entities=0, input_size=6)
    sd = _params(gm, "ckpt/param/")
    assert list(gm["ckpt/keys"]) == list(net.state_dict().keys())          # same names, same order
    net.load_state_dict(sd, strict=True)
    assert sum(p.numel() for p in net.parameters()) == 158153
    with torch.no_grad():
        v, lp, ent, _ = net.evaluate_actions(_t(gm, "ckpt/own"), None, _t(gm, "ckpt/opp"), None, _t(gm, "ckpt/act"))
    assert np.allclose(v.numpy(), gm["ckpt/value"], atol=5e-5, rtol=1e-4)
    assert np.allclose(lp.numpy(), gm["ckpt/logp"], atol=5e-5, rtol=1e-4)
    assert np.allclose(ent.numpy(), gm["ckpt/entropy"], atol=5e-5, rtol=1e-4)
    # team size is not baked into the parameters: the same weights drive a 3v3 policy
    net3 = MPNN(action_space=Shape(8), num_agents=3, num_opp_agents=3, num_entities=0, input_size=6)
    net3.load_state_dict(sd, strict=True)


def _rollouts(gp, device=None):
    T, P, n, m, hid = [int(x) for x in gp["ppo/meta"]]
    rs = []
    for k in range(n + m):
        r = storage.RolloutStorage(T, P, (6,), None, 1, device=device)
        for f in ("obs", "rewards", "value_preds", "action_log_probs", "actions", "masks", "returns"):
            getattr(r, f).copy_(torch.from_numpy(gp["in/%d/%s" % (k, f)]))
        rs.append(r)
    return rs[:n], rs[n:], (T, P, n, m, hid)


def test_rollout_storage_layout_and_insert():
    r = storage.RolloutStorage(5, 3, (6,), None, 1)
    shapes = {k: tuple(getattr(r, k).shape) for k in r._FIELDS}
    assert shapes == {"obs": (6, 3, 6), "recurrent_hidden_states": (6, 3, 1), "rewards": (5, 3, 1),
                      "value_preds": (6, 3, 1), "returns": (6, 3, 1), "action_log_probs": (5, 3, 1),
                      "actions": (5, 3, 1), "masks": (6, 3, 1)}
    assert r.actions.dtype == torch.int64 and bool((r.masks == 1).all()) and r.step == 0
    for s in range(5):
        r.insert(torch.full((3, 6), s + 1.0), torch.zeros(3, 1), torch.full((3, 1), s), torch.zeros(3, 1),
                 torch.full((3, 1), 10.0 + s), torch.full((3, 1), -1.0 * s), torch.zeros(3, 1))
    assert r.step == 0 and float(r.obs[5, 0, 0]) == 5 and float(r.value_preds[4, 0, 0]) == 14 and float(r.masks[5, 0, 0]) == 0
    r.after_update()
    assert float(r.obs[0, 0, 0]) == 5 and float(r.obs[1:].abs().sum()) == 0 and float(r.masks[0, 0, 0]) == 0


def test_gae_segments_match_reference(gp):
    team, _, (T, P, n, m, hid) = _rollouts(gp)
    gamma, tau = [float(x) for x in gp["gae/gamma_tau"]]
    end_pts = [int(e) for e in gp["gae/end_pts"]]
    team_b, _, _ = _rollouts(gp)
    start = 0
    for e in end_pts:
        for k, r in enumerate(team):
            r.compute_returns(torch.full((P, 1), 0.1 * (k + 1) + 0.01 * e), True, gamma, tau, start, e)
        start = e + 1
    for k, r in enumerate(team):
        assert np.allclose(r.returns.numpy(), gp["gae/%d/returns" % k], atol=1e-6)
        assert np.allclose(r.value_preds.numpy(), gp["gae/%d/value_preds" % k], atol=0)
    # batched sweep with per-env end flags: same numbers when value_preds[end] already hold the bootstraps
    ends = torch.zeros(T + 1, P, dtype=torch.bool)
    ends[end_pts] = True
    for k, r in enumerate(team_b):
        for e in end_pts[:-1]:
            r.value_preds[e] = 0.1 * (k + 1) + 0.01 * e
        r.compute_returns_batched(torch.full((P, 1), 0.1 * (k + 1) + 0.01 * end_pts[-1]), ends, gamma, tau)
        assert np.allclose(r.returns.numpy(), gp["gae/%d/returns" % k], atol=1e-6)
    # and with different boundaries per env it equals the per-env reference protocol
    r = _rollouts(gp)[0][0]
    q = _rollouts(gp)[0][0]
    per_env = [[5, 24], [9, 17, 24], [24]]
    ends = torch.zeros(T + 1, P, dtype=torch.bool)
    for p, pts in enumerate(per_env):
        ends[pts, p] = True
    r.compute_returns_batched(r.value_preds[T].clone(), ends, gamma, tau)
    for p, pts in enumerate(per_env):
        one = storage.RolloutStorage(T, 1, (6,), None, 1)
        for f in ("rewards", "value_preds", "masks", "returns"):
            getattr(one, f).copy_(getattr(q, f)[:, p:p + 1])
        start = 0
        for e in pts:
            one.compute_returns(one.value_preds[e].clone(), True, gamma, tau, start, e)
            start = e + 1
        assert torch.allclose(one.returns[:, 0], r.returns[:, p], atol=1e-6), p


@pytest.mark.parametrize("clipped", [True, False])
def test_joint_ppo_update_matches_reference(gp, clipped):
    team, opp, (T, P, n, m, hid) = _rollouts(gp)
    for k, r in enumerate(team):                                   # returns as the reference computed them
        r.returns.copy_(torch.from_numpy(gp["gae/%d/returns" % k]))
        r.value_preds.copy_(torch.from_numpy(gp["gae/%d/value_preds" % k]))
    clip, epochs, nmb, vcoef, ecoef, lr, mgn = [float(x) for x in gp["ppo/hparams"]]
    torch.manual_seed(21)
    pol = MPNN(action_space=Shape(8), num_agents=n, num_opp_agents=m, num_entities=0, input_size=6, hidden_dim=hid)
    for k, v in pol.state_dict().items():
        assert torch.equal(v, torch.from_numpy(gp["param0/" + k])), k
    ppo = algo.JointPPO(pol, clip, int(epochs), int(nmb), vcoef, ecoef, lr=lr, max_grad_norm=mgn,
                        use_clipped_value_loss=clipped)
    torch.manual_seed(1234)
    losses = ppo.update(team, opp)
    tag = "ppo_clip/" if clipped else "ppo_noclip/"
    assert np.allclose(losses, gp[tag + "losses"], rtol=2e-4, atol=1e-5), (losses, gp[tag + "losses"])
    worst = 0.0
    for k, v in pol.state_dict().items():
        ref = gp[tag + "param1/" + k]
        moved = np.abs(ref - gp["param0/" + k]).max()
        err = np.abs(v.numpy() - ref).max()
        worst = max(worst, err)
        assert err <= 2e-5 + 0.02 * moved, (k, err, moved)        # 8 Adam steps of 1e-3: |update| ~ 8e-3
    assert worst < 2e-4


def test_magent_generator_is_time_aligned(gp):
    team, opp, (T, P, n, m, hid) = _rollouts(gp)
    adv = [r.returns[:-1] - r.value_preds[:-1] for r in team]
    torch.manual_seed(0)
    perm = torch.randperm(T * P)
    torch.manual_seed(0)
    batches = list(algo.magent_feed_forward_generator(team, opp, adv, 4))
    assert len(batches) == 4
    mb = T * P // 4
    for b, smp in enumerate(batches):
        obs, mask, opp_obs = smp[0], smp[1], smp[2]
        assert obs.shape == (n * mb, 6) and opp_obs.shape == (m * mb, 6) and torch.equal(mask[:, 0], obs[:, 0])
        idx = perm[b * mb:(b + 1) * mb]
        for k in range(n):                                          # agent-major blocks over the same indices
            assert torch.equal(obs[k * mb:(k + 1) * mb], team[k].obs[:-1].view(-1, 6)[idx])
