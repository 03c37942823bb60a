"""Golden vectors for the rollout / learner half of the path, produced by RUNNING THE UNCHANGED REFERENCE
(mpnn.py, rlcore/storage.py, rlcore/algo/ppo.py, rlcore/distributions.py) in this container.

    python tests/golden/make_rl_golden.py      -> tests/golden/rl_mpnn.npz, rl_ppo.npz

rl_mpnn.npz   for several (n_team, n_opp, hidden) cases: the reference MPNN's parameters after
              torch.manual_seed(seed) construction, random agent-major inputs, and the outputs of
              evaluate_actions / get_value / act(deterministic) and act(sampled, seeded), attention matrices;
              plus one case with the shipped checkpoint marlsave/tmp_1/ep2520.pt (guard policy, 5v5).
rl_ppo.npz    synthetic rollouts of a 3v2 team pair (hidden 32): RolloutStorage.compute_returns over
              several episode segments (learner.py:191-211 protocol), then JointPPO.update with a fixed
              seed: returns, losses and every parameter after the update.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.import_reference()
from mpnn import MPNN  # noqa: E402  (reference)
from rlcore.algo import JointPPO  # noqa: E402
from rlcore.storage import RolloutStorage  # noqa: E402


class Shape(object):
    def __init__(self, *shape):
        self.shape = shape


def sd_np(module, prefix):
    return {prefix + k: v.detach().cpu().numpy().copy() for k, v in module.state_dict().items()}


def rand_obs(g, rows):
    o = torch.randn(rows, 6, generator=g)
    o[:, 0] = (torch.rand(rows, generator=g) > 0.25).float()      # alive flag
    o[:, 3] = o[:, 3].abs() * 20                                    # headings grow large
    return o


def mpnn_cases():
    out = {}
    cases = [("a", 5, 5, 128, 7, 3), ("b", 3, 3, 32, 16, 4), ("c", 1, 2, 32, 9, 5), ("d", 2, 1, 32, 5, 6)]
    for name, n, m, hid, B, seed in cases:
        torch.manual_seed(seed)
        net = MPNN(action_space=Shape(8), num_agents=n, num_opp_agents=m, num_entities=0, input_size=6,
                   hidden_dim=hid, pos_index=2, mask_dist=None, entity_mp=False, policy_layers=1)
        g = torch.Generator().manual_seed(100 + seed)
        own, opp = rand_obs(g, n * B), rand_obs(g, m * B)
        act = torch.randint(0, 8, (n * B, 1), generator=g)
        with torch.no_grad():
            v, lp, ent, _ = net.evaluate_actions(own, None, opp, None, act)
            attn, opp_attn = np.array(net.attn_mat), np.array(net.opp_attn_mat)
            gv = net.get_value(own, None, opp, None)
            _, a_det, lp_det, _ = net.act(own, None, opp, None, deterministic=True)
            torch.manual_seed(999)
            _, a_smp, lp_smp, _ = net.act(own, None, opp, None, deterministic=False)
        if hid == 128:       # init parity only needs checksums at full size
            for k, val in net.state_dict().items():
                out["%s/init_sum/%s" % (name, k)] = np.array([val.double().sum().item(), val.double().abs().sum().item()])
        else:
            out.update(sd_np(net, name + "/param/"))
        out.update({name + "/meta": np.array([n, m, hid, B, seed]), name + "/own": own.numpy(), name + "/opp": opp.numpy(),
                    name + "/act": act.numpy(), name + "/value": v.numpy(), name + "/logp": lp.numpy(),
                    name + "/entropy": ent.numpy(), name + "/attn": attn, name + "/opp_attn": opp_attn,
                    name + "/get_value": gv.numpy(), name + "/a_det": a_det.numpy(), name + "/lp_det": lp_det.numpy(),
                    name + "/a_smp": a_smp.numpy(), name + "/lp_smp": lp_smp.numpy()})
        if hid == 128:
            out[name + "/value_full_init"] = v.numpy()
    # shipped checkpoint: guard policy of marlsave/tmp_1/ep2520.pt (5v5), stored as fp16-safe float32
    ck = torch.load(os.path.join(ref_shim.REF, "marlsave", "tmp_1", "ep2520.pt"), map_location="cpu")
    net = MPNN(action_space=Shape(8), num_agents=5, num_opp_agents=5, num_entities=0, input_size=6)
    net.load_state_dict(ck["models"][0])
    g = torch.Generator().manual_seed(77)
    own, opp = rand_obs(g, 5 * 6), rand_obs(g, 5 * 6)
    act = torch.randint(0, 8, (30, 1), generator=g)
    with torch.no_grad():
        v, lp, ent, _ = net.evaluate_actions(own, None, opp, None, act)
    out.update(sd_np(net, "ckpt/param/"))
    out.update({"ckpt/own": own.numpy(), "ckpt/opp": opp.numpy(), "ckpt/act": act.numpy(), "ckpt/value": v.numpy(),
                "ckpt/logp": lp.numpy(), "ckpt/entropy": ent.numpy(),
                "ckpt/keys": np.array(list(ck["models"][0].keys()))})
    return out


def ppo_case():
    out = {}
    T, P, n, m, hid = 24, 3, 3, 2, 32
    torch.manual_seed(21)
    pol = MPNN(action_space=Shape(8), num_agents=n, num_opp_agents=m, num_entities=0, input_size=6, hidden_dim=hid)
    g = torch.Generator().manual_seed(5)
    team = [RolloutStorage(T, P, (6,), None, 1) for _ in range(n)]
    opp = [RolloutStorage(T, P, (6,), None, 1) for _ in range(m)]
    for k, r in enumerate(team + opp):
        r.obs.copy_(rand_obs(g, (T + 1) * P).view(T + 1, P, 6))
        r.rewards.copy_(torch.randn(T, P, 1, generator=g))
        r.value_preds.copy_(torch.randn(T + 1, P, 1, generator=g))
        r.action_log_probs.copy_(-torch.rand(T, P, 1, generator=g) * 2 - 0.5)
        r.actions.copy_(torch.randint(0, 8, (T, P, 1), generator=g))
        r.masks.copy_((torch.rand(T + 1, P, 1, generator=g) > 0.2).float())
        r.returns.copy_(torch.randn(T + 1, P, 1, generator=g) * 0.01)       # stale values the skipped indices keep
        for f in ("obs", "rewards", "value_preds", "action_log_probs", "actions", "masks", "returns"):
            out["in/%d/%s" % (k, f)] = getattr(r, f).numpy().copy()
    out.update(sd_np(pol, "param0/"))
    # Learner.wrap_horizon protocol (learner.py:191-211) with P envs sharing the end points (the reference
    # has P=1): segments [0,7) [8,15) [16,24)
    end_pts = [7, 15, 24]
    gamma, tau = 0.99, 0.95
    start = 0
    for e in end_pts:
        for k, r in enumerate(team):
            nv = torch.full((P, 1), 0.1 * (k + 1) + 0.01 * e)
            r.compute_returns(nv, True, gamma, tau, start, e)
        start = e + 1
    for k, r in enumerate(team):
        out["gae/%d/returns" % k] = r.returns.numpy().copy()
        out["gae/%d/value_preds" % k] = r.value_preds.numpy().copy()
    out["gae/end_pts"] = np.array(end_pts)
    out["gae/gamma_tau"] = np.array([gamma, tau])
    for clipped in (True, False):
        torch.manual_seed(21)
        pol = MPNN(action_space=Shape(8), num_agents=n, num_opp_agents=m, num_entities=0, input_size=6, hidden_dim=hid)
        algo = JointPPO(pol, 0.2, 2, 4, 0.5, 0.01, lr=1e-3, max_grad_norm=0.5, use_clipped_value_loss=clipped)
        torch.manual_seed(1234)
        losses = algo.update(team, opp)
        tag = "ppo_clip/" if clipped else "ppo_noclip/"
        out[tag + "losses"] = np.array(losses)
        out.update(sd_np(pol, tag + "param1/"))
    out["ppo/hparams"] = np.array([0.2, 2, 4, 0.5, 0.01, 1e-3, 0.5])
    out["ppo/meta"] = np.array([T, P, n, m, hid])
    return out


def main():
    np.savez_compressed(os.path.join(HERE, "rl_mpnn.npz"), **mpnn_cases())
    np.savez_compressed(os.path.join(HERE, "rl_ppo.npz"), **ppo_case())
    for f in ("rl_mpnn.npz", "rl_ppo.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
