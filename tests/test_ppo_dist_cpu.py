"""CPU, gloo, world_size 2: the multi-rank PPO step.  Two ranks, each holding half of the env
columns, must produce the same parameters and losses as one rank holding all of them."""
import os
import sys
from importlib import import_module

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = "emergent-multiagent-strategies_b200"
GOLD = os.path.join(ROOT, "tests", "golden", "rl_ppo.npz")


class Shape(object):
    def __init__(self, *shape):
        self.shape = shape


def _build(cols, process_group, clipped):
    sys.path.insert(0, ROOT)
    MPNN = import_module(PKG + ".mpnn").MPNN
    storage = import_module(PKG + ".rlcore.storage")
    algo = import_module(PKG + ".rlcore.algo")
    gp = dict(np.load(GOLD))
    T, P, n, m, hid = [int(x) for x in gp["ppo/meta"]]
    rs = []
    for k in range(n + m):
        r = storage.RolloutStorage(T, len(cols), (6,), None, 1)
        for f in ("obs", "rewards", "value_preds", "action_log_probs", "actions", "masks", "returns"):
            getattr(r, f).copy_(torch.from_numpy(gp["in/%d/%s" % (k, f)])[:, [c % P for c in cols]])
        rs.append(r)
    torch.manual_seed(21)
    pol = MPNN(action_space=Shape(8), num_agents=n, num_opp_agents=m, num_entities=0, input_size=6, hidden_dim=hid)
    ppo = algo.JointPPO(pol, 0.2, 2, 2, 0.5, 0.01, lr=1e-3, max_grad_norm=0.5, use_clipped_value_loss=clipped,
                        process_group=process_group)
    return pol, ppo, rs[:n], rs[n:]


def _worker(rank, world, port, clipped, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    # env columns are interleaved so that both ranks see the same time indices (same permutation of t)
    cols = [c for c in range(4) if c % world == rank]
    pol, ppo, team, opp = _build(cols, dist.group.WORLD, clipped)
    torch.manual_seed(100 + rank)                       # different local seeds: rank 0's permutation must win
    losses = ppo.update(team, opp)
    if rank == 0:
        q.put((losses, {k: v.numpy() for k, v in pol.state_dict().items()}))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("clipped", [True, False])
def test_two_ranks_equal_one_rank(clipped):
    gp = dict(np.load(GOLD))
    assert int(gp["ppo/meta"][1]) == 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000 + (7 if clipped else 0)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, clipped, q)) for r in range(2)]
    for p in procs:
        p.start()
    losses2, sd2 = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # One rank holding all four columns, fed the minibatches the two ranks form together: for every index
    # (t, e_local) of rank 0's permutation, the rows of both shards (global column = 2*e_local + rank).
    pol, ppo, team, opp = _build([0, 1, 2, 3], None, clipped)
    T = int(gp["ppo/meta"][0])
    torch.manual_seed(100)
    batches = []
    for _ in range(2):
        perm = torch.randperm(T * 2)
        mb = T * 2 // 2
        ep = []
        for i in range(0, T * 2, mb):
            idx = perm[i:i + mb]
            t, el = idx // 2, idx % 2
            ep.append(torch.cat([t * 4 + 2 * el, t * 4 + 2 * el + 1]))
        batches.append(ep)
    losses1 = ppo.update(team, opp, index_batches=batches)
    assert np.allclose(losses1, losses2, rtol=1e-4, atol=1e-6), (losses1, losses2)
    for k, v in pol.state_dict().items():
        assert np.allclose(v.numpy(), sd2[k], atol=5e-6), k
