"""Multi-GPU check (run under torchrun, one rank per GPU, NCCL):
  1. env shards: rank r's envs [r*E, (r+1)*E) equal the same slice of a single-GPU run with G*E envs
     (reset streams keyed by global env id; policy samples keyed by env_id0) -- bit-equal, no communication;
  2. BatchedTrainer with a process group: one all-reduce per optimizer step keeps the replicas' weights identical.
Usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/dist_train_gpu.py"""
import os
import sys
from importlib import import_module

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "emergent-multiagent-strategies_b200"


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ro = import_module(PKG + ".rollout")
    E, T = 256, 16
    torch.manual_seed(0)                                   # same initial weights everywhere (also broadcast by the trainer)
    tr = ro.BatchedTrainer(E, 3, 3, num_steps=T, max_episode_steps=12, device=dev, seed=7, env_id0=rank * E,
                           ppo_epoch=1, num_mini_batch=4, process_group=dist.group.WORLD)
    tr.collect()
    tr.wrap_horizon()
    # -- 1. shard invariance against one big single-GPU run, computed on every rank ------------------
    torch.manual_seed(0)
    ref = ro.BatchedTrainer(world * E, 3, 3, num_steps=T, max_episode_steps=12, device=dev, seed=7, env_id0=0,
                            ppo_epoch=1, num_mini_batch=4)
    ref.load_models(tr.state()["models"])
    ref.collect()
    ref.wrap_horizon()
    sl = slice(rank * E, (rank + 1) * E)
    ok = True
    for name in ("obs", "rewards", "actions", "value_preds", "action_log_probs", "masks", "returns"):
        a, b = getattr(tr.roll, name), getattr(ref.roll, name)[:, :, sl]
        same = torch.equal(a, b)
        ok &= same
        if not same:
            print("rank %d: %s differs, max |d| %.3e" % (rank, name, float((a.float() - b.float()).abs().max())), flush=True)
    # -- 2. synchronous update ------------------------------------------------------------------------
    vals = tr.update()
    flat = torch.cat([p.detach().reshape(-1) for pol in tr.policies for p in pol.parameters()])
    lo, hi = flat.clone(), flat.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    spread = float((hi - lo).abs().max())
    ok &= spread == 0.0 and all(torch.isfinite(torch.tensor(v)).all() for v in vals)
    # -- 3. graph-replayed optimizer steps: the joint two-team step (both teams' branches of one graph around ONE all-reduce,
    #       the default with a process group) against one team after the other (one captured all-reduce per team step) --------
    res = []
    for ov in (True, False):
        torch.manual_seed(0)
        tg = ro.BatchedTrainer(E, 3, 3, num_steps=T, max_episode_steps=12, device=dev, seed=7, env_id0=rank * E, ppo_epoch=2,
                               num_mini_batch=4, process_group=dist.group.WORLD, graph_update=True, overlap_teams=ov)
        tg.load_models(tr.state()["models"])
        gv = []
        for it in range(2):
            tg.collect(); tg.recompute_old(); tg.wrap_horizon()
            torch.manual_seed(100 + it)
            gv.append(tg.update())
            tg.after_update()
        w = torch.cat([p.detach().reshape(-1) for pol in tg.policies for p in pol.parameters()])
        lo, hi = w.clone(), w.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        captured = getattr(tg.trainers[0], "_joint", None) is not None if ov else tg.trainers[0]._g is not None
        res.append((w, gv, float((hi - lo).abs().max()), captured))
        for trn in tg.trainers:
            trn.release_graphs()
        del tg
    joint_diff = float((res[0][0] - res[1][0]).abs().max())
    loss_diff = max(abs(a - b) for va, vb in zip(res[0][1], res[1][1]) for ta, tb in zip(va, vb) for a, b in zip(ta, tb))
    # two ranks: a + b == b + a, so one collective over the shared buffer and one per team give the same bits; more ranks: the
    # collective's summation order may depend on the buffer layout, and Adam turns a last-bit gradient difference of a
    # cancelling component into a fraction of lr per step
    tol_w, tol_l = (0.0, 0.0) if world == 2 else (2e-3, 5e-3)
    ok &= res[0][2] == 0.0 and res[1][2] == 0.0 and res[0][3] and res[1][3] and joint_diff <= tol_w and loss_diff <= tol_l
    if rank == 0:
        print("dist_train_gpu: joint two-team step vs sequential: max weight difference %.2e, loss difference %.2e, replica "
              "spreads %.1e / %.1e, graphs captured %s / %s" % (joint_diff, loss_diff, res[0][2], res[1][2], res[0][3], res[1][3]), flush=True)
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("dist_train_gpu: world %d shard-invariant rollouts %s, replica weight spread after update %.1e, losses %s -> %s"
              % (world, ok, spread, vals, "OK" if flag.item() == 1.0 else "FAILED"), flush=True)
    # captured optimizer steps hold the group's NCCL kernels: drop them before the group goes away (a trainer that was given
    # graph_update=True and is still alive here makes destroy_process_group wait for ever)
    tr.release_graphs()
    ref.release_graphs()
    torch.cuda.synchronize(dev)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
