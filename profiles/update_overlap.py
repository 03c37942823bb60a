"""JointPPO update, graph-replayed, with the two teams' updates on two streams (BatchedTrainer(overlap_teams=True), the default
on one rank) against one team after the other.  Config 3 (3v3 x 16 384 envs) and one GPU's share of config 4 (5v5 x 8192), T=128.
  python profiles/update_overlap.py [--small]"""
import json
import os
import sys
from importlib import import_module

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ro = import_module("emergent-multiagent-strategies_b200.rollout")
fused = import_module("emergent-multiagent-strategies_b200.rlcore.fused")
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)


def models(ckpt=2520):
    path = os.path.join(ROOT, "baseline", "_ref", "reference", "marlsave", "tmp_1", "ep%d.pt" % ckpt)
    return torch.load(path, map_location="cpu")["models"] if os.path.exists(path) else None


def timed(fn, reps=3):
    best = None
    for _ in range(reps):
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    return best


def run(E, ng, na, T=128):
    torch.manual_seed(0)
    tr = ro.BatchedTrainer(E, ng, na, num_steps=T, device=dev, seed=0, graph_update=True)
    m = models()
    if m is not None and ng == 5:
        tr.load_models(m)
    elif m is not None:
        tr.load_models([m[0]] * ng + [m[-1]] * na)
    tr.collect(); tr.recompute_old(); tr.wrap_horizon()
    out = {"envs": E, "teams": "%dv%d" % (ng, na), "T": T}
    for name, ov in (("overlapped", True), ("sequential", False), ("overlapped_again", True)):
        tr.overlap_teams = ov
        tr.update()                          # (first call: eager steps + capture)
        out[name + "_ms"] = timed(tr.update)
    fused.tg_check_status(dev)
    return out


if __name__ == "__main__":
    small = "--small" in sys.argv
    res = [run(2048 if small else 16384, 3, 3, 32 if small else 128)]
    if not small:
        res.append(run(8192, 5, 5))
    print(json.dumps(res))
