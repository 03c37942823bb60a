"""PPO update time per GPU at fixed per-GPU work (weak scaling of BASELINE configs[3]'s update: 5v5, 8192 envs per GPU,
T=128, 4 epochs x 32 minibatches x 2 teams, one gradient all-reduce per optimizer step).  Run with and without torchrun:
  python profiles/update_scaling.py
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29561 profiles/update_scaling.py"""
import os
import sys
from importlib import import_module

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ro = import_module("emergent-multiagent-strategies_b200.rollout")
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
pg = None
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
    pg = dist.group.WORLD
E, T = 8192, 128
torch.manual_seed(0)
GRAPH = "--graph" in sys.argv            # JointPPO(graph_update=True): the optimizer step (incl. its all-reduces) replayed from a CUDA graph
tr = ro.BatchedTrainer(E, 5, 5, num_steps=T, device=dev, seed=0, env_id0=rank * E, process_group=pg, allow_tf32=True,
                       graph_update=GRAPH)
for _ in range(2):
    tr.collect(); tr.wrap_horizon(); tr.after_update()
tr.collect(); tr.wrap_horizon()
tr.update()
torch.cuda.synchronize(dev)
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
if world > 1:
    dist.barrier()
e[0].record(); tr.collect(); tr.wrap_horizon(); e[1].record(); vals = tr.update(); e[2].record()
torch.cuda.synchronize(dev)
t = torch.tensor([e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
spread = 0.0
if world > 1:                                # replicas must hold identical weights after the synchronous update
    flat = torch.cat([p.detach().reshape(-1) for pol in tr.policies for p in pol.parameters()])
    lo, hi = flat.clone(), flat.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    spread = float((hi - lo).abs().max())
if rank == 0:
    print("update_scaling: world %d%s, 5v5 x %d envs per GPU, T=%d: rollout+GAE %.1f ms, PPO update (TF32) %.0f ms -> %.3e agent-steps/s trained per GPU-set; losses %s; replica weight spread %.1e"
          % (world, " (graph-replayed optimizer step)" if GRAPH else "", E, T, t[0].item(), t[1].item(),
             world * E * 10 * T / ((t[0].item() + t[1].item()) * 1e-3), [tuple(round(x, 4) for x in v) for v in vals], spread), flush=True)
if world > 1:
    # captured graphs hold NCCL kernels of this process group: release them before the group goes away (the r1j 2-GPU run
    # printed its result and then sat in teardown until its timeout without this; not re-run since -- GPU budget spent)
    for trn in tr.trainers:
        trn._g = None
    import gc
    gc.collect()
    torch.cuda.synchronize(dev)
    dist.destroy_process_group()
