"""Profiling driver: a few fused policy forwards (3v3, E envs) and, with --ppo, one PPO minibatch step under torch.profiler."""
import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from importlib import import_module
import policy_util as pu
from test_policy_cpu import make
pk = import_module("emergent-multiagent-strategies_b200.policy_kernel")
E = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 16384
if "--ppo" in sys.argv:
    ro = import_module("emergent-multiagent-strategies_b200.rollout")
    if "--nofold" in sys.argv:       # message rounds with explicit Q/K/V/out projections (the pre-r1i training forward)
        import_module("emergent-multiagent-strategies_b200.mpnn").MPNN.fold_projections = False
    tr = ro.BatchedTrainer(E, 3, 3, num_steps=32, max_episode_steps=100, seed=0, ppo_epoch=1, num_mini_batch=8,
                           allow_tf32="--tf32" in sys.argv)
    tr.collect(); tr.wrap_horizon()
    tr.update()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); tr.update(); e1.record()
    torch.cuda.synchronize()
    print("update: %.2f ms for 16 optimizer steps (8 minibatches x 2 teams, %d rows per team minibatch) = %.2f ms per step"
          % (e0.elapsed_time(e1), 3 * E * 32 // 8, e0.elapsed_time(e1) / 16))
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        tr.update()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
    sys.exit(0)
net = make(3, 3, seed=33).cuda()
fp = pk.FusedPolicy(net, seed=3)
gen = torch.Generator().manual_seed(1)
own, opp = pu.random_obs(3, E, gen, "cuda"), pu.random_obs(3, E, gen, "cuda")
out = fp.forward(own, opp)
for _ in range(4):
    fp.forward(own, opp, out=out)
torch.cuda.synchronize()
fp.check_status()
print("ok")
