"""Persistent / single-step time of the group mapping at 3v3 x 4096 for 1, 2 and 4 warps per block (FA_GROUP_BLOCK_WARPS)."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fortattack_b200 as fab
dev = torch.device("cuda:0")
def timed(fn, reps=7):
    best = None
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); torch.cuda._sleep(400000); e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1); best = ms if best is None else min(best, ms)
    return best
for ng, na, E in ((3, 3, 4096), (5, 5, 4096), (3, 3, 8192)):
    for bw in (1, 2, 4):
        os.environ["FA_GROUP_BLOCK_WARPS"] = str(bw)
        env = fab.FortAttackBatch(E, ng, na, max_steps=100, seed=0, device=dev, mapping="group")
        env.reset()
        row = {"teams": "%dv%d" % (ng, na), "envs": E, "block_warps": bw, "info": env.kernel_info()}
        for T in (20, 1000):
            a = torch.randint(0, 8, (T, ng + na, E), device=dev, dtype=torch.int32)
            out = (torch.empty(T, ng + na, E, 6, device=dev), torch.empty(T, ng + na, E, device=dev),
                   torch.empty(T, E, dtype=torch.uint8, device=dev), torch.empty(T, E, dtype=torch.uint8, device=dev))
            env.step_many(a, out=out)
            row["persistent_T%d_us" % T] = round(1e3 * timed(lambda: env.step_many(a, out=out)) / T, 3)
            del a, out
        print(json.dumps(row), flush=True)
        del env
