"""us per step of the three thread mappings of the step kernel: CUDA graph of single-step launches, and the persistent
T-step launch (fa_step_many), at small batches."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fortattack_b200 as fab
dev = torch.device("cuda:0")

def timed(fn, reps=5):
    best = None
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); torch.cuda._sleep(400000); e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1); best = ms if best is None else min(best, ms)
    return best

for ng, na in ((3, 3), (5, 5)):
    A = ng + na
    for E in (4096, 16384, 65536):
        for mapping in ("group", "agent", "env"):
            env = fab.FortAttackBatch(E, ng, na, max_steps=100, seed=0, device=dev, mapping=mapping)
            env.reset()
            row = {"teams": "%dv%d" % (ng, na), "envs": E, "mapping": mapping, "info": env.kernel_info()}
            n = 40
            acts = torch.randint(0, 8, (n, A, E), device=dev, dtype=torch.int32)
            o = [(torch.empty(A, E, 6, device=dev), torch.empty(A, E, device=dev), torch.empty(E, dtype=torch.uint8, device=dev),
                  torch.empty(E, dtype=torch.uint8, device=dev)) for _ in range(8)]
            for t in range(3):
                env.step(acts[t], out=o[t % 8])
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph(); side = torch.cuda.Stream(dev); side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                with torch.cuda.graph(g, stream=side):
                    for t in range(n):
                        env.step(acts[t], out=o[t % 8])
            torch.cuda.current_stream().wait_stream(side)
            g.replay()
            row["graph_single_step_us"] = round(1e3 * timed(g.replay) / n, 3)
            for T in (20, 200, 1000):
                if T * A * E * 28 > 6e9:
                    continue
                a = torch.randint(0, 8, (T, A, E), device=dev, dtype=torch.int32)
                out = (torch.empty(T, A, E, 6, device=dev), torch.empty(T, A, E, device=dev),
                       torch.empty(T, E, dtype=torch.uint8, device=dev), torch.empty(T, E, dtype=torch.uint8, device=dev))
                env.step_many(a, out=out)
                row["persistent_T%d_us" % T] = round(1e3 * timed(lambda: env.step_many(a, out=out)) / T, 3)
                del a, out
            print(json.dumps(row), flush=True)
            del env, g
