"""us and effective HBM GB/s of tg_linear / tg_wgrad for every product shape of one PPO minibatch step (config 3: 3v3, hidden
128, 196 608 rows per team minibatch), against the bytes each product has to move."""
import os, sys, json
from importlib import import_module
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
fused = import_module("emergent-multiagent-strategies_b200.rlcore.fused")
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 196608
dev = "cuda:0"
if os.environ.get("TG_BENCH_STAGED") == "0":       # A/B: keep tg_linear on the register loaders
    import_module("emergent-multiagent-strategies_b200._capi").lib().tg_debug_staged(0)
if os.environ.get("TG_BENCH_TMA_OUT") == "0":      # A/B: keep tg_linear's epilogue on STG stores
    import_module("emergent-multiagent-strategies_b200._capi").lib().tg_debug_tma_out(0)
if os.environ.get("TG_BENCH_WGRAD_STAGED") == "0":   # A/B: keep tg_wgrad on the register loaders
    import_module("emergent-multiagent-strategies_b200._capi").lib().tg_debug_wgrad_staged(0)
if os.environ.get("TG_BENCH_WGRAD_ROWS"):          # A/B: force tg_wgrad's block height (32 / 64)
    import_module("emergent-multiagent-strategies_b200._capi").lib().tg_debug_wgrad_rows(int(os.environ["TG_BENCH_WGRAD_ROWS"]))

def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(400000); e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / n

flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
print("rows", rows)
for K, N, relu, acc in ((6, 64, True, False), (64, 64, False, False), (64, 128, False, False), (128, 128, False, False), (256, 128, True, False),
                        (128, 256, False, False), (128, 128, False, True), (128, 8, False, False), (8, 128, False, False), (128, 1, False, False),
                        (1, 128, False, False), (128, 64, False, False)):
    # several independent input/output sets larger than the L2 in total, used round-robin: every launch reads from HBM
    sets = max(2, int(400e6 / (rows * (K + N) * 4)) + 1)
    xs = [torch.randn(rows, K, device=dev) for _ in range(sets)]
    outs = [torch.empty(rows, N, device=dev) for _ in range(sets)]
    pack = fused.tg_pack(torch.randn(N, K, device=dev), False)
    bias = torch.randn(N, device=dev)
    it = [0]
    def run():
        i = it[0] % sets; it[0] += 1
        fused.tg_linear(xs[i], pack, bias, relu, out=outs[i], accumulate=acc)
    us = timed(run)
    b = rows * 4 * (K + N * (2 if acc else 1))
    print(json.dumps({"op": "tg_linear", "K": K, "N": N, "relu": relu, "acc": acc, "us": round(us, 1), "gbs": round(b / us / 1e3, 1),
                      "mb": round(b / 1e6, 1)}), flush=True)
    del xs, outs
for a, b_ in ((64, 6), (64, 64), (64, 128), (128, 128), (256, 128), (128, 256), (128, 8), (128, 1), (128, 64)):
    sets = max(2, int(400e6 / (rows * (a + b_) * 4)) + 1)
    xs = [torch.randn(rows, a, device=dev) for _ in range(sets)]
    ys = [torch.randn(rows, b_, device=dev) for _ in range(sets)]
    it = [0]
    def run():
        i = it[0] % sets; it[0] += 1
        fused.tg_wgrad(xs[i], ys[i])
    us = timed(run)
    by = rows * 4 * (a + b_)
    print(json.dumps({"op": "tg_wgrad", "a": a, "b": b_, "us": round(us, 1), "gbs": round(by / us / 1e3, 1), "mb": round(by / 1e6, 1)}), flush=True)
    del xs, ys
fused.tg_check_status(dev)
