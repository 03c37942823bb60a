"""Debug driver for the dense training functions (rlcore/fused.py): each piece against torch, synchronising after each."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from importlib import import_module
fused = import_module("emergent-multiagent-strategies_b200.rlcore.fused")
what = sys.argv[1]
torch.manual_seed(0)
dev = "cuda"
def sync(tag):
    torch.cuda.synchronize(); print("ok", tag, flush=True)
if what == "colsum":
    for rows, cols in ((6144, 128), (6144, 64), (2331, 128), (5, 256), (196608, 128), (1001, 32)):
        out = torch.relu(torch.randn(rows, cols, device=dev)); dout = torch.randn(rows, cols, device=dev)
        dpre, db = fused.relu_bwd_colsum(dout, out)
        sync("colsum %d x %d" % (rows, cols))
        ref = dout * (out > 0)
        print("   max err dpre %.2e db %.2e" % (float((dpre - ref).abs().max()), float((db - ref.sum(0)).abs().max())), flush=True)
elif what == "xtdy":
    for rows, a, b in ((6144, 1, 128), (6144, 8, 128), (6144, 64, 6), (6144, 128, 128), (6144, 128, 64), (196608, 128, 128), (2331, 128, 1)):
        x = torch.randn(rows, a, device=dev); dy = torch.randn(rows, b, device=dev)
        got = fused.xt_dy(x, dy); sync("xt_dy %d: %d x %d, split %d" % (rows, a, b, fused._split(rows)))
        ref = (x.double().t() @ dy.double()).float()
        print("   max rel err %.2e" % float((got - ref).abs().max() / ref.abs().max()), flush=True)
    hm = torch.randn(6144, 256, device=dev); dg = torch.randn(6144, 128, device=dev)
    got = fused.xt_dy(hm[:, :128], dg); sync("xt_dy strided")
    print("   max rel err %.2e" % float((got - hm[:, :128].t() @ dg).abs().max()), flush=True)
elif what == "linear":
    for rows, K, N, relu in ((6144, 6, 64, True), (6144, 128, 128, True), (6144, 128, 1, False), (6144, 128, 8, False)):
        x = torch.randn(rows, K, device=dev, requires_grad=K != 6); W = torch.randn(N, K, device=dev, requires_grad=True)
        b = torch.randn(N, device=dev, requires_grad=True); w = torch.randn(rows, N, device=dev)
        y = fused.linear(x, W, b, relu); sync("linear fwd %d %d %d" % (rows, K, N))
        (y * w).sum().backward(); sync("linear bwd %d %d %d" % (rows, K, N))
        ref = torch.nn.functional.linear(x.detach(), W.detach(), b.detach()); ref = torch.relu(ref) if relu else ref
        print("   fwd err %.2e" % float((y - ref).abs().max()), flush=True)
elif what.startswith("update"):
    if "nocolsum" in what:
        def _rb(dout, out):
            dpre = torch.ops.aten.threshold_backward(dout.contiguous(), out, 0)
            return dpre, dpre.sum(0)
        fused.relu_bwd_colsum = _rb
    if "nosplit" in what:
        fused.xt_dy = lambda x, dy: x.t() @ dy
    ro = import_module("emergent-multiagent-strategies_b200.rollout")
    tr = ro.BatchedTrainer(256, 3, 3, num_steps=32, max_episode_steps=25, hidden_dim=128, ppo_epoch=1, num_mini_batch=4)
    tr.collect(); tr.wrap_horizon(); sync("collect")
    print(tr.update()); sync("update")
