"""Per-parameter gradient error of MPNN.evaluate_actions: bmm mirror vs the fused path on cuBLAS vs the fused path on the
tcgen05 kernels (diagnostic for tests/test_rollout_gpu.py::test_fused_training_attention_matches_bmm_path)."""
import os, sys
from importlib import import_module
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PKG = "emergent-multiagent-strategies_b200"
ro, mp, fused = import_module(PKG + ".rollout"), import_module(PKG + ".mpnn"), import_module(PKG + ".rlcore.fused")
torch.manual_seed(0)
n = m = 3
net = mp.MPNN(action_space=ro._Shape(8), num_agents=n, num_opp_agents=m, input_size=6, hidden_dim=128).cuda()
with torch.no_grad():
    for p in net.parameters():
        if p.dim() == 1:
            p.uniform_(-0.3, 0.3)
for B in (777, 512):
    own, opp = torch.randn(n * B, 6, device="cuda"), torch.randn(m * B, 6, device="cuda")
    act = torch.randint(0, 8, (n * B, 1), device="cuda")
    w = torch.randn(n * B, 1, device="cuda")
    res = {}
    for name, fused_on, fold, dense in (("bmm", False, False, "cublas"), ("fused-nofold-cublas", True, False, "cublas"),
                                        ("fused-fold-cublas", True, True, "cublas"), ("fused-nofold-tg", True, False, "tcgen05"),
                                        ("fused-fold-tg", True, True, "tcgen05")):
        net.fused_attention, net.fold_projections, fused.DENSE = fused_on, fold, dense
        net.zero_grad()
        v, lp, ent, _ = net.evaluate_actions(own, None, opp, None, act)
        ((v * w).sum() + (lp * w).sum() * 0.7 + ent.sum() * 0.3).backward()
        res[name] = (v.detach().clone(), lp.detach().clone(), {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None})
    fused.DENSE = "tcgen05"
    fused.tg_check_status("cuda:0")
    a = res["bmm"]
    for name in list(res)[1:]:
        b = res[name]
        print("B=%d %-22s dv %.2e dlp %.2e" % (B, name, float((a[0] - b[0]).abs().max()), float((a[1] - b[1]).abs().max())))
        for k in a[2]:
            print("      %-28s err/scale %.2e" % (k, float((a[2][k] - b[2][k]).abs().max()) / (float(a[2][k].abs().max()) + 1e-9)))
