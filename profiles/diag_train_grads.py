"""Per-parameter gradient error of MPNN.evaluate_actions: bmm mirror vs the fused path on cuBLAS vs the fused path on the
tcgen05 kernels (diagnostic for tests/test_rollout_gpu.py::test_fused_training_attention_matches_bmm_path)."""
import os, sys
from importlib import import_module
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PKG = "emergent-multiagent-strategies_b200"
ro, mp, fused = import_module(PKG + ".rollout"), import_module(PKG + ".mpnn"), import_module(PKG + ".rlcore.fused")
SEED = int(sys.argv[1]) if len(sys.argv) > 1 else 0
torch.manual_seed(SEED)
n = m = 3
net = mp.MPNN(action_space=ro._Shape(8), num_agents=n, num_opp_agents=m, input_size=6, hidden_dim=128).cuda()
with torch.no_grad():
    for p in net.parameters():
        if p.dim() == 1:
            p.uniform_(-0.3, 0.3)
for B in (777,):
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

# how different are the encoder pre-activations (ReLU masks) between the two dense back ends?
with torch.no_grad():
    W, b = net.encoder[0].weight, net.encoder[0].bias
    pre_ref = torch.addmm(b, own, W.t())
    pre_tg = fused.tg_linear(own, fused.tg_pack(W, False), b, relu=False)
    print("encoder pre-activation: max |d| %.2e, ReLU-mask disagreements %d of %d, min |pre| %.2e"
          % (float((pre_ref - pre_tg).abs().max()), int(((pre_ref > 0) != (pre_tg > 0)).sum()), pre_ref.numel(), float(pre_ref.abs().min())))
    d64 = own.double() @ W.double().t() + b.double()
    print("   vs float64: cublas %.2e  tg %.2e" % (float((pre_ref.double() - d64).abs().max()), float((pre_tg.double() - d64).abs().max())))
    # the weight-gradient product of that layer on a random upstream gradient
    g = torch.randn(own.shape[0], 64, device="cuda")
    ref = g.double().t() @ own.double()
    got = fused.tg_wgrad(g, own)
    print("encoder wgrad on random g: err/max %.2e" % float((got.double() - ref).abs().max() / ref.abs().max()))

# the input-gradient product of dist.linear (K = 8) on the REAL upstream gradient rows
dist_mod = import_module(PKG + ".rlcore.distributions")
with torch.no_grad():
    net.fused_attention, fused.DENSE = True, "cublas"
    x = net._fwd(own, opp, None)
    p1 = torch.relu(torch.addmm(net.policy_head[0].bias, x, net.policy_head[0].weight.t()))
    logits0 = torch.addmm(net.dist.linear.bias, p1, net.dist.linear.weight.t())
    fused.DENSE = "tcgen05"
logits = logits0.clone().requires_grad_()
d = dist_mod.FixedCategorical(logits=logits)
((d.log_probs(act) * w).sum() * 0.7 + d.entropy().sum() * 0.3).backward()
dl = logits.grad.contiguous()
Wd = net.dist.linear.weight.detach()
ref = dl.double() @ Wd.double()
bound = dl.double().abs() @ Wd.double().abs()
got = fused.tg_linear(dl, fused.tg_pack(Wd, True))
err = (got.double() - ref).abs() / (bound + 1e-30)
r = int(err.max(dim=1).values.argmax())
print("dist.linear dx on real dlogits: rows %d, max err/bound %.3e at row %d; rows with err/bound > 1e-4: %d"
      % (dl.shape[0], float(err.max()), r, int((err.max(dim=1).values > 1e-4).sum())))
print("   worst row dlogits:", dl[r].tolist())
print("   its amax %.6e, bits %s" % (float(dl[r].abs().max()), hex(dl[r].abs().max().view(torch.int32).item())))
bad = (err.max(dim=1).values > 1e-4).nonzero().flatten()[:8].tolist()
for rr in bad:
    print("   bad row %d: amax %.6e  got[:4] %s  ref[:4] %s" % (rr, float(dl[rr].abs().max()), got[rr, :4].tolist(), ref[rr, :4].tolist()))
got2 = fused.tg_linear(torch.randn_like(dl), fused.tg_pack(Wd, True))
print("   same product on randn rows: ok" )

# in-situ capture: policy hidden layer p1 and the gradient arriving at it, per dense back end
cap = {}
orig_policy = net._policy
def run(mode):
    fused.DENSE = mode
    net.fused_attention, net.fold_projections = True, False
    net.zero_grad()
    store = {}
    def pol(x):
        out = orig_policy(x)
        store["p1"] = out.detach().clone()
        out.register_hook(lambda g: store.__setitem__("dp1", g.detach().clone()))
        store["h3"] = x.detach().clone()
        x.register_hook(lambda g: store.__setitem__("dh3", g.detach().clone()))
        return out
    net._policy = pol
    v, lp, ent, _ = net.evaluate_actions(own, None, opp, None, act)
    ((v * w).sum() + (lp * w).sum() * 0.7 + ent.sum() * 0.3).backward()
    net._policy = orig_policy
    store["gW"] = net.policy_head[0].weight.grad.clone(); store["gb"] = net.policy_head[0].bias.grad.clone()
    return store
A, Bt = run("cublas"), run("tcgen05")
fused.DENSE = "tcgen05"
for k in ("p1", "dp1", "h3", "dh3", "gW", "gb"):
    print("in situ %-4s max |d| %.3e  scale %.3e" % (k, float((A[k] - Bt[k]).abs().max()), float(A[k].abs().max())))
print("   ReLU mask disagreements in p1: %d; rows of dp1 differing > 1e-5: %d"
      % (int(((A["p1"] > 0) != (Bt["p1"] > 0)).sum()), int(((A["dp1"] - Bt["dp1"]).abs().max(dim=1).values > 1e-5).sum())))
badrows = ((A["dp1"] - Bt["dp1"]).abs().max(dim=1).values > 1e-5).nonzero().flatten()
print("   bad dp1 rows:", badrows[:20].tolist(), "... of", badrows.numel())
if badrows.numel():
    r = int(badrows[0]); print("   row", r, "cublas", A["dp1"][r, :6].tolist(), "tg", Bt["dp1"][r, :6].tolist())
