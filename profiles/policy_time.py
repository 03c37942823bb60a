"""us per fused team forward (mp_policy_kernel) at 3v3 / 5v5 for a few batch sizes; MP_OPT=0..3 switches the tile-overlap measures."""
import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from importlib import import_module
import policy_util as pu
from test_policy_cpu import make
pk = import_module("emergent-multiagent-strategies_b200.policy_kernel")
for n, E in ((3, 4096), (3, 16384), (3, 65536), (5, 8192)):
    net = make(n, n, seed=33).cuda()
    fp = pk.FusedPolicy(net, seed=3)
    gen = torch.Generator().manual_seed(1)
    own, opp = pu.random_obs(n, E, gen, "cuda"), pu.random_obs(n, E, gen, "cuda")
    out = fp.forward(own, opp)
    for _ in range(5):
        fp.forward(own, opp, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        fp.forward(own, opp, out=out)
    e1.record(); torch.cuda.synchronize(); fp.check_status()
    print("MP_OPT=%s %dv%d E=%d: %.2f us per team forward" % (os.environ.get("MP_OPT", "3"), n, n, E, 1e3 * e0.elapsed_time(e1) / 50), flush=True)
