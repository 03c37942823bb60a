import sys, os, ctypes, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from importlib import import_module
import policy_util as pu
from test_policy_cpu import make
pk = import_module("emergent-multiagent-strategies_b200.policy_kernel")
n = m = 3
net = make(n, m, seed=33).cuda()
fp = pk.FusedPolicy(net, seed=3)
L = fp._lib
L.mp_set_trace.argtypes = [ctypes.c_void_p]
gen = torch.Generator().manual_seed(1)
head = ["start", "enc+arrive", "wait(oppQKV)", "drainQ", "drainV", "bar", "dot+softmax", "mix+bar", "store+arrive", "wait(oout)", "drain eOpp", "arrive"]
rnd = ["wait(QK)", "drainK", "arrive", "bar", "dot+softmax", "bar", "wait(V)", "drainV", "bar", "mix+bar", "store+arrive", "wait(upd)", "drain h", "arrive"]
labels = head + rnd * 3 + ["wait(heads)", "heads math"]
for EE in (42, 16384):
    own, opp = pu.random_obs(n, EE, gen, "cuda"), pu.random_obs(m, EE, gen, "cuda")
    for _ in range(3): fp.forward(own, opp, pk.MODE_SAMPLE)
    tr = torch.zeros(96, dtype=torch.int64, device="cuda")
    L.mp_set_trace(tr.data_ptr())
    fp.forward(own, opp, pk.MODE_SAMPLE)
    torch.cuda.synchronize()
    L.mp_set_trace(None)
    t = tr.cpu().tolist()
    nz = [x for x in t if x]
    print("E=%d: %d stamps, total %d cycles" % (EE, len(nz), nz[-1] - nz[0]))
    for i in range(1, len(nz)):
        print("  %2d %-14s %7d" % (i, labels[i] if i < len(labels) else "?", nz[i] - nz[i - 1]))
