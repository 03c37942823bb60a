import ctypes
# needs the diagnostic build: make -C emergent-multiagent-strategies_b200/csrc trace; run with
#   FORTATTACK_B200_LIB=emergent-multiagent-strategies_b200/libfortattack_b200_trace.so python profiles/policy_phase_trace.py
import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from importlib import import_module
import policy_util as pu
from test_policy_cpu import make
pk = import_module("emergent-multiagent-strategies_b200.policy_kernel")
gen = torch.Generator().manual_seed(1)
# correctness over shapes
for n, m, E in ((3, 3, 42), (3, 3, 4096), (5, 5, 1000), (1, 1, 300), (2, 4, 777), (4, 1, 129), (5, 3, 26)):
    net = make(n, m, seed=n * 10 + m).cuda()
    fp = pk.FusedPolicy(net, seed=3)
    own, opp = pu.random_obs(n, E, gen, "cuda"), pu.random_obs(m, E, gen, "cuda")
    o = fp.forward(own, opp, pk.MODE_ARGMAX, want_logits=True)
    torch.cuda.synchronize()
    lg, v = o["logits"].double().cpu(), o["value"].double().cpu()
    lgq, vq = pu.emulate(fp.blob, own, opp, quantize=True)
    lgr, vr = pu.module_forward(net, own, opp)
    print(n, m, E, "status", int(fp.status.item()), "finite", bool(torch.isfinite(lg).all()), "vs quant-emul: logits %.3e value %.3e | vs fp32 module: logits %.3e value %.3e | scale %.2f %.2f" % (
        float((lg-lgq).abs().max()), float((v-vq).abs().max()), float((lg-lgr.double().cpu()).abs().max()), float((v-vr.double().cpu()).abs().max()),
        float(lgq.abs().max()), float(vq.abs().max())), flush=True)
n = m = 3
net = make(n, m, seed=33).cuda()
fp = pk.FusedPolicy(net, seed=3)
print(fp.kernel_info())
L = fp._lib
head = ["start", "enc+arrive", "wait(T')", "scores', wait z', drain, combine", "opp msg+arrive"]
rnd = ["wait(T)", "scores", "wait(ZY)+drain+combine+softmax", "update+arrive"]
labels = head + rnd * 3 + ["wait(heads)", "heads math"]
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
for EE in (42, 4096, 16384, 65536, 262144):
    own, opp = pu.random_obs(n, EE, gen, "cuda"), pu.random_obs(m, EE, gen, "cuda")
    out = fp.forward(own, opp, pk.MODE_SAMPLE)
    for _ in range(3): fp.forward(own, opp, pk.MODE_SAMPLE, out=out)
    torch.cuda.synchronize()
    t0.record()
    for _ in range(20): fp.forward(own, opp, pk.MODE_SAMPLE, out=out)
    t1.record(); torch.cuda.synchronize()
    us = t0.elapsed_time(t1) / 20 * 1e3
    print("E=%d: %.1f us per team forward, %.3e agent-forwards/s, %.1f TFLOP/s (0.7 MFLOP/row)  status %d" % (EE, us, n*EE/us*1e6, n*EE*0.7e6/us*1e6/1e12, int(fp.status.item())), flush=True)
    if EE in (42, 16384):
        tr = torch.zeros(96, dtype=torch.int64, device="cuda")
        L.mp_set_trace.argtypes = [ctypes.c_void_p]; L.mp_set_trace(tr.data_ptr())
        fp.forward(own, opp, pk.MODE_SAMPLE, out=out)
        torch.cuda.synchronize()
        L.mp_set_trace(None)
        nz = [x for x in tr.cpu().tolist() if x]
        print("  trace: %d stamps, total %d cycles" % (len(nz), nz[-1] - nz[0]))
        print("   " + "  ".join("%s=%d" % (labels[i] if i < len(labels) else "?", nz[i] - nz[i - 1]) for i in range(1, len(nz))))
