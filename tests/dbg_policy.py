import sys, os, subprocess, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
if len(sys.argv) == 1:
    for v in ("0", "1", "2"):
        env = dict(os.environ, MP_VARIANT=v)
        print("==== MP_VARIANT", v, flush=True)
        subprocess.call([sys.executable, __file__, "run"], env=env)
    sys.exit(0)
from importlib import import_module
import policy_util as pu
from test_policy_cpu import make
pk = import_module("emergent-multiagent-strategies_b200.policy_kernel")
n = m = 3
net = make(n, m, seed=33).cuda()
fp = pk.FusedPolicy(net, seed=3)
print(fp.kernel_info(), flush=True)
pr = torch.cuda.get_device_properties(0)
print("smem/SM", pr.shared_memory_per_multiprocessor, "regs/SM", pr.regs_per_multiprocessor, flush=True)
gen = torch.Generator().manual_seed(1)
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
for EE in (4096, 16384, 65536, 262144):
    own2, opp2 = pu.random_obs(n, EE, gen, "cuda"), pu.random_obs(m, EE, gen, "cuda")
    out = fp.forward(own2, opp2, pk.MODE_SAMPLE, want_logits=True)
    torch.cuda.synchronize()
    lgq, vq = pu.emulate(fp.blob, own2[:, :500].contiguous(), opp2[:, :500].contiguous(), quantize=True)
    err = float((out["logits"][:, :500].double().cpu() - lgq).abs().max())
    t0.record()
    for _ in range(20): fp.forward(own2, opp2, pk.MODE_SAMPLE, out=out)
    t1.record(); torch.cuda.synchronize()
    us = t0.elapsed_time(t1) / 20 * 1e3
    print("E=%d: %.1f us per team forward, %.3e agent-forwards/s, %.1f TFLOP/s  status %d err %.2e" % (EE, us, n*EE/us*1e6, n*EE*0.7e6/us*1e6/1e12, int(fp.status.item()), err), flush=True)
