"""GPU: the fused MPNN rollout forward (csrc/mp_policy.cu) against (a) a torch evaluation of the same
blob with fp16 rounding at the kernel's rounding points (tight) and (b) the fp32 module (the
reference's arithmetic, mpnn.py:117-205; loose: fp16 tensor-core operands), plus the sampling contract."""
from importlib import import_module

import pytest
import torch

import policy_util as pu
from test_policy_cpu import make

pytestmark = pytest.mark.gpu
PKG = "emergent-multiagent-strategies_b200"
pk = import_module(PKG + ".policy_kernel")
_capi = import_module(PKG + "._capi")


def test_tcgen05_gemm_probe():
    """One 128 x N x K tile through tcgen05.mma with the canonical no-swizzle K-major descriptors."""
    L = _capi.probe_lib()
    g = torch.Generator().manual_seed(0)
    for K, N in ((64, 64), (128, 64), (128, 256), (256, 128), (64, 16)):
        a = torch.randn(128, K, generator=g).half().cuda()
        w = (torch.randn(N, K, generator=g) / K ** 0.5).half().cuda()
        out = torch.full((128, N), float("nan"), device="cuda")
        err = torch.zeros(1, dtype=torch.int32, device="cuda")
        rc = L.mp_probe_gemm(a.data_ptr(), pk._canonical(w).data_ptr(), out.data_ptr(), K, N, 128, K * 16, 0, err.data_ptr(), None)
        torch.cuda.synchronize()
        assert rc == 0 and int(err.item()) == 0
        ref = a.double() @ w.double().t()
        assert (out.double() - ref).abs().max() < 1e-4, (K, N)


@pytest.mark.parametrize("n,m,E", [(3, 3, 4096), (3, 3, 42), (3, 3, 43), (3, 3, 5), (3, 3, 1), (5, 5, 1000), (1, 1, 300), (2, 4, 777),
                                   (4, 1, 129), (5, 3, 26), (3, 3, 16384)])
def test_forward_matches_blob_arithmetic(n, m, E):
    net = make(n, m, seed=n * 10 + m).cuda()
    fp = pk.FusedPolicy(net, seed=3)
    gen = torch.Generator().manual_seed(E)
    own, opp = pu.random_obs(n, E, gen, "cuda"), pu.random_obs(m, E, gen, "cuda")
    o = fp.forward(own, opp, pk.MODE_ARGMAX, want_logits=True, want_entropy=True)
    fp.check_status()
    lg, v = o["logits"].double().cpu(), o["value"].double().cpu()
    assert torch.isfinite(lg).all() and torch.isfinite(v).all()
    lgq, vq = pu.emulate(fp.blob, own, opp, quantize=True)
    s = max(1.0, float(lgq.abs().max()))
    # same roundings, different fp32 summation order; an fp16 rounding flip costs up to 1 ulp(fp16) of one activation
    assert (lg - lgq).abs().max() < 4e-3 * s, float((lg - lgq).abs().max())
    assert (v - vq).abs().max() < 4e-3 * max(1.0, float(vq.abs().max())), float((v - vq).abs().max())
    lgr, vr = pu.module_forward(net, own, opp)
    assert (lg - lgr.double().cpu()).abs().max() < 3e-2 * s
    assert (v - vr.double().cpu()).abs().max() < 3e-2 * max(1.0, float(vr.abs().max()))
    # log-prob / entropy / argmax are consistent with the kernel's own logits
    lsm = torch.log_softmax(lg, dim=-1)
    act = o["action"].cpu()
    top = lg.max(dim=-1).values
    assert torch.equal(lg.gather(-1, act.unsqueeze(-1)).squeeze(-1), top)
    assert (o["logp"].double().cpu() - lsm.gather(-1, act.unsqueeze(-1)).squeeze(-1)).abs().max() < 1e-5
    assert (o["entropy"].double().cpu() + (lsm.exp() * lsm).sum(-1)).abs().max() < 1e-5
    assert torch.equal(o["action_i32"].cpu().long(), act)


# Tolerances of the fp16-operand rollout kernel with TRAINED weights (logits O(10), values O(20)).  They are
# what the kernel's arithmetic gives, stated as gates: 2.5e-2 on log-probs / logits, 6e-2 on values, against both the
# reference's own outputs (golden rows) and the fp32 module on realistic observations.
TRAINED_TOL_LOGP, TRAINED_TOL_VALUE = 2.5e-2, 6e-2


def test_trained_checkpoint_against_reference_outputs():
    """Kernel with the shipped ep2520 guard policy vs what the UNCHANGED reference computed for the same rows
    (tests/golden/rl_mpnn.npz ckpt/*: MPNN.evaluate_actions of /root/reference on 6 envs x 5 agents)."""
    sd, d = pu.trained_state_dict()
    net = pu.trained_net(5, 5, "cuda")
    fp = pk.FusedPolicy(net, seed=0)
    own = torch.from_numpy(d["ckpt/own"]).cuda().view(5, 6, 6).contiguous()      # agent-major rows [n*B, 6]
    opp = torch.from_numpy(d["ckpt/opp"]).cuda().view(5, 6, 6).contiguous()
    act = torch.from_numpy(d["ckpt/act"]).cuda().view(5, 6)
    o = fp.forward(own, opp, pk.MODE_EVAL, action_in=act, want_entropy=True)
    fp.check_status()
    ref_v = torch.from_numpy(d["ckpt/value"]).view(5, 6)
    ref_lp = torch.from_numpy(d["ckpt/logp"]).view(5, 6)
    ref_en = torch.from_numpy(d["ckpt/entropy"]).view(5, 6)
    ev, elp, een = [float((a.cpu() - b).abs().max()) for a, b in ((o["value"], ref_v), (o["logp"], ref_lp), (o["entropy"], ref_en))]
    print("trained ep2520, kernel vs reference rows: |dvalue| %.2e (scale %.1f)  |dlogp| %.2e  |dentropy| %.2e"
          % (ev, float(ref_v.abs().max()), elp, een))
    assert elp < TRAINED_TOL_LOGP and een < TRAINED_TOL_LOGP and ev < TRAINED_TOL_VALUE * max(1.0, float(ref_v.abs().max()) / 10)
    # the fp32 module itself reproduces the reference rows (state_dict + arithmetic parity, tight)
    with torch.no_grad():
        v, lp, en, _ = net.evaluate_actions(own.view(-1, 6), None, opp.view(-1, 6), None, act.view(-1, 1))
    assert (v.cpu().view(5, 6) - ref_v).abs().max() < 2e-4 and (lp.cpu().view(5, 6) - ref_lp).abs().max() < 2e-4


@pytest.mark.parametrize("n,m", [(5, 5), (3, 3)])
def test_trained_checkpoint_on_rollout_observations(n, m):
    """Same policy on observations of real rollouts driven by it (E = 2048, 60 steps in): kernel vs fp32 module."""
    fab = import_module("fortattack_b200")
    net = pu.trained_net(n, m, "cuda")
    onet = pu.trained_net(m, n, "cuda")
    fp, fo = pk.FusedPolicy(net, seed=1), pk.FusedPolicy(onet, seed=2)
    E = 2048
    env = fab.FortAttackBatch(E, n, m, max_steps=100, seed=5, device="cuda:0")
    obs = env.reset()
    acts = torch.empty(n + m, E, dtype=torch.int32, device="cuda")
    for t in range(60):
        a = fp.forward(obs[:n].contiguous(), obs[n:].contiguous(), pk.MODE_SAMPLE)["action_i32"]
        b = fo.forward(obs[n:].contiguous(), obs[:n].contiguous(), pk.MODE_SAMPLE)["action_i32"]
        acts[:n], acts[n:] = a, b
        obs = env.step(acts, auto_reset=True)[0]
    own, opp = obs[:n].contiguous(), obs[n:].contiguous()
    o = fp.forward(own, opp, pk.MODE_SAMPLE, want_logits=True)
    fp.check_status()
    lgr, vr = pu.module_forward(net, own, opp)
    lp_ref = torch.log_softmax(lgr, -1).gather(-1, o["action"].unsqueeze(-1)).squeeze(-1)
    e_lg, e_lp, e_v = [float(x) for x in ((o["logits"] - lgr).abs().max(), (o["logp"] - lp_ref).abs().max(), (o["value"] - vr).abs().max())]
    m_lp = float((o["logp"] - lp_ref).abs().mean())
    print("trained ep2520 %dv%d on rollout obs: logits scale %.1f  |dlogits| %.2e  |dlogp| max %.2e mean %.2e  |dvalue| %.2e (scale %.1f)"
          % (n, m, float(lgr.abs().max()), e_lg, e_lp, m_lp, e_v, float(vr.abs().max())))
    assert e_lp < TRAINED_TOL_LOGP and m_lp < 3e-3 and e_lg < 2 * TRAINED_TOL_LOGP
    assert e_v < TRAINED_TOL_VALUE * max(1.0, float(vr.abs().max()) / 10)


def test_sampling_distribution_and_determinism():
    n = m = 3
    E = 60000
    net = make(n, m, seed=1).cuda()
    fp = pk.FusedPolicy(net, seed=11)
    gen = torch.Generator().manual_seed(0)
    own1, opp1 = pu.random_obs(n, 1, gen, "cuda"), pu.random_obs(m, 1, gen, "cuda")
    own, opp = own1.expand(n, E, 6).contiguous(), opp1.expand(m, E, 6).contiguous()   # every env identical
    o = fp.forward(own, opp, pk.MODE_SAMPLE, want_logits=True)
    fp.check_status()
    probs = torch.softmax(o["logits"][:, 0].double(), dim=-1).cpu()                  # [n, 8]
    for a in range(n):
        freq = torch.bincount(o["action"][a].cpu(), minlength=8).double() / E
        assert (freq - probs[a]).abs().max() < 5 * (0.25 / E) ** 0.5 + 1e-3, (freq, probs[a])
    lsm = torch.log_softmax(o["logits"].double(), dim=-1)
    assert (o["logp"].double() - lsm.gather(-1, o["action"].unsqueeze(-1)).squeeze(-1)).abs().max() < 1e-5
    # same (seed, call counter) -> same draw; next call -> a different one
    fp2 = pk.FusedPolicy(net, seed=11)
    o2 = fp2.forward(own, opp, pk.MODE_SAMPLE)
    assert torch.equal(o2["action"], o["action"])
    o3 = fp2.forward(own, opp, pk.MODE_SAMPLE)
    assert not torch.equal(o3["action"], o["action"])
    # shard invariance: envs [E/2, E) drawn by a second shard with env_id0 = E/2
    fp4 = pk.FusedPolicy(net, seed=11, env_id0=E // 2)
    o4 = fp4.forward(own[:, E // 2:].contiguous(), opp[:, E // 2:].contiguous(), pk.MODE_SAMPLE)
    assert torch.equal(o4["action"], o["action"][:, E // 2:])
    # evaluate mode reproduces the log-probs of given actions
    o5 = fp.forward(own, opp, pk.MODE_EVAL, action_in=o["action"], want_entropy=True)
    assert (o5["logp"] - o["logp"]).abs().max() < 1e-6


def test_ensemble_launch_with_empty_and_tiny_lists():
    """mp_forward_ensemble with more checkpoints than environments (some lists empty, every list shorter than a tile)
    and with one environment: finishes (no pipeline timeout) and equals the single-checkpoint forwards."""
    nets = [make(2, 3, seed=50 + k).cuda() for k in range(8)]
    fps = [pk.FusedPolicy(n_, seed=4) for n_ in nets]
    gen = torch.Generator().manual_seed(9)
    for E in (1, 10, 300):
        own, opp = pu.random_obs(2, E, gen, "cuda"), pu.random_obs(3, E, gen, "cuda")
        ids = torch.randint(0, 8, (E,), generator=gen).to("cuda")
        ids[0] = 5
        order = torch.argsort(ids, stable=True).to(torch.int32)
        offsets = torch.zeros(9, dtype=torch.int32, device="cuda")
        offsets[1:] = torch.cumsum(torch.bincount(ids, minlength=8), 0).to(torch.int32)
        one = pk.forward_ensemble(fps, own, opp, order, offsets, pk.MODE_ARGMAX)
        fps[0].check_status()
        for k in range(8):
            ref = fps[k].forward(own, opp, pk.MODE_ARGMAX)
            sel = ids == k
            for key in ("value", "action", "logp"):
                assert torch.equal(one[key][:, sel], ref[key][:, sel]), (E, k, key)


def test_act_surface_matches_module_statistics():
    """FusedPolicy.act has MPNN.act's signature and shapes (mpnn.py:180-188)."""
    net = make(3, 3, seed=2).cuda()
    fp = pk.FusedPolicy(net)
    gen = torch.Generator().manual_seed(5)
    E = 512
    own, opp = pu.random_obs(3, E, gen, "cuda"), pu.random_obs(3, E, gen, "cuda")
    value, action, logp, state = fp.act(own.view(-1, 6), None, opp.view(-1, 6), None, deterministic=True)
    v_ref, a_ref, lp_ref, _ = net.act(own.view(-1, 6), None, opp.view(-1, 6), None, deterministic=True)
    assert value.shape == v_ref.shape and action.shape == a_ref.shape and logp.shape == lp_ref.shape and action.dtype == a_ref.dtype
    assert (value - v_ref).abs().max() < 3e-2 * max(1.0, float(v_ref.abs().max()))
    assert (action == a_ref).float().mean() > 0.97          # arg-max flips only where two logits nearly tie
    assert (fp.get_value(own.view(-1, 6), None, opp.view(-1, 6), None) - value).abs().max() == 0


def test_forward_rejects_buffers_the_kernel_would_misread():
    """mp_forward takes raw pointers: wrong element types, sizes or devices must be refused before the launch."""
    n, m, E = 3, 3, 64
    net = make(n, m, seed=5).cuda()
    fp = pk.FusedPolicy(net, seed=1)
    gen = torch.Generator().manual_seed(9)
    own, opp = pu.random_obs(n, E, gen, "cuda"), pu.random_obs(m, E, gen, "cuda")
    order = torch.arange(E, dtype=torch.int32, device="cuda")
    offsets = torch.tensor([0, E], dtype=torch.int32, device="cuda")
    bad = [dict(own=own.cpu()), dict(out={"value": torch.empty(n, E, dtype=torch.float64, device="cuda")}),
           dict(out={"action": torch.empty(n, E, dtype=torch.int32, device="cuda")}),
           dict(out={"logp": torch.empty(n, E)}), dict(out={"values": torch.empty(n, E, device="cuda")}),
           dict(mode=pk.MODE_EVAL), dict(mode=pk.MODE_EVAL, action_in=torch.zeros(n * E - 1, dtype=torch.int64)),
           dict(env_order=order), dict(env_order=order.long(), env_offsets=offsets),
           dict(env_order=order[:-1].contiguous(), env_offsets=offsets), dict(env_order=order, env_offsets=offsets, sel_value=1),
           dict(env_sel=torch.zeros(E, dtype=torch.int32))]
    for kw in bad:
        args = dict(own=own, opp=opp)
        args.update(kw)
        with pytest.raises(ValueError):
            fp.forward(**args)
    with pytest.raises(ValueError):
        pk.forward_ensemble([fp], own, opp, order, offsets[:1].contiguous())
    with pytest.raises(ValueError):
        pk.forward_ensemble([fp], own, opp, order.cpu(), offsets)
    ok = fp.forward(own, opp, pk.MODE_ARGMAX, env_order=order, env_offsets=offsets, sel_value=0)
    assert torch.isfinite(ok["value"]).all()
    fp.check_status()
