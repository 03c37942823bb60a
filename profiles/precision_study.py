"""CPU study for the fused training kernel of DESIGN section 10: which operand precision can the tensor-core GEMMs of the
TRAINING forward/backward use and still hold the gradient gate of the PPO update (2e-4 * scale per parameter,
tests/test_rollout_gpu.py)?

The folded training forward of rlcore/fused.py is restated in float64 with ONE change: every row-sized GEMM -- forward,
data gradient and weight gradient -- rounds its two operands to the format under study and accumulates exactly (the
tensor cores' fp32 accumulation is not the question here).  Everything else (attention softmax, ReLU, heads' log-softmax)
stays exact.  Weights: the shipped 5v5 checkpoint marlsave/tmp_1/ep2520.pt when the reference tree is present (this
container), else the module's own initialisation; observations: observation-like random rows.  Prints, per format, the
worst relative error of value / log-prob outputs and of every parameter gradient.

    python profiles/precision_study.py > profiles/r1j_training_precision_study.txt
"""
import math
import os
import sys
from importlib import import_module

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
mp = import_module("emergent-multiagent-strategies_b200.mpnn")
import policy_util as pu

torch.set_default_dtype(torch.float64)


def chop(x, bits):
    """Round to `bits` explicit mantissa bits (round to nearest even), exponent range unlimited."""
    m, e = torch.frexp(x)
    s = 2.0 ** (bits + 1)
    return torch.ldexp(torch.round(m * s) / s, e)


FORMATS = {
    "exact (float64)": lambda x: x,
    "fp32": lambda x: x.float().double(),
    "tf32 (10-bit mantissa)": lambda x: chop(x, 10),
    "fp16 (10-bit mantissa, 5-bit exponent)": lambda x: x.half().double(),
    "bf16 (7-bit mantissa)": lambda x: x.bfloat16().double(),
    "fp16 hi + lo split (2 MMAs per product side)": lambda x: x.half().double() + (x - x.half().double()).half().double(),
    "fp16 hi + lo split, operand scaled to max 2^12": lambda x: _scaled(x, lambda y: y.half().double() + (y - y.half().double()).half().double()),
    "tf32 hi + lo split (3xTF32)": lambda x: chop(x, 10) + chop(x - chop(x, 10), 10),
}


def _scaled(x, q):
    """Per-tensor power-of-two scaling before the split (what loss scaling does for fp16 gradients: without it the lo parts
    and small gradients fall into fp16's subnormal range)."""
    m = float(x.abs().max())
    if m == 0.0:
        return x
    sc = 2.0 ** (12 - math.ceil(math.log2(m)))
    return q(x * sc) / sc


def make_qmm(q):
    class QMM(torch.autograd.Function):
        @staticmethod
        def forward(ctx, a, b):
            ctx.save_for_backward(a, b)
            return q(a) @ q(b)

        @staticmethod
        def backward(ctx, g):
            a, b = ctx.saved_tensors
            return q(g) @ q(b).t(), q(a).t() @ q(g)
    return QMM.apply


def rows3(x, n):
    B = x.shape[0] // n
    return x.view(n, B, -1).transpose(0, 1)                      # agent-major rows -> [B, n, k]


def attention(a, b, v, n, m, norm, mask_diag):
    A3, B3, V3 = rows3(a, n), rows3(b, m), rows3(v, m)
    s = norm * A3 @ B3.transpose(1, 2)
    if mask_diag:
        s = s.masked_fill(torch.eye(n, dtype=torch.bool), -math.inf)
    return (torch.softmax(s, -1) @ V3).transpose(0, 1).reshape(a.shape[0], -1)


def forward(net, own, opp, act, mm, quantise_encoders):
    n, m, d = net.num_agents, net.num_opp_agents, net.h_dim
    lin = lambda x, l: mm(x, l.weight.t()) + l.bias
    oa, ms = net.oppAttn, net.messages
    # the K = 6 input encoders run on the CUDA cores in fp32 in mp_policy_kernel (the heading feature reaches 40 rad and
    # beyond: a 10-bit mantissa cannot hold it); cuBLAS with allow_tf32 rounds them like every other GEMM
    enc = lin if quantise_encoders else (lambda x, l: x @ l.weight.t() + l.bias)
    h0, hO = torch.relu(enc(own, net.encoder[0])), torch.relu(enc(opp, net.oppEncoder[0]))
    bv = mm(hO, torch.cat((oa.W_query[0], oa.W_val[0]), 1))
    e = attention(mm(h0, oa.W_key[0]), bv[:, :d // 2], bv[:, d // 2:], n, m, oa.norm_factor, False)
    h = torch.cat((h0, mm(e, oa.W_out[0])), 1)
    W = net.update[0].weight
    Mqk = ms.W_query[0] @ ms.W_key[0].t()
    Wc = torch.cat((W[:, :d].t(), ms.W_val[0] @ ms.W_out[0] @ W[:, d:].t()), 0)
    for _ in range(net.K):
        mixed = attention(mm(h, Mqk), h, h, n, n, ms.norm_factor, True)
        h = torch.relu(mm(torch.cat((h, mixed), 1), Wc) + net.update[0].bias)
    value = lin(torch.relu(lin(h, net.value_head[0])), net.value_head[2])
    logp = torch.log_softmax(lin(torch.relu(lin(h, net.policy_head[0])), net.dist.linear), -1)
    ent = -(logp.exp() * logp).sum(-1)
    return value[:, 0], logp.gather(1, act)[:, 0], ent


def main():
    n = m = 5
    B = 2048

    class Shape(object):
        shape = (8,)
    torch.manual_seed(0)
    net = mp.MPNN(action_space=Shape(), num_agents=n, num_opp_agents=m, input_size=6, hidden_dim=128).double()
    ck = "/root/reference/marlsave/tmp_1/ep2520.pt"
    src = "module initialisation"
    if os.path.exists(ck):
        net.load_state_dict({k: v.double() for k, v in torch.load(ck, map_location="cpu", weights_only=False)["models"][0].items()})
        src = "shipped checkpoint marlsave/tmp_1/ep2520.pt (guards)"
    gen = torch.Generator().manual_seed(1)
    own = pu.random_obs(n, B, gen).double().reshape(-1, 6)
    opp = pu.random_obs(m, B, gen).double().reshape(-1, 6)
    act = torch.randint(0, 8, (n * B, 1), generator=gen)
    wv, wl = torch.randn(n * B, generator=gen), torch.randn(n * B, generator=gen)
    params = [p for k, p in net.named_parameters() if not k.startswith("oppUpdate")]
    names = [k for k, _ in net.named_parameters() if not k.startswith("oppUpdate")]
    print("weights: %s; %dv%d, %d envs (%d rows per team), hidden 128; loss = sum(w_v v + w_l logp + 0.01 entropy) / rows, w ~ N(0, 1)"
          % (src, n, m, B, n * B))
    print("gate of the GPU parity test: every parameter gradient within 2e-4 * max|gradient|\n")
    ref = None
    runs = [(k, q, True) for k, q in FORMATS.items()]
    runs += [(k + ", encoders exact", q, False) for k, q in FORMATS.items() if k.startswith(("tf32 (", "fp16 (", "bf16"))]
    for name, q, qe in runs:
        v, lp, ent = forward(net, own, opp, act, make_qmm(q), qe)
        loss = ((wv * v).sum() + (wl * lp).sum() + 0.01 * ent.sum()) / (n * B)
        grads = torch.autograd.grad(loss, params)
        if ref is None:
            ref = (v.detach(), lp.detach(), grads)
            continue
        ev = float((v.detach() - ref[0]).abs().max() / ref[0].abs().max())
        el = float((lp.detach() - ref[1]).abs().max())
        errs = [(float((g - r).abs().max() / r.abs().max()), k) for g, r, k in zip(grads, ref[2], names)]
        worst = max(errs)
        flat = lambda gs: torch.cat([g.reshape(-1) for g in gs])
        l2 = float((flat(grads) - flat(ref[2])).norm() / flat(ref[2]).norm())
        cos = float(torch.nn.functional.cosine_similarity(flat(grads), flat(ref[2]), dim=0))
        print("%-62s value rel %.1e  log-prob abs %.1e | gradient: worst tensor %.1e (%s), whole vector L2 %.1e, 1 - cos %.1e | %s"
              % (name, ev, el, worst[0], worst[1], l2, 1 - cos,
                 "holds the 2e-4 gate" if worst[0] < 2e-4 else "misses the 2e-4 gate by %.0fx" % (worst[0] / 2e-4)))


if __name__ == "__main__":
    main()
