"""CPU: the hand-written forward/backward of one folded message-passing round (rlcore/fused.py _MessageRound) against
autograd on the mirror of mpnn.py:157-159,249-331.  The two device steps (rl_attn_mix_forward / rl_attn_mix_backward)
are replaced by torch restatements here, so what is checked is the folding algebra and the manual gradient chain; the
kernels themselves are compared with the bmm path on the GPU (tests/test_rollout_gpu.py)."""
from importlib import import_module

import pytest
import torch

PKG = "emergent-multiagent-strategies_b200"
mp = import_module(PKG + ".mpnn")
fused = import_module(PKG + ".rlcore.fused")


class Shape(object):
    def __init__(self, *shape):
        self.shape = shape


def _mix(g, h, n, norm):
    B, k = h.shape[0] // n, h.shape[1]
    G, H = g.view(n, B, k).transpose(0, 1), h.view(n, B, k).transpose(0, 1)
    s = norm * G @ H.transpose(1, 2)
    s = s.masked_fill(torch.eye(n, dtype=torch.bool), float("-inf"))
    attn = torch.softmax(s, dim=-1)
    return (attn @ H).transpose(0, 1).reshape(n * B, k), attn


def mix_forward(g, h, n, norm):
    mixed, attn = _mix(g, h, n, norm)
    return torch.cat((h, mixed), dim=1), attn


def mix_backward(dhm, g, hm, attn, n, norm):
    k = g.shape[1]
    with torch.enable_grad():
        g2, h2 = g.detach().clone().requires_grad_(), hm[:, :k].detach().clone().requires_grad_()
        mixed, _ = _mix(g2, h2, n, norm)
        dg, dh = torch.autograd.grad(mixed, (g2, h2), dhm[:, k:])
    return dg, dh + dhm[:, :k]


def relu_bwd_colsum(dout, out):
    dpre = dout * (out > 0)
    return dpre, dpre.sum(0)


def test_dense_functions_match_autograd(monkeypatch):
    """fused.linear / fused.matmul (manual backward: ReLU-backward + bias gradient, split weight gradient) == F.linear."""
    monkeypatch.setattr(fused, "relu_bwd_colsum", relu_bwd_colsum)
    torch.manual_seed(0)
    for rows in (7, 1024 * 3, 5120):                    # 3072 -> 6 chunks, 5120 -> 10 chunks, 7 -> plain product
        x = torch.randn(rows, 24, dtype=torch.float64, requires_grad=True)
        W = torch.randn(16, 24, dtype=torch.float64, requires_grad=True)
        b = torch.randn(16, dtype=torch.float64, requires_grad=True)
        M = torch.randn(16, 8, dtype=torch.float64, requires_grad=True)
        w = torch.randn(rows, 8, dtype=torch.float64)
        for relu in (True, False):
            ref = torch.nn.functional.linear(x, W, b)
            ref = (torch.relu(ref) if relu else ref) @ M
            got = fused.matmul(fused.linear(x, W, b, relu), M)
            assert torch.allclose(ref, got, rtol=1e-12, atol=1e-12)
            for a, c in zip(torch.autograd.grad((ref * w).sum(), (x, W, b, M)), torch.autograd.grad((got * w).sum(), (x, W, b, M))):
                assert torch.allclose(a, c, rtol=1e-10, atol=1e-10)
    assert fused._split(5120) == 10 and fused._split(196608) == 96 and fused._split(7) == 1 and fused._split(40960 * 4) == 80


@pytest.mark.parametrize("n,hid", [(3, 128), (5, 128), (2, 64)])
def test_folded_rounds_match_the_mirror(n, hid, monkeypatch):
    monkeypatch.setattr(fused, "mix_forward", mix_forward)
    monkeypatch.setattr(fused, "mix_backward", mix_backward)
    monkeypatch.setattr(fused, "relu_bwd_colsum", relu_bwd_colsum)
    torch.manual_seed(n + hid)
    net = mp.MPNN(action_space=Shape(8), num_agents=n, num_opp_agents=3, input_size=6, hidden_dim=hid).double()
    with torch.no_grad():
        net.update[0].bias.uniform_(-0.3, 0.3)
    ms, B = net.messages, 37
    h_in = torch.randn(n * B, hid, dtype=torch.float64)
    w = torch.randn(n * B, hid, dtype=torch.float64)
    params = [ms.W_query, ms.W_key, ms.W_val, ms.W_out, net.update[0].weight, net.update[0].bias]
    res = []
    for folded in (False, True):
        h = h_in.clone().requires_grad_()
        x = h
        if folded:
            W, bias = net.update[0].weight, net.update[0].bias
            Mqk = ms.W_query[0] @ ms.W_key[0].t()
            Wc = torch.cat((W[:, :hid].t(), ms.W_val[0] @ ms.W_out[0] @ W[:, hid:].t()), dim=0)
            for _ in range(3):
                x, attn = fused.message_round(x, Mqk, Wc, bias, n, ms.norm_factor)
        else:
            x3 = x.view(n, B, hid).transpose(0, 1)
            for _ in range(3):
                msg, attn = ms(x3, return_attn=True)
                x3 = net.update(torch.cat((x3, msg), dim=2))
            x = x3.transpose(0, 1).reshape(n * B, hid)
            attn = attn.squeeze(0)
        grads = torch.autograd.grad((x * w).sum(), [h] + params)
        res.append((x.detach(), attn.detach(), grads))
    (xa, aa, ga), (xb, ab, gb) = res
    assert torch.allclose(xa, xb, rtol=1e-10, atol=1e-12) and torch.allclose(aa.reshape(ab.shape), ab, atol=1e-12)
    for a, b in zip(ga, gb):
        assert torch.allclose(a, b, rtol=1e-9, atol=1e-11), float((a - b).abs().max())


# ---- the tcgen05 wiring (DENSE = "tcgen05"): which operand is packed how, on exact CPU stand-ins of the three primitives ----
def _tg_pack(w, transposed):
    B = (w.detach().t() if transposed else w.detach()).clone()
    return (B, B.shape[0], B.shape[1])


def _tg_linear(x, pack, bias=None, relu=False, out=None, accumulate=False, res=None):
    B, N, K = pack
    assert x.shape[1] == K, (tuple(x.shape), K)
    y = x.detach() @ B.t()
    if bias is not None:
        y = y + bias.detach()
    if relu:
        y = torch.relu(y)
    if res is not None:
        y = y + res.detach()
    if out is None:
        return y
    assert out.shape == y.shape
    if accumulate:
        out += y
    else:
        out.copy_(y)
    return out


def _cross_forward(kk, qv, n, m, norm):                  # torch restatement of rl_attn_forward (mpnn.py:409-437), agent-major rows
    Bsz, k = kk.shape[0] // n, kk.shape[1]
    A = kk.view(n, Bsz, k).transpose(0, 1)
    Q, V = qv[:, :k].reshape(m, Bsz, k).transpose(0, 1), qv[:, k:].reshape(m, Bsz, k).transpose(0, 1)
    attn = torch.softmax(norm * A @ Q.transpose(1, 2), dim=-1)
    return (attn @ V).transpose(0, 1).reshape(n * Bsz, k), attn


def _cross_backward(de, kk, qv, attn, n, m, norm):       # ... and of rl_attn_backward, by autograd on the restatement
    with torch.enable_grad():
        kk2, qv2 = kk.detach().requires_grad_(), qv.detach().requires_grad_()
        e, _ = _cross_forward(kk2, qv2, n, m, norm)
        return torch.autograd.grad((e * de).sum(), (kk2, qv2))


def _tg_wgrad(x, y):
    return x.detach().t() @ y.detach()


def _patch_tg(monkeypatch):
    monkeypatch.setattr(fused, "tg_pack", _tg_pack)
    monkeypatch.setattr(fused, "tg_linear", _tg_linear)
    monkeypatch.setattr(fused, "tg_wgrad", _tg_wgrad)
    monkeypatch.setattr(fused, "_use_tg", lambda x: True)
    monkeypatch.setattr(fused, "mix_forward", mix_forward)
    monkeypatch.setattr(fused, "mix_backward", mix_backward)
    monkeypatch.setattr(fused, "relu_bwd_colsum", relu_bwd_colsum)
    monkeypatch.setattr(fused, "cross_forward", _cross_forward)
    monkeypatch.setattr(fused, "cross_backward", _cross_backward)


def test_tcgen05_wiring_of_dense_functions(monkeypatch):
    """fused.linear / matmul / matmul_nt routed through tg_pack / tg_linear / tg_wgrad (stand-ins with the kernels' contract:
    y = act(x B^T + b), dW = x^T y) reproduce autograd: every operand is packed in the right orientation."""
    _patch_tg(monkeypatch)
    torch.manual_seed(1)
    x = torch.randn(50, 24, dtype=torch.float64, requires_grad=True)
    W = torch.randn(16, 24, dtype=torch.float64, requires_grad=True)
    b = torch.randn(16, dtype=torch.float64, requires_grad=True)
    M = torch.randn(16, 8, dtype=torch.float64, requires_grad=True)
    Bn = torch.randn(5, 8, dtype=torch.float64, requires_grad=True)
    w = torch.randn(50, 5, dtype=torch.float64)
    for relu in (True, False):
        ref = torch.nn.functional.linear(x, W, b)
        ref = ((torch.relu(ref) if relu else ref) @ M) @ Bn.t()
        got = fused.matmul_nt(fused.matmul(fused.linear(x, W, b, relu), M), Bn)
        assert torch.allclose(ref, got, rtol=1e-12, atol=1e-12)
        for a, c in zip(torch.autograd.grad((ref * w).sum(), (x, W, b, M, Bn)), torch.autograd.grad((got * w).sum(), (x, W, b, M, Bn))):
            assert torch.allclose(a, c, rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize("n,m", [(3, 3), (1, 2), (4, 2)])
def test_tcgen05_wiring_of_the_whole_network(n, m, monkeypatch):
    """MPNN.evaluate_actions through the fused path with the dense products on the (stand-in) tcgen05 primitives ==
    the bmm mirror of mpnn.py:117-205: outputs and every parameter gradient, in float64."""
    _patch_tg(monkeypatch)

    def cross(a, bv, n_, m_, norm):                       # torch restatement of rl_attn_forward/backward (mpnn.py:409-437)
        Bsz, k = a.shape[0] // n_, a.shape[1]
        A = a.view(n_, Bsz, k).transpose(0, 1)
        Bq, V = bv[:, :k].reshape(m_, Bsz, k).transpose(0, 1), bv[:, k:].reshape(m_, Bsz, k).transpose(0, 1)
        attn = torch.softmax(norm * A @ Bq.transpose(1, 2), dim=-1)
        return (attn @ V).transpose(0, 1).reshape(n_ * Bsz, k), attn
    monkeypatch.setattr(fused, "cross_attention", cross)
    torch.manual_seed(n * 7 + m)
    net = mp.MPNN(action_space=Shape(8), num_agents=n, num_opp_agents=m, input_size=6, hidden_dim=128).double()
    with torch.no_grad():
        for p in net.parameters():
            if p.dim() == 1:
                p.uniform_(-0.3, 0.3)
    monkeypatch.setattr(mp.MPNN, "_use_fused", lambda self, x: self.fused_attention)
    Bsz = 23
    own, opp = torch.randn(n * Bsz, 6, dtype=torch.float64), torch.randn(m * Bsz, 6, dtype=torch.float64)
    act = torch.randint(0, 8, (n * Bsz, 1))
    w = torch.randn(n * Bsz, 1, dtype=torch.float64)
    res = []
    for fused_on in (False, True):
        net.zero_grad()
        net.fused_attention = fused_on
        if fused_on:
            # the production path of the update: front_end + folded rounds + stacked heads (what evaluate_actions /
            # evaluate_logits run on CUDA)
            x = net._fwd_fused(own, opp)
            v, logits = net._heads(x)
            dist = import_module(PKG + ".rlcore.distributions").FixedCategorical(logits=logits)
            lp, ent = dist.log_probs(act), dist.entropy()
        else:
            v, lp, ent, _ = net.evaluate_actions(own, None, opp, None, act)
        ((v * w).sum() + (lp * w).sum() * 0.7 + ent.sum() * 0.3).backward()
        res.append((v.detach(), lp.detach(), {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}))
        net.fused_attention = False
    (v0, lp0, g0), (v1, lp1, g1) = res
    assert torch.allclose(v0, v1, rtol=1e-10, atol=1e-10) and torch.allclose(lp0, lp1, rtol=1e-10, atol=1e-10)
    assert set(g0) == set(g1)
    for k in g0:
        assert torch.allclose(g0[k], g1[k], rtol=1e-8, atol=1e-9), (k, float((g0[k] - g1[k]).abs().max()))


def test_scratch_scope_and_flat_buffer_host_logic():
    """Host logic of the overlapped team updates that needs no GPU: scratch storage is keyed by (device, scope) and scopes
    nest; JointPPO.use_flat_buffer hands out views of a caller-owned segment and refuses a wrong size."""
    assert fused.scratch_scope(0) == 0
    k0 = fused._stream_key("cpu")
    prev = fused.scratch_scope(1234)
    assert prev == 0 and fused._stream_key("cpu") == (torch.device("cpu"), 1234) and fused._stream_key("cpu") != k0
    inner = fused.scratch_scope(99)
    assert inner == 1234 and fused._stream_key("cpu")[1] == 99
    fused.scratch_scope(inner)
    fused.scratch_scope(prev)
    assert fused._stream_key("cpu") == k0
    ppo = import_module(PKG + ".rlcore.algo.ppo")
    net = mp.MPNN(action_space=Shape(8), num_agents=2, num_opp_agents=2, num_entities=0, input_size=6, hidden_dim=32, pos_index=2)
    tr = ppo.JointPPO(net, 0.2, 1, 1, 0.5, 0.01, lr=1e-4, max_grad_norm=0.5)
    params = [p for p in net.parameters()]
    total = sum(p.numel() for p in params)
    shared = torch.zeros(2 * (total + 5))
    tr.use_flat_buffer(shared[total + 5:])
    flat, views = tr._flat_grads(params)
    assert flat.data_ptr() == shared[total + 5:].data_ptr() and flat.numel() == total + 5
    assert [tuple(v.shape) for v in views] == [tuple(p.shape) for p in params]
    views[0].fill_(1.0)
    assert float(shared[total + 5:total + 5 + params[0].numel()].sum()) == params[0].numel() and float(shared[:total + 5].sum()) == 0.0
    with pytest.raises(ValueError):
        tr.use_flat_buffer(shared[:total])
    tr.release_graphs()
    assert tr._g is None and tr._joint is None


def test_flat_buffer_halves_of_the_several_ranks_step():
    """JointPPO._ranks_fill_flat / _ranks_apply_flat around a SIMULATED all-reduce (the sum of two ranks' flat buffers), with two
    trainers sharing one buffer as in BatchedTrainer's joint step: every parameter ends where one process holding both ranks'
    data and the reference's normalisation (sums / global alive count, ppo.py:150-187) puts it."""
    ppo = import_module(PKG + ".rlcore.algo.ppo")

    def make(seed):
        torch.manual_seed(seed)
        return mp.MPNN(action_space=Shape(8), num_agents=2, num_opp_agents=2, num_entities=0, input_size=6, hidden_dim=32, pos_index=2)

    def local_sums(net, x, mask):
        """un-normalised local sums of a toy loss that touches every parameter the same way on all ranks"""
        s = sum((p * p).sum() for p in net.parameters())
        per_row = (x * mask).sum(1)
        return per_row, s

    gen = torch.Generator().manual_seed(5)
    data = [(torch.randn(7, 3, generator=gen), (torch.rand(7, 1, generator=gen) > 0.4).float()) for _ in range(2)]    # two ranks
    teams = []
    for seed in (1, 2):                                   # two teams = two trainers on ONE shared flat buffer
        nets = [make(seed) for _ in range(3)]             # rank 0, rank 1, the single-process reference
        trs = [ppo.JointPPO(n, 0.2, 1, 1, 0.5, 0.01, lr=1e-3, max_grad_norm=0.5) for n in nets]
        teams.append((nets, trs))
    sizes = [sum(p.numel() for p in teams[t][0][0].parameters()) + 5 for t in range(2)]
    shared = [torch.zeros(sum(sizes)) for _ in range(2)]  # one buffer per simulated rank
    for t, (nets, trs) in enumerate(teams):
        off = sum(sizes[:t])
        for r in range(2):
            trs[r].use_flat_buffer(shared[r][off:off + sizes[t]])
    for t, (nets, trs) in enumerate(teams):
        for r in range(2):
            x, mask = data[r]
            net, params = nets[r], [p for p in nets[r].parameters()]

            def loss_of(norm, net=net, x=x, mask=mask):
                per_row, s = local_sums(net, x, mask)
                total = (per_row.sum() * s) / norm.reshape(())
                return total, torch.stack([total.detach(), total.detach() * 2, total.detach() * 3, total.detach()])
            trs[r]._ranks_fill_flat(loss_of, mask, mask.sum().view(1), mask.new_full((1,), float(mask.numel())), params)
    reduced = shared[0] + shared[1]                       # the ONE all-reduce of the joint step
    for r in range(2):
        shared[r].copy_(reduced)
    for t, (nets, trs) in enumerate(teams):
        totals = [torch.zeros(3) for _ in range(2)]
        for r in range(2):
            trs[r]._ranks_apply_flat(totals[r], [p for p in nets[r].parameters()])
        # the single-process reference: both ranks' rows, divided by the global alive count, clip + Adam
        ref, rtr = nets[2], trs[2]
        alive = sum(float(m.sum()) for _x, m in data)
        loss = sum(local_sums(ref, x, m)[0].sum() * local_sums(ref, x, m)[1] for x, m in data) / alive
        rtr.optimizer.zero_grad()
        loss.backward()
        rtr._clip_and_step()
        for pa, pb, pr in zip(nets[0].parameters(), nets[1].parameters(), ref.parameters()):
            assert torch.equal(pa, pb)                                        # replicas identical
            assert torch.allclose(pa, pr, rtol=1e-5, atol=1e-7), float((pa - pr).abs().max())
        assert torch.allclose(totals[0], torch.stack([loss.detach(), loss.detach() * 2, loss.detach() * 3]), rtol=1e-5)
