"""Fused pieces of the PPO update on the shared rollout blocks (include/fortattack_rollout.h):
`gather_minibatch` replaces magent_feed_forward_generator's ~10 x n index + cat launches per minibatch
(rlcore/algo/ppo.py:207-246) by one kernel, `ppo_loss` evaluates the masked clipped-PPO loss of
ppo.py:150-187 and its gradient with respect to (values, log-probs, entropy) in one pass."""

import torch

try:
    from .. import _capi
except ImportError:          # drop-in mode (this directory's parent is on sys.path as top-level)
    import _capi


def _lib():
    return _capi.lib()           # every prototype is declared there


def gather_minibatch(shared, idx, a0, n, o0, m, adv):
    """shared: rollout.SharedRollouts; idx int64 [mb] on the device (flat t*E+e); adv float [n, T, E].
    Returns (obs_own [n*mb,6], alive [n*mb,1], obs_opp [m*mb,6], actions [n*mb,1], value_preds, returns, masks,
    old_logp, adv  (each [n*mb,1]), alive_sum [1])."""
    T, A, E = shared.rewards.shape
    mb, dev = idx.numel(), idx.device
    f = lambda *s: torch.empty(*s, device=dev)
    obs_own, obs_opp = f(n * mb, 6), f(m * mb, 6)
    actions = torch.empty(n * mb, 1, dtype=torch.int64, device=dev)
    vp_, ret, msk, olp, ad, alive = (f(n * mb, 1) for _ in range(6))
    alive_sum = torch.zeros(1, device=dev)
    assert adv.is_contiguous() and adv.shape == (n, T, E) and idx.dtype == torch.int64 and idx.is_contiguous()
    _capi.check(_lib().rl_gather_minibatch(
        idx.data_ptr(), mb, T, A, E, a0, n, o0, m, shared.obs.data_ptr(), shared.actions.data_ptr(),
        shared.value_preds.data_ptr(), shared.returns.data_ptr(), shared.masks.data_ptr(), shared.action_log_probs.data_ptr(),
        adv.data_ptr(), obs_own.data_ptr(), obs_opp.data_ptr(), actions.data_ptr(), vp_.data_ptr(), ret.data_ptr(),
        msk.data_ptr(), olp.data_ptr(), ad.data_ptr(), alive.data_ptr(), alive_sum.data_ptr(),
        torch.cuda.current_stream(dev).cuda_stream))
    return obs_own, alive, obs_opp, actions, vp_, ret, msk, olp, ad, alive_sum


class _PPOLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, values, logp, entropy, old_values, returns, old_logp, adv, mask, norm, clip, vcoef, ecoef):
        c = lambda t: t.detach().reshape(-1).contiguous().float()
        v, lp, en = c(values), c(logp), c(entropy)
        N = v.numel()
        out = torch.zeros(4, device=v.device)
        gv, glp, gen = torch.empty_like(v), torch.empty_like(v), torch.empty_like(v)
        _capi.check(_lib().rl_ppo_loss(v.data_ptr(), lp.data_ptr(), en.data_ptr(), c(old_values).data_ptr(), c(returns).data_ptr(),
                                       c(old_logp).data_ptr(), c(adv).data_ptr(), c(mask).data_ptr(), c(norm).data_ptr(), N,
                                       float(clip), float(vcoef), float(ecoef), out.data_ptr(), gv.data_ptr(), glp.data_ptr(),
                                       gen.data_ptr(), torch.cuda.current_stream(v.device).cuda_stream))
        ctx.save_for_backward(gv, glp, gen)
        ctx.shapes = (values.shape, logp.shape, entropy.shape)
        ctx.mark_non_differentiable(out)
        return out[3].clone(), out

    @staticmethod
    def backward(ctx, g_total, _g_out):
        gv, glp, gen = ctx.saved_tensors
        sv, slp, sen = ctx.shapes
        return ((gv * g_total).view(sv), (glp * g_total).view(slp), (gen * g_total).view(sen)) + (None,) * 9


def ppo_loss(values, logp, entropy, old_values, returns, old_logp, adv, mask, norm, clip, vcoef, ecoef):
    """-> (total loss with autograd, stats [value_loss, action_loss, entropy, total] detached)."""
    return _PPOLoss.apply(values, logp, entropy, old_values, returns, old_logp, adv, mask, norm, clip, vcoef, ecoef)


# ---- attention over a handful of agents: forward + backward kernels behind autograd ---------------------------------
_Opnd = _capi.RlAttnOperand


def _attn_lib():
    return _capi.lib()


def _op(t, col, width, batch):
    """Agent-major flat rows [agents * batch, width]: agent i of batch entry b is row i * batch + b; take the columns
    from `col` on."""
    return _Opnd(t.data_ptr() + 4 * col, width, batch * width)


class _SelfAttention(torch.autograd.Function):
    """qkv [n*B, 3k] (agent-major rows, Q | K | V) -> out [n*B, k], attn [B, n, n]; no self-messages (mpnn.py:297-298)."""

    @staticmethod
    def forward(ctx, qkv, n, norm):
        qkv = qkv.contiguous()
        B, k = qkv.shape[0] // n, qkv.shape[1] // 3
        out = torch.empty(n * B, k, device=qkv.device)
        attn = torch.empty(B, n, n, device=qkv.device)
        st = torch.cuda.current_stream(qkv.device).cuda_stream
        _capi.check(_attn_lib().rl_attn_forward(_op(qkv, 0, 3 * k, B), _op(qkv, k, 3 * k, B), _op(qkv, 2 * k, 3 * k, B),
                                                _op(out, 0, k, B), attn.data_ptr(), B, n, n, k, float(norm), 1, st))
        ctx.save_for_backward(qkv, attn)
        ctx.meta = (n, float(norm))
        ctx.mark_non_differentiable(attn)
        return out, attn

    @staticmethod
    def backward(ctx, g, _ga):
        qkv, attn = ctx.saved_tensors
        n, norm = ctx.meta
        B, k = qkv.shape[0] // n, qkv.shape[1] // 3
        g = g.contiguous()
        d = torch.empty_like(qkv)
        st = torch.cuda.current_stream(qkv.device).cuda_stream
        _capi.check(_attn_lib().rl_attn_backward(_op(g, 0, k, B), _op(qkv, 0, 3 * k, B), _op(qkv, k, 3 * k, B),
                                                 _op(qkv, 2 * k, 3 * k, B), attn.data_ptr(), _op(d, 0, 3 * k, B),
                                                 _op(d, k, 3 * k, B), _op(d, 2 * k, 3 * k, B), B, n, n, k, norm, st))
        return d, None, None


class _CrossAttention(torch.autograd.Function):
    """a [n*B, k] (rows of the own team), bv [m*B, 2k] (B | V of the other team) -> out [n*B, k], attn [B, n, m]
    (mpnn.py:409-437: compatibility = K(own) . Q(opp), values from the opponents)."""

    @staticmethod
    def forward(ctx, a, bv, n, m, norm):
        a, bv = a.contiguous(), bv.contiguous()
        B, k = a.shape[0] // n, a.shape[1]
        out = torch.empty(n * B, k, device=a.device)
        attn = torch.empty(B, n, m, device=a.device)
        st = torch.cuda.current_stream(a.device).cuda_stream
        _capi.check(_attn_lib().rl_attn_forward(_op(a, 0, k, B), _op(bv, 0, 2 * k, B), _op(bv, k, 2 * k, B), _op(out, 0, k, B),
                                                attn.data_ptr(), B, n, m, k, float(norm), 0, st))
        ctx.save_for_backward(a, bv, attn)
        ctx.meta = (n, m, float(norm))
        ctx.mark_non_differentiable(attn)
        return out, attn

    @staticmethod
    def backward(ctx, g, _ga):
        a, bv, attn = ctx.saved_tensors
        n, m, norm = ctx.meta
        B, k = a.shape[0] // n, a.shape[1]
        g = g.contiguous()
        da, dbv = torch.empty_like(a), torch.empty_like(bv)
        st = torch.cuda.current_stream(a.device).cuda_stream
        _capi.check(_attn_lib().rl_attn_backward(_op(g, 0, k, B), _op(a, 0, k, B), _op(bv, 0, 2 * k, B), _op(bv, k, 2 * k, B),
                                                 attn.data_ptr(), _op(da, 0, k, B), _op(dbv, 0, 2 * k, B), _op(dbv, k, 2 * k, B),
                                                 B, n, m, k, norm, st))
        return da, dbv, None, None, None


def self_attention(qkv, n, norm):
    return _SelfAttention.apply(qkv, n, norm)


def cross_attention(a, bv, n, m, norm):
    return _CrossAttention.apply(a, bv, n, m, norm)


# ---- dense layers of the training forward with a hand-written backward ------------------------------------------------
def _relu_bwd_colsum_cuda(dout, out):
    """(dpre = dout * [out > 0], column sums of dpre) in one pass over [rows, cols] (rl_relu_bwd_colsum)."""
    L = _lib()
    rows, cols = out.shape
    blocks = L.rl_relu_bwd_colsum_blocks(rows, cols)
    if blocks < 1:                                   # widths the kernel does not cover
        dpre = torch.ops.aten.threshold_backward(dout, out, 0)
        return dpre, dpre.sum(0)
    dout = dout.contiguous()
    dpre = torch.empty_like(out)
    partial = torch.empty(blocks, cols, device=out.device)
    _capi.check(L.rl_relu_bwd_colsum(dout.data_ptr(), out.data_ptr(), dpre.data_ptr(), partial.data_ptr(), rows, cols,
                                     torch.cuda.current_stream(out.device).cuda_stream))
    return dpre, partial.sum(0)


relu_bwd_colsum = _relu_bwd_colsum_cuda               # tests substitute a torch restatement on the CPU

_SPLITS = {}


def _split(rows):
    """Number of row chunks for xt_dy: the largest divisor of `rows` that is <= 96 and leaves chunks of >= 512 rows."""
    if rows not in _SPLITS:
        s = 1
        for c in range(min(96, rows // 512), 1, -1):
            if rows % c == 0:
                s = c
                break
        _SPLITS[rows] = s
    return _SPLITS[rows]


def xt_dy(x, dy):
    """x^T @ dy for tall operands ([rows, a]^T [rows, b] -> [a, b], rows ~ 2e5): the weight-gradient product of every
    dense layer.  As a plain GEMM its whole reduction dimension lands on a handful of CTAs (measured 90-170 us for
    200 MB of operands); as a batched GEMM over row chunks plus a sum of the [chunks, a, b] partials every SM works."""
    rows = x.shape[0]
    S = _split(rows)
    if S == 1:
        return x.t() @ dy
    r = rows // S
    return torch.bmm(x.unflatten(0, (S, r)).transpose(1, 2), dy.unflatten(0, (S, r))).sum(0)


class _Linear(torch.autograd.Function):
    """y = x W^T + b, optionally followed by ReLU (bias and activation in the GEMM epilogue); backward = fused ReLU-backward
    + bias gradient, split-K weight gradient, one GEMM for dx (skipped for network inputs)."""

    @staticmethod
    def forward(ctx, x, W, b, relu):
        if relu:
            y = torch._addmm_activation(b, x, W.t()) if x.is_cuda else torch.relu(torch.addmm(b, x, W.t()))
        else:
            y = torch.addmm(b, x, W.t())
        ctx.relu = relu
        ctx.save_for_backward(x, W, y if relu else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, W, y = ctx.saved_tensors
        if ctx.relu:
            dpre, db = relu_bwd_colsum(dy, y)
        else:
            dpre = dy.contiguous()
            db = dpre.sum(0)
        dW = xt_dy(dpre, x)
        dx = dpre @ W if ctx.needs_input_grad[0] else None
        return dx, dW, db, None


class _Matmul(torch.autograd.Function):
    """y = x M with the split-K weight gradient."""

    @staticmethod
    def forward(ctx, x, M):
        ctx.save_for_backward(x, M)
        return x @ M

    @staticmethod
    def backward(ctx, dy):
        x, M = ctx.saved_tensors
        dy = dy.contiguous()
        return (dy @ M.t() if ctx.needs_input_grad[0] else None), xt_dy(x, dy)


def linear(x, W, b, relu=False):
    return _Linear.apply(x, W, b, relu)


def matmul(x, M):
    return _Matmul.apply(x, M)


# ---- one message-passing round with the projections folded into the weights --------------------------------------------
def _col(t, col, batch):
    """Columns `col`... of a row-major [agents * batch, width] tensor as an attention operand (agent-major rows)."""
    assert t.stride(1) == 1
    return _Opnd(t.data_ptr() + 4 * col, t.stride(0), batch * t.stride(0))


def _mix_forward_cuda(g, h, n, norm):
    B, k = h.shape[0] // n, h.shape[1]
    hm = torch.empty(n * B, 2 * k, device=h.device)
    attn = torch.empty(B, n, n, device=h.device)
    _capi.check(_attn_lib().rl_attn_mix_forward(_col(g, 0, B), _col(h, 0, B), _col(hm, k, B), _col(hm, 0, B), attn.data_ptr(),
                                                B, n, n, k, float(norm), 1, torch.cuda.current_stream(h.device).cuda_stream))
    return hm, attn


def _mix_backward_cuda(dhm, g, hm, attn, n, norm):
    """dhm [n*B, 2k] = [dh_direct | dmixed] -> (dg [n*B, k], dh [n*B, k] = dh_direct + attention gradient)."""
    B, k = g.shape[0] // n, g.shape[1]
    dg, dh = torch.empty_like(g), torch.empty_like(g)
    _capi.check(_attn_lib().rl_attn_mix_backward(_col(dhm, k, B), _col(g, 0, B), _col(hm, 0, B), attn.data_ptr(), _col(dg, 0, B),
                                                 _col(dh, 0, B), _col(dhm, 0, B), B, n, n, k, float(norm),
                                                 torch.cuda.current_stream(g.device).cuda_stream))
    return dg, dh


# the two device steps of a round; tests substitute torch restatements to check the hand-written backward on the CPU
mix_forward, mix_backward = _mix_forward_cuda, _mix_backward_cuda


class _MessageRound(torch.autograd.Function):
    """h' = relu(cat(h, mixed) @ Wc + bias),  mixed_i = sum_{j != i} softmax_j(norm <(h Mqk)_i, h_j>) h_j.

    With Mqk = W_query W_key^T and Wc = [U1^T ; W_val W_out U2^T] this is one round of mpnn.py:157-159
    (h = update(cat(h, messages(h)))) with every [rows, d] x [d, d] projection of the attention folded into [d, d] products
    of the weights: three row-sized GEMMs per round (G, the update, and none for Q/K/V/out) instead of six, and half
    the activation traffic.  Forward and backward are written out by hand so that nothing but hm, G, the attention
    matrix and the output is kept."""

    @staticmethod
    def forward(ctx, h, Mqk, Wc, bias, n, norm):
        h = h.contiguous()
        g = h @ Mqk
        hm, attn = mix_forward(g, h, n, norm)
        out = torch._addmm_activation(bias, hm, Wc) if h.is_cuda else torch.relu(torch.addmm(bias, hm, Wc))
        ctx.save_for_backward(g, hm, attn, out, Mqk, Wc)
        ctx.meta = (n, float(norm))
        ctx.mark_non_differentiable(attn)
        return out, attn

    @staticmethod
    def backward(ctx, dout, _dattn):
        g, hm, attn, out, Mqk, Wc = ctx.saved_tensors
        n, norm = ctx.meta
        k = g.shape[1]
        dpre, dbias = relu_bwd_colsum(dout, out)
        dWc = xt_dy(hm, dpre)
        dhm = dpre @ Wc.t()                                  # [dh through U1 | dmixed]
        dg, dh = mix_backward(dhm, g, hm, attn, n, norm)
        dMqk = xt_dy(hm[:, :k], dg)
        dh.addmm_(dg, Mqk.t())
        return dh, dMqk, dWc, dbias, None, None


def message_round(h, Mqk, Wc, bias, n, norm):
    return _MessageRound.apply(h, Mqk, Wc, bias, n, norm)
