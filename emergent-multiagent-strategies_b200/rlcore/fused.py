"""Fused pieces of the PPO update on the shared rollout blocks (include/fortattack_rollout.h):
`gather_minibatch` replaces magent_feed_forward_generator's ~10 x n index + cat launches per minibatch
(rlcore/algo/ppo.py:207-246) by one kernel, `ppo_loss` evaluates the masked clipped-PPO loss of
ppo.py:150-187 and its gradient with respect to (values, log-probs, entropy) in one pass."""
import ctypes

import torch

try:
    from .. import _capi
except ImportError:          # drop-in mode (this directory's parent is on sys.path as top-level)
    import _capi


def _lib():
    L = _capi.lib()
    if not getattr(L, "_rl2_bound", False):
        vp, i32, f32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
        L.rl_gather_minibatch.argtypes = [vp] + [i32] * 8 + [vp] * 18
        L.rl_ppo_loss.argtypes = [vp] * 9 + [i32, f32, f32, f32] + [vp] * 5
        L._rl2_bound = True
    return L


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
class _Opnd(ctypes.Structure):
    _fields_ = [("ptr", ctypes.c_void_p), ("batch_stride", ctypes.c_int64), ("row_stride", ctypes.c_int64)]


def _attn_lib():
    L = _lib()
    if not getattr(L, "_rl3_bound", False):
        P, vp, i32, f32 = ctypes.POINTER(_Opnd), ctypes.c_void_p, ctypes.c_int, ctypes.c_float
        L.rl_attn_forward.argtypes = [P, P, P, P, vp, i32, i32, i32, i32, f32, i32, vp]
        L.rl_attn_backward.argtypes = [P, P, P, P, vp, P, P, P, i32, i32, i32, i32, f32, vp]
        L._rl3_bound = True
    return L


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
