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
