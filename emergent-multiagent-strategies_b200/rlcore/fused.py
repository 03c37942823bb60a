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


# ---- the loss taken from the action head's logits: Categorical log-prob / entropy and their backward inside the loss kernel --
LOSS_ACTIONS = 8
_LOSS_SCRATCH = {}


_SCOPE = 0


def scratch_scope(token):
    """Scratch buffers (partial sums, counters, status words) are kept per (device, scope).  JointPPO brackets its optimizer
    steps with scratch_scope(id(self)) / scratch_scope(0): the two teams' updates may run on different streams at the same time
    (rollout.BatchedTrainer.update) and must not share that storage.  A scope, unlike the current stream, is the same for the
    eager warm-up steps and the captured graph, so nothing is allocated (or zeroed) inside a capture."""
    global _SCOPE
    prev, _SCOPE = _SCOPE, token
    return prev


def _stream_key(dev):
    return (torch.device(dev), _SCOPE)


def _loss_scratch(dev):
    key = _stream_key(dev)
    s = _LOSS_SCRATCH.get(key)
    if s is None:
        s = _LOSS_SCRATCH[key] = torch.zeros(int(_lib().rl_ppo_loss_logits_scratch_floats()), device=dev)
    return s


def categorical_eval(logits, actions):
    """(log-prob of `actions` [N], entropy [N]) of Categorical(logits=logits [N, 8]) in the loss kernel's arithmetic, no
    autograd (BatchedTrainer.recompute_old: the behaviour log-probs the update's ratio starts from)."""
    lg = logits.detach().contiguous().float()
    act = actions.detach().reshape(-1).contiguous()
    N = lg.shape[0]
    if lg.shape[1] != LOSS_ACTIONS or act.dtype != torch.int64 or act.numel() != N:
        raise ValueError("categorical_eval: logits [N, %d] float and int64 actions [N] expected" % LOSS_ACTIONS)
    lp, ent = torch.empty(N, device=lg.device), torch.empty(N, device=lg.device)
    if N:
        _capi.check(_lib().rl_ppo_loss_logits(None, lg.data_ptr(), act.data_ptr(), None, None, None, None, None, None, N, LOSS_ACTIONS,
                                              0.0, 0.0, 0.0, None, None, None, lp.data_ptr(), ent.data_ptr(), None,
                                              torch.cuda.current_stream(lg.device).cuda_stream))
    return lp, ent


class _PPOLossLogits(torch.autograd.Function):
    @staticmethod
    def forward(ctx, values, logits, actions, old_values, returns, old_logp, adv, mask, norm, clip, vcoef, ecoef):
        c = lambda t: t.detach().reshape(-1).contiguous().float()
        v, lg = c(values), logits.detach().contiguous().float()
        act = actions.detach().reshape(-1).contiguous()
        N = v.numel()
        if lg.shape != (N, LOSS_ACTIONS) or act.dtype != torch.int64 or act.numel() != N:
            raise ValueError("ppo_loss_logits: logits [N, %d] float and int64 actions [N] expected" % LOSS_ACTIONS)
        out = torch.empty(4, device=v.device)
        gv, glg = torch.empty_like(v), torch.empty_like(lg)
        _capi.check(_lib().rl_ppo_loss_logits(v.data_ptr(), lg.data_ptr(), act.data_ptr(), c(old_values).data_ptr(),
                                              c(returns).data_ptr(), c(old_logp).data_ptr(), c(adv).data_ptr(), c(mask).data_ptr(),
                                              c(norm).data_ptr(), N, LOSS_ACTIONS, float(clip), float(vcoef), float(ecoef),
                                              out.data_ptr(), gv.data_ptr(), glg.data_ptr(), None, None,
                                              _loss_scratch(v.device).data_ptr(), torch.cuda.current_stream(v.device).cuda_stream))
        ctx.save_for_backward(gv, glg)
        ctx.shapes = (values.shape, logits.shape)
        ctx.mark_non_differentiable(out)
        return out[3].clone(), out

    @staticmethod
    def backward(ctx, g_total, _g_out):
        gv, glg = ctx.saved_tensors
        sv, sl = ctx.shapes
        return ((gv * g_total).view(sv), (glg * g_total).view(sl)) + (None,) * 10


def ppo_loss_logits(values, logits, actions, old_values, returns, old_logp, adv, mask, norm, clip, vcoef, ecoef):
    """ppo_loss with the Categorical head inside: -> (total loss with autograd through values and logits, stats
    [value_loss, action_loss, entropy, total] detached)."""
    return _PPOLossLogits.apply(values, logits, actions, old_values, returns, old_logp, adv, mask, norm, clip, vcoef, ecoef)


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


# ---- the dense products themselves: hand-written tcgen05 kernels (csrc/tg_gemm.cu, include/fortattack_train.h) ---------
# DENSE = "tcgen05": every product of the training forward / backward runs on this repo's kernels in fp32-grade split
# arithmetic (fp16 hi/lo with power-of-two row scales; bf16 three-term split for the weight gradients).
# DENSE = "cublas": the same functions on torch's library GEMMs -- kept as the checker the tests compare against.
DENSE = "tcgen05"
_TG = {}


def _tg_state(dev):
    key = _stream_key(dev)
    st = _TG.get(key)
    if st is None:
        st = _TG[key] = {"status": torch.zeros(1, dtype=torch.int32, device=dev),
                         "scratch": torch.empty(_lib().tg_wgrad_scratch_bytes(256, 128), dtype=torch.uint8, device=dev)}
    return st


def tg_check_status(dev):
    """Synchronising: raises if any dense kernel so far reported a pipeline timeout (a protocol bug, never expected)."""
    dev = torch.device(dev)
    for (d, _scope), st in list(_TG.items()):
        if d == dev or (d.type == dev.type and dev.index is None):
            code = int(st["status"].item())
            if code:
                raise _capi.FaError("tg_gemm pipeline timeout (wait site %d)" % code)


def _row_major(t):
    """(tensor usable as a row-major operand, ld): last stride 1, row stride >= width."""
    if t.dim() != 2:
        raise ValueError("2-D operand expected")
    if t.dtype != torch.float32 or (t.shape[1] > 1 and t.stride(1) != 1) or (t.shape[0] > 1 and t.stride(0) < t.shape[1]):
        t = t.contiguous().float()
    # (the stride of a size-1 dimension is arbitrary: a [1, n] or [n, 1] tensor reports whatever its history left there)
    return t, (t.stride(0) if t.shape[0] > 1 else max(int(t.shape[1]), 1))


# Within one optimizer step the same weight is packed several times (the folded round weights: once per round and
# orientation).  JointPPO brackets a step with pack_cache(True) / pack_cache(False); inside the bracket a pack is reused while
# the tensor it was made from is alive and unchanged (the entry keeps the tensor alive, so its storage cannot be handed to
# another weight).  Outside the bracket nothing is cached: parameters change in place between steps.
_PACKS = None


def pack_cache(on):
    global _PACKS
    _PACKS = {} if on else None


def tg_pack(w, transposed):
    """Pack B (the [N][K] operand of y = x B^T) for tg_linear.  w: [N, K] (transposed=False) or [K, N] (transposed=True),
    fp32, row-major with any row stride.  Returns (packed uint8 tensor, N, K)."""
    w, ld = _row_major(w.detach())
    key = None
    if _PACKS is not None:
        key = (w.data_ptr(), tuple(w.shape), ld, bool(transposed), w._version)
        hit = _PACKS.get(key)
        if hit is not None:
            return hit[0]
    N, K = (w.shape[1], w.shape[0]) if transposed else (w.shape[0], w.shape[1])
    packed = torch.empty(_lib().tg_packed_bytes(N, K), dtype=torch.uint8, device=w.device)
    _capi.check(_lib().tg_pack_weight(w.data_ptr(), N, K, ld, int(bool(transposed)), packed.data_ptr(),
                                      torch.cuda.current_stream(w.device).cuda_stream))
    if key is not None:
        _PACKS[key] = ((packed, N, K), w)
    return packed, N, K


def tg_linear(x, pack, bias=None, relu=False, out=None, accumulate=False, res=None):
    """act(x B^T + bias) (+ out | + res) with B packed by tg_pack; x [rows, K] fp32 row-major (row stride free).
    res: float32 [rows, N] with unit column stride, added in the epilogue (accumulate=True is res = out)."""
    packed, N, K = pack
    x, ldx = _row_major(x.detach())
    rows = x.shape[0]
    if x.shape[1] != K:
        raise ValueError("x has %d columns, the packed weight expects %d" % (x.shape[1], K))
    if out is None:
        out = torch.empty(rows, N, device=x.device)
    elif out.shape != (rows, N) or out.stride(1) != 1 or out.dtype != torch.float32:
        raise ValueError("out must be a float32 [rows, N] tensor with unit column stride")
    if rows == 0:
        return out
    st = _tg_state(x.device)
    if accumulate:
        res = out
    if res is not None and (res.shape != (rows, N) or res.stride(1) != 1 or res.dtype != torch.float32):
        raise ValueError("res must be a float32 [rows, N] tensor with unit column stride")
    _capi.check(_lib().tg_linear_res(x.data_ptr(), ldx, rows, K, packed.data_ptr(), N, None if bias is None else bias.detach().data_ptr(),
                                     int(bool(relu)), None if res is None else res.data_ptr(), 0 if res is None else res.stride(0),
                                     out.data_ptr(), out.stride(0), st["status"].data_ptr(),
                                     torch.cuda.current_stream(x.device).cuda_stream))
    return out


def tg_wgrad(x, y):
    """x^T y for tall operands: [rows, a]^T [rows, b] -> [a, b]."""
    x, ldx = _row_major(x.detach())
    y, ldy = _row_major(y.detach())
    a, b, rows = x.shape[1], y.shape[1], x.shape[0]
    out = torch.empty(a, b, device=x.device)
    if rows == 0:
        return out.zero_()
    st = _tg_state(x.device)
    need = _lib().tg_wgrad_scratch_bytes(a, b)
    if st["scratch"].numel() < need:
        st["scratch"] = torch.empty(need, dtype=torch.uint8, device=x.device)
    _capi.check(_lib().tg_wgrad(x.data_ptr(), ldx, a, y.data_ptr(), ldy, b, rows, out.data_ptr(), b, 0, st["scratch"].data_ptr(),
                                st["status"].data_ptr(), torch.cuda.current_stream(x.device).cuda_stream))
    return out


def _use_tg(x):
    return DENSE == "tcgen05" and x.is_cuda and x.dtype == torch.float32


# ---- dense layers of the training forward with a hand-written backward ------------------------------------------------
_COLSUM_SCRATCH = {}


def colsum(x):
    """x.sum(0) of a tall float32 CUDA matrix [rows, cols] (cols a power of two <= 256, unit column stride) in one launch with a
    fixed summation order (rl_colsum); anything else falls back to torch."""
    if not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.shape[0] > 0 and (x.shape[1] == 1 or x.stride(1) == 1)):
        return x.sum(0)
    rows, cols = x.shape
    L = _lib()
    blocks = L.rl_colsum_blocks(rows, cols)
    if blocks < 1:
        return x.sum(0)
    ld = x.stride(0) if rows > 1 else cols
    if ld < cols:
        return x.sum(0)
    need = 4 + 1024 * 256
    key = _stream_key(x.device)
    sc = _COLSUM_SCRATCH.get(key)
    if sc is None:
        sc = _COLSUM_SCRATCH[key] = torch.zeros(need, device=x.device)
    out = torch.empty(cols, device=x.device)
    _capi.check(L.rl_colsum(x.data_ptr(), rows, cols, ld, out.data_ptr(), sc.data_ptr(), torch.cuda.current_stream(x.device).cuda_stream))
    return out


def _relu_bwd_colsum_cuda(dout, out):
    """(dpre = dout * [out > 0], column sums of dpre) in one pass over [rows, cols] (rl_relu_bwd_colsum)."""
    L = _lib()
    rows, cols = out.shape
    blocks = L.rl_relu_bwd_colsum_blocks(rows, cols)
    strided = lambda t: t.stride(1) == 1 and t.stride(0) >= cols and t.stride(0) % 4 == 0 and t.data_ptr() % 16 == 0
    if blocks < 1:                                   # widths the kernel does not cover
        dpre = torch.ops.aten.threshold_backward(dout, out, 0)
        return dpre, dpre.sum(0)
    if not strided(dout):
        dout = dout.contiguous()
    if not strided(out):
        out = out.contiguous()
    dpre = torch.empty(rows, cols, device=out.device)
    partial = torch.empty(blocks, cols, device=out.device)
    _capi.check(L.rl_relu_bwd_colsum_ld(dout.data_ptr(), dout.stride(0), out.data_ptr(), out.stride(0), dpre.data_ptr(), cols,
                                        partial.data_ptr(), rows, cols, torch.cuda.current_stream(out.device).cuda_stream))
    return dpre, colsum(partial)


relu_bwd_colsum = _relu_bwd_colsum_cuda               # tests substitute a torch restatement on the CPU

# Test hook.  Two fp32-grade evaluations of the same layer can disagree on the SIGN of a pre-activation that is within
# rounding (~1e-7) of zero; the backward then masks one row's gradient differently and every upstream weight gradient
# moves by a whole row's term.  To compare the arithmetic of two dense back ends the tests pin the ReLU decisions:
# RELU_TRACE = {"mode": "record", "outs": []} collects the activations every ReLU backward masks with (in call order),
# RELU_TRACE = {"mode": "replay", "outs": [...], "pos": 0, "flips": 0} substitutes them and counts the disagreements.
RELU_TRACE = None


def _relu_mask_source(y):
    tr = RELU_TRACE
    if tr is None:
        return y
    if tr["mode"] == "record":
        tr["outs"].append(y)
        return y
    ref = tr["outs"][tr["pos"]]
    tr["pos"] += 1
    tr["flips"] += int(((ref > 0) != (y > 0)).sum())
    return ref

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
    if _use_tg(x):
        return tg_wgrad(x, dy)
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
        if _use_tg(x):
            y = tg_linear(x, tg_pack(W, False), b, relu)
        elif relu:
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
            dpre, db = relu_bwd_colsum(dy, _relu_mask_source(y))
        else:
            dpre = dy.contiguous()
            db = colsum(dpre)
        dW = xt_dy(dpre, x)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = tg_linear(dpre, tg_pack(W, True)) if _use_tg(dpre) else dpre @ W
        return dx, dW, db, None


class _Matmul(torch.autograd.Function):
    """y = x M with the split-K weight gradient."""

    @staticmethod
    def forward(ctx, x, M):
        ctx.save_for_backward(x, M)
        return tg_linear(x, tg_pack(M, True)) if _use_tg(x) else x @ M

    @staticmethod
    def backward(ctx, dy):
        x, M = ctx.saved_tensors
        dy = dy.contiguous()
        dx = None
        if ctx.needs_input_grad[0]:
            dx = tg_linear(dy, tg_pack(M, False)) if _use_tg(dy) else dy @ M.t()
        return dx, (xt_dy(x, dy) if ctx.needs_input_grad[1] else None)


class _MatmulNT(torch.autograd.Function):
    """y = x B^T for B [N, K] (the folds of the attention projections: W_query W_key^T, ... U_2^T)."""

    @staticmethod
    def forward(ctx, x, B):
        ctx.save_for_backward(x, B)
        return tg_linear(x, tg_pack(B, False)) if _use_tg(x) else x @ B.t()

    @staticmethod
    def backward(ctx, dy):
        x, B = ctx.saved_tensors
        dy = dy.contiguous()
        dx = dB = None
        if ctx.needs_input_grad[0]:
            dx = tg_linear(dy, tg_pack(B, True)) if _use_tg(dy) else dy @ B
        if ctx.needs_input_grad[1]:
            dB = xt_dy(dy, x)
        return dx, dB


def linear(x, W, b, relu=False):
    return _Linear.apply(x, W, b, relu)


def matmul(x, M):
    return _Matmul.apply(x, M)


def matmul_nt(x, B):
    return _MatmulNT.apply(x, B)


# ---- encoders + attention over the opponents + the [h0 | eOpp] feature block as one function ---------------------------------
def _cross_forward(kk, qv, n, m, norm):
    """kk [n*B, k] (keys of the own team), qv [m*B, 2k] (queries | values of the opponents) -> (e [n*B, k], attn [B, n, m])."""
    B, k = kk.shape[0] // n, kk.shape[1]
    e, attn = torch.empty(n * B, k, device=kk.device), torch.empty(B, n, m, device=kk.device)
    kk, qv = kk.contiguous(), qv.contiguous()
    _capi.check(_attn_lib().rl_attn_forward(_op(kk, 0, k, B), _op(qv, 0, 2 * k, B), _op(qv, k, 2 * k, B), _op(e, 0, k, B),
                                            attn.data_ptr(), B, n, m, k, float(norm), 0, torch.cuda.current_stream(kk.device).cuda_stream))
    return e, attn


def _cross_backward(de, kk, qv, attn, n, m, norm):
    B, k = kk.shape[0] // n, kk.shape[1]
    dkk, dqv = torch.empty_like(kk), torch.empty_like(qv)
    de = de.contiguous()
    _capi.check(_attn_lib().rl_attn_backward(_op(de, 0, k, B), _op(kk, 0, k, B), _op(qv, 0, 2 * k, B), _op(qv, k, 2 * k, B),
                                             attn.data_ptr(), _op(dkk, 0, k, B), _op(dqv, 0, 2 * k, B), _op(dqv, k, 2 * k, B),
                                             B, n, m, k, float(norm), torch.cuda.current_stream(kk.device).cuda_stream))
    return dkk, dqv


def _prod(x, B, transposed, bias=None, relu=False, out=None, res=None):
    """act(x B'^T + bias) (+ res), B' = B^T if transposed: tg_linear, or the same on torch's library GEMM when DENSE = "cublas"
    (the checker the tests compare the tcgen05 path against)."""
    if _use_tg(x):
        return tg_linear(x, tg_pack(B, transposed), bias, relu, out=out, res=res)
    y = x @ (B if transposed else B.t())
    if bias is not None:
        y = y + bias
    if relu:
        y = torch.relu(y)
    if res is not None:
        y = y + res
    if out is None:
        return y
    out.copy_(y)
    return out


class _FrontEnd(torch.autograd.Function):
    """h = [ReLU(encoder own) | (softmax(K(h0) Q(hOpp)^T / sqrt(dk)) V(hOpp)) W_out]   (mpnn.py:127-142, 376-443) on this repo's dense
    kernels.  The two halves of the feature block are written in place by the products that make them (no concatenation),
    backward reads the halves of dh in place, the encoder's gradient dh0 = dh[:, :d] + dK W_key^T comes out of ONE product with
    the first term added in its epilogue, and the ReLU masks are read from the stored block -- no slice copies, zero fills or
    gradient sums between the kernels."""

    @staticmethod
    def forward(ctx, inp, oppInp, We, be, Wo, bo, Wk, Wq, Wv, Wout, n, m, norm):
        d = We.shape[0]
        We, be, Wo, bo, Wk, Wout = (t.detach() for t in (We, be, Wo, bo, Wk, Wout))
        Wqv = torch.cat((Wq.detach(), Wv.detach()), dim=1)                          # [d, 2k]
        h = torch.empty(inp.shape[0], d + Wout.shape[1], device=inp.device, dtype=inp.dtype)
        _prod(inp, We, False, be, True, out=h[:, :d])
        hO = _prod(oppInp, Wo, False, bo, True)
        kk = _prod(h[:, :d], Wk, True)
        qv = _prod(hO, Wqv, True)
        e, attn = cross_forward(kk, qv, n, m, norm)
        _prod(e, Wout, True, out=h[:, d:])
        ctx.save_for_backward(inp, oppInp, h, hO, kk, qv, e, attn, Wk, Wqv, Wout)
        ctx.meta = (n, m, float(norm), d)
        ctx.mark_non_differentiable(attn)
        return h, attn

    @staticmethod
    def backward(ctx, dh, _dattn):
        inp, oppInp, h, hO, kk, qv, e, attn, Wk, Wqv, Wout = ctx.saved_tensors
        n, m, norm, d = ctx.meta
        k = Wk.shape[1]
        if dh.stride(1) != 1 or dh.stride(0) % 4 or dh.data_ptr() % 16:
            dh = dh.contiguous()
        dE, h0 = dh[:, d:], h[:, :d]
        de = _prod(dE, Wout, False)
        dWout = xt_dy(e, dE)
        dkk, dqv = cross_backward(de, kk, qv, attn, n, m, norm)
        dWk, dWqv = xt_dy(h0, dkk), xt_dy(hO, dqv)
        dh0 = _prod(dkk, Wk, False, res=dh[:, :d])                                  # dh[:, :d] + dK W_key^T
        dpre0, dbe = relu_bwd_colsum(dh0, _relu_mask_source(h0))
        dWe = xt_dy(dpre0, inp)
        dhO = _prod(dqv, Wqv, False)
        dpreO, dbo = relu_bwd_colsum(dhO, _relu_mask_source(hO))
        dWo = xt_dy(dpreO, oppInp)
        return None, None, dWe, dbe, dWo, dbo, dWk, dWqv[:, :k], dWqv[:, k:], dWout, None, None, None


def front_end(inp, oppInp, We, be, Wo, bo, Wk, Wq, Wv, Wout, n, m, norm):
    return _FrontEnd.apply(inp, oppInp, We, be, Wo, bo, Wk, Wq, Wv, Wout, n, m, norm)


# ---- value head + policy head + action head as one function of the final features -----------------------------------------
class _Heads(torch.autograd.Function):
    """(value, logits) = (value_head.2(ReLU(value_head.0 x)), dist.linear(ReLU(policy_head.0 x)))  (mpnn.py:174-205 with
    policy_layers = 1).  The two 128 -> 128 hidden layers read the same x: they run as ONE product with the weights stacked
    ([2d, d], ReLU in the epilogue), and backward as ONE K = 2d product for dx and ONE weight-gradient product, instead of two
    of each plus the sum of the two dx; the small heads read / write their halves of the stacked activation in place."""

    @staticmethod
    def forward(ctx, x, Wv0, bv0, Wp0, bp0, Wv2, bv2, Wd, bd):
        x = x.contiguous()
        d = Wv0.shape[0]
        Wcat, bcat = torch.cat((Wv0, Wp0), 0), torch.cat((bv0, bp0), 0)
        if _use_tg(x):
            hv = tg_linear(x, tg_pack(Wcat, False), bcat, True)
            value = tg_linear(hv[:, :d], tg_pack(Wv2, False), bv2)
            logits = tg_linear(hv[:, d:], tg_pack(Wd, False), bd)
        else:
            hv = torch.relu(torch.addmm(bcat, x, Wcat.t()))
            value, logits = torch.addmm(bv2, hv[:, :d], Wv2.t()), torch.addmm(bd, hv[:, d:], Wd.t())
        ctx.save_for_backward(x, hv, Wcat, Wv2, Wd)
        return value, logits

    @staticmethod
    def backward(ctx, dvalue, dlogits):
        x, hv, Wcat, Wv2, Wd = ctx.saved_tensors
        d = Wcat.shape[0] // 2
        dvalue, dlogits = dvalue.contiguous(), dlogits.contiguous()
        dhv = torch.empty_like(hv)
        if _use_tg(x):
            tg_linear(dvalue, tg_pack(Wv2, True), out=dhv[:, :d])
            tg_linear(dlogits, tg_pack(Wd, True), out=dhv[:, d:])
        else:
            dhv[:, :d] = dvalue @ Wv2
            dhv[:, d:] = dlogits @ Wd
        dWv2, dWd = xt_dy(dvalue, hv[:, :d]), xt_dy(dlogits, hv[:, d:])
        dpre, dbcat = relu_bwd_colsum(dhv, _relu_mask_source(hv))
        dWcat = xt_dy(dpre, x)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = tg_linear(dpre, tg_pack(Wcat, True)) if _use_tg(dpre) else dpre @ Wcat
        return dx, dWcat[:d], dbcat[:d], dWcat[d:], dbcat[d:], dWv2, colsum(dvalue), dWd, colsum(dlogits)


def heads(x, Wv0, bv0, Wp0, bp0, Wv2, bv2, Wd, bd):
    return _Heads.apply(x, Wv0, bv0, Wp0, bp0, Wv2, bv2, Wd, bd)


# ---- the [d, d] folds of the attention projections: a handful of small products per launch ------------------------------------
def small_matmul(items):
    """items: list of (C, A, B, trans_a, trans_b) with C = op(A) op(B), all float32 CUDA matrices (unit column stride, every
    dimension <= 256) that do not depend on each other: ONE launch (rl_small_matmul)."""
    arr = (_capi.RlSmallMatmul * len(items))()
    for k, (C, A, B, ta, tb) in enumerate(items):
        for t in (C, A, B):
            if not (t.is_cuda and t.dtype == torch.float32 and t.dim() == 2 and t.stride(1) == 1):
                raise ValueError("small_matmul: float32 CUDA matrices with unit column stride expected")
        M, N = C.shape
        K = A.shape[0] if ta else A.shape[1]
        if (A.shape[1] if ta else A.shape[0]) != M or (B.shape[1] if tb else B.shape[0]) != K or (B.shape[0] if tb else B.shape[1]) != N:
            raise ValueError("small_matmul: shapes do not chain")
        arr[k] = _capi.RlSmallMatmul(A.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, K, A.stride(0), B.stride(0), C.stride(0),
                                     int(ta), int(tb), 0)
    _capi.check(_lib().rl_small_matmul(arr, len(items), torch.cuda.current_stream(items[0][0].device).cuda_stream))


class _FoldWeights(torch.autograd.Function):
    """(Mqk, Wz) = (W_query W_key^T, W_val W_out U2^T): the projections of the message attention folded into two [d, d] weights
    (see _MessageRound).  Forward is two launches of small products, backward two more -- before, each of the three products and
    of their six backward products was a persistent tcgen05 launch with its own weight pack and partial-sum reduction."""

    @staticmethod
    def forward(ctx, Wq, Wk, Wv, Wout, U2):
        Wq, Wk, Wv, Wout, U2 = (t.detach() for t in (Wq, Wk, Wv, Wout, U2))
        d = Wq.shape[0]
        new = lambda *sh: torch.empty(*sh, device=Wq.device)
        Mqk, Wvo, Wz = new(d, Wk.shape[0]), new(Wv.shape[0], Wout.shape[1]), new(Wv.shape[0], U2.shape[0])
        small_matmul([(Mqk, Wq, Wk, False, True), (Wvo, Wv, Wout, False, False)])
        small_matmul([(Wz, Wvo, U2, False, True)])
        ctx.save_for_backward(Wq, Wk, Wv, Wout, U2, Wvo)
        return Mqk, Wz

    @staticmethod
    def backward(ctx, dMqk, dWz):
        Wq, Wk, Wv, Wout, U2, Wvo = ctx.saved_tensors
        dMqk, dWz = dMqk.contiguous(), dWz.contiguous()
        like = torch.empty_like
        dWq, dWk, dWv, dWout, dU2, dWvo = like(Wq), like(Wk), like(Wv), like(Wout), torch.empty(U2.shape, device=U2.device), like(Wvo)
        small_matmul([(dWvo, dWz, U2, False, False), (dU2, dWz, Wvo, True, False), (dWq, dMqk, Wk, False, False),
                      (dWk, dMqk, Wq, True, False)])
        small_matmul([(dWv, dWvo, Wout, False, True), (dWout, Wv, dWvo, True, False)])
        return dWq, dWk, dWv, dWout, dU2


def fold_weights(Wq, Wk, Wv, Wout, U2):
    """-> (W_query W_key^T, W_val W_out U2^T) with autograd; one small-product launch per dependency level on CUDA float32,
    plain torch otherwise."""
    if Wq.is_cuda and Wq.dtype == torch.float32 and max(Wq.shape + Wk.shape + Wv.shape + Wout.shape + U2.shape) <= 256:
        return _FoldWeights.apply(Wq, Wk, Wv, Wout, U2)
    return Wq @ Wk.t(), (Wv @ Wout) @ U2.t()


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
cross_forward, cross_backward = _cross_forward, _cross_backward          # likewise for the attention over the opponents


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
        tg = _use_tg(h)
        g = tg_linear(h, tg_pack(Mqk, True)) if tg else h @ Mqk
        hm, attn = mix_forward(g, h, n, norm)
        if tg:
            out = tg_linear(hm, tg_pack(Wc, True), bias, True)
        else:
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
        dpre, dbias = relu_bwd_colsum(dout, _relu_mask_source(out))
        tg = _use_tg(dpre)
        dWc = xt_dy(hm, dpre)
        dhm = tg_linear(dpre, tg_pack(Wc, False)) if tg else dpre @ Wc.t()        # [dh through U1 | dmixed]
        dg, dh = mix_backward(dhm, g, hm, attn, n, norm)
        dMqk = xt_dy(hm[:, :k], dg)
        if tg:
            tg_linear(dg, tg_pack(Mqk, False), out=dh, accumulate=True)
        else:
            dh.addmm_(dg, Mqk.t())
        return dh, dMqk, dWc, dbias, None, None


def message_round(h, Mqk, Wc, bias, n, norm):
    return _MessageRound.apply(h, Mqk, Wc, bias, n, norm)


# ---- optimizer step: gradient-norm clip + Adam for all parameters in two launches --------------------------------------
class TgAdam(object):
    """nn.utils.clip_grad_norm_(params, max_norm) followed by optim.Adam(params, lr).step() (rlcore/algo/ppo.py:114,191-192)
    as tg_adam_step (csrc/tg_gemm.cu): same arithmetic as torch's, step counter on the device (graph-capturable)."""

    def __init__(self, params, lr, betas=(0.9, 0.999), eps=1e-8):
        self.params = [p for p in params]
        if not self.params or not all(p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() for p in self.params):
            raise ValueError("TgAdam needs contiguous float32 CUDA parameters")
        if len(self.params) > 32:
            raise ValueError("TgAdam handles up to 32 parameter tensors (TG_MAX_TENSORS)")
        dev = self.params[0].device
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.exp_avg = [torch.zeros_like(p) for p in self.params]
        self.exp_avg_sq = [torch.zeros_like(p) for p in self.params]
        self.step_count = torch.zeros(1, dtype=torch.int64, device=dev)
        self.scratch = torch.zeros(256, device=dev)
        self.total_norm = torch.zeros(1, device=dev)
        self.param_groups = [{"params": self.params, "lr": self.lr, "betas": self.betas, "eps": self.eps}]

    def zero_grad(self, set_to_none=True):
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    def step(self, max_norm=0.0, grad_scale=None, grads=None):
        """grad_scale: optional device scalar tensor every gradient is multiplied by first (in place).
        grads: optional list (one per parameter, None = skip) to read the gradients from instead of p.grad -- the views of
        a flat all-reduced buffer in a multi-rank step."""
        arr = (_capi.TgTensor * len(self.params))()
        for k, p in enumerate(self.params):
            g = p.grad if grads is None else grads[k]
            if g is not None and (not g.is_contiguous() or g.dtype != torch.float32):
                if grads is not None:
                    raise ValueError("explicit gradients must be contiguous float32")
                g = p.grad = g.contiguous().float()
            arr[k] = _capi.TgTensor(p.data_ptr(), None if g is None else g.data_ptr(), self.exp_avg[k].data_ptr(),
                                    self.exp_avg_sq[k].data_ptr(), p.numel())
        dev = self.params[0].device
        _capi.check(_lib().tg_adam_step(arr, len(self.params), self.param_groups[0]["lr"], self.betas[0], self.betas[1], self.eps,
                                        float(max_norm or 0.0), None if grad_scale is None else grad_scale.data_ptr(),
                                        self.step_count.data_ptr(), self.scratch.data_ptr(), self.total_norm.data_ptr(),
                                        torch.cuda.current_stream(dev).cuda_stream))
