"""Team policy / value network of the reference (mpnn.py:17-443), written for batched rollouts.

Same parameters, names and shapes as the reference's MPNN so that its checkpoints load unchanged
(`state_dict` keys: encoder.0, oppEncoder.0, oppAttn.W_{query,key,val,out}, oppUpdate.0 (present, never
used: mpnn.py:44-45), messages.W_{query,key,val,out}, update.0, value_head.{0,2}, policy_head.0,
dist.linear), same call surface (`act`, `evaluate_actions`, `get_value`, `attn_mat`, `opp_attn_mat`,
`is_recurrent`), same arithmetic:

    h    = ReLU(encoder(own))            own rows are agent-major [n*B, d]       mpnn.py:127,132
    hOpp = ReLU(oppEncoder(opp))                                                 mpnn.py:128,133
    eOpp = softmax(K(h) Q(hOpp)^T / sqrt(dk)) V(hOpp) W_out   over opponents     mpnn.py:376-443
    h    = [h, eOpp]; 3 x { m = selfattn(h) without the diagonal ; h = ReLU(update([h, m])) }  :142-159
    value = value_head(h) ; logits = dist.linear(policy_head(h))                 mpnn.py:174-205

Differences in form, not in results: one fused QKV product per attention instead of three, the
diagonal mask is a constant tensor instead of a Python loop (mpnn.py:297-298), and the attention
matrices are kept on the device and only copied to the host when `attn_mat` / `opp_attn_mat` is read
(the reference synchronises the device twice per forward, mpnn.py:140,166).
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

try:
    from .rlcore.distributions import Categorical
except ImportError:          # imported top-level (drop-in mode: this directory is on sys.path)
    from rlcore.distributions import Categorical


def weights_init(m):
    # orthogonal Linear weights, zero bias (mpnn.py:9-14) -- also overrides Categorical's gain
    if isinstance(m, nn.Linear):
        nn.init.orthogonal_(m.weight.data)
        if m.bias is not None:
            m.bias.data.fill_(0)


class _Attention(nn.Module):
    """Parameter container shared by both attention blocks (mpnn.py:208-247, 334-374)."""

    def __init__(self, n_heads, input_dim, embed_dim):
        super().__init__()
        if n_heads != 1:
            raise NotImplementedError("the reference only ever builds single-head attention (mpnn.py:42,47)")
        self.n_heads, self.input_dim, self.embed_dim = n_heads, input_dim, embed_dim
        self.key_dim = self.val_dim = embed_dim // n_heads
        self.norm_factor = 1.0 / math.sqrt(self.key_dim)
        self.W_query = nn.Parameter(torch.empty(n_heads, input_dim, self.key_dim))
        self.W_key = nn.Parameter(torch.empty(n_heads, input_dim, self.key_dim))
        self.W_val = nn.Parameter(torch.empty(n_heads, input_dim, self.val_dim))
        self.W_out = nn.Parameter(torch.empty(n_heads, self.key_dim, embed_dim))
        for p in self.parameters():                       # U(-1/sqrt(d), 1/sqrt(d)) (mpnn.py:243-247)
            p.data.uniform_(-1.0 / math.sqrt(p.size(-1)), 1.0 / math.sqrt(p.size(-1)))


class MultiHeadAttention(_Attention):
    """Intra-team attention without self-messages (mpnn.py:249-331). q: [B, n, d] -> ([B, n, e], [1,B,n,n])."""

    def forward(self, q, h=None, mask=None, return_attn=False):
        if h is not None or mask is not None:
            raise NotImplementedError("the FortAttack path calls messages(h, mask=None) only (mpnn.py:157)")
        B, n, d = q.shape
        if n == 1:                                        # a lone agent receives a zero message (mpnn.py:262-270)
            out, attn = q.new_zeros(B, 1, self.embed_dim), q.new_zeros(1, B, 1, 1)
            return (out, attn) if return_attn else out
        w = torch.cat((self.W_query[0], self.W_key[0], self.W_val[0]), dim=1)       # [d, 3k]
        qkv = q.reshape(B * n, d) @ w
        Q, K, V = qkv.view(B, n, 3, -1).unbind(2)
        comp = self.norm_factor * torch.matmul(Q, K.transpose(1, 2))                # [B, n, n]
        comp = comp.masked_fill(torch.eye(n, dtype=torch.bool, device=q.device), -math.inf)
        attn = F.softmax(comp, dim=-1)
        out = (torch.matmul(attn, V).reshape(B * n, -1) @ self.W_out[0]).view(B, n, self.embed_dim)
        return (out, attn.unsqueeze(0)) if return_attn else out


class MultiHeadOppAttention(_Attention):
    """Attention of each team member over the opponents; note the reference's naming: keys come from the
    own team, queries and values from the opponents (mpnn.py:409-417)."""

    def __init__(self, n_heads, input_dim, opp_input_dim, embed_dim):
        super().__init__(n_heads, input_dim, embed_dim)
        self.opp_input_dim = opp_input_dim

    def forward(self, h, hOpp, mask=None, return_attn=False):
        B, n, d = h.shape
        m = hOpp.shape[1]
        qv = hOpp.reshape(B * m, -1) @ torch.cat((self.W_query[0], self.W_val[0]), dim=1)
        Q, V = qv.view(B, m, 2, -1).unbind(2)
        K = (h.reshape(B * n, d) @ self.W_key[0]).view(B, n, -1)
        attn = F.softmax(self.norm_factor * torch.matmul(K, Q.transpose(1, 2)), dim=-1)   # [B, n, m]
        out = (torch.matmul(attn, V).reshape(B * n, -1) @ self.W_out[0]).view(B, n, self.embed_dim)
        return (out, attn) if return_attn else out


class MPNN(nn.Module):
    def __init__(self, action_space, num_agents, num_opp_agents, num_entities=0, input_size=16, hidden_dim=128,
                 embed_dim=None, pos_index=2, norm_in=False, nonlin=nn.ReLU, n_heads=1, mask_dist=None,
                 entity_mp=False, policy_layers=1):
        super().__init__()
        if entity_mp or norm_in:
            raise NotImplementedError("entity message passing / input batch-norm are not on the FortAttack path "
                                      "(learner.py:36-41 passes entity_mp=False; norm_in defaults to False)")
        self.h_dim, self.nonlin = hidden_dim, nonlin
        self.num_agents, self.num_opp_agents, self.num_entities = num_agents, num_opp_agents, num_entities
        self.K = 3                                                                 # message passing rounds
        self.embed_dim = hidden_dim if embed_dim is None else embed_dim
        self.n_heads, self.mask_dist, self.input_size = n_heads, mask_dist, input_size
        self.entity_mp, self.policy_layers, self.pos_index = entity_mp, policy_layers, pos_index
        half = hidden_dim // 2
        self.encoder = nn.Sequential(nn.Linear(input_size, half), nonlin(inplace=True))
        self.oppEncoder = nn.Sequential(nn.Linear(input_size, half), nonlin(inplace=True))
        self.oppAttn = MultiHeadOppAttention(n_heads, half, half, self.embed_dim // 2)
        self.oppUpdate = nn.Sequential(nn.Linear(half + self.embed_dim // 2, half), nonlin(inplace=True))
        self.messages = MultiHeadAttention(n_heads, hidden_dim, self.embed_dim)
        self.update = nn.Sequential(nn.Linear(hidden_dim + self.embed_dim, hidden_dim), nonlin(inplace=True))
        self.value_head = nn.Sequential(nn.Linear(hidden_dim, hidden_dim), nonlin(inplace=True),
                                        nn.Linear(hidden_dim, 1))
        if policy_layers == 1:
            self.policy_head = nn.Sequential(nn.Linear(hidden_dim, hidden_dim), nonlin(inplace=True))
        elif policy_layers == 2:
            self.policy_head = nn.Sequential(nn.Linear(hidden_dim, hidden_dim), nonlin(inplace=True),
                                             nn.Linear(hidden_dim, hidden_dim))
        else:
            raise ValueError("policy_layers must be 1 or 2")
        self.dist = Categorical(hidden_dim, action_space.shape[0])                 # mpnn.py:73-74
        self.is_recurrent = False
        self.in_fn = lambda x: x
        self.apply(weights_init)
        self._attn = self._opp_attn = None
        self.dropout_mask = self.dead_mask = None

    # attention matrices of the last forward, as the reference exposes them (numpy, first env only
    # when batched: `.squeeze(0).squeeze(0)` of [1,B,n,n], mpnn.py:140,166); copied lazily
    @property
    def attn_mat(self):
        if self._attn is None:
            import numpy as np
            return np.ones((self.num_agents, self.num_agents))                     # mpnn.py:86
        return self._attn.squeeze(0).squeeze(0).detach().cpu().numpy()

    @property
    def opp_attn_mat(self):
        return None if self._opp_attn is None else self._opp_attn.squeeze(0).squeeze(0).detach().cpu().numpy()

    # Training-time forward on CUDA (rlcore/fused.py, csrc/rl_kernels.cu): same function as _fwd, but everything stays in the
    # agent-major row layout (no transposes), the [batch, n, n] bmm -> mask -> softmax -> bmm chains are one kernel each
    # way, every dense layer has a hand-written backward (ReLU-backward + bias gradient in one kernel, split-K weight
    # gradients), and -- with fold_projections -- the Q/K/V/out projections of the message rounds are folded into
    # [d, d] products of the weights, so a round is G = h Mqk -> rl_attn_mix -> ONE K = 2d GEMM with bias + ReLU in its
    # epilogue.  Opt-in: `fused_attention = True` (BatchedTrainer(fused_update=True) sets it).
    fused_attention = False
    fused_no_grad = False       # also take the fused path under torch.no_grad() (BatchedTrainer.recompute_old: same arithmetic as the update)
    fold_projections = True     # within the fused path: message rounds with the attention projections folded (see _fwd_fused)
    front_end_fused = True      # within the fused path: encoders + opponent attention + feature block as rlcore/fused.front_end

    def _fwd_fused(self, inp, oppInp):
        try:
            from .rlcore import fused
        except ImportError:
            from rlcore import fused
        n, m = self.num_agents, self.num_opp_agents
        oa, ms = self.oppAttn, self.messages
        if self.front_end_fused:
            # encoders, opponent attention and the [h0 | eOpp] block as one function with a hand-written backward
            h, oattn = fused.front_end(inp, oppInp, self.encoder[0].weight, self.encoder[0].bias, self.oppEncoder[0].weight,
                                       self.oppEncoder[0].bias, oa.W_key[0], oa.W_query[0], oa.W_val[0], oa.W_out[0], n, m,
                                       oa.norm_factor)
        else:
            h0 = fused.linear(inp, self.encoder[0].weight, self.encoder[0].bias, True)  # [n*B, 64] agent-major rows
            hO = fused.linear(oppInp, self.oppEncoder[0].weight, self.oppEncoder[0].bias, True)   # [m*B, 64]
            e, oattn = fused.cross_attention(fused.matmul(h0, oa.W_key[0]),
                                             fused.matmul(hO, torch.cat((oa.W_query[0], oa.W_val[0]), dim=1)), n, m, oa.norm_factor)
            h = torch.cat((h0, fused.matmul(e, oa.W_out[0])), dim=1)               # [n*B, 128]
        W, bias = self.update[0].weight, self.update[0].bias
        U1t, U2t = W[:, :self.h_dim].t(), W[:, self.h_dim:].t()
        attn = None
        if self.fold_projections and n > 1 and self.h_dim <= 128:
            # Q/K/V/out projections folded into [d, d] products of the weights (rlcore/fused.py _MessageRound)
            # the [d, d] folds go through the same dense functions (own kernels, hand-written backward): W_query ... W_out
            # and update.0.weight receive their exact gradients through them
            Mqk, Wz = fused.fold_weights(ms.W_query[0], ms.W_key[0], ms.W_val[0], ms.W_out[0], W[:, self.h_dim:])
            Wc = torch.cat((U1t, Wz), dim=0)                                                    # [U1^T ; W_val W_out U2^T]
            for _ in range(self.K):
                h, attn = fused.message_round(h, Mqk, Wc, bias, n, ms.norm_factor)
            self._opp_attn = oattn
            self._attn = attn.unsqueeze(0)
            return h
        wqkv = torch.cat((ms.W_query[0], ms.W_key[0], ms.W_val[0]), dim=1)
        for _ in range(self.K):
            if n > 1:
                msg, attn = fused.self_attention(h @ wqkv, n, ms.norm_factor)
                h = torch.relu(torch.addmm(torch.addmm(bias, h, U1t), msg @ ms.W_out[0], U2t))
            else:                                                                  # a lone agent receives a zero message
                h = fused.linear(h, W[:, :self.h_dim], bias, True)
        self._opp_attn = oattn
        # a lone agent: the reference reports a zero attention matrix (mpnn.py:262-270)
        self._attn = attn.unsqueeze(0) if attn is not None else h.new_zeros(1, h.shape[0] // n, 1, 1)
        return h

    def _fwd(self, inp, oppInp, masks=None):
        if (self.fused_attention and inp.is_cuda and inp.dtype == torch.float32 and self.h_dim % 64 == 0
                and self.nonlin is nn.ReLU and self.h_dim <= 256 and self.num_agents <= 5 and self.num_opp_agents <= 5):
            return self._fwd_fused(inp, oppInp)
        n, m, half = self.num_agents, self.num_opp_agents, self.h_dim // 2
        h = self.encoder(inp).view(n, -1, half).transpose(0, 1)                    # [B, n, 64]
        hOpp = self.oppEncoder(oppInp).view(m, -1, half).transpose(0, 1)           # [B, m, 64]
        eOpp, self._opp_attn = self.oppAttn(h, hOpp, masks, return_attn=True)
        h = torch.cat((h, eOpp), dim=2)                                            # [B, n, 128]
        attn = None
        for _ in range(self.K):
            msg, attn = self.messages(h, return_attn=True)
            h = self.update(torch.cat((h, msg), dim=2))
        self._attn = attn
        return h.transpose(0, 1).reshape(-1, self.h_dim)                           # agent-major rows again

    def forward(self, inp, state, mask=None):
        raise NotImplementedError

    def _use_fused(self, x):
        return (self.fused_attention and x.is_cuda and x.dtype == torch.float32
                and (torch.is_grad_enabled() or self.fused_no_grad) and self.nonlin is nn.ReLU)

    def _dist(self, p):
        if self._use_fused(p):
            try:
                from .rlcore import fused
                from .rlcore.distributions import FixedCategorical
            except ImportError:
                from rlcore import fused
                from rlcore.distributions import FixedCategorical
            logits = fused.linear(p, self.dist.linear.weight, self.dist.linear.bias)
            return FixedCategorical(logits=logits, validate_args=False if (p.is_cuda and torch.cuda.is_current_stream_capturing()) else None)
        return self.dist(p)

    def _value(self, x):
        if self._use_fused(x):               # the same layers through rlcore/fused.py's dense functions (training backward)
            try:
                from .rlcore import fused
            except ImportError:
                from rlcore import fused
            v0, v2 = self.value_head[0], self.value_head[2]
            return fused.linear(fused.linear(x, v0.weight, v0.bias, True), v2.weight, v2.bias)
        return self.value_head(x)

    def _policy(self, x):
        if self._use_fused(x) and self.policy_layers == 1:
            try:
                from .rlcore import fused
            except ImportError:
                from rlcore import fused
            return fused.linear(x, self.policy_head[0].weight, self.policy_head[0].bias, True)
        return self.policy_head(x)

    def act(self, inp, state, oppInp, mask=None, deterministic=False):
        x = self._fwd(inp, oppInp, mask)
        value = self._value(x)
        dist = self.dist(self._policy(x))
        action = dist.mode() if deterministic else dist.sample()
        return value, action, dist.log_probs(action).view(-1, 1), state

    def evaluate_actions(self, inp, state, oppInp, mask, action):
        x = self._fwd(inp, oppInp, mask)
        if self._logits_path(x):
            value, logits = self._heads(x)
            if not torch.is_grad_enabled():
                # no autograd wanted (BatchedTrainer.recompute_old): log-prob and entropy in the arithmetic of the update's loss kernel
                lp, ent = self._fused_module().categorical_eval(logits, action)
                return value, lp.view(-1, 1), ent, state
            try:
                from .rlcore.distributions import FixedCategorical
            except ImportError:
                from rlcore.distributions import FixedCategorical
            dist = FixedCategorical(logits=logits, validate_args=False if torch.cuda.is_current_stream_capturing() else None)
            return value, dist.log_probs(action), dist.entropy(), state
        value = self._value(x)
        dist = self._dist(self._policy(x))
        return value, dist.log_probs(action), dist.entropy(), state

    @staticmethod
    def _fused_module():
        try:
            from .rlcore import fused
        except ImportError:
            from rlcore import fused
        return fused

    def _logits_path(self, x):
        return (self._use_fused(x) and self.policy_layers == 1
                and self.dist.linear.out_features == self._fused_module().LOSS_ACTIONS)

    def _heads(self, x):
        """(value [N, 1], logits [N, 8]) through rlcore/fused.heads: the two hidden layers of the heads as one stacked product."""
        v0, v2, p0, dl = self.value_head[0], self.value_head[2], self.policy_head[0], self.dist.linear
        return self._fused_module().heads(x, v0.weight, v0.bias, p0.weight, p0.bias, v2.weight, v2.bias, dl.weight, dl.bias)

    def evaluate_logits(self, inp, oppInp):
        """(value [N, 1], action-head logits [N, 8]) of the fused training forward, or None when that path does not apply:
        JointPPO feeds them to rlcore/fused.ppo_loss_logits, which holds the Categorical (mpnn.py:199-200) and its backward."""
        if not (self.fused_attention and inp.is_cuda and inp.dtype == torch.float32 and self.nonlin is nn.ReLU
                and self.policy_layers == 1 and self.dist.linear.out_features == self._fused_module().LOSS_ACTIONS):
            return None
        return self._heads(self._fwd(inp, oppInp, None))

    def get_value(self, inp, state, oppInp, mask):
        return self._value(self._fwd(inp, oppInp, mask))
