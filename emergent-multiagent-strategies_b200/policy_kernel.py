"""Host side of the fused MPNN rollout forward (csrc/mp_policy.cu, include/fortattack_policy.h).

`pack_mpnn(module)` turns an MPNN's parameters (the reference's state_dict names, mpnn.py:25-84) into
the one packed blob the kernel streams; `FusedPolicy` wraps an MPNN and offers the rollout half of its
call surface -- `act`, `get_value` (mpnn.py:180-205) -- on device tensors in the step kernel's
agent-major layout.  Training (`evaluate_actions` with autograd) stays on the module itself; call
`refresh()` after the optimizer changed the weights.

There is no fallback: without the CUDA library these calls raise.
"""
import ctypes

import torch

from . import _capi

HIDDEN, OBS_DIM, ACTIONS = 128, 6, 8
BLOB_F16_BYTES, BLOB_CONST_FLOATS = 180224, 2576
MODE_SAMPLE, MODE_ARGMAX, MODE_EVAL = 0, 1, 2


def _canonical(w):
    """[N][K] -> UMMA canonical K-major core-matrix order [N/8][K/8][8][8] (csrc/mp_umma.cuh)."""
    n, k = w.shape
    return w.reshape(n // 8, 8, k // 8, 8).permute(0, 2, 1, 3).contiguous().reshape(-1)


def check_supported(m):
    ok = (m.h_dim == HIDDEN and m.embed_dim == HIDDEN and m.input_size == OBS_DIM and m.K == 3 and m.n_heads == 1
          and not m.entity_mp and m.policy_layers == 1 and m.dist.linear.out_features == ACTIONS
          and 1 <= m.num_agents <= 5 and 1 <= m.num_opp_agents <= 5)
    if not ok:
        raise _capi.FaError("the fused policy kernel covers the reference's FortAttack configuration only: hidden 128, "
                            "6-float observations, 8 actions, 1 head, policy_layers=1, teams of 1..5 (learner.py:57-69)")


@torch.no_grad()
def pack_mpnn(m, device=None):
    """MPNN parameters -> uint8 blob of MP_BLOB_BYTES (layout: csrc/mp_policy.cu header)."""
    check_supported(m)
    device = device or next(m.parameters()).device
    d = lambda t: t.detach().to(device=device, dtype=torch.float64)
    oa, ms = m.oppAttn, m.messages
    # attention parameters are stored [in][out] (mpnn.py:229-232); B operands are [N out][K in] = W^T for y = x W.
    # Folding (float64): scores (h0 Wkey)(hOpp Wquery)^T = (h0 [Wkey Wquery^T]) . hOpp ; eOpp = sum p (hOpp [Wval Wout])
    f16 = [_canonical((d(oa.W_key[0]) @ d(oa.W_query[0]).t()).t().contiguous()),
           _canonical((d(oa.W_val[0]) @ d(oa.W_out[0])).t().contiguous())]
    U = d(m.update[0].weight)                          # [128 out][256 in] = [U1 | U2] over cat(h, msg)
    U1, U2 = U[:, :HIDDEN].contiguous(), U[:, HIDDEN:]
    G = d(ms.W_query[0]) @ d(ms.W_key[0]).t()          # (h_a Wq)(h_b Wk)^T = (h_a G) . h_b
    Wz = d(ms.W_val[0]) @ d(ms.W_out[0]) @ U2.t()      # msg U2^T = sum_b p_b (h_b Wv Wout U2^T)
    f16 += [_canonical(G.t().contiguous()), _canonical(Wz.t().contiguous()), _canonical(U1)]
    for w in (m.value_head[0].weight, m.policy_head[0].weight):
        w = d(w)
        f16 += [_canonical(w[:64]), _canonical(w[64:])]
    f16 = torch.cat(f16).to(torch.float16)
    assert f16.numel() * 2 == BLOB_F16_BYTES, f16.numel()

    c = torch.zeros(BLOB_CONST_FLOATS, dtype=torch.float64, device=device)

    def enc(lin, off):
        blk = torch.zeros(64, 8, dtype=torch.float64, device=device)
        blk[:, :6] = d(lin.weight)
        blk[:, 6] = d(lin.bias)
        c[off:off + 512] = blk.reshape(-1)
    enc(m.encoder[0], 0)
    enc(m.oppEncoder[0], 512)
    c[1024:1152] = d(m.update[0].bias)
    c[1152:1280] = d(m.value_head[0].bias)
    c[1280:1408] = d(m.value_head[2].weight).reshape(-1)
    c[1408:1536] = d(m.policy_head[0].bias)
    c[1536:2560] = d(m.dist.linear.weight).t().contiguous().reshape(-1)      # [128][8]
    c[2560:2568] = d(m.dist.linear.bias)
    c[2568] = d(m.value_head[2].bias)[0]
    blob = torch.cat((f16.view(torch.uint8), c.to(torch.float32).view(torch.uint8)))
    return blob.contiguous()


_OUT_DTYPES = {"value": torch.float32, "action": torch.int64, "action_i32": torch.int32, "logp": torch.float32,
               "entropy": torch.float32, "logits": torch.float32}


def _check_outputs(out, n, E, dev):
    """The kernel writes through raw pointers: every output must have the element type, device and size it assumes."""
    for key, t in out.items():
        if key not in _OUT_DTYPES:
            raise ValueError("unknown output %r (expected one of %s)" % (key, sorted(_OUT_DTYPES)))
        if t.dtype != _OUT_DTYPES[key] or t.device != dev:
            raise ValueError("output %r must be %s on %s, got %s on %s" % (key, _OUT_DTYPES[key], dev, t.dtype, t.device))
        if not t.is_contiguous() or t.numel() != n * E * (ACTIONS if key == "logits" else 1):
            raise ValueError("output %r must be contiguous with %d rows" % (key, n * E))


def _check_env_lists(env_order, env_offsets, E, K, dev):
    if (env_order.dtype != torch.int32 or env_offsets.dtype != torch.int32 or env_order.numel() != E
            or (K is not None and env_offsets.numel() != K + 1) or env_offsets.numel() < 2
            or not env_order.is_contiguous() or not env_offsets.is_contiguous()
            or env_order.device != dev or env_offsets.device != dev):
        raise ValueError("env_order (int32 [E]) and env_offsets (int32 [K+1]) must be contiguous tensors on %s" % (dev,))


class FusedPolicy(object):
    """Rollout-time forward of one team's MPNN in a single kernel launch."""

    def __init__(self, module, seed=0, env_id0=0):
        check_supported(module)
        self.module = module
        self.device = next(module.parameters()).device
        if self.device.type != "cuda":
            raise _capi.FaError("FusedPolicy needs the module on a CUDA device; there is no CPU path")
        self._lib = _capi.lib()
        self.n, self.m = module.num_agents, module.num_opp_agents
        self.seed, self.env_id0, self.calls = int(seed), int(env_id0), 0
        self.status = torch.zeros(1, dtype=torch.int32, device=self.device)
        # device-resident call counter {count, ticket}: the kernel advances it, so CUDA-graph replays stay fresh
        self.counter = torch.zeros(2, dtype=torch.int64, device=self.device)
        self.launches = 0
        self.refresh()

    def refresh(self):
        """Re-pack the weights (after an optimizer step / load_state_dict) IN PLACE: launches captured in a CUDA
        graph keep pointing at the same blob."""
        blob = pack_mpnn(self.module, self.device)
        if getattr(self, "blob", None) is None:
            self.blob = blob
        else:
            self.blob.copy_(blob)

    def _ptr(self, t):
        return None if t is None else t.data_ptr()

    def forward(self, own, opp, mode=MODE_SAMPLE, action_in=None, out=None, want_logits=False, want_entropy=False,
                env_sel=None, sel_value=0, env_order=None, env_offsets=None):
        """own float32 [n, E, 6], opp float32 [m, E, 6] (contiguous, agent-major).
        Returns dict(value [n,E], action int64 [n,E], action_i32 [n,E], logp [n,E], entropy?, logits?).
        `out` may hold preallocated tensors under the same keys (e.g. views into the rollout storage).
        env_sel (int32 [E]) / sel_value: write outputs only for environments with env_sel[e] == sel_value.
        env_order (int32 [E]) / env_offsets (int32 [K+1]): compacted form -- process only the environments
        env_order[env_offsets[sel_value] : env_offsets[sel_value + 1]] (device tensors; no host synchronisation)."""
        n, m = self.n, self.m
        E = own.shape[1]
        if own.shape != (n, E, OBS_DIM) or opp.shape != (m, E, OBS_DIM):
            raise ValueError("own/opp must be [%d,E,6] / [%d,E,6], got %s / %s" % (n, m, tuple(own.shape), tuple(opp.shape)))
        if own.dtype != torch.float32 or opp.dtype != torch.float32 or not own.is_contiguous() or not opp.is_contiguous():
            raise ValueError("observations must be contiguous float32")
        dev = self.device
        if own.device != dev or opp.device != dev:
            raise ValueError("observations must live on %s (the kernel reads raw device pointers)" % (dev,))
        out = dict(out or {})
        for key, dt, shape in (("value", torch.float32, (n, E)), ("action", torch.int64, (n, E)),
                               ("action_i32", torch.int32, (n, E)), ("logp", torch.float32, (n, E))):
            if key not in out:
                out[key] = torch.empty(shape, dtype=dt, device=dev)
        if want_entropy and "entropy" not in out:
            out["entropy"] = torch.empty((n, E), dtype=torch.float32, device=dev)
        if want_logits and "logits" not in out:
            out["logits"] = torch.empty((n, E, ACTIONS), dtype=torch.float32, device=dev)
        _check_outputs(out, n, E, dev)
        if mode == MODE_EVAL:
            if action_in is None or action_in.numel() != n * E:
                raise ValueError("MODE_EVAL needs action_in with one action per (agent, env) row (%d)" % (n * E))
            action_in = action_in.to(device=dev, dtype=torch.int64).contiguous()
        if env_sel is not None and (env_sel.dtype != torch.int32 or env_sel.numel() != E or not env_sel.is_contiguous()
                                    or env_sel.device != dev):
            raise ValueError("env_sel must be a contiguous int32 tensor on %s with one entry per environment" % (dev,))
        if (env_order is None) != (env_offsets is None):
            raise ValueError("env_order and env_offsets come together")
        if env_order is not None:
            _check_env_lists(env_order, env_offsets, E, None, dev)
            if not 0 <= int(sel_value) < env_offsets.numel() - 1:
                raise ValueError("sel_value %d outside the %d groups of env_offsets" % (sel_value, env_offsets.numel() - 1))
        stream = torch.cuda.current_stream(dev).cuda_stream
        _capi.check(self._lib.mp_forward(self.blob.data_ptr(), own.data_ptr(), opp.data_ptr(), n, m, E, mode,
                                         self.seed, self.calls, self.counter.data_ptr(), self.env_id0, self._ptr(action_in),
                                         out["value"].data_ptr(), out["action"].data_ptr(), out["action_i32"].data_ptr(),
                                         out["logp"].data_ptr(), self._ptr(out.get("entropy")), self._ptr(out.get("logits")),
                                         self._ptr(env_sel), int(sel_value), self._ptr(env_order), self._ptr(env_offsets),
                                         self.status.data_ptr(), stream))
        self.calls += 1
        self.launches += 1
        return out

    def check_status(self):
        """Synchronising: raises if any launch so far reported an internal pipeline timeout."""
        code = int(self.status.item())
        if code:
            raise _capi.FaError("mp_policy_kernel pipeline timeout (wait site %d)" % code)

    # -- MPNN's rollout call surface on flat agent-major rows [n*E, 6] (mpnn.py:180-205) ------------------
    def act(self, inp, state, oppInp, mask=None, deterministic=False):
        E = inp.shape[0] // self.n
        o = self.forward(inp.view(self.n, E, OBS_DIM), oppInp.view(self.m, E, OBS_DIM),
                         MODE_ARGMAX if deterministic else MODE_SAMPLE)
        return o["value"].view(-1, 1), o["action"].view(-1, 1), o["logp"].view(-1, 1), state

    def get_value(self, inp, state, oppInp, mask=None):
        E = inp.shape[0] // self.n
        return self.forward(inp.view(self.n, E, OBS_DIM), oppInp.view(self.m, E, OBS_DIM), MODE_ARGMAX)["value"].view(-1, 1)

    def kernel_info(self):
        v = [ctypes.c_int32() for _ in range(5)]
        _capi.check(self._lib.mp_kernel_info(self.n, self.m, *[ctypes.byref(x) for x in v]))
        return dict(zip(("regs", "block", "smem", "blocks_per_sm", "envs_per_tile"), [x.value for x in v]))


def forward_ensemble(policies, own, opp, env_order, env_offsets, mode=MODE_SAMPLE, out=None):
    """ONE launch for an ensemble of frozen checkpoints (mp_forward_ensemble): `policies` is a list of FusedPolicy of the
    same team shape; env_order int32 [E] / env_offsets int32 [K+1] (device) group the environments by the checkpoint they
    play.  Sampling stream, status word and seed are those of policies[0].  Returns the same dict as FusedPolicy.forward."""
    lead = policies[0]
    n, m, dev = lead.n, lead.m, lead.device
    E = own.shape[1]
    if own.shape != (n, E, OBS_DIM) or opp.shape != (m, E, OBS_DIM) or not own.is_contiguous() or not opp.is_contiguous():
        raise ValueError("own/opp must be contiguous [%d,E,6] / [%d,E,6]" % (n, m))
    if own.device != dev or opp.device != dev or own.dtype != torch.float32 or opp.dtype != torch.float32:
        raise ValueError("observations must be float32 on %s" % (dev,))
    if any(f.n != n or f.m != m or f.device != dev for f in policies):
        raise ValueError("every checkpoint of an ensemble must have the team shape and device of the first")
    _check_env_lists(env_order, env_offsets, E, len(policies), dev)
    out = dict(out or {})
    for key, dt in (("value", torch.float32), ("action", torch.int64), ("action_i32", torch.int32), ("logp", torch.float32)):
        if key not in out:
            out[key] = torch.empty((n, E), dtype=dt, device=dev)
    _check_outputs(out, n, E, dev)
    blobs = (ctypes.c_void_p * len(policies))(*[f.blob.data_ptr() for f in policies])
    _capi.check(lead._lib.mp_forward_ensemble(blobs, len(policies), own.data_ptr(), opp.data_ptr(), n, m, E, mode, lead.seed,
                                              lead.calls, lead.counter.data_ptr(), lead.env_id0, out["value"].data_ptr(),
                                              out["action"].data_ptr(), out["action_i32"].data_ptr(), out["logp"].data_ptr(),
                                              env_order.data_ptr(), env_offsets.data_ptr(), lead.status.data_ptr(),
                                              torch.cuda.current_stream(dev).cuda_stream))
    lead.calls += 1
    lead.launches += 1
    return out
