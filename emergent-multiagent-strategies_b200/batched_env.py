"""Device-resident batch of E independent FortAttack environments.

Host-side mirror of the reference's env seam for the batched configurations:
`reset()` / `step(actions)` have the meaning of FortAttackGlobalEnv.reset / .step
(gym_fortattack/fortattack.py:127-186) applied to every env at once, with agent-major tensors
(`obs[i]`, `reward[i]` index an agent, exactly what Learner.update_rollout / RolloutStorage.insert
take: learner.py:239-243, rlcore/storage.py:33-43).  All arithmetic happens in the CUDA library
(csrc/, through include/fortattack.h); PyTorch only owns the memory and the stream.
"""
import ctypes
import os

import torch

from . import _capi


class FortAttackBatch(object):
    def __init__(self, n_envs, n_guards=3, n_attackers=3, max_steps=100, seed=0, env_id0=0, device="cuda:0",
                 dtype=torch.float32, mapping="auto"):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise _capi.FaError("FortAttackBatch needs a CUDA device (got %r); there is no CPU path" % (device,))
        if dtype not in (torch.float32, torch.float64):
            raise ValueError("dtype must be torch.float32 or torch.float64")
        self.device, self.dtype = dev, dtype
        self.E, self.n_guards, self.n_attackers = int(n_envs), int(n_guards), int(n_attackers)
        self.A = self.n_guards + self.n_attackers
        self._lib = _capi.lib()
        self.cfg = _capi.FaConfig(self.E, self.n_guards, self.n_attackers, int(max_steps),
                                  _capi.FA_F64 if dtype == torch.float64 else _capi.FA_F32,
                                  dev.index if dev.index is not None else torch.cuda.current_device(),
                                  {"auto": _capi.FA_MAP_AUTO, "env": _capi.FA_MAP_ENV, "agent": _capi.FA_MAP_AGENT, "group": _capi.FA_MAP_GROUP}[mapping],
                                  0, int(seed), int(env_id0))
        nbytes = ctypes.c_size_t()
        _capi.check(self._lib.fa_workspace_bytes(ctypes.byref(self.cfg), ctypes.byref(nbytes)))
        with torch.cuda.device(dev):
            # torch's caching allocator hands out 512-byte aligned blocks
            self.workspace = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
            h = ctypes.c_void_p()
            torch.cuda.synchronize(dev)
            _capi.check(self._lib.fa_create(ctypes.byref(self.cfg), self.workspace.data_ptr(), ctypes.byref(h)))
        self._h = h
        self.max_steps = int(max_steps)

    # -- helpers ---------------------------------------------------------------------------------
    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _new(self, *shape, dtype=None):
        return torch.empty(shape, dtype=dtype or self.dtype, device=self.device)

    def _dev(self, name, t, shape, dtype, optional=False, numel=None):
        """The library reads / writes through raw pointers: refuse anything that is not a contiguous tensor of exactly
        the documented shape (or, with shape None, element count) and type on this handle's device."""
        if t is None and optional:
            return None
        if (t is None or (shape is not None and tuple(t.shape) != tuple(shape)) or (numel is not None and t.numel() != numel)
                or (dtype is not None and t.dtype != dtype) or t.device != self.workspace.device or not t.is_contiguous()):
            raise ValueError("%s must be a contiguous %s tensor of %s on %s"
                             % (name, dtype, "shape %r" % (tuple(shape),) if shape is not None else "%d elements" % numel, self.device))
        return t

    def _guard(self):
        """Launches, streams and events of this handle belong to its device whatever device is current."""
        return torch.cuda.device(self.device)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.fa_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- the env seam -----------------------------------------------------------------------------
    def reset(self, mask=None, out=None):
        """Reset every env (or those with mask != 0); returns obs [A, E, 6] of all envs."""
        obs = self._dev("out", out, (self.A, self.E, 6), self.dtype) if out is not None else self._new(self.A, self.E, 6)
        m = None
        if mask is not None:
            if mask.numel() != self.E:
                raise ValueError("mask must have one entry per env (%d), got %d" % (self.E, mask.numel()))
            m = mask.reshape(self.E).to(device=self.device, dtype=torch.uint8).contiguous()
        with self._guard():
            _capi.check(self._lib.fa_reset(self._h, m.data_ptr() if m is not None else None, obs.data_ptr(),
                                           self._stream()))
        return obs

    def step(self, actions, auto_reset=True, out=None, bookkeeping=None):
        """actions int32 [A, E] -> (obs [A,E,6], reward [A,E], done u8 [E], result u8 [E]).
        bookkeeping = (mask_next float32 [A, E], end_next uint8/bool [E], episode_reward float32 [A, E]) (entries may be None):
        the rollout loop's per-step glue written by the same launch (fa_set_rollout_outputs; train_fortattack.py:53,97-104)."""
        a = self._actions(actions, (self.A, self.E))
        if out is None:
            out = (self._new(self.A, self.E, 6), self._new(self.A, self.E),
                   self._new(self.E, dtype=torch.uint8), self._new(self.E, dtype=torch.uint8))
        else:
            if len(out) != 4:
                raise ValueError("out = (obs, reward, done, result)")
            self._dev("out[0] (obs)", out[0], (self.A, self.E, 6), self.dtype)
            self._dev("out[1] (reward)", out[1], (self.A, self.E), self.dtype)
            self._dev("out[2] (done)", out[2], (self.E,), torch.uint8)
            self._dev("out[3] (result)", out[3], (self.E,), torch.uint8)
        self._check_alive_end(1)
        obs, rew, done, result = out
        bk = None
        if bookkeeping is not None:
            if len(bookkeeping) != 3:
                raise ValueError("bookkeeping = (mask_next, end_next, episode_reward)")
            self._dev("bookkeeping[0] (mask_next)", bookkeeping[0], None, torch.float32, optional=True, numel=self.A * self.E)
            if bookkeeping[1] is not None and bookkeeping[1].dtype not in (torch.uint8, torch.bool):
                raise ValueError("bookkeeping[1] (end_next) must be uint8 or bool")
            self._dev("bookkeeping[1] (end_next)", bookkeeping[1], None, None, optional=True, numel=self.E)
            self._dev("bookkeeping[2] (episode_reward)", bookkeeping[2], None, torch.float32, optional=True, numel=self.A * self.E)
            bk = [None if t is None else t.data_ptr() for t in bookkeeping]
        with self._guard():
            if bk is not None:
                _capi.check(self._lib.fa_set_rollout_outputs(self._h, *bk))
            try:
                _capi.check(self._lib.fa_step(self._h, a.data_ptr(), obs.data_ptr(), rew.data_ptr(), done.data_ptr(),
                                              result.data_ptr(), int(bool(auto_reset)), self._stream()))
            finally:
                if bk is not None:
                    _capi.check(self._lib.fa_set_rollout_outputs(self._h, None, None, None))
        return obs, rew, done, result

    def step_many(self, actions, out=None, store_obs=True):
        """actions int32 [T, A, E]: T steps in one persistent launch (auto-reset on)."""
        T = int(actions.shape[0])
        a = self._actions(actions, (T, self.A, self.E))
        if out is None:
            out = (self._new(T, self.A, self.E, 6) if store_obs else None, self._new(T, self.A, self.E),
                   self._new(T, self.E, dtype=torch.uint8), self._new(T, self.E, dtype=torch.uint8))
        else:
            if len(out) != 4:
                raise ValueError("out = (obs or None, reward, done, result)")
            self._dev("out[0] (obs)", out[0], (T, self.A, self.E, 6), self.dtype, optional=True)
            self._dev("out[1] (reward)", out[1], (T, self.A, self.E), self.dtype, optional=True)
            self._dev("out[2] (done)", out[2], (T, self.E), torch.uint8, optional=True)
            self._dev("out[3] (result)", out[3], (T, self.E), torch.uint8, optional=True)
        self._check_alive_end(T)
        obs, rew, done, result = out
        ptr = lambda t: t.data_ptr() if t is not None else None
        with self._guard():
            _capi.check(self._lib.fa_step_many(self._h, T, a.data_ptr(), ptr(obs), ptr(rew), ptr(done), ptr(result),
                                               self._stream()))
        return obs, rew, done, result

    def _check_alive_end(self, T):
        buf = getattr(self, "_alive_end", None)
        if buf is not None and buf.numel() < T * self.E:
            raise ValueError("the alive-end buffer holds %d entries; %d step(s) of %d envs write %d (set_alive_end_buffer "
                             "with a [T, E] tensor, or None)" % (buf.numel(), T, self.E, T * self.E))

    def _actions(self, actions, shape):
        a = actions
        if a.device != self.device or a.dtype != torch.int32 or not a.is_contiguous():
            a = a.to(device=self.device, dtype=torch.int32).contiguous()
        if tuple(a.shape) != shape:
            raise ValueError("actions must have shape %r, got %r" % (shape, tuple(a.shape)))
        return a

    # -- host-buffer step (numpy-facing env.step) --------------------------------------------------
    def make_host_buffers(self):
        """Page-locked host buffers for step_host: actions, obs, reward, done, result.  The four result
        buffers are views of ONE pinned block laid out as fa_host_layout() says, so the staged path
        returns them in a single copy and the mapped path writes them in place."""
        off = [ctypes.c_size_t() for _ in range(4)]
        _capi.check(self._lib.fa_host_layout(self._h, *[ctypes.byref(o) for o in off]))
        o_rew, o_done, o_res, total = [o.value for o in off]
        block = torch.zeros(total, dtype=torch.uint8).pin_memory()
        rs = 8 if self.dtype == torch.float64 else 4
        n = self.A * self.E
        obs = block[:n * 6 * rs].view(self.dtype).view(self.A, self.E, 6)
        rew = block[o_rew:o_rew + n * rs].view(self.dtype).view(self.A, self.E)
        done, res = block[o_done:o_done + self.E], block[o_res:o_res + self.E]
        self._host_block = block
        return (torch.zeros(self.A, self.E, dtype=torch.int32).pin_memory(), obs, rew, done, res)

    def step_host(self, h_actions, h_obs, h_rew, h_done, h_result, auto_reset=True):
        """One env.step() with HOST tensors: H2D actions, fused step, D2H results, synchronous."""
        for name, t, shape, dtype in (("h_actions", h_actions, (self.A, self.E), torch.int32),
                                      ("h_obs", h_obs, (self.A, self.E, 6), self.dtype), ("h_rew", h_rew, (self.A, self.E), self.dtype),
                                      ("h_done", h_done, (self.E,), torch.uint8), ("h_result", h_result, (self.E,), torch.uint8)):
            if tuple(t.shape) != shape or t.dtype != dtype or t.is_cuda or not t.is_contiguous():
                raise ValueError("%s must be a contiguous host %s tensor of shape %r" % (name, dtype, shape))
        self._check_alive_end(1)
        with self._guard():
            _capi.check(self._lib.fa_step_host(self._h, h_actions.data_ptr(), h_obs.data_ptr(), h_rew.data_ptr(),
                                               h_done.data_ptr(), h_result.data_ptr(), int(bool(auto_reset)),
                                               self._stream()))
        return h_obs, h_rew, h_done, h_result

    def make_host_streams(self, T, store_obs=True):
        """Page-locked host streams for step_many_host: actions int32 [T,A,E] and (obs [T,A,E,6] or None,
        reward [T,A,E], done u8 [T,E], result u8 [T,E])."""
        pin = lambda *shape, dtype: torch.zeros(shape, dtype=dtype).pin_memory()
        return (pin(T, self.A, self.E, dtype=torch.int32),
                pin(T, self.A, self.E, 6, dtype=self.dtype) if store_obs else None,
                pin(T, self.A, self.E, dtype=self.dtype), pin(T, self.E, dtype=torch.uint8),
                pin(T, self.E, dtype=torch.uint8))

    def step_many_host(self, h_actions, h_obs, h_rew, h_done, h_result, chunk_steps=None):
        """T env.step() calls of every env with HOST streams (fa_step_many_host), synchronous.
        chunk_steps=None: chunks of about 16 MB of results (measured: profiles/r3g_e2e_chunks.log -- smaller chunks pay the
        per-chunk copy / event calls, larger ones the pipeline fill), copied by the DMA engines while the neighbouring chunks
        compute; chunk_steps=0: no staging, one persistent launch working through mapped pinned memory."""
        T = self._check_host_streams(h_actions, h_obs, h_rew, h_done, h_result)
        self._check_alive_end(T)
        ptr = lambda t: t.data_ptr() if t is not None else None
        stage, nbytes = None, 0
        if chunk_steps is None:
            rs = 8 if self.dtype == torch.float64 else 4
            target = int(float(os.environ.get("FA_HOST_CHUNK_MB", "16")) * (1 << 20))      # (experiments: chunk size of the pipeline)
            chunk_steps = max(1, min(T, target // (self.A * self.E * 7 * rs + 2 * self.E)))
        if chunk_steps > 0:
            need = ctypes.c_size_t()
            _capi.check(self._lib.fa_host_stage_bytes(self._h, int(chunk_steps), ctypes.byref(need)))
            if getattr(self, "_stage", None) is None or self._stage.numel() < need.value:
                self._stage = torch.empty(need.value, dtype=torch.uint8, device=self.device)
            stage, nbytes = self._stage.data_ptr(), need.value
        with self._guard():
            _capi.check(self._lib.fa_step_many_host(self._h, T, h_actions.data_ptr(), ptr(h_obs), ptr(h_rew), ptr(h_done),
                                                    ptr(h_result), stage, nbytes, self._stream()))
        return h_obs, h_rew, h_done, h_result

    def _check_host_streams(self, h_actions, h_obs, h_rew, h_done, h_result):
        """The library writes T steps through these raw pointers: refuse anything that is not a contiguous host tensor of
        exactly the documented shape and type.  Returns T."""
        T = int(h_actions.shape[0]) if h_actions.dim() == 3 else -1
        want = (("h_actions", h_actions, (T, self.A, self.E), torch.int32, False),
                ("h_obs", h_obs, (T, self.A, self.E, 6), self.dtype, True),
                ("h_rew", h_rew, (T, self.A, self.E), self.dtype, True),
                ("h_done", h_done, (T, self.E), torch.uint8, True),
                ("h_result", h_result, (T, self.E), torch.uint8, True))
        for name, t, shape, dtype, optional in want:
            if t is None and optional:
                continue
            if (t is None or T < 1 or tuple(t.shape) != shape or t.dtype != dtype or t.is_cuda or not t.is_contiguous()):
                raise ValueError("%s must be a contiguous host %s tensor of shape %r" % (name, dtype, shape))
        return T

    # -- state exchange (canonical float64 layout, include/fortattack.h FaState) -------------------
    def get_state(self):
        st_f = self._new(self.E, self.A, 6, dtype=torch.float64)
        st_i = self._new(self.E, self.A, 6, dtype=torch.uint8)
        t = self._new(self.E, dtype=torch.int32)
        ep = self._new(self.E, dtype=torch.int32)       # bit pattern of uint32
        s = _capi.FaState(st_f.data_ptr(), st_i.data_ptr(), t.data_ptr(), ep.data_ptr())
        with self._guard():
            _capi.check(self._lib.fa_get_state(self._h, ctypes.byref(s), self._stream()))
        return st_f, st_i, t, ep

    def set_state(self, st_f, st_i, time_step, episode):
        dev = self.device
        st_f = torch.as_tensor(st_f).to(device=dev, dtype=torch.float64).contiguous()
        st_i = torch.as_tensor(st_i).to(device=dev, dtype=torch.uint8).contiguous()
        t = torch.as_tensor(time_step).to(device=dev, dtype=torch.int32).contiguous()
        ep = torch.as_tensor(episode).to(device=dev).to(torch.int32).contiguous()
        assert tuple(st_f.shape) == (self.E, self.A, 6) and tuple(st_i.shape) == (self.E, self.A, 6)
        assert tuple(t.shape) == (self.E,) and tuple(ep.shape) == (self.E,)
        s = _capi.FaState(st_f.data_ptr(), st_i.data_ptr(), t.data_ptr(), ep.data_ptr())
        with self._guard():
            _capi.check(self._lib.fa_set_state(self._h, ctypes.byref(s), self._stream()))
        torch.cuda.current_stream(dev).synchronize()    # the temporaries above may be freed on return

    def alive_counts(self):
        """(numAliveGuards [E], numAliveAttackers [E]) int32 (core.py:113-114)."""
        c = self._new(2, self.E, dtype=torch.int32)
        with self._guard():
            _capi.check(self._lib.fa_alive_counts(self._h, c.data_ptr(), self._stream()))
        return c[0], c[1]

    def set_max_steps(self, max_steps):
        _capi.check(self._lib.fa_set_max_steps(self._h, int(max_steps)))
        self.max_steps = int(max_steps)

    def set_alive_end_buffer(self, buf):
        """uint8 [E] (or [T, E] for step_many) device tensor that every later step fills with
        numAliveGuards | numAliveAttackers << 4 as the step leaves them (before an auto-reset); None switches it off."""
        if buf is not None and (buf.dtype != torch.uint8 or not buf.is_contiguous() or buf.device != self.workspace.device):
            raise ValueError("alive-end buffer must be a contiguous uint8 tensor on %s" % self.device)
        self._alive_end = buf                     # keep it alive
        _capi.check(self._lib.fa_set_alive_end_buffer(self._h, None if buf is None else buf.data_ptr()))

    def launch_count(self):
        n = ctypes.c_uint64()
        _capi.check(self._lib.fa_launch_count(self._h, ctypes.byref(n)))
        return n.value

    def kernel_info(self):
        v = [ctypes.c_int32() for _ in range(5)]
        _capi.check(self._lib.fa_kernel_info(self._h, *[ctypes.byref(x) for x in v]))
        d = dict(zip(("regs", "block", "grid", "smem", "mapping"), [x.value for x in v]))
        d["mapping"] = {_capi.FA_MAP_ENV: "env", _capi.FA_MAP_AGENT: "agent", _capi.FA_MAP_GROUP: "group"}[d["mapping"]]
        return d
