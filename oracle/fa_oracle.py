"""ctypes binding of the CPU oracle (oracle/fa_oracle.c).  TEST INFRASTRUCTURE, NOT PRODUCT.

Allowed importers: tests/, __graft_entry__.smoke(), bench.py (cpu_baseline / --impl reference).
The product package never imports this module.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libfa_oracle.so")

_c = ctypes
_pd = _c.POINTER(_c.c_double)
_pu8 = _c.POINTER(_c.c_uint8)
_pi32 = _c.POINTER(_c.c_int32)
_pu32 = _c.POINTER(_c.c_uint32)


def build(force=False):
    """Compile oracle/fa_oracle.c with the committed Makefile (gcc, no FMA contraction)."""
    src = os.path.join(HERE, "fa_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-B", "libfa_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(LIB_PATH)
        L.fa_oracle_reset.argtypes = [_c.c_int, _c.c_int, _c.c_int, _c.c_uint64, _c.c_uint64, _pu8,
                                      _pd, _pu8, _pi32, _pu32, _pd]
        L.fa_oracle_step.argtypes = [_c.c_int, _c.c_int, _c.c_int, _c.c_int, _pd, _pu8, _pi32, _pi32,
                                     _pd, _pd, _pu8, _pu8, _pd, _c.c_int, _c.c_uint64, _c.c_uint64,
                                     _pu32, _c.c_int]
        L.fa_oracle_step_many.argtypes = [_c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _pd, _pu8,
                                          _pi32, _pi32, _pd, _pd, _pu8, _pu8, _c.c_uint64,
                                          _c.c_uint64, _pu32, _c.c_int]
        L.fa_oracle_philox.argtypes = [_c.c_uint32, _c.c_uint32, _pu32]
        L.fa_oracle_philox.restype = None
        _lib = L
    return _lib


def _p(a, t):
    return None if a is None else a.ctypes.data_as(t)


def philox(k0, k1, ctr):
    c = np.array(ctr, np.uint32)
    lib().fa_oracle_philox(k0, k1, _p(c, _pu32))
    return c


class OracleEnv(object):
    """E independent FortAttack envs in float64, stepped by the C oracle.

    st_f [E,A,6] f64: x, y, vx, vy, ang, prevDist(NaN=None); st_i [E,A,6] u8: alive, justDied, hit,
    wasHit, numHit, numWasHit; time_step [E] i32; episode [E] u32.
    """

    def __init__(self, n_envs, n_guards=3, n_attackers=3, max_steps=100, seed=0, env_id0=0,
                 n_threads=1):
        self.E, self.ng, self.na = int(n_envs), int(n_guards), int(n_attackers)
        self.A = self.ng + self.na
        self.max_steps, self.seed, self.env_id0 = int(max_steps), int(seed), int(env_id0)
        self.n_threads = int(n_threads)
        self.st_f = np.zeros((self.E, self.A, 6), np.float64)
        self.st_f[:, :, 5] = np.nan                      # prevDist = None (core.py:104)
        self.st_i = np.zeros((self.E, self.A, 6), np.uint8)
        self.st_i[:, :, 0] = 1
        self.time_step = np.zeros(self.E, np.int32)
        self.episode = np.zeros(self.E, np.uint32)

    def reset(self, mask=None):
        obs = np.zeros((self.E, self.A, 6), np.float64)
        if mask is not None:
            mask = np.ascontiguousarray(mask, np.uint8)
            obs[:] = self.observe()
        rc = lib().fa_oracle_reset(self.E, self.ng, self.na, self.seed, self.env_id0, _p(mask, _pu8),
                                   _p(self.st_f, _pd), _p(self.st_i, _pu8), _p(self.time_step, _pi32),
                                   _p(self.episode, _pu32), _p(obs, _pd))
        assert rc == 0
        return obs

    def observe(self):
        o = np.empty((self.E, self.A, 6), np.float64)
        o[:, :, 0] = self.st_i[:, :, 0]
        o[:, :, 1:3] = self.st_f[:, :, 0:2]
        o[:, :, 3] = self.st_f[:, :, 4]
        o[:, :, 4:6] = self.st_f[:, :, 2:4]
        return o

    def step(self, actions, auto_reset=False, want_margin=False):
        actions = np.ascontiguousarray(actions, np.int32).reshape(self.E, self.A)
        obs = np.empty((self.E, self.A, 6), np.float64)
        rew = np.empty((self.E, self.A), np.float64)
        done = np.empty(self.E, np.uint8)
        result = np.empty(self.E, np.uint8)
        margin = np.empty(self.E, np.float64) if want_margin else None
        rc = lib().fa_oracle_step(self.E, self.ng, self.na, self.max_steps, _p(self.st_f, _pd),
                                  _p(self.st_i, _pu8), _p(self.time_step, _pi32), _p(actions, _pi32),
                                  _p(obs, _pd), _p(rew, _pd), _p(done, _pu8), _p(result, _pu8),
                                  _p(margin, _pd), int(auto_reset), self.seed, self.env_id0,
                                  _p(self.episode, _pu32), self.n_threads)
        assert rc == 0
        if want_margin:
            return obs, rew, done, result, margin
        return obs, rew, done, result

    def alloc_out(self, T):
        """Reusable output streams for step_many (touch them once so timing excludes page faults)."""
        out = (np.zeros((T, self.E, self.A, 6), np.float64), np.zeros((T, self.E, self.A), np.float64),
               np.zeros((T, self.E), np.uint8), np.zeros((T, self.E), np.uint8))
        return out

    def step_many(self, actions, store=True, out=None):
        actions = np.ascontiguousarray(actions, np.int32)
        T = actions.shape[0]
        assert actions.shape == (T, self.E, self.A)
        obs = rew = done = result = None
        if out is not None:
            obs, rew, done, result = out
            assert obs.shape == (T, self.E, self.A, 6) and obs.flags.c_contiguous
        elif store:
            obs = np.empty((T, self.E, self.A, 6), np.float64)
            rew = np.empty((T, self.E, self.A), np.float64)
            done = np.empty((T, self.E), np.uint8)
            result = np.empty((T, self.E), np.uint8)
        rc = lib().fa_oracle_step_many(T, self.E, self.ng, self.na, self.max_steps, _p(self.st_f, _pd),
                                       _p(self.st_i, _pu8), _p(self.time_step, _pi32),
                                       _p(actions, _pi32), _p(obs, _pd), _p(rew, _pd), _p(done, _pu8),
                                       _p(result, _pu8), self.seed, self.env_id0,
                                       _p(self.episode, _pu32), self.n_threads)
        assert rc == 0
        return obs, rew, done, result
