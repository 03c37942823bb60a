"""Import the UNCHANGED reference from /root/reference in this container.

Test infrastructure only (golden-vector generation).  Nothing under tests/ that
runs on the GPU box imports this file: /root/reference does not exist there.

The reference needs five packages that are absent from this image and that do no
arithmetic on the path: gym, pygame, pyglet, gym_vecenv,
tensorboardX.  They are replaced by the minimal stubs below; every number the
golden files hold is produced by the reference's own numpy/torch code
(gym_fortattack/core.py, gym_fortattack/envs/fortattack_env_v1.py,
gym_fortattack/fortattack.py, mpnn.py, rlcore/*).
"""
import importlib
import io
import contextlib
import os
import sys
import types

REF = os.environ.get("FA_REFERENCE_DIR", "/root/reference")


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_stubs():
    if "gym" in sys.modules and getattr(sys.modules["gym"], "_fa_stub", False):
        return

    class Space(object):
        def __init__(self, shape=None, dtype=None):
            self.shape = shape
            self.dtype = dtype

    class Discrete(Space):
        def __init__(self, n):
            Space.__init__(self, (), int)
            self.n = n

    class Box(Space):
        def __init__(self, low=None, high=None, shape=None, dtype=None):
            Space.__init__(self, tuple(shape) if shape is not None else (), dtype)
            self.low, self.high = low, high

    class Tuple(Space):
        def __init__(self, spaces):
            Space.__init__(self)
            self.spaces = spaces

    class Dict(Space):
        def __init__(self, spaces=None):
            Space.__init__(self)
            self.spaces = spaces

    class Env(object):
        metadata = {}

        def seed(self, seed=None):
            return []

    registry = {}

    def register(id, entry_point=None, **kw):
        registry[id] = entry_point

    def make(id, **kw):
        modname, cls = registry[id].split(":")
        return getattr(importlib.import_module(modname), cls)(**kw)

    spaces = _mod("gym.spaces", Space=Space, Discrete=Discrete, Box=Box, Tuple=Tuple, Dict=Dict)
    seeding = _mod("gym.utils.seeding")
    gutils = _mod("gym.utils", seeding=seeding)
    registration = _mod("gym.envs.registration", register=register, EnvSpec=object)
    envs = _mod("gym.envs", registration=registration)
    error = _mod("gym.error")
    wrappers = _mod("gym.wrappers", Monitor=object)
    _mod("gym", Env=Env, Space=Space, make=make, spaces=spaces, utils=gutils, envs=envs,
         error=error, wrappers=wrappers, _fa_stub=True)

    music = types.SimpleNamespace(load=lambda *a, **k: None, play=lambda *a, **k: None)
    mixer = _mod("pygame.mixer", init=lambda *a, **k: None, music=music)
    _mod("pygame", mixer=mixer)
    gl = _mod("pyglet.gl")
    _mod("pyglet", gl=gl)
    _mod("gym_vecenv")

    class SummaryWriter(object):
        def __init__(self, *a, **k):
            pass

        def add_scalar(self, *a, **k):
            pass

        def close(self):
            pass

    _mod("tensorboardX", SummaryWriter=SummaryWriter)


def import_reference():
    """Put /root/reference first on sys.path and return its modules (unchanged code)."""
    if not os.path.isdir(REF):
        raise RuntimeError("reference tree %s is not present (golden generation runs only "
                           "in the build container)" % REF)
    install_stubs()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import gym_fortattack  # noqa: F401  (registers fortattack-v1)
    from gym_fortattack.fortattack import make_fortattack_env, FortAttackGlobalEnv
    from gym_fortattack.envs.fortattack_env_v1 import FortAttackEnvV1
    return types.SimpleNamespace(make_fortattack_env=make_fortattack_env,
                                 FortAttackGlobalEnv=FortAttackGlobalEnv,
                                 FortAttackEnvV1=FortAttackEnvV1)


def make_ref_env(n_guards, n_attackers, max_steps):
    """Reference env with the team sizes trimmed from the hard-coded 5v5
    (gym_fortattack/envs/fortattack_env_v1.py:18-19) to n_guards v n_attackers."""
    ref = import_reference()
    assert 1 <= n_guards <= 5 and 1 <= n_attackers <= 5
    with contextlib.redirect_stdout(io.StringIO()):
        scen = ref.FortAttackEnvV1()
    w = scen.world
    guards = [a for a in w.agents if not a.attacker][:n_guards]
    attackers = [a for a in w.agents if a.attacker][:n_attackers]
    w.agents = guards + attackers
    w.numGuards, w.numAttackers = n_guards, n_attackers
    w.numAgents = n_guards + n_attackers
    w.max_time_steps = max_steps
    with contextlib.redirect_stdout(io.StringIO()):
        scen.reset_world()
        env = ref.FortAttackGlobalEnv(w, scen.reset_world, scen.reward, scen.observation)
    return env, scen


@contextlib.contextmanager
def quiet():
    """The reference prints on every episode end (gym_fortattack/fortattack.py:208,214,220)."""
    with contextlib.redirect_stdout(io.StringIO()):
        yield
