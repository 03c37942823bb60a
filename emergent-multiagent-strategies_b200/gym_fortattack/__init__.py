"""Drop-in for the reference's `gym_fortattack` package: only the env factory is on the path
(gym_fortattack/__init__.py:1-8 registers 'fortattack-v1'; fortattack.py:17 is the factory callers use)."""
from .fortattack import FortAttackGlobalEnv, make_fortattack_env  # noqa: F401
