"""B200-native batched FortAttack simulator behind the reference's env / rollout API.

The directory name contains hyphens, so import it with
    importlib.import_module("emergent-multiagent-strategies_b200")
(or through the repo-root alias module `fortattack_b200`).
"""
from . import _capi  # noqa: F401
from ._capi import FaError  # noqa: F401
from .batched_env import FortAttackBatch  # noqa: F401
