"""Helpers shared by the CPU (oracle) and GPU (CUDA) replays of tests/golden/env_*.npz."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CAP = 100      # common episode cap used when replaying; t_pre is shifted so `done` is unchanged


def load(name):
    g = dict(np.load(os.path.join(GOLDEN, name)))
    g["n_guards"] = int(g["n_guards"])
    g["n_attackers"] = int(g["n_attackers"])
    # done only depends on time_step == cap-1 (fortattack.py:218): shift t so one cap serves all rows
    g["t_shift"] = (g["t_pre"] + (CAP - g["cap"])).astype(np.int32)
    return g
