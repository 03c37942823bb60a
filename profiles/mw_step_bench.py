"""mw_step roofline: simple_spread-shaped worlds (3 agents + 3 landmarks), float32, E worlds per launch.
Algorithmic bytes per world-step: positions 6 x 8 B read + 3 x 8 B written, velocities 3 x (8 + 8) B, actions 3 x 8 B = 144 B."""
import json
import os
import sys
from importlib import import_module

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
mw = import_module("emergent-multiagent-strategies_b200.mape_world")
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
for E in (4096, 65536, 1 << 20, 1 << 22):
    w = mw.MapeWorldBatch(E, [dict(size=0.15)] * 3, [dict(collide=False)] * 3)
    w.pos.uniform_(-1, 1)
    u = torch.randn(3, E, 2, device="cuda")
    for _ in range(5):
        w.step(u)
    n = 50
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        w.step(u)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    gbs = E * 144 / (us * 1e-6) / 1e9
    print("mw_step E=%8d: %8.2f us/step, %.3e world-steps/s, %.0f GB/s algorithmic = %.2f of the measured %.0f GB/s" % (E, us, E / us * 1e6, gbs, gbs / peak, peak))
