"""GPU: BASELINE configs[4] with the reference's SHIPPED checkpoints -- guards marlsave/tmp_1/ep2520.pt (models[0]) against
the attacker ensemble {ep220, ep650, ep1240, ep1600, ep2520} (arguments.py:62; learner.py:119-140).

The batched ensemble path (one mp_forward_ensemble launch per step, per-env per-episode checkpoint draw) must reproduce
the reference's own evaluation table (test_fortattack_v2.py:50,94-124) per attacker checkpoint.  The fixture
tests/golden/ensemble_ref_stats.json is that table as the UNCHANGED reference script computes it on its numpy env
(200 episodes per checkpoint, tests/golden/make_ensemble_golden.py); the GPU side plays ~25 000 episodes, so the
tolerance is the fixture's own sampling noise (+ the fp16-operand policy kernel's ~1e-3 mean log-prob deviation)."""
import json
import os
import shutil
import subprocess
import sys
from importlib import import_module

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
PKG = "emergent-multiagent-strategies_b200"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CKDIR = os.path.join(ROOT, "baseline", "_ref", "reference", "marlsave", "tmp_1")
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "ensemble_ref_stats.json")))
needs_ckpts = pytest.mark.skipif(not os.path.exists(os.path.join(CKDIR, "ep2520.pt")),
                                 reason="shipped checkpoints absent: python baseline/install_ref.py (needs /root/reference)")


def _load(k):
    return torch.load(os.path.join(CKDIR, "ep%d.pt" % k), map_location="cpu")["models"]


@needs_ckpts
def test_ensemble_table_matches_reference_evaluation():
    ro = import_module(PKG + ".rollout")
    ckpts = GOLD["attacker_ckpts"]
    torch.manual_seed(0)
    tr = ro.BatchedTrainer(2048, 5, 5, num_steps=100, max_episode_steps=100, seed=11,
                           attacker_ensemble=[_load(k)[-1] for k in ckpts])
    tr.load_models(_load(2520))                       # guards <- models[0] (learner.py:245-249)
    for _ in range(4):
        tr.collect()
        tr.after_update()
    for f in tr.fused + tr.ensemble:
        f.check_status()
    table = tr.ensemble_table()                        # [K, 8]
    episodes = tr.ensemble_sums[:, 0].cpu().numpy()
    ref = np.array(GOLD["table"])
    print("episodes per checkpoint:", episodes.astype(int).tolist())
    for k, row, r in zip(ckpts, table, ref):
        print("ep%-5d gpu %s\n        ref %s" % (k, np.round(row, 2).tolist(), r.tolist()))
    assert episodes.min() > 2000
    n = GOLD["episodes_per_ckpt"]
    p = ref[:, :4]
    tol_p = 3.0 * np.sqrt(np.maximum(p * (1 - p), 0.02) / n) + 0.02            # 3 sigma of the fixture + policy-kernel slack
    assert (np.abs(table[:, :4] - p) <= tol_p).all(), (np.abs(table[:, :4] - p), tol_p)
    assert (np.abs(table[:, 4] - ref[:, 4]) <= 0.45).all() and (np.abs(table[:, 5] - ref[:, 5]) <= 0.25).all()
    assert (np.abs(table[:, 6] - ref[:, 6]) <= 1.6).all() and (np.abs(table[:, 7] - ref[:, 7]) <= 0.8).all()
    # the ordering the paper's ensemble argument rests on: the early strategy (ep220) is the easiest for these guards, the
    # late sneak strategy (ep1600) the hardest
    assert table[0, 2] > table[4, 2] > table[3, 2]


@needs_ckpts
def test_reference_eval_script_runs_on_cuda_env(tmp_path):
    """The reference's UNCHANGED test_fortattack_v2.py (guards ep2520 vs attacker ckpts 220 and 2520, 25 episodes each) on
    the CUDA env facade; its own stats file must land near the fixture.  (--no-cuda: the script's np.average over a torch
    tensor, :100-101, only works with the policies on the CPU; the env is the CUDA engine either way.)"""
    os.makedirs(tmp_path / "marlsave" / "stats")
    shutil.copytree(CKDIR, tmp_path / "marlsave" / "tmp_1")
    cmd = [sys.executable, os.path.join(ROOT, "baseline", "run_config1.py"), "--env", "ours", "--rl", "ours", "--teams", "5v5",
           "--workdir", str(tmp_path), "--script", "test_fortattack_v2.py", "--", "--test", "--train-guards-only", "--no-cuda",
           "--num-eval-episodes", "25", "--load-dir", "tmp_1", "--ckpt", "2520", "--attacker-load-dir", "tmp_1",
           "--attacker-ckpts", "220", "2520", "--seed", "2"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-3000:]
    res = json.loads(out.stdout.strip().splitlines()[-1])
    assert res["env_module"].startswith("emergent-multiagent-strategies_b200/gym_fortattack")
    stats = np.loadtxt(tmp_path / "marlsave" / "stats" / "stats_ensemble_strategies.csv", delimiter=",")
    ref = np.array(GOLD["table"])[[0, 4]]
    print("reference script on the CUDA env:", stats.tolist(), "fixture:", ref.tolist())
    assert stats.shape == (2, 8) and np.isfinite(stats).all()
    assert abs(stats[0, 2] - ref[0, 2]) < 0.2 and abs(stats[1, 2] - ref[1, 2]) < 0.35      # 25 episodes: loose
    assert (np.abs(stats[:, 0] + stats[:, 1] + stats[:, 3] - 1.0) < 1e-6).all()
