"""BASELINE configs[0]: the reference's UNCHANGED train_fortattack.py (run as a script from baseline/_ref, the byte copy
baseline/install_ref.py makes) on this repo's modules -- sys.path order is the only thing that changes
(baseline/run_config1.py).  CPU: the reference's numpy env under this repo's rlcore + mpnn.  GPU: this repo's CUDA env
(gym facade -> fa_step_host) under (a) this repo's rlcore + mpnn and (b) the reference's own rlcore + mpnn.
Checked: the script runs its updates, every logged loss is finite, ep0.pt is written in the reference's format and the
REFERENCE's MPNN loads it strictly (train_fortattack.py:121-128, learner.py:245-249)."""
import json
import math
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUNNER = os.path.join(ROOT, "baseline", "run_config1.py")
REF = os.path.join(ROOT, "baseline", "_ref", "reference")

needs_ref = pytest.mark.skipif(not os.path.isdir(REF) and not os.path.isdir("/root/reference"),
                               reason="baseline/_ref is absent (python baseline/install_ref.py needs /root/reference)")


def _run(tmp_path, env, rl, teams, flags, timeout=900):
    cmd = [sys.executable, RUNNER, "--env", env, "--rl", rl, "--teams", teams, "--workdir", str(tmp_path), "--"] + flags
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    assert out.returncode == 0, out.stderr[-3000:]
    res = json.loads(out.stdout.strip().splitlines()[-1])
    chk = subprocess.run([sys.executable, RUNNER, "--teams", teams, "--check-ckpt", os.path.join(res["save_dir"], "ep0.pt")],
                         capture_output=True, text=True, timeout=300)
    assert chk.returncode == 0, chk.stderr[-3000:]
    res["ckpt"] = json.loads(chk.stdout.strip().splitlines()[-1])
    return res


def _check(res, n_agents, updates, env_mod, rl_mod):
    assert res["env_module"].startswith(env_mod) and res["mpnn_module"].startswith(rl_mod) and res["algo_module"].startswith(rl_mod)
    assert "ep0.pt" in res["files"] and "params.json" in res["files"]
    assert res["ckpt"] == {"keys": 24, "params": 158153}
    for tag in ("all/value_loss", "all/action_loss", "all/dist_entropy"):
        assert len(res["scalars"][tag]) == updates and all(math.isfinite(v) for v in res["scalars"][tag]), res["scalars"]
    # a fresh policy is near uniform; the reference divides the per-minibatch sum by ppo_epoch * 32 even when T is not a
    # multiple of 32 and the sampler yields more minibatches (ppo.py:198-202), hence the slack above ln 8
    assert 0.5 < res["scalars"]["all/dist_entropy"][0] <= math.log(8) * 1.25
    for i in range(n_agents):
        assert len(res["scalars"]["agent%d/training_reward" % i]) == updates


@needs_ref
def test_reference_train_script_on_this_rlcore_cpu(tmp_path):
    res = _run(tmp_path, "ref", "ours", "3v3", ["--no-cuda", "--num-steps", "100", "--num-frames", "200", "--seed", "3",
                                                "--save-dir", "c1"])
    _check(res, 6, 2, "baseline/_ref/reference/gym_fortattack", "emergent-multiagent-strategies_b200")


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("teams,rl", [("5v5", "ours"), ("3v3", "ours"), ("5v5", "ref")])
def test_reference_train_script_on_cuda_env(tmp_path, teams, rl):
    """`python train_fortattack.py --num-steps 300 --num-frames 600 --seed 3`, env = the CUDA engine."""
    res = _run(tmp_path, "ours", rl, teams, ["--num-steps", "300", "--num-frames", "600", "--seed", "3", "--save-dir", "c1"])
    n = sum(int(x) for x in teams.split("v"))
    _check(res, n, 2, "emergent-multiagent-strategies_b200/gym_fortattack",
           "emergent-multiagent-strategies_b200" if rl == "ours" else "baseline/_ref/reference")
