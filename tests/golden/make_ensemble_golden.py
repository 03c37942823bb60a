"""Golden fixture for BASELINE configs[4]: the reference's own evaluation table of the shipped guard policy
(marlsave/tmp_1/ep2520.pt, models[0]) against each shipped attacker checkpoint {220, 650, 1240, 1600, 2520}.

    python tests/golden/make_ensemble_golden.py [episodes]

Runs the UNCHANGED reference script test_fortattack_v2.py (:25-124) on the reference's own numpy env, CPU, through
baseline/run_config1.py (stub modules for gym / pygame / pyglet only), and stores the table it writes to
marlsave/stats/stats_ensemble_strategies.csv -- columns [P(all attackers dead), P(time limit), P(guards win),
P(fort reached), alive guards, alive attackers, mean guard return, mean attacker return], one row per attacker
checkpoint -- in tests/golden/ensemble_ref_stats.json.  Only runs where /root/reference (or baseline/_ref) exists."""
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CKPTS = [220, 650, 1240, 1600, 2520]


def main():
    episodes = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    ref = os.path.join(ROOT, "baseline", "_ref", "reference")
    if not os.path.isdir(ref):
        subprocess.check_call([sys.executable, os.path.join(ROOT, "baseline", "install_ref.py")])
    work = tempfile.mkdtemp(prefix="fa_ens_")
    os.makedirs(os.path.join(work, "marlsave", "stats"))
    shutil.copytree(os.path.join(ref, "marlsave", "tmp_1"), os.path.join(work, "marlsave", "tmp_1"))
    seed = 1
    cmd = [sys.executable, os.path.join(ROOT, "baseline", "run_config1.py"), "--env", "ref", "--rl", "ref", "--teams", "5v5",
           "--workdir", work, "--script", "test_fortattack_v2.py", "--", "--test", "--train-guards-only", "--no-cuda",
           "--num-eval-episodes", str(episodes), "--load-dir", "tmp_1", "--ckpt", "2520", "--attacker-load-dir", "tmp_1",
           "--attacker-ckpts"] + [str(c) for c in CKPTS] + ["--seed", str(seed)]
    subprocess.check_call(cmd, stdout=subprocess.DEVNULL)
    table = np.loadtxt(os.path.join(work, "marlsave", "stats", "stats_ensemble_strategies.csv"), delimiter=",")
    out = {"source": "unchanged /root/reference test_fortattack_v2.py, numpy env, CPU, sampled actions",
           "guards": "marlsave/tmp_1/ep2520.pt models[0]", "attacker_ckpts": CKPTS, "episodes_per_ckpt": episodes, "seed": seed,
           "columns": ["p_all_attackers_dead", "p_time_limit", "p_guards_win", "p_fort_reached", "alive_guards",
                       "alive_attackers", "guard_return", "attacker_return"],
           "table": table.tolist()}
    with open(os.path.join(ROOT, "tests", "golden", "ensemble_ref_stats.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))
    shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
