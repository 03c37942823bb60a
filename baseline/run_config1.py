"""BASELINE configs[0]: the reference's UNCHANGED train_fortattack.py, run as a script, on this repo's engine.

    python baseline/run_config1.py --env ours --rl ours --teams 5v5 -- --num-steps 300 --num-frames 600 --seed 3

What is swapped is only what `sys.path` order swaps (north_star: "train_fortattack.py drops in unchanged"):
    --env ours   `gym_fortattack` resolves to emergent-multiagent-strategies_b200/gym_fortattack (the CUDA engine behind
                 make_fortattack_env / env.reset / env.step); `ref` = the reference's numpy env
    --rl  ours   `rlcore` (RolloutStorage, JointPPO) and `mpnn` (MPNN) resolve to this repo's modules; `ref` = the
                 reference's own
train_fortattack.py, learner.py, rlagent.py, arguments.py, utils.py, eval.py always come from baseline/_ref/reference
(byte copies of /root/reference, baseline/install_ref.py) and are executed with runpy as `__main__`.  Packages the
reference imports but that do no arithmetic (gym, pygame, pyglet, gym_vecenv, tensorboardX) are the stubs of
baseline/_ref/ref_shim.py; the tensorboardX stub here records the scalars so the caller can check them.
Team sizes: the reference hard-codes 5v5 (gym_fortattack/envs/fortattack_env_v1.py:18-19); `--teams 3v3` sets
FORTATTACK_TEAMS, which this repo's make_fortattack_env reads (with --env ref the reference world is trimmed the way
tests/golden/ref_shim.make_ref_env does it).
Prints one JSON line: scalars per tag, checkpoint files written, env/rl modules actually used.
"""
import argparse
import contextlib
import io
import json
import os
import runpy
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REFDIR = os.path.join(HERE, "_ref", "reference")
PKG = os.path.join(ROOT, "emergent-multiagent-strategies_b200")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--env", choices=("ours", "ref"), default="ours")
    ap.add_argument("--rl", choices=("ours", "ref"), default="ours")
    ap.add_argument("--teams", default="5v5")
    ap.add_argument("--workdir", default=None, help="cwd for the run (arguments.py writes ./marlsave/<save-dir> there)")
    ap.add_argument("--script", default="train_fortattack.py")
    ap.add_argument("--check-ckpt", default=None, help="instead of training: load this checkpoint the reference's way "
                    "(learner.load_models -> Neo.load_model -> MPNN.load_state_dict, learner.py:245-249, rlagent.py:20-21) "
                    "into the REFERENCE's own MPNN and run one act() with it")
    ap.add_argument("rest", nargs=argparse.REMAINDER, help="-- then the reference script's own flags")
    a = ap.parse_args()
    rest = a.rest[1:] if a.rest[:1] == ["--"] else a.rest
    if not os.path.isdir(REFDIR):
        sys.path.insert(0, HERE)
        import install_ref
        install_ref.install()
    ng, na = (int(x) for x in a.teams.split("v"))
    os.environ["FORTATTACK_TEAMS"] = "%dv%d" % (ng, na)
    os.environ["FA_REFERENCE_DIR"] = REFDIR
    if a.workdir:
        os.makedirs(a.workdir, exist_ok=True)
        os.chdir(a.workdir)

    sys.path.insert(0, os.path.join(HERE, "_ref"))
    import ref_shim
    ref_shim.install_stubs()
    scalars = {}

    class SummaryWriter(object):
        def __init__(self, *args, **kw):
            pass

        def add_scalar(self, tag, value, step=None):
            scalars.setdefault(tag, []).append(float(value))

        def close(self):
            pass
    sys.modules["tensorboardX"].SummaryWriter = SummaryWriter

    if a.check_ckpt:
        sys.path.insert(0, REFDIR)
        import torch
        from mpnn import MPNN                          # the reference's module
        from malib.spaces import Box
        ck = torch.load(a.check_ckpt, map_location="cpu")
        assert set(ck) == {"models", "ob_rms"} and ck["ob_rms"] == (None, None) and len(ck["models"]) == ng + na
        out = {"keys": len(ck["models"][0]), "params": int(sum(v.numel() for v in ck["models"][0].values()))}
        for sd, n, m in ((ck["models"][0], ng, na), (ck["models"][-1], na, ng)):
            net = MPNN(input_size=6, num_agents=n, num_opp_agents=m, num_entities=0, action_space=Box(0., 1., (8,)),
                       pos_index=2, mask_dist=None, entity_mp=False)
            net.load_state_dict(sd)                    # strict: names and shapes must be the reference's
            with torch.no_grad():
                v, act, lp, _ = net.act(torch.randn(n * 2, 6), None, torch.randn(m * 2, 6), None)
            assert bool(torch.isfinite(v).all()) and bool(torch.isfinite(lp).all())
        print(json.dumps(out))
        return

    # import order decides who provides gym_fortattack / rlcore / mpnn
    sys.path.insert(0, REFDIR)
    if a.env == "ours" and a.rl == "ours":
        sys.path.insert(0, PKG)                       # our directory shadows the three names, nothing else
    elif a.env == "ours":
        sys.path.insert(0, PKG)
        import gym_fortattack                          # noqa: F401  ours; pinned in sys.modules ...
        import gym_fortattack.fortattack               # noqa: F401
        sys.path.remove(PKG)                           # ... while rlcore / mpnn fall through to the reference
    elif a.rl == "ours":
        import gym_fortattack                          # noqa: F401  the reference's env
        import gym_fortattack.fortattack               # noqa: F401
        sys.path.insert(0, PKG)
    if a.env == "ref" and (ng, na) != (5, 5):
        # trim the hard-coded 5v5 world (tests/golden/ref_shim.make_ref_env)
        import gym_fortattack.fortattack as gf
        def make_trimmed(num_steps, benchmark=False):
            env, _ = ref_shim.make_ref_env(ng, na, num_steps)
            return env
        gf.make_fortattack_env = make_trimmed

    sys.argv = [os.path.join(REFDIR, a.script)] + rest
    t0 = time.time()
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):              # the script prints per step / per episode
        runpy.run_path(os.path.join(REFDIR, a.script), run_name="__main__")
    dt = time.time() - t0
    import gym_fortattack.fortattack as gf_used
    import mpnn as mpnn_used
    import rlcore.algo as algo_used
    save_dir = None
    for i, tok in enumerate(rest):
        if tok == "--save-dir":
            save_dir = os.path.join("marlsave", rest[i + 1])
    ckpts = sorted(f for f in os.listdir(save_dir)) if save_dir and os.path.isdir(save_dir) else []
    out = {"scalars": scalars, "seconds": dt, "save_dir": os.path.abspath(save_dir) if save_dir else None, "files": ckpts,
           "env_module": os.path.relpath(gf_used.__file__, ROOT), "mpnn_module": os.path.relpath(mpnn_used.__file__, ROOT),
           "algo_module": os.path.relpath(algo_used.__file__, ROOT),
           "fps_lines": [l for l in buf.getvalue().splitlines() if l.startswith("Updates ")][-2:]}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
