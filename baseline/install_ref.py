"""Install the UNMODIFIED reference into baseline/_ref/ so that it travels to the GPU box.

    python baseline/install_ref.py            (also run by __graft_entry__.build() when /root/reference exists)

The reference (Ankur-Deka/Emergent-Multiagent-Strategies) is plain Python without setup.py / pyproject.toml,
so `pip install --target baseline/_ref /root/reference` cannot work (pip: "neither 'setup.py' nor
'pyproject.toml' found"; the attempt and its message are recorded in baseline/_ref/INSTALL.json).  The
install is therefore a byte copy of the files the hot path and its callers need:

    baseline/_ref/reference/   *.py of the top level (train_fortattack*.py, learner.py, rlagent.py, mpnn.py,
                               arguments.py, utils.py, eval.py, test_fortattack*.py), gym_fortattack/ (without the
                               6.7 MB image archive), rlcore/, multiagent/, malib/spaces + malib/core (the only
                               malib modules gym_fortattack/fortattack.py:6 imports), marlsave/tmp_1/ep*.pt and
                               marlsave/tmp_2/ep5050.pt (the shipped checkpoints), out_files/1.npy
    baseline/_ref/ref_shim.py  copy of tests/golden/ref_shim.py: stub modules for gym / pygame / pyglet /
                               gym_vecenv / tensorboardX, none of which does arithmetic on the path
                               

baseline/_ref is git-ignored and NOT gpurun-ignored.  Nothing of the product imports it: users are
bench.py's reference arm / cpu_baseline leg and tests/ (config-1 and checkpoint tests).
"""
import hashlib
import json
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
DST = os.path.join(HERE, "_ref")
SRC = os.environ.get("FA_REFERENCE_DIR", "/root/reference")

TOP_DIRS = ("gym_fortattack", "rlcore", "multiagent")
SKIP_NAMES = {"__pycache__", "Pygame-Images.zip", "Game", "logs"}
EXTRA = ("malib/__init__.py", "malib/spaces", "malib/core", "marlsave/tmp_1", "marlsave/tmp_2/ep5050.pt",
         "marlsave/tmp_2/params.json", "out_files/1.npy", "requirements.txt", "LICENSE")


def _copy(src, dst):
    if os.path.isdir(src):
        for name in sorted(os.listdir(src)):
            if name in SKIP_NAMES or name.endswith(".pyc"):
                continue
            _copy(os.path.join(src, name), os.path.join(dst, name))
    else:
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)


def _digest(root):
    h, n = hashlib.sha256(), 0
    for d, dirs, files in os.walk(root):
        dirs.sort()
        for f in sorted(files):
            p = os.path.join(d, f)
            h.update(os.path.relpath(p, root).encode())
            with open(p, "rb") as fh:
                h.update(fh.read())
            n += 1
    return h.hexdigest()[:16], n


def try_pip():
    """The contract's install command; expected to fail (no setup.py) -- outcome recorded, not fatal."""
    cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--find-links", "/opt/wheelhouse",
           "--target", os.path.join(DST, "_pip"), SRC]
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
        msg = (out.stderr.strip().splitlines() or out.stdout.strip().splitlines() or [""])[-1]
        return {"cmd": " ".join(cmd[2:]), "rc": out.returncode, "last_line": msg[:300]}
    except Exception as e:      # pragma: no cover
        return {"cmd": " ".join(cmd[2:]), "rc": -1, "last_line": repr(e)[:300]}


def install(force=False, pip_attempt=True):
    if not os.path.isdir(SRC):
        if os.path.isdir(os.path.join(DST, "reference")):
            return DST                                   # GPU box: already installed, nothing to copy from
        raise RuntimeError("reference tree %s not present and baseline/_ref is empty" % SRC)
    ref = os.path.join(DST, "reference")
    meta_path = os.path.join(DST, "INSTALL.json")
    if os.path.isdir(ref) and os.path.exists(meta_path) and not force:
        return DST
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(ref)
    pip = try_pip() if pip_attempt else None
    shutil.rmtree(os.path.join(DST, "_pip"), ignore_errors=True)
    for name in sorted(os.listdir(SRC)):
        if name.endswith(".py"):
            _copy(os.path.join(SRC, name), os.path.join(ref, name))
    for d in TOP_DIRS + EXTRA:
        _copy(os.path.join(SRC, d), os.path.join(ref, d))
    shutil.copyfile(os.path.join(ROOT, "tests", "golden", "ref_shim.py"), os.path.join(DST, "ref_shim.py"))
    digest, n = _digest(ref)
    with open(meta_path, "w") as f:
        json.dump({"source": SRC, "files": n, "sha256_16": digest, "pip_attempt": pip,
                   "method": "byte copy of the unmodified reference files (no setup.py/pyproject.toml to pip-install)"}, f, indent=1)
    return DST


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))
    print(open(os.path.join(DST, "INSTALL.json")).read())
