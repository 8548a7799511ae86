"""TEST / MEASUREMENT INFRASTRUCTURE ONLY -- stages the UNMODIFIED reference for the GPU box.

The reference (kyle-he/gym-comm) is pure Python; `/root/reference` exists only in the build container.
`gpurun` ships untracked files (that is how the built `.so` travels), so this script copies the files
the hot path imports -- byte for byte, nothing edited -- into the git-ignored `oracle/_ref/`:

    gym_comm/**.py                         OvercookedMultiEnv (gym_comm/envs/overcooked_env.py:15-297)
    gym_cooking/**.py (minus misc/)        OvercookedEnvironment + utils + planners it imports
    gym_cooking/utils/levels/*.txt         the 19 level files (overcooked_environment.py:103)
    pantheonrl/common/*.py                 MultiAgentEnv / SimultaneousEnv base classes

`oracle/ref_harness.py` then finds the reference at `oracle/_ref` when `/root/reference` is absent,
and `bench.py`'s CPU arm (`cpu_baseline`, `--impl reference`) times it as `kind: "reference"`.
Nothing under `oracle/_ref/` is tracked by git, imported by the product package, or edited.

    python -m oracle.stage_ref            # build container; also run by __graft_entry__.build()
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SOURCE = "/root/reference"

TREES = [("gym_comm", (".py",)), ("gym_cooking", (".py", ".txt")), ("pantheonrl/common", (".py",))]
SKIP_DIRS = ("gym_cooking/misc",)


def stage(source: str = SOURCE, dest: str = DEST, quiet: bool = False) -> dict:
    if not os.path.isdir(os.path.join(source, "gym_cooking")):
        raise FileNotFoundError("%s does not hold the reference" % source)
    manifest = {}
    for tree, exts in TREES:
        top = os.path.join(source, tree)
        for root, dirs, files in os.walk(top):
            rel_root = os.path.relpath(root, source)
            if any(rel_root == s or rel_root.startswith(s + os.sep) for s in SKIP_DIRS):
                dirs[:] = []
                continue
            for f in sorted(files):
                if not f.endswith(exts):
                    continue
                rel = os.path.join(rel_root, f)
                out = os.path.join(dest, rel)
                os.makedirs(os.path.dirname(out), exist_ok=True)
                src = os.path.join(root, f)
                with open(src, "rb") as fh:
                    data = fh.read()
                manifest[rel] = hashlib.sha256(data).hexdigest()
                if not (os.path.exists(out) and open(out, "rb").read() == data):
                    shutil.copyfile(src, out)
    with open(os.path.join(dest, "MANIFEST.json"), "w") as fh:
        json.dump({"source": source, "files": manifest}, fh, indent=0, sort_keys=True)
    if not quiet:
        print("staged %d reference files into %s" % (len(manifest), dest))
    return manifest


def verify(dest: str = DEST) -> bool:
    """True iff every staged file still has the hash recorded when it was copied (unmodified)."""
    try:
        man = json.load(open(os.path.join(dest, "MANIFEST.json")))["files"]
    except Exception:
        return False
    for rel, h in man.items():
        try:
            if hashlib.sha256(open(os.path.join(dest, rel), "rb").read()).hexdigest() != h:
                return False
        except OSError:
            return False
    return bool(man)


if __name__ == "__main__":
    stage(sys.argv[1] if len(sys.argv) > 1 else SOURCE)
