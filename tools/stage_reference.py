"""Stages the UNMODIFIED upstream reference into the git-ignored baseline/_ref/ (build container only).

    python tools/stage_reference.py            # copies /root/reference/*.py (+ model_file/one_layer checkpoints)

baseline/_ref/ is listed in .gitignore (reference sources never enter the history) but not in .gpurunignore, so it
travels to the GPU box with the repository snapshot. It is used by
  * tests/test_gpu_reference_driver.py: the reference's own CPPO_main.py (train_pursuer_network / test_network,
    CPPO_main.py:94-161, 233-282) imported unmodified and run on top of the drop-in modules, and
  * bench.py's cpu_baseline "reference" rows: the reference's own Python env step / RK4 / actor / GAE / update timed on the
    box's host cores (C-env-1, C-rk4-scalar, C-rk4-batch, C-actor, C-gae, C-update).
Nothing in the product package reads it; when the directory is absent the test skips and the rows are omitted.
__graft_entry__.build() calls this when /root/reference is present.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("SAT_REFERENCE_DIR", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")


def stage(verbose=True):
    if not os.path.isfile(os.path.join(SRC, "CPPO_main.py")):
        if verbose:
            print("reference tree not present at", SRC, "- nothing staged")
        return None
    os.makedirs(DST, exist_ok=True)
    manifest = {}
    for name in sorted(os.listdir(SRC)):
        p = os.path.join(SRC, name)
        if os.path.isfile(p) and name.endswith(".py"):
            shutil.copyfile(p, os.path.join(DST, name))
            manifest[name] = hashlib.sha256(open(p, "rb").read()).hexdigest()
    sub = os.path.join(SRC, "single_pluse_model")          # environment.py:5 imports single_pluse_model.real_time_data_process
    if os.path.isdir(sub):
        os.makedirs(os.path.join(DST, "single_pluse_model"), exist_ok=True)
        for name in sorted(os.listdir(sub)):
            if name.endswith(".py"):
                shutil.copyfile(os.path.join(sub, name), os.path.join(DST, "single_pluse_model", name))
                manifest["single_pluse_model/" + name] = hashlib.sha256(open(os.path.join(sub, name), "rb").read()).hexdigest()
    ck = os.path.join(SRC, "model_file", "one_layer")
    if os.path.isdir(ck):
        os.makedirs(os.path.join(DST, "model_file", "one_layer"), exist_ok=True)
        for name in sorted(os.listdir(ck)):
            shutil.copyfile(os.path.join(ck, name), os.path.join(DST, "model_file", "one_layer", name))
            manifest["model_file/one_layer/" + name] = hashlib.sha256(open(os.path.join(ck, name), "rb").read()).hexdigest()
    json.dump({"source": SRC, "files": manifest}, open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1)
    if verbose:
        print(f"staged {len(manifest)} files into {DST}")
    return DST


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
