"""TEST INFRASTRUCTURE ONLY — recipe for ``oracle/_ref/``: the UNMODIFIED reference files of the hot path, copied verbatim
from ``/root/reference`` so that ``bench.py --impl reference`` / ``cpu_baseline`` can time the reference itself on the GPU
box (where ``/root/reference`` does not exist) instead of the port.

``oracle/_ref/`` is a build output: git-ignored (never part of the history, like the compiled ``.so``), NOT
gpurun-ignored (it travels with the snapshot).  ``__graft_entry__.build()`` runs this when ``/root/reference`` is
present.  Nothing under the product package imports it; ``oracle/ref_harness.py`` loads it through the same two shims as
the original (stub torchrec KJT, ``cuda:0`` -> cpu).

Files (SURVEY.md §8(c) "files that need importing"), paths relative to the reference root:
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
SRC = os.environ.get("TT_REFERENCE_SRC", "/root/reference")

FILES = [
    "src/towers/two_tower_train_task.py",
    "src/towers/two_tower_model.py",
    "src/towers/cat_embed.py",
    "src/towers/tower/base_tower.py",
    "src/towers/tower/notice_tower.py",
    "src/towers/tower/company_tower.py",
    "src/torchrec_preprocess/schema.py",
    "data/column_classifier.py",
    "meta/metadata.csv",
    "meta/notice_categorical_config.json",
    "meta/company_categorical_config.json",
]
PACKAGES = ["src", "src/towers", "src/towers/tower", "src/torchrec_preprocess", "data"]


def make(verbose: bool = True) -> bool:
    if not os.path.isfile(os.path.join(SRC, FILES[0])):
        if verbose:
            print(f"[make_ref] {SRC} not present: keeping whatever is in {DST}")
        return False
    for rel in FILES:
        src = os.path.join(SRC, rel)
        if not os.path.isfile(src):
            continue
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
    for pkg in PACKAGES:  # package markers exactly where the reference has them
        if os.path.isfile(os.path.join(SRC, pkg, "__init__.py")):
            d = os.path.join(DST, pkg)
            os.makedirs(d, exist_ok=True)
            shutil.copyfile(os.path.join(SRC, pkg, "__init__.py"), os.path.join(d, "__init__.py"))
    if verbose:
        print(f"[make_ref] {len(FILES)} reference files -> {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if make() else 1)
