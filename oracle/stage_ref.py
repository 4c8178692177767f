"""Stage the UNMODIFIED reference module into oracle/_ref/ -- TEST / BASELINE INFRASTRUCTURE ONLY.

The reference's hot path is one pure-Python file, /root/reference/KDE_distance.py (147 lines over the sklearn /
scipy wheels); there is nothing to compile.  `stage()` copies that file byte for byte from where it lies into
oracle/_ref/ (git-ignored: it never enters this repository's history, but it travels to the GPU box with the
snapshot), records its sha256, and bench.py's `--impl reference` / `cpu_baseline` legs then time the reference
itself (`kind: "reference"`) instead of its port.  Called by __graft_entry__.build() in the build container; on
the GPU box /root/reference does not exist and the staged copy is simply used as found.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")
SOURCE = os.path.join(os.environ.get("KDEJS_REFERENCE_DIR", "/root/reference"), "KDE_distance.py")


def stage() -> bool:
    """Copy the reference file if the reference tree is present; True when a staged copy exists afterwards."""
    dst = os.path.join(REF_DIR, "KDE_distance.py")
    if os.path.exists(SOURCE):
        os.makedirs(REF_DIR, exist_ok=True)
        shutil.copyfile(SOURCE, dst)
        with open(dst, "rb") as f:
            digest = hashlib.sha256(f.read()).hexdigest()
        with open(os.path.join(REF_DIR, "STAGED.json"), "w") as f:
            json.dump({"source": SOURCE, "sha256": digest, "note": "unmodified copy, staged by oracle/stage_ref.py"}, f)
    return os.path.exists(dst)


if __name__ == "__main__":
    print("staged" if stage() else "reference not present; nothing staged")
