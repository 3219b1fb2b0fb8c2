"""
ORACLE SIDE (test / bench infrastructure): stage the UNMODIFIED reference package for the CPU arm.

    python oracle/build_ref.py        # /root/reference/simfire/**/*.py -> oracle/_ref/simfire/

The reference (mitrefireline/simfire, pure Python) is mounted read-only at /root/reference in the
dev container and does not exist on the GPU box.  Its Python sources are copied, byte for byte,
into ``oracle/_ref/`` -- listed in .gitignore (no reference source enters the history) but not in
.gpurunignore (the copy travels with the snapshot, like the built .so) -- together with a
MANIFEST.json of sha256 sums, so that ``bench.py --impl reference`` / ``cpu_baseline`` can time
``RothermelFireManager.update`` (simfire/game/managers/fire.py:616) itself on the GPU box's host
cores, and the same-run parity checks compare the device with the reference, not with a port.

A ``pip install --target`` of the reference is not possible here: its build backend (poetry-core)
and most of its dependencies (pygame, landfire, geopandas, ...) are not in the offline wheelhouse;
the hot path needs none of them (oracle/ref_shim.py registers inert stand-ins).  Tests, assets and
data files are not copied (the path under test needs none).

``__graft_entry__.build()`` calls ``stage()`` when /root/reference is present.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/simfire"
DST = os.path.join(HERE, "_ref", "simfire")


def stage(force: bool = False) -> str | None:
    """Copy the package's .py files; returns the staged root or None when the reference is not mounted."""
    if not os.path.isdir(SRC):
        return None
    manifest_path = os.path.join(HERE, "_ref", "MANIFEST.json")
    if os.path.exists(manifest_path) and not force:
        return os.path.dirname(DST)
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    manifest = {}
    for root, dirs, files in os.walk(SRC):
        dirs[:] = [d for d in dirs if d not in ("_tests", "__pycache__")]  # .py files only: images and data files stay behind
        for f in files:
            if not f.endswith(".py"):
                continue
            src = os.path.join(root, f)
            rel = os.path.relpath(src, SRC)
            dst = os.path.join(DST, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(src, dst)
            with open(src, "rb") as fh:
                manifest[rel] = hashlib.sha256(fh.read()).hexdigest()
    with open(manifest_path, "w") as fh:
        json.dump({"source": SRC, "files": manifest}, fh, indent=1, sort_keys=True)
    return os.path.dirname(DST)


if __name__ == "__main__":
    out = stage(force=True)
    print(f"staged {out}" if out else f"{SRC} is not mounted: nothing staged")
