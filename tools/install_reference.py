#!/usr/bin/env python
"""Put the UNMODIFIED reference (researchmm/MM-Diffusion) under baseline/_ref/ so that the reference arms of bench.py
and the script-level drop-in tests can run it on the GPU box, where /root/reference does not exist.

baseline/_ref/ is git-ignored (reference sources never enter this repository's history) but NOT gpurun-ignored, so the
tree travels with the snapshot like the built libmmdiff.so does.

Recipe (the task contract's one offline install):
  1. `python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref <src>`
     from a /tmp copy (the source tree is read-only).  The reference ships no setup.py / pyproject.toml, so pip has
     nothing to build; that outcome is recorded in baseline/_ref/INSTALL.json and DESIGN.md.
  2. Fallback that always works for a pure-Python tree: copy the package directories byte for byte
     (mm_diffusion/, py_scripts/, ssh_scripts/, evaluations/ minus binary assets) and record the sha256 of every file.

    python tools/install_reference.py [--src /root/reference] [--force]
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "baseline", "_ref")
TREES = ["mm_diffusion", "py_scripts", "ssh_scripts", "evaluations"]
KEEP_EXT = {".py", ".sh", ".txt", ".md", ".yaml", ".yml", ".json", ".cfg"}


def try_pip(src: str) -> dict:
    tmp = tempfile.mkdtemp(prefix="mmd_ref_")
    work = os.path.join(tmp, "src")
    shutil.copytree(src, work, ignore=shutil.ignore_patterns(".git", "fig"))
    cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links",
           "/opt/wheelhouse", "--target", DEST, work]
    res = subprocess.run(cmd, capture_output=True, text=True)
    shutil.rmtree(tmp, ignore_errors=True)
    tail = (res.stdout + res.stderr).strip().splitlines()[-3:]
    return {"cmd": " ".join(cmd[:-1] + ["<tmp copy of the reference>"]), "returncode": res.returncode, "tail": tail}


def copy_tree(src: str) -> dict:
    files = {}
    for tree in TREES:
        s = os.path.join(src, tree)
        if not os.path.isdir(s):
            continue
        for dirpath, _, names in os.walk(s):
            for n in names:
                if os.path.splitext(n)[1] not in KEEP_EXT:
                    continue
                p = os.path.join(dirpath, n)
                rel = os.path.relpath(p, src)
                d = os.path.join(DEST, rel)
                os.makedirs(os.path.dirname(d), exist_ok=True)
                shutil.copyfile(p, d)
                with open(p, "rb") as f:
                    files[rel] = hashlib.sha256(f.read()).hexdigest()
    return files


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    ap.add_argument("--force", action="store_true")
    args = ap.parse_args()
    marker = os.path.join(DEST, "INSTALL.json")
    if os.path.exists(marker) and not args.force:
        print(f"{DEST} already present (use --force to redo)")
        return 0
    if not os.path.isdir(os.path.join(args.src, "mm_diffusion")):
        print(f"reference not found at {args.src}; nothing installed", file=sys.stderr)
        return 1
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    os.makedirs(DEST)
    pip = try_pip(args.src)
    pip_ok = pip["returncode"] == 0 and os.path.isdir(os.path.join(DEST, "mm_diffusion"))
    files = {}
    if not pip_ok:
        # pip left nothing usable behind (no build metadata in the reference): plain byte-for-byte copy
        for n in os.listdir(DEST):
            p = os.path.join(DEST, n)
            shutil.rmtree(p) if os.path.isdir(p) else os.remove(p)
        files = copy_tree(args.src)
    with open(marker, "w") as f:
        json.dump({"source": args.src, "pip": pip, "method": "pip" if pip_ok else "copy", "n_files": len(files),
                   "sha256": files}, f, indent=1)
    print(f"installed the reference into {DEST} via {'pip' if pip_ok else 'copy'} ({len(files)} files)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
