#!/usr/bin/env python
"""Stage the UNMODIFIED reference checkout for the GPU box.

/root/reference exists only in the authoring container; gpurun ships /root/repo (git-ignored files included).  The
live-model overlay tests (tests/test_reference_overlay.py) and `bench.py --raft` need the reference's own `optical_flow`
package and `methods/raft/model` on the box, so this copies them -- byte for byte, nothing edited -- into
`baseline/_ref/reference/`, which `.gitignore` keeps out of history (the sanctioned place for a reference install;
it is the baseline the overlay is compared with, never product source).  `__graft_entry__.build()` calls this when
/root/reference is present; without it the overlay tests skip with a message that says what is missing.
"""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref", "reference")
TREES = ("optical_flow", os.path.join("methods", "raft", "model"), os.path.join("tests", "operator"))


def stage(src: str = SRC, dst: str = DST) -> bool:
    if not os.path.isdir(src):
        return False
    manifest = {}
    for tree in TREES:
        s, d = os.path.join(src, tree), os.path.join(dst, tree)
        if os.path.isdir(d):
            shutil.rmtree(d)
        shutil.copytree(s, d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        for base, _, files in os.walk(d):
            for f in sorted(files):
                p = os.path.join(base, f)
                manifest[os.path.relpath(p, dst)] = hashlib.sha256(open(p, "rb").read()).hexdigest()
    with open(os.path.join(dst, "MANIFEST.json"), "w") as fh:
        json.dump({"source": src, "files": manifest}, fh, indent=1, sort_keys=True)
    return True


if __name__ == "__main__":
    ok = stage()
    print(f"staged {SRC} -> {DST}" if ok else f"{SRC} not present: nothing staged")
    sys.exit(0)
