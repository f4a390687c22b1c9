#!/usr/bin/env python
"""Container only: place the UNMODIFIED reference where it travels to the GPU box.

    python scripts/vendor_reference.py            # /root/reference/PriOr-RAFT -> baseline/_ref/PriOr-RAFT
    python scripts/vendor_reference.py --check    # verify baseline/_ref against baseline/ref_manifest.json

The reference is a pure-Python repo without setup.py / pyproject.toml, so the "install" of the base contract
(`pip install --target baseline/_ref /root/reference`) has nothing to build: the install IS a byte-for-byte copy of
its .py files.  `baseline/_ref/` is git-ignored (no reference source enters the history) but not gpurun-ignored, so the
copy rides along with the snapshot.  `baseline/ref_manifest.json` (committed: file names + sha256 only) lets the GPU-side
tests and `bench.py` prove that what they import is the unmodified reference.
"""
import argparse
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("PRIORFLOW_REFERENCE_SRC", "/root/reference/PriOr-RAFT")
DST = os.path.join(ROOT, "baseline", "_ref", "PriOr-RAFT")
MANIFEST = os.path.join(ROOT, "baseline", "ref_manifest.json")


def py_files(base):
    out = []
    for d, _, files in os.walk(base):
        for f in files:
            if f.endswith(".py"):
                out.append(os.path.relpath(os.path.join(d, f), base))
    return sorted(out)


def sha(path):
    with open(path, "rb") as fh:
        return hashlib.sha256(fh.read()).hexdigest()


def check(base=DST):
    """True when every file of the manifest exists under `base` with the recorded hash."""
    if not os.path.isfile(MANIFEST):
        return False
    with open(MANIFEST) as fh:
        man = json.load(fh)["files"]
    return all(os.path.isfile(os.path.join(base, rel)) and sha(os.path.join(base, rel)) == h for rel, h in man.items())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    if a.check:
        ok = check()
        print("baseline/_ref matches the manifest" if ok else "baseline/_ref is missing or differs from the manifest")
        sys.exit(0 if ok else 1)
    if not os.path.isdir(os.path.join(SRC, "core")):
        sys.exit(f"reference not found at {SRC}")
    files = py_files(SRC)
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    for rel in files:
        os.makedirs(os.path.dirname(os.path.join(DST, rel)), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, rel), os.path.join(DST, rel))
    man = {"source": "longliangLiu/PriOr-Flow, PriOr-RAFT/ (read-only mount /root/reference)", "files": {rel: sha(os.path.join(SRC, rel)) for rel in files}}
    os.makedirs(os.path.dirname(MANIFEST), exist_ok=True)
    with open(MANIFEST, "w") as fh:
        json.dump(man, fh, indent=1, sort_keys=True)
    print(f"vendored {len(files)} files into {DST}; manifest {MANIFEST}")


if __name__ == "__main__":
    main()
