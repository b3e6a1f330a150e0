"""Stage the UNMODIFIED reference tree into oracle/_ref/reference (git-ignored, but it travels to the GPU box
with the gpurun snapshot like the built .so files) so that

  * `bench.py --impl reference` / `cpu_baseline` can time the reference's own CPU implementation of the path
    (`cpu_baseline.kind = "reference"`), and
  * tests/test_gpu_level1.py can run the unmodified `test_fullframework.main()` with this package patched in.

Nothing is edited; only Python sources and YAML configs are staged (the reference is pure Python: there is
nothing to compile). Run in the build container, where /root/reference exists:

    python oracle/stage_reference.py
"""
from __future__ import annotations

import hashlib
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("MOCHA_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "oracle", "_ref", "reference")
KEEP_EXT = (".py", ".yaml", ".yml", ".md", ".sh", ".txt")


def staged_path():
    """Path of the staged tree, or None when it has not been staged."""
    return DST if os.path.exists(os.path.join(DST, "test_fullframework.py")) else None


def reference_root():
    """The reference tree to import from: the live one in the build container, else the staged copy."""
    if os.path.exists(os.path.join(SRC, "test_fullframework.py")):
        return SRC
    return staged_path()


def stage(force: bool = False) -> str | None:
    if not os.path.exists(os.path.join(SRC, "test_fullframework.py")):
        return staged_path()
    if force and os.path.exists(DST):
        shutil.rmtree(DST)
    digest = hashlib.sha256()
    n = 0
    for d, dirs, files in os.walk(SRC):
        dirs[:] = sorted(x for x in dirs if x not in (".git", "__pycache__"))
        for f in sorted(files):
            if not f.endswith(KEEP_EXT) and f != "LICENSE":
                continue
            s = os.path.join(d, f)
            rel = os.path.relpath(s, SRC)
            t = os.path.join(DST, rel)
            os.makedirs(os.path.dirname(t), exist_ok=True)
            with open(s, "rb") as fh:
                data = fh.read()
            digest.update(rel.encode())
            digest.update(data)
            if not os.path.exists(t) or open(t, "rb").read() != data:
                with open(t, "wb") as fh:
                    fh.write(data)
            n += 1
    with open(os.path.join(os.path.dirname(DST), "MANIFEST"), "w") as fh:
        fh.write(f"staged from {SRC}: {n} files, sha256 {digest.hexdigest()}\n")
    return DST


if __name__ == "__main__":
    p = stage(force="--force" in sys.argv)
    print(p or "reference tree not available")
