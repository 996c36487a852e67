"""Install the UNMODIFIED reference package under baseline/_ref (git-ignored, NOT gpurun-ignored: it travels to the
GPU box with the snapshot).  Test / benchmark infrastructure: nothing under theia_b200/ imports it.

    python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
        --target baseline/_ref <copy of /root/reference>

(--no-deps: the reference pins tensorflow / hydra / a webdataset fork that are absent here and that the model
classes do not need; the install is done from a /tmp copy because /root/reference is read-only.)"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
SRC = "/root/reference"


def present() -> bool:
    return os.path.exists(os.path.join(REF, "theia", "models", "rvfm.py"))


def ensure(force: bool = False) -> bool:
    """True when baseline/_ref holds the reference package (installing it if the source tree is available)."""
    if present() and not force:
        return True
    if not os.path.exists(os.path.join(SRC, "pyproject.toml")):
        return False
    tmp = tempfile.mkdtemp(prefix="theia_ref_src_")
    try:
        src = os.path.join(tmp, "reference")
        shutil.copytree(SRC, src, ignore=shutil.ignore_patterns("media", ".git"))
        if os.path.isdir(REF):
            shutil.rmtree(REF)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
               "--find-links", "/opt/wheelhouse", "--target", REF, src]
        r = subprocess.run(cmd, capture_output=True, text=True, cwd=tmp)
        if r.returncode != 0:
            raise RuntimeError("pip install of the reference failed:\n" + r.stdout[-2000:] + r.stderr[-2000:])
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return present()


if __name__ == "__main__":
    print("baseline/_ref present:", ensure(force="--force" in sys.argv))
