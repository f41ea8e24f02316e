#!/usr/bin/env python
"""Install the unmodified reference into baseline/_ref (offline, no dependencies resolved):

    python baseline/install_ref.py

= ``pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --no-deps --target baseline/_ref <copy>``
from a copy of /root/reference under /tmp (the source tree is read-only and setuptools writes build files into it).
--no-deps because matplotlib / pyvista / tifffile / scikit-image / ipython are not in the wheelhouse; the solver path
does not use them (baseline/shim).  A no-op when /root/reference is absent (the GPU box) or baseline/_ref exists.
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference"
REF_DIR = os.path.join(HERE, "_ref")


def install(force=False):
    if os.path.isfile(os.path.join(REF_DIR, "taufactor", "taufactor.py")) and not force:
        return REF_DIR
    if not os.path.isdir(REF_SRC):
        return None
    tmp = tempfile.mkdtemp(prefix="taufactor_ref_")
    try:
        src = os.path.join(tmp, "reference")
        shutil.copytree(REF_SRC, src)
        if os.path.isdir(REF_DIR):
            shutil.rmtree(REF_DIR)
        subprocess.check_call([sys.executable, "-m", "pip", "install", "-q", "--no-index", "--no-build-isolation",
                               "--find-links", "/opt/wheelhouse", "--no-deps", "--target", REF_DIR, src])
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return REF_DIR


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))
