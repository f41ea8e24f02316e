"""The UNMODIFIED reference (tldr-group/taufactor v1.2.1) as a baseline / parity oracle.

``baseline/_ref`` is a plain ``pip install --target`` of /root/reference (see ``install_ref.py``; git-ignored,
shipped to the GPU box by gpurun).  ``load_reference()`` imports it -- or /root/reference itself when that exists --
with three empty stand-in packages (``baseline/shim``: IPython, matplotlib, skimage) for third-party modules the
reference imports but never touches on the solver path and that this image does not have.  No reference file is
modified or copied into the repository's history.

Test / bench infrastructure only: nothing under ``taufactor_b200/`` imports this.
"""
import importlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
SHIM_DIR = os.path.join(HERE, "shim")


def reference_path():
    """Directory that holds the reference's ``taufactor`` package, or None (TAUB_NO_REFERENCE=1 hides it)."""
    if os.environ.get("TAUB_NO_REFERENCE") == "1":
        return None
    for cand in (REF_DIR, "/root/reference"):
        if os.path.isfile(os.path.join(cand, "taufactor", "taufactor.py")):
            return cand
    return None


def load_reference():
    """The reference's ``taufactor`` module (unmodified), or None when it is not available here."""
    mod = sys.modules.get("taufactor")
    if mod is not None:
        return mod
    path = reference_path()
    if path is None:
        return None
    for pkg in ("IPython", "matplotlib", "skimage"):      # only stand in for what is really missing
        try:
            importlib.import_module(pkg)
        except Exception:
            if SHIM_DIR not in sys.path:
                sys.path.append(SHIM_DIR)
    sys.path.insert(0, path)
    try:
        return importlib.import_module("taufactor")
    except Exception:
        return None
    finally:
        sys.path.remove(path)
