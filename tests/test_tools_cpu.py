"""The measurement helpers under tools/ are only ever run on the GPU box: at least make sure they parse."""
import glob
import os
import py_compile
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(ROOT, "tools", "*.py")) + [os.path.join(ROOT, "bench.py"),
                                                                                        os.path.join(ROOT, "__graft_entry__.py")]))
def test_python_helpers_compile(path, tmp_path):
    py_compile.compile(path, cfile=str(tmp_path / "out.pyc"), doraise=True)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(ROOT, "tools", "*.sh"))))
def test_shell_helpers_parse(path):
    if shutil.which("bash") is None:
        pytest.skip("no bash")
    subprocess.check_call(["bash", "-n", path])
    # every script a helper calls exists
    for line in open(path):
        for tok in line.split():
            if tok.startswith("tools/") and tok.endswith((".py", ".sh")):
                assert os.path.exists(os.path.join(ROOT, tok)), tok
