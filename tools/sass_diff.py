"""Which kernels changed between two revisions?  Builds csrc/ of a git revision (default HEAD~1) and of the working
tree with the product's nvcc flags and compares the SASS instruction stream of every kernel (encodings and
addresses ignored, trailing template bools normalised).  Runs without a GPU: a refactor whose kernels come out
"identical" needs no new parity run.
    python tools/sass_diff.py [REV]"""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from taufactor_b200.build import NVCC_FLAGS, nvcc  # noqa: E402


def kernels(so):
    out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    d, cur = {}, None
    for line in out.splitlines():
        ls = line.strip()
        if ls.startswith("Function :"):
            cur = re.sub(r"ELb[01]EEEv", "EEEv", ls.split(":", 1)[1].strip())
            while cur in d:
                cur += "#"
            d[cur] = []
        elif cur and re.match(r"/\*[0-9a-f]{4}\*/", ls):
            ins = re.sub(r"/\* 0x[0-9a-f]+ \*/", "", ls)
            d[cur].append(re.sub(r"\s+", " ", re.sub(r"^/\*[0-9a-f]+\*/", "", ins)).strip())
    return d


def build_rev(rev, tmp):
    subprocess.run(f"git -C {ROOT} archive {rev} taufactor_b200/csrc include | tar -x -C {tmp}", shell=True, check=True)
    so = os.path.join(tmp, "rev.so")
    src = os.path.join(tmp, "taufactor_b200", "csrc")
    subprocess.check_call([nvcc()] + NVCC_FLAGS + ["-I", os.path.join(tmp, "include"), "-I", src]
                          + sorted(os.path.join(src, f) for f in os.listdir(src) if f.endswith(".cu")) + ["-o", so])
    return so


def main():
    rev = sys.argv[1] if len(sys.argv) > 1 else "HEAD~1"
    from taufactor_b200 import build as tb
    new = kernels(tb.build())
    with tempfile.TemporaryDirectory() as tmp:
        old = kernels(build_rev(rev, tmp))
    # jump targets move when code in front of them grows: compare with branch targets masked as well
    mask = lambda L: [re.sub(r"0x[0-9a-f]+", "ADDR", x) if re.search(r"\b(BRA|BSSY|CALL|JMP)\b", x) else x for x in L]
    for name, body in old.items():
        cands = [k for k in new if k.rstrip("#") == name.rstrip("#")]
        if any(new[k] == body for k in cands):
            verdict = "identical"
        elif any(mask(new[k]) == mask(body) for k in cands):
            verdict = "identical up to branch targets"
        else:
            verdict = "CHANGED" if cands else "REMOVED"
        print(f"{len(body):6d}  {verdict:32s} {name[:120]}")
    for k in new:
        if not any(k.rstrip("#") == n.rstrip("#") for n in old):
            print(f"{len(new[k]):6d}  {'NEW':32s} {k[:120]}")


if __name__ == "__main__":
    main()
