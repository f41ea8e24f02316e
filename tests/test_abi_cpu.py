"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, computes the storage geometry, and the Python surface raises the reference's errors."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from taufactor_b200 import build, _lib
    build.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    from taufactor_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "taub200.h")).read()
    declared = set(re.findall(r"\b(taub_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in taub200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes prototype"
    assert lib.taub_abi_version() == 13


def test_struct_layout_matches_header(lib):
    import ctypes
    from taufactor_b200 import _lib
    assert ctypes.sizeof(_lib.Geom) == 10 * 4 + 2 * 8
    assert _lib.Problem.field.offset == ctypes.sizeof(_lib.Geom) + 8
    assert ctypes.sizeof(_lib.Problem) == ctypes.sizeof(_lib.Geom) + 8 + 5 * 8 + 8 + 8 + 4 * 8 + 8 + 8 + 8
    assert _lib.Problem.stop.offset == ctypes.sizeof(_lib.Problem) - 8 * 8
    assert _lib.Problem.sync_ws.offset == ctypes.sizeof(_lib.Problem) - 3 * 8
    assert _lib.Problem.redo_ws.offset == ctypes.sizeof(_lib.Problem) - 8


@pytest.mark.parametrize("shape", [(1, 512, 512, 512), (3, 11, 13, 9), (2, 30, 28, 1), (1, 8, 8, 2049)])
def test_geometry(lib, shape):
    from taufactor_b200 import _lib
    g = _lib.Geom()
    bs, Nx, Ny, Nz = shape
    assert lib.taub_geom_init(g, bs, Nx, Ny, Nz, Nx, 0, 1) == 0
    assert (g.planes, g.rows) == (Nx + 4, Ny + 4)
    assert g.pitch % 32 == 0 and g.pitch >= Nz + 8 and g.pitch < Nz + 40
    assert g.plane_stride == g.rows * g.pitch and g.image_stride == g.planes * g.plane_stride
    assert lib.taub_field_elems(g) == bs * g.image_stride
    assert lib.taub_codes_elems(g) * 4 == lib.taub_field_elems(g)
    assert lib.taub_sums_ws_bytes(g) >= 16 * bs * Nx


def test_geometry_rejects_bad_slab(lib):
    from taufactor_b200 import _lib
    g = _lib.Geom()
    assert lib.taub_geom_init(g, 1, 10, 4, 4, 8, 0, 0) == _lib.ERR_ARG
    assert b"slab" in lib.taub_last_error()
    assert lib.taub_geom_init(g, 0, 10, 4, 4, 10, 0, 0) == _lib.ERR_ARG


def test_reference_error_conventions():
    """ref: taufactor.py:196-203 (TypeError / ValueError on the image), :391-397 (binary labels),
    :539-546 (diffusivities); raised before any device work, like the reference."""
    import taufactor_b200 as tau
    with pytest.raises(TypeError):
        tau.Solver([[0, 1], [1, 0]])
    with pytest.raises(ValueError):
        tau.Solver(np.ones((2, 2, 2, 2, 2)))
    with pytest.raises(ValueError, match="only contain 0s and 1s"):
        tau.Solver(np.full((4, 4, 4), 2))
    with pytest.raises(ValueError):
        tau.MultiPhaseSolver(np.zeros([6, 6, 6]), {0: 1.0, 1: -0.1})
    with pytest.raises(TypeError):
        tau.MultiPhaseSolver(np.zeros([6, 6, 6]), [1.0])
    with pytest.raises(TypeError):
        tau.MultiPhaseSolver(np.zeros([6, 6, 6]), {"a": 1.0})


def test_no_cpu_fallback():
    import torch
    import taufactor_b200 as tau
    with pytest.raises(RuntimeError, match="CUDA"):
        tau.Solver(np.ones((4, 4, 4)), device="cpu")
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            tau.Solver(np.ones((4, 4, 4)))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "taufactor_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_harmonic_table_matches_oracle():
    from taufactor_b200.solvers import MultiPhaseSolver
    from oracle import sor_numpy as orc
    D = np.array([0.0, 1.0, 0.3, 2.0, 0.5, 1e-3, 0.0], np.float32)
    t = MultiPhaseSolver.harmonic_table(D)
    ref = orc.harmonic_mean(np.repeat(D[:, None], len(D), 1), np.repeat(D[None, :], len(D), 0))
    assert np.array_equal(t, ref) and np.array_equal(t, t.T)


def test_face_conductance_tensors_equal_the_reference_state():
    """MultiPhaseSolver.D_x / D_y / D_z / factor (rebuilt on demand) against the tensors of the unmodified reference
    (tests/golden/multiphase_state.npz, taufactor.py:585-604 and :626-650), bit for bit -- the host logic on CPU tensors."""
    import cases
    from taufactor_b200.solvers import face_conductance_tensors, _expand_to_4d
    gold = np.load(os.path.join(ROOT, "tests", "golden", "multiphase_state.npz"))
    for name in ("odd3_mp", "odd3_pmp", "ref_mp_batched", "ref_pmp_slanted_odd"):
        cls, build, ckw, _, _ = cases.CASES[name]
        got = face_conductance_tensors(_expand_to_4d(build()), ckw["diffusivities"], cls.startswith("Periodic"), "cpu")
        for a, t in zip(("D_x", "D_y", "D_z", "factor"), got):
            assert np.array_equal(t.numpy(), gold[f"{name}@{a}"]), (name, a)


def test_python_surface_matches_reference_signatures():
    """Drop-in check against tests/golden/api.json (generated from the unmodified reference by
    tests/golden/make_golden.py): same constructor / solve() parameter names, order and defaults, same
    class hierarchy names."""
    import inspect
    import json
    import taufactor_b200 as tau
    api = json.load(open(os.path.join(ROOT, "tests", "golden", "api.json")))
    for cls, ref in api.items():
        if cls.startswith("solved_attributes"):
            continue
        C = getattr(tau, cls)
        for meth in ("init", "solve"):
            f = C.__init__ if meth == "init" else C.solve
            got = [[n, None if p.default is inspect._empty else repr(p.default)]
                   for n, p in inspect.signature(f).parameters.items() if n != "self"]
            for (gn, gd), (rn, rd) in zip(got, ref[meth]):
                assert gn == rn, (cls, meth, gn, rn)
                if rn != "device":        # the reference's AnisotropicSolver default is a torch.device object
                    assert gd == rd, (cls, meth, gn, gd, rd)
            assert len(got) == len(ref[meth]), (cls, meth)
        names = [b.__name__ for b in C.__mro__[1:-1]]
        assert [n for n in ref["bases"] if n != "ABC"] == names, (cls, names)


def test_header_is_c99_and_layout_matches_ctypes(lib, tmp_path):
    """A plain C99 caller: the header compiles with gcc -std=c99 -pedantic, links against the library, and the C
    compiler's struct layout (sizeof / offsetof) is the one the ctypes binding assumes."""
    import ctypes
    import shutil
    import subprocess
    from taufactor_b200 import _lib
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = tmp_path / "caller.c"
    src.write_text(r'''
#include <stddef.h>
#include <stdio.h>
#include "taub200.h"
int main(void)
{
    taub_geom g;
    if (taub_abi_version() != TAUB_ABI_VERSION) return 2;
    if (taub_geom_init(&g, 2, 30, 28, 1, 30, 0, 1) != TAUB_OK) return 3;
    if (taub_geom_init(&g, 0, 30, 28, 1, 30, 0, 1) != TAUB_ERR_ARG || taub_last_error()[0] == 0) return 4;
    taub_geom_init(&g, 2, 30, 28, 1, 30, 0, 1);
    printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(taub_geom), sizeof(taub_problem),
           offsetof(taub_geom, plane_stride), offsetof(taub_problem, kind), offsetof(taub_problem, field),
           offsetof(taub_problem, codes), offsetof(taub_problem, lut), offsetof(taub_problem, omega),
           offsetof(taub_problem, stop), offsetof(taub_problem, peer_lo), offsetof(taub_problem, peer_hi),
           offsetof(taub_problem, sync_ws), offsetof(taub_problem, sync_epoch), offsetof(taub_problem, redo_ws));
    printf("%d %d %d %lld %lld %zu\n", g.planes, g.rows, g.pitch, (long long)g.plane_stride, (long long)g.image_stride,
           taub_field_elems(&g));
    return 0;
}
''')
    exe = tmp_path / "caller"
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                           str(src), "-o", str(exe), "-L", libdir, "-l:libtaub200.so", f"-Wl,-rpath,{libdir}"])
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split("\n")
    P, Gm = _lib.Problem, _lib.Geom
    expect = [ctypes.sizeof(Gm), ctypes.sizeof(P), Gm.plane_stride.offset, P.kind.offset, P.field.offset, P.codes.offset,
              P.lut.offset, P.omega.offset, P.stop.offset, P.peer_lo.offset, P.peer_hi.offset, P.sync_ws.offset,
              P.sync_epoch.offset, P.redo_ws.offset]
    assert [int(x) for x in out[0].split()] == expect
    g = Gm()
    assert lib.taub_geom_init(g, 2, 30, 28, 1, 30, 0, 1) == 0
    assert [int(x) for x in out[1].split()] == [g.planes, g.rows, g.pitch, g.plane_stride, g.image_stride,
                                                lib.taub_field_elems(g)]
