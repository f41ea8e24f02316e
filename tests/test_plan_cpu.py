"""CPU checks of the fused pass's launch plan (taub_fused_plan: pure host arithmetic, no CUDA call): tile shape, the
two plane-chunk models, the strided chunk numbering and the cluster width, on the shapes the measurements in
profiles/r2_chunks_*.txt were made on -- and, as a property over many shapes, that the chunks cover every plane of
the pass exactly once whatever the numbering."""
import ctypes
import itertools
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from taufactor_b200 import _lib  # noqa: E402

BINARY, CLASS, ANISO = 0, 3, 2
KEYS = ("LR", "OR", "OG", "tiles_j", "tiles_k", "chunk_len", "chunks", "grid_rows", "perm_R", "perm_S", "cluster", "elastic")


@pytest.fixture(scope="module")
def lib():
    return _lib.load()


def plan(lib, shape, kind=BINARY, periodic=False, planes=None, resident=296, L=300):
    bs, Nx, Ny, Nz = shape
    g = _lib.Geom()
    assert lib.taub_geom_init(g, bs, Nx, Ny, Nz, Nx, 0, int(periodic)) == 0
    p = _lib.Problem()
    p.g, p.kind, p.L = g, kind, L
    dummy = ctypes.c_void_p(4096)          # never dereferenced: the plan only looks at extents and kinds
    p.field[0], p.field[1], p.codes, p.lut = dummy, dummy, dummy, dummy
    out = (ctypes.c_int32 * 12)()
    lo, hi = planes or (0, Nx)
    rc = lib.taub_fused_plan(p, lo, hi, resident, out)
    assert rc == 0, lib.taub_last_error()
    return dict(zip(KEYS, out))


def chunk_of(f, y):
    return (y % f["perm_R"]) * f["perm_S"] + y // f["perm_R"] if f["perm_R"] > 1 else y


def test_headline_volume_takes_short_chunks_a_strided_numbering_and_clusters(lib):
    f = plan(lib, (1, 512, 512, 512))
    assert (f["LR"], f["OR"], f["OG"], f["tiles_j"], f["tiles_k"]) == (30, 26, 32, 20, 4)
    assert f["elastic"] == 1 and f["cluster"] == 2
    assert (f["chunk_len"], f["chunks"]) == (26, 20)
    # 80 tiles per chunk row, 296 slots: four chunk rows in flight, five chunks apart
    assert (f["perm_R"], f["perm_S"], f["grid_rows"]) == (4, 5, 20)
    assert [chunk_of(f, y) for y in range(8)] == [0, 5, 10, 15, 1, 6, 11, 16]


def test_volume_in_l2_takes_the_list_model_with_a_short_last_chunk(lib):
    f = plan(lib, (1, 256, 256, 256))
    assert f["elastic"] == 0 and f["perm_R"] == 1 and f["grid_rows"] == f["chunks"]
    # 20 tiles x 15 chunks = 300 CTAs on 296 slots: the last chunk is 4 planes long and ends with the others
    assert (f["tiles_j"] * f["tiles_k"], f["chunk_len"], f["chunks"]) == (20, 18, 15)
    assert 256 - 14 * 18 == 4


def test_class_and_anisotropic_kinds_keep_the_list_model_and_plain_launches(lib):
    for kind in (CLASS, ANISO):
        f = plan(lib, (1, 512, 512, 512), kind=kind)
        assert f["elastic"] == 0 and f["cluster"] == 0 and f["perm_R"] == 1, kind
        assert f["tiles_j"] * f["tiles_k"] * f["chunks"] <= 2 * 296, kind     # whole waves, not many short CTAs


def test_large_passes_fill_the_device_with_one_chunk_row(lib):
    for shape, planes in (((1, 1024, 1024, 1024), None), ((1, 1024, 2048, 2048), (2, 1022))):
        f = plan(lib, shape, planes=planes)
        assert f["elastic"] == 1 and f["perm_R"] == 1 and f["grid_rows"] == f["chunks"], shape
        assert 22 <= f["chunk_len"] <= 30, (shape, f)


def test_boundary_planes_of_a_slab_are_one_short_chunk(lib):
    f = plan(lib, (1, 1024, 2048, 2048), planes=(0, 2))
    assert (f["chunk_len"], f["chunks"], f["elastic"]) == (2, 1, 0)


@pytest.mark.parametrize("model", ["0", "1", None])
def test_chunks_cover_every_plane_once_under_every_numbering(lib, monkeypatch, model):
    if model is None:
        monkeypatch.delenv("TAUB_CHUNK_MODEL", raising=False)
    else:
        monkeypatch.setenv("TAUB_CHUNK_MODEL", model)
    rng = np.random.default_rng(5)
    shapes = [(1, 2, 8, 8), (1, 3, 30, 28), (2, 49, 63, 50), (1, 100, 100, 100), (8, 384, 384, 384), (1, 320, 320, 320),
              (1, 448, 448, 448), (1, 640, 640, 640), (1, 768, 768, 768), (3, 513, 70, 1030)]
    shapes += [(1,) + tuple(int(v) for v in rng.integers(2, 700, size=3)) for _ in range(40)]
    for shape, resident in itertools.product(shapes, (296, 264, 2)):
        Nx = shape[1]
        lo = int(rng.integers(0, max(1, Nx // 3)))
        hi = int(rng.integers(lo + 1, Nx + 1))
        for planes in ((0, Nx), (lo, hi)):
            f = plan(lib, shape, planes=planes, resident=resident)
            n, cl, ch = planes[1] - planes[0], f["chunk_len"], f["chunks"]
            assert cl % 2 == 0 and cl >= 2 and (ch - 1) * cl < n <= ch * cl, (shape, planes, f)
            assert f["grid_rows"] >= ch and f["grid_rows"] - ch < max(f["perm_R"], 1), (shape, planes, f)
            got = sorted(c for c in (chunk_of(f, y) for y in range(f["grid_rows"])) if c * cl < n)
            assert got == list(range(ch)), (shape, planes, f)
            assert f["cluster"] in (0, 2) and (f["cluster"] == 0 or (f["tiles_j"] * f["tiles_k"]) % 2 == 0)


def test_measurement_switches_override_the_plan(lib, monkeypatch):
    monkeypatch.setenv("TAUB_FUSED_CHUNKS", "7")
    f = plan(lib, (1, 512, 512, 512))
    assert (f["chunk_len"], f["chunks"]) == (74, 7) and f["perm_R"] == 4 and f["grid_rows"] == 8
    monkeypatch.setenv("TAUB_FUSED_PERM", "0")
    f = plan(lib, (1, 512, 512, 512))
    assert f["perm_R"] == 1 and f["grid_rows"] == 7
