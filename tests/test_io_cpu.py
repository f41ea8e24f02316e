"""TIFF ingestion (taufactor_b200.io.imread): the reader against Pillow / libtiff on files written
here in every layout it claims to support, against hand-assembled files for the layouts Pillow cannot write
(BigTIFF, big endian, tiles, planar samples, ImageJ contiguous stacks), and against committed fixtures."""
import io
import os
import struct
import zlib

import numpy as np
import pytest

from taufactor_b200.io import TiffError, imread

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def labels(shape, seed=0):
    rng = np.random.default_rng(seed)
    return (rng.integers(0, 3, size=shape) * 85).astype(np.uint8)


# ----------------------------------------------------------------------------- against Pillow (libtiff)
def pil_bytes(vol, **kw):
    from PIL import Image
    ims = [Image.fromarray(p) for p in vol]
    b = io.BytesIO()
    ims[0].save(b, format="TIFF", save_all=True, append_images=ims[1:], **kw)
    return b.getvalue()


@pytest.mark.parametrize("compression", ["raw", "tiff_deflate", "tiff_adobe_deflate", "packbits", "tiff_lzw"])
@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.int32, np.float32])
def test_reads_what_libtiff_writes(compression, dtype):
    rng = np.random.default_rng(7)
    vol = (rng.random((4, 33, 27)) * 200).astype(dtype)          # noisy: long LZW tables, width changes
    got = imread(pil_bytes(vol, compression=compression))
    assert got.dtype == vol.dtype and np.array_equal(got, vol)


def test_label_volume_and_single_page():
    vol = labels((6, 20, 31))
    assert np.array_equal(imread(pil_bytes(vol, compression="tiff_lzw")), vol)
    one = imread(pil_bytes(vol[:1]))
    assert one.shape == (20, 31) and np.array_equal(one, vol[0])


def test_horizontal_predictor():
    vol = np.cumsum(labels((2, 40, 64), 3), axis=2, dtype=np.uint8)
    data = pil_bytes(vol, compression="tiff_lzw", tiffinfo={317: 2})
    assert np.array_equal(imread(data), vol)


def test_bilevel_and_rgb():
    rng = np.random.default_rng(1)
    mask = rng.random((3, 19, 21)) < 0.5
    got = imread(pil_bytes(mask))
    assert got.dtype == np.bool_ and np.array_equal(got, mask)
    rgb = (rng.random((2, 9, 11, 3)) * 255).astype(np.uint8)
    assert np.array_equal(imread(pil_bytes(rgb, compression="tiff_lzw")), rgb)


def test_file_name_and_file_object(tmp_path):
    vol = labels((3, 8, 9))
    path = tmp_path / "v.tif"
    path.write_bytes(pil_bytes(vol))
    assert np.array_equal(imread(str(path)), vol)
    with open(path, "rb") as fh:
        assert np.array_equal(imread(fh), vol)


# ----------------------------------------------------------------------------- hand-assembled layouts
def build_tiff(pages, bo="<", big=False, compress=None, tile=None, planar=False, description=None):
    """Minimal TIFF writer for the tests: ``pages`` = list of [H, W] or [H, W, S] arrays of one dtype."""
    pages = [p if p.ndim == 3 else p[..., None] for p in pages]
    H, W, S = pages[0].shape
    dt = pages[0].dtype
    fmt = {"u": 1, "i": 2, "f": 3}[dt.kind]
    osz, ofmt = (8, "Q") if big else (4, "I")
    out = bytearray(b"II" if bo == "<" else b"MM")
    out += struct.pack(bo + "H", 43 if big else 42)
    if big:
        out += struct.pack(bo + "HH", 8, 0)
    first_ptr = len(out)
    out += struct.pack(bo + ofmt, 0)
    enc = (lambda b: zlib.compress(b)) if compress == 8 else (lambda b: b)
    prev_ptr = first_ptr
    for n, page in enumerate(pages):
        chunks = []
        planes = [page[..., s:s + 1] for s in range(S)] if planar else [page]
        for pl in planes:
            if tile:
                tl, tw = tile
                for r in range(0, H, tl):
                    for c in range(0, W, tw):
                        t = np.zeros((tl, tw, pl.shape[2]), dt)
                        blk = pl[r:r + tl, c:c + tw]
                        t[: blk.shape[0], : blk.shape[1]] = blk
                        chunks.append(enc(t.astype(dt.newbyteorder(bo)).tobytes()))
            else:
                rps = max(1, H // 3)
                for r in range(0, H, rps):
                    chunks.append(enc(pl[r:r + rps].astype(dt.newbyteorder(bo)).tobytes()))
        offs = []
        for ch in chunks:
            offs.append(len(out))
            out += ch
        entries = []           # (tag, type, values)
        entries += [(256, 3, [W]), (257, 3, [H]), (258, 3, [8 * dt.itemsize] * S), (259, 3, [compress or 1]),
                    (262, 3, [1]), (277, 3, [S]), (284, 3, [2 if planar else 1]), (339, 3, [fmt] * S)]
        lt = 16 if big else 4
        if tile:
            entries += [(322, 3, [tile[1]]), (323, 3, [tile[0]]), (324, lt, offs), (325, lt, [len(c) for c in chunks])]
        else:
            entries += [(273, lt, offs), (278, 3, [max(1, H // 3)]), (279, lt, [len(c) for c in chunks])]
        if description and n == 0:
            entries.append((270, 2, description.encode() + b"\0"))
        entries.sort()
        tsz = {2: 1, 3: 2, 4: 4, 16: 8}
        tfm = {3: "H", 4: "I", 16: "Q"}
        blobs = []
        for tag, typ, vals in entries:
            raw = bytes(vals) if typ == 2 else struct.pack(bo + tfm[typ] * len(vals), *vals)
            blobs.append((tag, typ, len(vals), raw))
        for i, (tag, typ, cnt, raw) in enumerate(blobs):      # out-of-line values go before the IFD
            if len(raw) > osz:
                if len(out) % 2:
                    out += b"\0"
                blobs[i] = (tag, typ, cnt, struct.pack(bo + ofmt, len(out)))
                out += raw
        if len(out) % 2:
            out += b"\0"
        ifd_at = len(out)
        struct.pack_into(bo + ofmt, out, prev_ptr, ifd_at)
        out += struct.pack(bo + ("Q" if big else "H"), len(blobs))
        for tag, typ, cnt, raw in blobs:
            out += struct.pack(bo + "HH" + ofmt, tag, typ, cnt)[: 4 + osz] + raw.ljust(osz, b"\0")
        prev_ptr = len(out)
        out += struct.pack(bo + ofmt, 0)
    return bytes(out)


@pytest.mark.parametrize("bo", ["<", ">"])
@pytest.mark.parametrize("big", [False, True])
@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.float32])
def test_byte_orders_and_bigtiff(bo, big, dtype):
    rng = np.random.default_rng(5)
    vol = (rng.random((3, 14, 10)) * 250).astype(dtype)
    for compress in (None, 8):
        assert np.array_equal(imread(build_tiff(list(vol), bo=bo, big=big, compress=compress)), vol)


def test_tiles_and_planar_samples():
    rng = np.random.default_rng(6)
    vol = (rng.random((2, 37, 45)) * 250).astype(np.uint16)
    assert np.array_equal(imread(build_tiff(list(vol), tile=(16, 16), compress=8)), vol)
    rgb = (rng.random((2, 12, 20, 3)) * 250).astype(np.uint8)
    assert np.array_equal(imread(build_tiff(list(rgb), planar=True)), rgb)
    assert np.array_equal(imread(build_tiff(list(rgb), planar=True, tile=(16, 16), bo=">")), rgb)


def test_imagej_contiguous_stack():
    vol = labels((5, 12, 8), 9)
    desc = "ImageJ=1.53t\nimages=5\nslices=5\nloop=false\n"
    got = imread(imagej_file(vol, desc))
    assert np.array_equal(got, vol)
    # the same description on a file that does carry one IFD per plane is read page by page
    assert np.array_equal(imread(build_tiff(list(vol), description=desc)), vol)


def imagej_file(vol, desc):
    """What ImageJ writes for a stack: header, ONE IFD (one strip), description with images=N, then all planes."""
    n, H, W = vol.shape
    d = desc.encode() + b"\0"
    entries = [(256, 4, 1, W), (257, 4, 1, H), (258, 3, 1, 8), (259, 3, 1, 1), (262, 3, 1, 1), (270, 2, len(d), None),
               (273, 4, 1, None), (277, 3, 1, 1), (278, 3, 1, H), (279, 4, 1, H * W)]
    ifd_at = 8
    ifd_len = 2 + 12 * len(entries) + 4
    desc_at = ifd_at + ifd_len
    data_at = desc_at + len(d) + (len(d) % 2)
    out = bytearray(b"MM" + struct.pack(">HI", 42, ifd_at))
    out += struct.pack(">H", len(entries))
    for tag, typ, cnt, val in entries:
        if tag == 270:
            val = desc_at
        if tag == 273:
            val = data_at
        out += struct.pack(">HHI", tag, typ, cnt)
        out += struct.pack(">HH", val, 0) if typ == 3 else struct.pack(">I", val)
    out += struct.pack(">I", 0)
    out += d + (b"\0" if len(d) % 2 else b"")
    assert len(out) == data_at
    out += vol.tobytes()
    return bytes(out)


# ----------------------------------------------------------------------------- errors and fixtures
def test_rejects_what_it_does_not_understand():
    with pytest.raises(TiffError):
        imread(b"not a tiff at all")
    good = build_tiff([labels((4, 4))[0:4]])
    with pytest.raises(TiffError):
        imread(good[: len(good) // 2])                       # truncated
    jpeg = build_tiff([np.zeros((4, 4), np.uint8)], compress=7)
    with pytest.raises(TiffError, match="compression"):
        imread(jpeg)
    assert issubclass(TiffError, ValueError)


def test_damaged_files_raise_tifferror_only():
    """Byte flips and truncation of valid files: the reader either returns an array or raises TiffError."""
    rng = np.random.default_rng(0)
    vol = labels((3, 9, 11), seed=11)
    seeds = [pil_bytes(vol, compression=c) for c in ("raw", "tiff_lzw", "packbits", "tiff_adobe_deflate")]
    seeds += [build_tiff(list(vol), bo=">", compress=8), build_tiff(list(vol), big=True), build_tiff(list(vol), tile=(16, 16)),
              imagej_file(vol, "ImageJ=1.53t\nimages=3\n")]
    for it in range(4000):
        b = bytearray(seeds[it % len(seeds)])
        for _ in range(rng.integers(1, 4)):
            b[rng.integers(0, len(b))] = rng.integers(0, 256)
        if rng.random() < 0.1:
            b = b[: rng.integers(0, len(b))]
        try:
            imread(bytes(b))
        except TiffError:
            pass


@pytest.mark.parametrize("name", ["tiny_lzw.tif", "tiny_deflate_be.tif", "tiny_packbits.tif"])
def test_committed_fixtures(name):
    """Files written once by tests/golden/make_tiff_fixtures.py (Pillow / the writer above) and kept in the repo:
    the expected volume is regenerated from its seed."""
    vol = labels((4, 9, 11), seed=11)
    assert np.array_equal(imread(os.path.join(GOLDEN, name)), vol)


# ----------------------------------------------------------------------------- native decoders (csrc/taub_tiff.cu)
def _strips(data):
    """(compressed strip, decoded size) pairs of a TIFF written by the helpers above."""
    import taufactor_b200.io as tio
    buf = memoryview(data)
    ifds, bo = tio._read_ifds(buf)
    out = []
    for t in ifds:
        W, H, rps = t[256][0], t[257][0], t.get(278, (t[257][0],))[0]
        bps = t[258][0] // 8 * t.get(277, (1,))[0]
        for k, (off, cnt) in enumerate(zip(t[273], t[279])):
            rows = min(rps, H - k * rps)
            out.append((bytes(buf[off:off + cnt]), rows * W * bps))
    return out


@pytest.mark.parametrize("compression,fn,code", [("packbits", "taub_unpackbits", 32773), ("tiff_lzw", "taub_unlzw", 5)])
def test_native_decoders_equal_the_python_ones(compression, fn, code):
    import taufactor_b200.io as tio
    if tio._native() is None:
        pytest.skip("libtaub200.so not built")
    rng = np.random.default_rng(4)
    vols = [labels((3, 40, 50), 2), (rng.random((2, 64, 96)) * 255).astype(np.uint8),          # runs / noise (12-bit codes)
            np.zeros((1, 300, 300), np.uint8), (rng.random((1, 128, 257)) * 65535).astype(np.uint16)]
    for v in vols:
        for strip, n in _strips(pil_bytes(v, compression=compression)):
            py = (tio._unpackbits if code == 32773 else tio._unlzw)(strip, n)
            assert tio._native_decode(fn, strip, n) == py and len(py) == n
            # a destination smaller than the stream: both stop when it is full
            assert tio._native_decode(fn, strip, n // 3) == py[: n // 3]
    # end to end through imread, both ways
    v = vols[1]
    data = pil_bytes(v, compression=compression)
    try:
        tio.USE_NATIVE = False
        a = imread(data)
    finally:
        tio.USE_NATIVE = True
    assert np.array_equal(a, v) and np.array_equal(imread(data), v)


def test_native_decoders_survive_damaged_streams():
    import taufactor_b200.io as tio
    if tio._native() is None:
        pytest.skip("libtaub200.so not built")
    rng = np.random.default_rng(5)
    v = (rng.random((1, 64, 96)) * 255).astype(np.uint8)
    for compression, fn, py in (("packbits", "taub_unpackbits", tio._unpackbits), ("tiff_lzw", "taub_unlzw", tio._unlzw)):
        strip, n = _strips(pil_bytes(v, compression=compression))[0]
        for it in range(3000):
            b = bytearray(strip)
            for _ in range(rng.integers(1, 6)):
                b[rng.integers(0, len(b))] = rng.integers(0, 256)
            if it % 7 == 0:
                b = b[: rng.integers(0, len(b))]
            try:
                got = tio._native_decode(fn, bytes(b), n)
            except TiffError:
                got = None
            try:
                want = py(bytes(b), n)
            except TiffError:
                want = None
            assert (got is None) == (want is None) and (got is None or bytes(got) == want), (compression, it)


def test_deflate_strip_cannot_expand_beyond_its_declared_size():
    """A crafted Deflate strip (1000 : 1) is cut off at the strip's declared size: TiffError, and the decoder never holds
    more than that size -- not the 64 MB the stream would inflate to."""
    import tracemalloc
    import zlib
    from taufactor_b200 import io as tio
    bomb = zlib.compress(b"\0" * (64 << 20), 9)
    assert len(bomb) < 100_000
    tracemalloc.start()
    with pytest.raises(tio.TiffError):
        tio._decompress(bomb, 8, 4096)
    peak = tracemalloc.get_traced_memory()[1]
    tracemalloc.stop()
    assert peak < (2 << 20), peak
    good = zlib.compress(bytes(range(256)) * 16)
    assert tio._decompress(good, 8, 4096) == bytes(range(256)) * 16
