"""TIFF ingestion for the solvers ("formats").

Every notebook of the reference and its README (README.md:51-54) load the segmented volume with
``tifffile.imread(path)`` and hand the array to ``Solver``.  ``tifffile`` is a third-party package that is
not part of this image, so this module reads the TIFF flavours those volumes come in with NumPy + zlib only:

* classic TIFF and BigTIFF, little or big endian, any number of pages (one IFD per x-plane);
* strips or tiles, chunky or planar samples, 8 / 16 / 32 / 64-bit unsigned, signed and float samples, and
  1-bit (bilevel) images, returned as ``bool``;
* compression: none, Deflate (8 and the old 32946 -- what tifffile wrote ``docs/notebooks/electrode.tiff``
  with), PackBits, LZW (MSB-first, early change); horizontal predictor;
* ImageJ hyperstacks whose planes follow the first page contiguously (``images=N`` in the description: how
  ImageJ / Fiji store stacks, and the only form they use above 4 GB).

PackBits and LZW are decoded by the C routines of ``libtaub200.so`` (``csrc/taub_tiff.cu``) when the library is
built, by the interpreter loops below otherwise.  ``imread`` returns what ``tifffile.imread`` returns for these files: ``[pages, height, width]`` (a single
page gives ``[height, width]``; several samples per pixel add a trailing axis), in the file's sample type.
Host-side input decoding only; nothing of it is on the solve path.
"""
from __future__ import annotations

import re
import struct
import zlib

import numpy as np

__all__ = ["imread", "TiffError"]


class TiffError(ValueError):
    """The file is not a TIFF this reader understands."""


# tag ids (TIFF 6.0)
_WIDTH, _LENGTH, _BITS, _COMPRESSION, _PHOTOMETRIC, _DESCRIPTION = 256, 257, 258, 259, 262, 270
_STRIP_OFFSETS, _SAMPLES, _ROWS_PER_STRIP, _STRIP_COUNTS, _PLANAR = 273, 277, 278, 279, 284
_PREDICTOR, _TILE_W, _TILE_L, _TILE_OFFSETS, _TILE_COUNTS, _SAMPLE_FORMAT = 317, 322, 323, 324, 325, 339

# field type -> (struct code, size)
_TYPES = {1: ("B", 1), 2: ("c", 1), 3: ("H", 2), 4: ("I", 4), 5: ("II", 8), 6: ("b", 1), 7: ("B", 1), 8: ("h", 2),
          9: ("i", 4), 10: ("ii", 8), 11: ("f", 4), 12: ("d", 8), 13: ("I", 4), 16: ("Q", 8), 17: ("q", 8), 18: ("Q", 8)}


def _read_ifds(buf):
    """All image file directories of the file as ``{tag: tuple of values}`` dicts, in file order."""
    if len(buf) < 8 or buf[:2] not in (b"II", b"MM"):
        raise TiffError("not a TIFF file (bad byte-order mark)")
    bo = "<" if buf[:2] == b"II" else ">"
    magic = struct.unpack_from(bo + "H", buf, 2)[0]
    if magic == 42:
        big, (off,) = False, struct.unpack_from(bo + "I", buf, 4)
    elif magic == 43:
        big, (off,) = True, struct.unpack_from(bo + "Q", buf, 8)
    else:
        raise TiffError(f"not a TIFF file (magic {magic})")
    n_fmt, e_fmt, e_size, inline = (("Q", "HHQ", 20, 8) if big else ("H", "HHI", 12, 4))
    ifds, seen = [], set()
    while off:
        if off in seen or off + struct.calcsize(n_fmt) > len(buf):
            raise TiffError("corrupt IFD chain")
        seen.add(off)
        (n,) = struct.unpack_from(bo + n_fmt, buf, off)
        pos = off + struct.calcsize(n_fmt)
        tags = {}
        for _ in range(n):
            tag, typ, count = struct.unpack_from(bo + e_fmt, buf, pos)
            code, size = _TYPES.get(typ, (None, 0))
            if code is not None:
                nbytes = size * count
                at = pos + e_size - inline
                if nbytes > inline:
                    (at,) = struct.unpack_from(bo + ("Q" if big else "I"), buf, at)
                if at + nbytes > len(buf):
                    raise TiffError(f"tag {tag} points outside the file")
                if typ == 2:
                    tags[tag] = bytes(buf[at:at + nbytes]).split(b"\0")[0].decode("latin-1")
                elif typ in (5, 10):
                    v = struct.unpack_from(bo + code[0] * (2 * count), buf, at)
                    tags[tag] = tuple(v[2 * k] / v[2 * k + 1] if v[2 * k + 1] else 0.0 for k in range(count))
                else:
                    tags[tag] = tuple(np.frombuffer(buf, dtype=np.dtype(bo + _np_code(code, size)), count=count, offset=at)
                                      .tolist())
            pos += e_size
        ifds.append(tags)
        (off,) = struct.unpack_from(bo + ("Q" if big else "I"), buf, pos)
    if not ifds:
        raise TiffError("TIFF file without an image directory")
    return ifds, bo


def _np_code(code, size):
    return {"B": "u1", "b": "i1", "H": "u2", "h": "i2", "I": "u4", "i": "i4", "Q": "u8", "q": "i8", "f": "f4",
            "d": "f8"}[code]


def _one(tags, tag, default=None):
    v = tags.get(tag)
    if v is None:
        if default is None:
            raise TiffError(f"required TIFF tag {tag} is missing")
        return default
    return v[0] if isinstance(v, tuple) else v


def _sample_dtype(tags, bo):
    bits = tags.get(_BITS, (1,))
    if len(set(bits)) != 1:
        raise TiffError(f"samples of different widths are not supported: {bits}")
    fmt = _one(tags, _SAMPLE_FORMAT, 1)
    b = bits[0]
    if b == 1:
        return None, 1
    kind = {1: "u", 2: "i", 3: "f", 4: "u"}.get(fmt)
    if kind is None or b not in (8, 16, 32, 64) or (kind == "f" and b < 16):
        raise TiffError(f"unsupported sample type: {b} bits, SampleFormat {fmt}")
    return np.dtype(f"{bo}{kind}{b // 8}"), b


# ----------------------------------------------------------------------------- decompression
def _unpackbits(data, expected):
    """PackBits (TIFF 6.0 section 9): n in 0..127 copies n+1 literal bytes, n in 129..255 repeats the next
    byte 257-n times, 128 is a no-op."""
    out = bytearray()
    i, n = 0, len(data)
    while i < n and len(out) < expected:
        c = data[i]
        i += 1
        if c < 128:
            out += data[i:i + c + 1]
            i += c + 1
        elif c > 128:
            out += data[i:i + 1] * (257 - c)
            i += 1
    return bytes(out[:expected])


def _unlzw(data, expected):
    """TIFF LZW (section 13): MSB-first variable-width codes 9..12 bits, ClearCode 256, EndOfInformation 257,
    "early change" of the code width (the width grows one code before the table is full)."""
    out = bytearray()
    table = [bytes([i]) for i in range(256)] + [b"", b""]
    bitbuf = nbits = 0
    width, prev = 9, None
    pos, n = 0, len(data)
    while len(out) < expected:
        while nbits < width and pos < n:
            bitbuf = (bitbuf << 8) | data[pos]
            pos += 1
            nbits += 8
        if nbits < width:
            break
        code = (bitbuf >> (nbits - width)) & ((1 << width) - 1)
        nbits -= width
        bitbuf &= (1 << nbits) - 1
        if code == 257:
            break
        if code == 256:
            del table[258:]
            width, prev = 9, None
            continue
        if prev is None:
            if code >= 256:
                raise TiffError("corrupt LZW stream")
            entry = table[code]
        elif code < len(table):
            entry = table[code]
            if len(table) < 4096:
                table.append(prev + entry[:1])
        elif code == len(table) and len(table) < 4096:
            entry = prev + prev[:1]
            table.append(entry)
        else:
            raise TiffError("corrupt LZW stream")
        out += entry
        prev = entry
        if len(table) >= (1 << width) - 1 and width < 12:
            width += 1
    return bytes(out[:expected])


def _native():
    """The C decoders of libtaub200.so (csrc/taub_tiff.cu: taub_unpackbits / taub_unlzw, ~100x the interpreter
    loops below), or None when the library is not built -- reading a file does not need a GPU."""
    global _NATIVE
    if _NATIVE is False:
        try:
            from . import _lib
            _NATIVE = _lib.load()
        except (ImportError, OSError, AttributeError):
            _NATIVE = None
    return _NATIVE


_NATIVE = False
USE_NATIVE = True      # False: always use the pure-Python decoders (tests compare the two)


def _native_decode(fn_name, data, expected):
    import ctypes
    lib = _native()
    src = np.frombuffer(data, np.uint8)
    dst = np.empty(expected, np.uint8)
    n = getattr(lib, fn_name)(src.ctypes.data_as(ctypes.c_void_p), src.size, dst.ctypes.data_as(ctypes.c_void_p), expected)
    if n < 0:
        raise TiffError(lib.taub_last_error().decode(errors="replace"))
    return dst[:n].tobytes() if n < expected else dst.data


def _decompress(chunk, compression, expected):
    if compression == 1:
        return chunk
    if compression in (8, 32946):
        # bounded inflate: a strip may expand to its declared size and no further (a crafted strip could otherwise
        # allocate ~1000x its size before any length check sees it)
        d = zlib.decompressobj()
        out = d.decompress(bytes(chunk), max(int(expected), 1))
        if d.unconsumed_tail:
            raise TiffError("Deflate strip expands beyond its declared size")
        return out
    if compression in (32773, 5):
        name = "taub_unpackbits" if compression == 32773 else "taub_unlzw"
        if USE_NATIVE and _native() is not None:
            return _native_decode(name, bytes(chunk), expected)
        return (_unpackbits if compression == 32773 else _unlzw)(bytes(chunk), expected)
    raise TiffError(f"unsupported TIFF compression {compression} (supported: none, Deflate, PackBits, LZW)")


# ----------------------------------------------------------------------------- one page
def _chunk_to_array(raw, rows, cols, spp, dtype, bits, predictor):
    """Decoded bytes of one strip / tile -> [rows, cols, spp] samples."""
    if bits == 1:
        row_bytes = (cols * spp + 7) // 8
        a = np.frombuffer(raw, np.uint8, count=rows * row_bytes).reshape(rows, row_bytes)
        return np.unpackbits(a, axis=1)[:, : cols * spp].reshape(rows, cols, spp)
    a = np.frombuffer(raw, dtype, count=rows * cols * spp).reshape(rows, cols, spp)
    if predictor == 2:                       # horizontal differencing, modular in the sample type
        if dtype.kind == "f":
            raise TiffError("predictor 2 on floating-point samples")
        a = np.cumsum(a.astype(dtype.newbyteorder("=")), axis=1, dtype=dtype.newbyteorder("="))
    elif predictor != 1:
        raise TiffError(f"unsupported TIFF predictor {predictor}")
    return a


def _read_page(buf, tags, bo):
    W, H = _one(tags, _WIDTH), _one(tags, _LENGTH)
    spp = _one(tags, _SAMPLES, 1)
    planar = _one(tags, _PLANAR, 1)
    comp = _one(tags, _COMPRESSION, 1)
    pred = _one(tags, _PREDICTOR, 1)
    dtype, bits = _sample_dtype(tags, bo)
    out_dtype = np.dtype(np.bool_) if bits == 1 else dtype.newbyteorder("=")       # bilevel -> bool like tifffile
    if min(W, H, spp) < 1 or W * H * spp * max(bits // 8, 1) > 1100 * len(buf) + (1 << 20):
        # no supported codec expands beyond ~1032 : 1 (Deflate): a damaged header, not an image
        raise TiffError(f"implausible image size {W} x {H} x {spp} for a file of {len(buf)} bytes")
    page = np.empty((H, W, spp), out_dtype)
    tiled = _TILE_OFFSETS in tags
    if tiled:
        tw, tl = _one(tags, _TILE_W), _one(tags, _TILE_L)
        offsets, counts = tags[_TILE_OFFSETS], tags[_TILE_COUNTS]
        across, down = -(-W // tw), -(-H // tl)
    else:
        tw, tl = W, min(_one(tags, _ROWS_PER_STRIP, H), H)
        offsets, counts = tags[_STRIP_OFFSETS], tags.get(_STRIP_COUNTS)
        across, down = 1, -(-H // tl)
        if counts is None:                   # legal for a single uncompressed strip
            if comp != 1 or len(offsets) != 1:
                raise TiffError("StripByteCounts is missing")
            counts = (len(buf) - offsets[0],)
    per_plane = across * down
    groups = spp if planar == 2 else 1
    if len(offsets) != per_plane * groups or len(counts) != len(offsets):
        raise TiffError("strip / tile table does not match the image size")
    s_chunk = 1 if planar == 2 else spp
    for gidx in range(groups):
        for t in range(per_plane):
            r0, c0 = (t // across) * tl, (t % across) * tw
            rows = tl if tiled else min(tl, H - r0)          # tiles are always stored whole
            k = gidx * per_plane + t
            if offsets[k] + counts[k] > len(buf):
                raise TiffError("strip / tile data outside the file")
            nbytes = rows * ((tw * s_chunk + 7) // 8) if bits == 1 else rows * tw * s_chunk * dtype.itemsize
            raw = _decompress(buf[offsets[k]: offsets[k] + counts[k]], comp, nbytes)
            if len(raw) < nbytes:
                raise TiffError("strip / tile holds fewer bytes than its size needs")
            a = _chunk_to_array(raw, rows, tw, s_chunk, dtype, bits, pred)
            a = a[: H - r0, : W - c0]
            if planar == 2:
                page[r0:r0 + a.shape[0], c0:c0 + a.shape[1], gidx] = a[:, :, 0]
            else:
                page[r0:r0 + a.shape[0], c0:c0 + a.shape[1], :] = a
    return page


def _imagej_stack(buf, tags, bo):
    """ImageJ writes ``images=N`` into the description; when it stores a single IFD (files above 4 GB) the
    N planes follow the first one contiguously and uncompressed."""
    desc = tags.get(_DESCRIPTION, "")
    if not (isinstance(desc, str) and desc.startswith("ImageJ=")):
        return None
    m = re.search(r"^images=(\d+)", desc, re.M)
    n = int(m.group(1)) if m else 1
    if n <= 1 or _one(tags, _COMPRESSION, 1) != 1 or _TILE_OFFSETS in tags:
        return None
    dtype, bits = _sample_dtype(tags, bo)
    if bits == 1:
        return None
    W, H, spp = _one(tags, _WIDTH), _one(tags, _LENGTH), _one(tags, _SAMPLES, 1)
    start = tags[_STRIP_OFFSETS][0]
    count = n * H * W * spp
    if start + count * dtype.itemsize > len(buf):
        raise TiffError("ImageJ stack is shorter than its description says")
    a = np.frombuffer(buf, dtype, count=count, offset=start).reshape(n, H, W, spp)
    return a.astype(dtype.newbyteorder("="))


def imread(path):
    """Read a (multi-page) TIFF into a NumPy array shaped like ``tifffile.imread`` shapes it:
    ``[pages, height, width]``; one page -> ``[height, width]``; several samples per pixel -> trailing axis.

    ``path`` is a file name, a ``bytes`` object or an open binary file.  Raises :class:`TiffError` (a
    ``ValueError``) for files this reader does not understand -- nothing is guessed.
    """
    if isinstance(path, (bytes, bytearray, memoryview)):
        buf = memoryview(bytes(path))
    elif hasattr(path, "read"):
        buf = memoryview(path.read())
    else:
        buf = memoryview(np.fromfile(path, dtype=np.uint8))     # one read, no second copy for raw strips
    try:
        ifds, bo = _read_ifds(buf)
        stack = _imagej_stack(buf, ifds[0], bo) if len(ifds) == 1 else None
        if stack is None:
            first = ifds[0]
            key = lambda t: (_one(t, _WIDTH), _one(t, _LENGTH), _one(t, _SAMPLES, 1), t.get(_BITS, (1,)),
                             _one(t, _SAMPLE_FORMAT, 1))
            # pages of another shape or type (thumbnails, masks) are not part of the volume: tifffile's first series
            pages = [t for t in ifds if key(t) == key(first)]
            stack = np.stack([_read_page(buf, t, bo) for t in pages])
    except TiffError:
        raise
    except (struct.error, zlib.error, KeyError, IndexError, TypeError, ValueError, ArithmeticError, MemoryError) as e:
        # truncated directory, damaged stream, tag of the wrong type / count, zero sizes ...
        raise TiffError(f"corrupt TIFF file: {type(e).__name__}: {e}") from e
    if stack.shape[-1] == 1:
        stack = stack[..., 0]
    if stack.shape[0] == 1:
        stack = stack[0]
    return np.ascontiguousarray(stack)
