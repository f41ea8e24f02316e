"""Writes the three tiny TIFF fixtures of tests/test_io_cpu.py::test_committed_fixtures (run once, from the repo
root: ``python tests/golden/make_tiff_fixtures.py``).  LZW and PackBits come from Pillow / libtiff, the
big-endian Deflate file from the test module's own minimal writer."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(HERE), os.path.dirname(os.path.dirname(HERE))]

import test_io_cpu as t  # noqa: E402

vol = t.labels((4, 9, 11), seed=11)
open(os.path.join(HERE, "tiny_lzw.tif"), "wb").write(t.pil_bytes(vol, compression="tiff_lzw"))
open(os.path.join(HERE, "tiny_packbits.tif"), "wb").write(t.pil_bytes(vol, compression="packbits"))
open(os.path.join(HERE, "tiny_deflate_be.tif"), "wb").write(t.build_tiff(list(vol), bo=">", compress=8))
print("written")
