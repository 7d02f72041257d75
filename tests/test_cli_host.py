"""CPU tests of the CLI's host-side support code (the `.svt` parser of utils.hpp:55-94 with its quirks, SURVEY Q25,
and the in-tree TIFF reader/writer that stands in for libtiff) through pgure-svt_b200/host_tools."""
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

TOOL = os.path.join(ROOT, "pgure-svt_b200", "host_tools")
pytestmark = pytest.mark.skipif(not os.path.exists(TOOL), reason="host_tools not built (run __graft_entry__.build())")


def _parse(text, tmp_path):
    f = tmp_path / "p.svt"
    f.write_bytes(text.encode())
    out = subprocess.run([TOOL, "parse", str(f)], capture_output=True, text=True, check=True).stdout
    return dict(line.split("=[", 1)[0:1] + [line.split("=[", 1)[1][:-1]] for line in out.splitlines() if line)


def test_parser_quirks(tmp_path):
    d = _parse("a : 1 2 3 # c\nc : 0.1#x\nb:5\nf : true", tmp_path)
    assert d == {"a": "123", "c": "", "b:5": "", "f": "true "}  # SURVEY Q25, verified against the reference's function
    d = _parse("# comment\n\nkey : value   # trailing\nother = 7 \n", tmp_path)
    assert d == {"key": "value", "other": "7"}


def test_parser_reference_example(tmp_path):
    d = _parse(open(os.path.join(GOLDEN, "param_example.svt")).read(), tmp_path)
    assert len(d) == 15
    assert d["filename"] == "./example.tif" and d["patch_size"] == "16" and d["lambda"] == "0.15"
    assert d["optimize_pgure"] == "false" and d["random_seed"] == "1" and "hot_pixel" not in d


def test_tiff_reader_matches_opencv_and_roundtrips(tmp_path):
    cv2 = pytest.importorskip("cv2")
    src = os.path.join(GOLDEN, "example.tif")
    out = subprocess.run([TOOL, "tiffinfo", src], capture_output=True, text=True, check=True).stdout.split("\n")
    n, w, h, bits = [int(x) for x in out[0].split()]
    sums = [int(x) for x in out[1:] if x]
    ok, pages = cv2.imreadmulti(src, flags=cv2.IMREAD_UNCHANGED)
    assert ok and (n, w, h, bits) == (len(pages), 128, 128, 16)
    assert sums == [int(p.astype(np.uint64).sum()) for p in pages]
    dst = str(tmp_path / "copy.tif")
    subprocess.run([TOOL, "tiffcopy", src, dst], check=True)
    ok, pages2 = cv2.imreadmulti(dst, flags=cv2.IMREAD_UNCHANGED)
    assert ok and len(pages2) == len(pages)
    assert all(np.array_equal(a, b) for a, b in zip(pages, pages2))


def test_cli_rejects_pages_of_different_size_and_empty_tags(tmp_path):
    """the TIFF reader trusted page 0's geometry for every page and indexed v[0] of zero-count tags."""
    cv2 = pytest.importorskip("cv2")
    exe = os.path.join(ROOT, "pgure-svt_b200", "PGURE-SVT")
    if not os.path.exists(exe):
        pytest.skip("CLI not built")
    rng = np.random.RandomState(0)
    pages = [rng.randint(0, 1000, (32, 32)).astype(np.uint16) for _ in range(16)] + [rng.randint(0, 1000, (16, 16)).astype(np.uint16)]
    assert cv2.imwritemulti(str(tmp_path / "mixed.tif"), pages, [cv2.IMWRITE_TIFF_COMPRESSION, 1])
    (tmp_path / "p.svt").write_text("filename : ./mixed.tif\nstart_frame : 1\nend_frame : 17\noptimize_pgure : false\nlambda : 0.1\n")
    r = subprocess.run([exe, "p.svt"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode != 0 and "page 17 is 16x16" in (r.stdout + r.stderr)
    # a tag with a zero value count (ImageWidth, count 0) is refused by the reader instead of read out of bounds
    raw = bytearray((tmp_path / "mixed.tif").read_bytes())
    import struct

    ifd = struct.unpack_from("<I", raw, 4)[0]
    nent = struct.unpack_from("<H", raw, ifd)[0]
    for e in range(nent):
        off = ifd + 2 + 12 * e
        if struct.unpack_from("<H", raw, off)[0] == 256:
            struct.pack_into("<I", raw, off + 4, 0)
    (tmp_path / "bad.tif").write_bytes(bytes(raw))
    r = subprocess.run([TOOL, "tiffinfo", str(tmp_path / "bad.tif")], capture_output=True, text=True)
    assert r.returncode != 0
