"""Compressed TIFF stacks (LZW, PackBits, deflate, horizontal differencing): utils/tiffio.py over libspimtiff.so
(csrc/tiff_codecs.c, include/spimtiff.h).  The reference reads these through tifffile (imgutils.py:18-23,
data_model.py:178-218); here the decoders are checked against libtiff's encoders (through PIL), against a
straightforward Python encoder of TIFF 6.0 section 13 for the corners libtiff's encoder never produces, and on
damaged streams.  CPU only."""
import ctypes
import os
import re
import struct
import zlib

import numpy as np
import pytest

from spimagine_b200 import frames
from spimagine_b200.utils import tiffio

PIL_Image = pytest.importorskip("PIL.Image")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _smooth(shape, dtype, seed=0):
    """a blob plus a little noise: compressible, long LZW strings, several code widths"""
    rng = np.random.default_rng(seed)
    grid = np.meshgrid(*[np.linspace(-1, 1, n) for n in shape], indexing="ij")
    r2 = sum(g * g for g in grid)
    a = 1000. * np.exp(-3 * r2) + rng.integers(0, 6, shape)
    if np.dtype(dtype).itemsize == 1:
        a = a / 4
    return a.astype(dtype)


def test_header_symbols_are_exported():
    text = open(os.path.join(ROOT, "include", "spimtiff.h")).read()
    names = sorted(set(re.findall(r"SPT_API[^;(]*?\b(spt_[a-z0-9_]+)\s*\(", text)))
    assert names == ["spt_lzw_decode", "spt_packbits_decode", "spt_undo_differencing", "spt_version"]
    lib = tiffio.load_codecs()
    for n in names:
        assert hasattr(lib, n)
    assert lib.spt_version() == 100


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.int32, np.float32])
@pytest.mark.parametrize("compression", ["tiff_lzw", "tiff_adobe_deflate", "tiff_deflate", "packbits"])
@pytest.mark.parametrize("predictor", [1, 2])
def test_we_read_what_libtiff_compresses(tmp_path, dtype, compression, predictor):
    if predictor == 2 and (np.dtype(dtype).kind == "f" or compression == "packbits"):
        pytest.skip("no horizontal differencing for this combination")
    a = _smooth((3, 150, 517), dtype, seed=1)
    fn = str(tmp_path / "c.tif")
    pages = [PIL_Image.fromarray(x) for x in a]
    kw = {"tiffinfo": {317: 2}} if predictor == 2 else {}
    pages[0].save(fn, compression=compression, save_all=True, append_images=pages[1:], **kw)
    t = tiffio.TiffFile(fn)
    assert t.pages[0].compression != 1 and t.pages[0].predictor == predictor and t._flat is None
    got = t.asarray()
    assert got.dtype == a.dtype and np.array_equal(got, a)
    t.decode_threads = 1                      # pages decoded one after the other instead of one per worker
    assert np.array_equal(t.asarray(), a)
    # one image out of the middle, into caller memory
    one = np.empty((1,) + a.shape[1:], t.dtype)
    t.read_into(one, first=1, count=1)
    assert np.array_equal(one[0], a[1])
    d = frames.TiffData(fn)
    assert d.size() == (1,) + a.shape and np.array_equal(d[0], a)


def _lzw_encode(data, eoi=True):
    """TIFF 6.0 section 13 as published: clear code first, codes written MSB first, the width grows one code early,
    the table is cleared when it holds 4094 entries"""
    out, acc, have = bytearray(), 0, 0

    def put(code, width):
        nonlocal acc, have
        acc = (acc << width) | code
        have += width
        while have >= 8:
            out.append((acc >> (have - 8)) & 255)
            have -= 8

    table = {bytes([i]): i for i in range(256)}
    nxt, width = 258, 9
    put(256, width)
    w = b""
    for b in data:
        wb = w + bytes([b])
        if wb in table:
            w = wb
            continue
        put(table[w], width)
        table[wb] = nxt
        nxt += 1
        if nxt == 512 or nxt == 1024 or nxt == 2048:     # the reader is one entry behind: it switches at 511
            width += 1
        if nxt == 4094:
            put(256, width)
            table = {bytes([i]): i for i in range(256)}
            nxt, width = 258, 9
        w = bytes([b])
    if w:
        put(table[w], width)
        # the reader has added one more entry by now
        if nxt in (511, 1023, 2047):
            width += 1
    if eoi:
        put(257, width)
    if have:
        out.append((acc << (8 - have)) & 255)
    return bytes(out)


def _decode(fn_name, raw, cap):
    lib = tiffio.load_codecs()
    dst = np.full(cap + 8, 0xAB, np.uint8)  # 8 guard bytes behind the output
    n = ctypes.c_size_t(0)
    src = np.frombuffer(raw, np.uint8) if len(raw) else np.zeros(1, np.uint8)
    rc = getattr(lib, fn_name)(src.ctypes.data, len(raw), dst.ctypes.data, cap, ctypes.byref(n))
    assert np.all(dst[cap:] == 0xAB), "wrote behind the output"
    return rc, bytes(dst[:n.value])


@pytest.mark.parametrize("case", ["empty", "one", "kwkwk", "runs", "random", "full_table", "no_eoi"])
def test_lzw_known_streams(case):
    rng = np.random.default_rng(7)
    data = {
        "empty": b"",
        "one": b"\x07",
        "kwkwk": b"ababababababababa" * 3,          # codes that name the entry being built
        "runs": b"\x00" * 70000,                     # one string growing to hundreds of bytes
        "random": rng.integers(0, 256, 9000, dtype=np.uint8).tobytes(),          # every width, no repeats
        "full_table": rng.integers(0, 4, 60000, dtype=np.uint8).tobytes(),       # fills the table: clear codes
        "no_eoi": b"hello hello hello hello",
    }[case]
    raw = _lzw_encode(data, eoi=case != "no_eoi")
    rc, got = _decode("spt_lzw_decode", raw, len(data))
    assert rc == 0 and got == data
    if len(data) > 4:
        # a destination smaller than the stream: the first bytes are delivered, nothing behind them is touched
        rc, got = _decode("spt_lzw_decode", raw, len(data) - 3)
        assert rc == -3 and got == data[:-3]


def test_lzw_spec_example():
    """TIFF 6.0 section 13, the worked example: 7 7 7 8 8 7 7 6 6 -> 256 7 258 8 8 258 6 6 257 in 9-bit codes"""
    codes = [256, 7, 258, 8, 8, 258, 6, 6, 257]
    bits = "".join(format(c, "09b") for c in codes)
    bits += "0" * (-len(bits) % 8)
    raw = bytes(int(bits[i:i + 8], 2) for i in range(0, len(bits), 8))
    assert raw == _lzw_encode(bytes([7, 7, 7, 8, 8, 7, 7, 6, 6]))
    rc, got = _decode("spt_lzw_decode", raw, 9)
    assert rc == 0 and got == bytes([7, 7, 7, 8, 8, 7, 7, 6, 6])


def test_damaged_streams_are_refused():
    good = _lzw_encode(b"abcabcabcabc")
    # a code beyond the next free entry
    bits = format(256, "09b") + format(97, "09b") + format(300, "09b")
    bits += "0" * (-len(bits) % 8)
    bad = bytes(int(bits[i:i + 8], 2) for i in range(0, len(bits), 8))
    assert _decode("spt_lzw_decode", bad, 64)[0] == -2
    # a table code right after a clear code
    bits = format(256, "09b") + format(258, "09b")
    bits += "0" * (-len(bits) % 8)
    bad = bytes(int(bits[i:i + 8], 2) for i in range(0, len(bits), 8))
    assert _decode("spt_lzw_decode", bad, 64)[0] == -2
    # truncated: fewer bytes than the image needs
    rc, got = _decode("spt_lzw_decode", good[:4], 12)
    assert rc == 0 and len(got) < 12 and b"abcabcabcabc".startswith(got)
    # PackBits: a literal run that leaves the stream, a repeat without its byte
    assert _decode("spt_packbits_decode", b"\x05ab", 64)[0] == -2
    assert _decode("spt_packbits_decode", b"\xfe", 64)[0] == -2
    lib = tiffio.load_codecs()
    n = ctypes.c_size_t(0)
    assert lib.spt_lzw_decode(None, 0, None, 0, ctypes.byref(n)) == -1
    assert lib.spt_undo_differencing(None, 2, 2, 3, 0) == -1


def test_packbits_known_stream():
    """TIFF 6.0 section 9 / Apple's example"""
    raw = bytes.fromhex("FE AA 02 80 00 2A FD AA 03 80 00 2A 22 F7 AA")
    want = bytes.fromhex("AA AA AA 80 00 2A AA AA AA AA 80 00 2A 22 AA AA AA AA AA AA AA AA AA AA")
    rc, got = _decode("spt_packbits_decode", raw, len(want))
    assert rc == 0 and got == want
    rc, got = _decode("spt_packbits_decode", b"\x80" + raw, len(want))   # -128 is a no-op
    assert rc == 0 and got == want
    rc, got = _decode("spt_packbits_decode", raw, 20)
    assert rc == -3 and got == want[:20]


@pytest.mark.parametrize("dtype", ["u1", "<u2", ">u2", "<i4", ">u4", "<u8", ">i8"])
def test_differencing_wraps_in_the_sample_width(dtype):
    dt = np.dtype(dtype)
    rng = np.random.default_rng(3)
    info = np.iinfo(dt)
    a = rng.integers(info.min, int(info.max) + 1, size=(5, 33), dtype=np.int64 if dt.kind == "i" else np.uint64).astype(dt)
    native = a.astype(dt.newbyteorder("="))
    diff = native.copy()
    diff[:, 1:] = native[:, 1:] - native[:, :-1]          # wraps
    stored = np.ascontiguousarray(diff.astype(dt))
    rc = tiffio.load_codecs().spt_undo_differencing(stored.ctypes.data, 5, 33, dt.itemsize, 0 if dt.isnative else 1)
    assert rc == 0 and np.array_equal(stored, a)


def _tiff_with_strips(fn, a, compression, predictor, bo=">", rows_per_strip=4, pad_last=False):
    """one page, classic TIFF, strips compressed here (deflate through zlib, LZW through the encoder above)"""
    ny, nx = a.shape
    stored = a.astype(a.dtype.newbyteorder(bo))
    if predictor == 2:
        native = a.astype(a.dtype.newbyteorder("="))
        d = native.copy()
        d[:, 1:] = native[:, 1:] - native[:, :-1]
        stored = d.astype(a.dtype.newbyteorder(bo))
    strips = []
    for r in range(0, ny, rows_per_strip):
        part = stored[r:r + rows_per_strip]
        if pad_last and part.shape[0] < rows_per_strip:   # some writers fill the last strip up to RowsPerStrip
            part = np.concatenate([part, np.zeros((rows_per_strip - part.shape[0], nx), part.dtype)]).astype(part.dtype)
        raw = part.tobytes()
        strips.append(zlib.compress(raw) if compression == 8 else _lzw_encode(raw))
    with open(fn, "wb") as f:
        f.write((b"MM" if bo == ">" else b"II") + struct.pack(bo + "HI", 42, 8))
        n = len(strips)
        entries = 10
        data_at = 8 + 2 + 12 * entries + 4
        offs_at = data_at
        cnts_at = offs_at + 4 * n
        pix_at = cnts_at + 4 * n
        offsets, pos = [], pix_at
        for s in strips:
            offsets.append(pos)
            pos += len(s)
        f.write(struct.pack(bo + "H", entries))

        def ent(tag, typ, count, value):
            f.write(struct.pack(bo + "HHI", tag, typ, count))
            f.write(struct.pack(bo + "HH", value, 0) if typ == 3 else struct.pack(bo + "I", value))

        ent(256, 4, 1, nx)
        ent(257, 4, 1, ny)
        ent(258, 3, 1, a.dtype.itemsize * 8)
        ent(259, 3, 1, compression)
        ent(262, 3, 1, 1)
        ent(273, 4, n, offs_at if n > 1 else offsets[0])
        ent(277, 3, 1, 1)
        ent(278, 4, 1, rows_per_strip)
        ent(279, 4, n, cnts_at if n > 1 else len(strips[0]))
        ent(317, 3, 1, predictor)
        f.write(struct.pack(bo + "I", 0))
        f.write(struct.pack(bo + "%dI" % n, *offsets))
        f.write(struct.pack(bo + "%dI" % n, *[len(s) for s in strips]))
        for s in strips:
            f.write(s)


@pytest.mark.parametrize("compression", [5, 8])
@pytest.mark.parametrize("predictor", [1, 2])
@pytest.mark.parametrize("bo", ["<", ">"])
def test_strips_byte_orders_and_padded_last_strip(tmp_path, compression, predictor, bo):
    a = _smooth((23, 41), np.uint16, seed=4)
    fn = str(tmp_path / "s.tif")
    _tiff_with_strips(fn, a, compression, predictor, bo=bo, rows_per_strip=4, pad_last=True)
    t = tiffio.TiffFile(fn)
    assert len(t.pages[0].offsets) == 6 and t.dtype == np.dtype(bo + "u2")
    got = t.asarray()
    assert got.dtype.isnative and np.array_equal(got[0], a)
    if bo == "<" or compression == 8:
        # libtiff agrees (PIL reads big-endian 16-bit LZW strips through a different path: skip that one)
        assert np.array_equal(np.array(PIL_Image.open(fn)), a)


def test_truncated_strip_says_so(tmp_path):
    a = _smooth((16, 16), np.uint8, seed=2)
    fn = str(tmp_path / "t.tif")
    _tiff_with_strips(fn, a[:9], 5, 1, rows_per_strip=16)
    # claim more rows than the strip holds
    raw = bytearray(open(fn, "rb").read())
    at = 8 + 2 + 12 * 1 + 8            # value of tag 257
    raw[at:at + 4] = struct.pack(">I", 16)
    open(fn, "wb").write(raw)
    with pytest.raises(tiffio.TiffError, match="holds"):
        tiffio.TiffFile(fn).asarray()


def _tiff_with_tiles(fn, pages, tile, compression=1, predictor=1, bo="<"):
    """classic TIFF, one directory per page, TileWidth x TileLength tiles (edge tiles zero-padded, TIFF 6.0 s. 15)"""
    tw, tl = tile
    ny, nx = pages[0].shape
    dt = pages[0].dtype
    blobs = []
    for a in pages:
        native = a.astype(dt.newbyteorder("="))
        tiles = []
        for y in range(0, ny, tl):
            for x in range(0, nx, tw):
                t = np.zeros((tl, tw), native.dtype)
                part = native[y:y + tl, x:x + tw]
                t[:part.shape[0], :part.shape[1]] = part
                if predictor == 2:
                    d = t.copy()
                    d[:, 1:] = t[:, 1:] - t[:, :-1]
                    t = d
                raw = t.astype(dt.newbyteorder(bo)).tobytes()
                tiles.append({1: raw, 8: zlib.compress(raw), 5: _lzw_encode(raw)}[compression])
        blobs.append(tiles)
    with open(fn, "wb") as f:
        f.write((b"MM" if bo == ">" else b"II") + struct.pack(bo + "HI", 42, 0))
        ifds = []
        for tiles in blobs:
            n = len(tiles)
            offsets = []
            for t in tiles:
                offsets.append(f.tell())
                f.write(t)
            f.write(b"\0" * (f.tell() % 2))
            offs_at = f.tell()
            f.write(struct.pack(bo + "%dI" % n, *offsets))
            cnts_at = f.tell()
            f.write(struct.pack(bo + "%dI" % n, *[len(t) for t in tiles]))
            ifds.append(f.tell())
            entries = [(256, 4, 1, nx), (257, 4, 1, ny), (258, 3, 1, dt.itemsize * 8), (259, 3, 1, compression),
                       (262, 3, 1, 1), (277, 3, 1, 1), (317, 3, 1, predictor), (322, 4, 1, tw), (323, 4, 1, tl),
                       (324, 4, n, offs_at if n > 1 else offsets[0]), (325, 4, n, cnts_at if n > 1 else len(tiles[0])),
                       (339, 3, 1, {"u": 1, "i": 2, "f": 3}[dt.kind])]
            f.write(struct.pack(bo + "H", len(entries)))
            for tag, typ, count, value in entries:
                f.write(struct.pack(bo + "HHI", tag, typ, count))
                f.write(struct.pack(bo + "HH", value, 0) if typ == 3 else struct.pack(bo + "I", value))
            f.write(struct.pack(bo + "I", 0))
        for i, at in enumerate(ifds):
            f.seek(at + 2 + 12 * 12)
            f.write(struct.pack(bo + "I", ifds[i + 1] if i + 1 < len(ifds) else 0))
        f.seek(4)
        f.write(struct.pack(bo + "I", ifds[0]))


@pytest.mark.parametrize("compression,predictor", [(1, 1), (8, 1), (8, 2), (5, 1), (5, 2)])
@pytest.mark.parametrize("bo", ["<", ">"])
@pytest.mark.parametrize("tile", [(16, 16), (32, 16), (64, 48)])
def test_tiled_pages(tmp_path, compression, predictor, bo, tile):
    """tiles smaller than, ragged against and larger than the 37 x 23 image"""
    a = _smooth((3, 23, 37), np.uint16, seed=6)
    fn = str(tmp_path / "tiles.tif")
    _tiff_with_tiles(fn, list(a), tile, compression, predictor, bo)
    t = tiffio.TiffFile(fn)
    assert t.shape == a.shape and t.pages[0].tile == tile and t._flat is None
    got = t.asarray()
    assert got.dtype.isnative and np.array_equal(got, a)
    one = np.empty((1, 23, 37), t.dtype)
    t.read_into(one, first=2, count=1)
    assert np.array_equal(one[0].astype(np.uint16), a[2])
    if bo == "<":
        assert np.array_equal(_pil_page(fn, 1), a[1])     # libtiff reads the same pixels


def _pil_page(fn, i):
    im = PIL_Image.open(fn)
    im.seek(i)
    return np.array(im)


def test_tiled_float_pages_and_bad_directories(tmp_path):
    a = _smooth((2, 20, 20), np.float32, seed=8)
    fn = str(tmp_path / "f.tif")
    _tiff_with_tiles(fn, list(a), (16, 16), 8, 1, "<")
    assert np.array_equal(tiffio.read3dTiff(fn), a)
    # a directory that lists fewer tiles than the image needs
    raw = bytearray(open(fn, "rb").read())
    ifd = struct.unpack("<I", raw[4:8])[0]
    at = ifd + 2 + 12 * 7 + 8          # value of tag 322 (TileWidth)
    assert struct.unpack("<H", raw[at - 8:at - 6])[0] == 322
    raw[at:at + 4] = struct.pack("<I", 8)
    open(fn, "wb").write(raw)
    with pytest.raises(tiffio.TiffError, match="324"):
        tiffio.TiffFile(fn)


def test_files_read_like_the_references_tifffile_reads_them(tmp_path):
    """tests/golden/tiff_ref.json: the reference's vendored tifffile.imread (what read3dTiff / TiffData call there) on
    the files of tests/golden/tiff_inputs.py -- stacks written by this package's writer (2-d, 3-d, ImageJ 4-d,
    BigTIFF) and strip / tile files with LZW, deflate and the predictor in both byte orders"""
    import hashlib
    import json
    import sys
    golden = os.path.join(ROOT, "tests", "golden")
    sys.path.insert(0, golden)
    try:
        import tiff_inputs
    finally:
        sys.path.remove(golden)
    with open(os.path.join(golden, "tiff_ref.json")) as f:
        ref = json.load(f)["files"]
    files = tiff_inputs.build(str(tmp_path))
    assert sorted(files) == sorted(ref) and len(ref) >= 20
    for name, fn in files.items():
        got = np.ascontiguousarray(tiffio.read3dTiff(fn))
        want = ref[name]
        assert got.dtype.name == want["dtype"] and got.dtype.isnative, name
        assert list(np.squeeze(got).shape) == list(np.squeeze(np.empty(want["shape"])).shape), name
        assert hashlib.sha1(got.tobytes()).hexdigest() == want["sha1"], name


def test_frame_source_prefetches_compressed_time_points(tmp_path):
    """a folder of LZW-compressed stacks (one file per time point, as Fiji's "save as image sequence" with compression
    writes them) and a CZI timelapse go through FrameSource's reader thread like plain files"""
    data = _smooth((4, 5, 30, 41), np.uint16, seed=9)
    folder = tmp_path / "series"
    folder.mkdir()
    for t in range(4):
        pages = [PIL_Image.fromarray(x) for x in data[t]]
        pages[0].save(str(folder / ("t%03d.tif" % t)), compression="tiff_lzw", save_all=True, append_images=pages[1:],
                      tiffinfo={317: 2})
    d = frames.TiffFolderData(str(folder))
    assert tuple(d.size()) == data.shape
    src = frames.FrameSource(d, frames=[0, 1, 2, 3], depth=2, pinned=False)
    try:
        for t in range(4):
            assert np.array_equal(src[t], data[t])
    finally:
        src.close()
    import sys
    golden = os.path.join(ROOT, "tests", "golden")
    sys.path.insert(0, golden)
    try:
        import czi_inputs
    finally:
        sys.path.remove(golden)
    fn = str(tmp_path / "t.czi")
    czi_inputs.write_czi(fn, data, "TZYX", "YX", lzw=_lzw_encode)
    c = frames.CZIData(fn)
    src = frames.FrameSource(c, frames=[3, 1], depth=2, pinned=False)
    try:
        assert np.array_equal(src[3], data[3]) and np.array_equal(src[1], data[1])
    finally:
        src.close()


def test_decoders_survive_arbitrary_bytes():
    """files are untrusted input: random and mutated streams never write outside [dst, dst + cap) and never crash"""
    from hypothesis import given, settings, strategies as st

    good = _lzw_encode(np.random.default_rng(1).integers(0, 7, 5000, dtype=np.uint8).tobytes())

    @settings(max_examples=300, deadline=None)
    @given(st.binary(max_size=600), st.integers(0, 700), st.integers(0, len(good) - 1), st.integers(0, 255))
    def run(raw, cap, where, value):
        for fn in ("spt_lzw_decode", "spt_packbits_decode"):
            rc, got = _decode(fn, raw, cap)
            assert rc in (0, -2, -3) and len(got) <= cap
        mutated = bytearray(good)
        mutated[where] = value
        rc, got = _decode("spt_lzw_decode", bytes(mutated), 5000)
        assert rc in (0, -2, -3) and len(got) <= 5000

    run()


def test_block_sizes_are_checked_against_the_file(tmp_path):
    a = _smooth((16, 16), np.uint8, seed=2)
    fn = str(tmp_path / "t.tif")
    _tiff_with_strips(fn, a, 8, 1, rows_per_strip=16)
    raw = bytearray(open(fn, "rb").read())
    at = 8 + 2 + 12 * 8 + 8            # value of tag 279 (the single strip's byte count)
    assert struct.unpack(">H", raw[at - 8:at - 6])[0] == 279
    raw[at:at + 4] = struct.pack(">I", 0x7FFFFFFF)
    open(fn, "wb").write(raw)
    with pytest.raises(tiffio.TiffError, match="leaves the file"):
        tiffio.TiffFile(fn).asarray()
