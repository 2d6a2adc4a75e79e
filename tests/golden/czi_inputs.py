"""Synthetic CZI (ZISRAW) files for tests/test_czi.py and tests/golden/make_czi_golden.py: uncompressed greyscale
sub-blocks with a sub-block directory, laid out by the published ZISRAW segment structure (32-byte segment headers
`id[16], allocated, used`; file header with the directory position; `DV` directory entries of 32 bytes + 20 per
dimension; sub-block segments `metadata_size, attachment_size, data_size, entry, fill to 256, metadata, data`)."""
import struct

import numpy as np

PIXEL = {"uint8": 0, "uint16": 1, "float32": 2, "int32": 12}


def _segment(sid, payload, align=32):
    alloc = -(-len(payload) // align) * align
    return struct.pack("<16sqq", sid, alloc, len(payload)) + payload + b"\0" * (alloc - len(payload))


def _entry(pixel, file_position, dims, compression=0, pyramid=0):
    """dims: [(name, start, size)] in file order (X first)"""
    e = struct.pack("<2siqiiBB4si", b"DV", pixel, file_position, 0, compression, pyramid, 0, b"\0" * 4, len(dims))
    for name, start, size in dims:
        e += struct.pack("<4siifi", name, start, size, float(start), size)
    return e


def write_czi(fName, data, axes, block_axes, starts=None, metadata=b"<METADATA/>", with_mosaic=False, lzw=None):
    """data: array whose axes are named by `axes` (e.g. "TZYX"); one sub-block per index of the axes NOT in
    block_axes (e.g. block_axes "YX": one plane per (T, Z)).  starts: {axis: offset} added to every start index
    (acquisitions that do not begin at 0)."""
    starts = starts or {}
    outer = [a for a in axes if a not in block_axes]
    outer_shape = [data.shape[axes.index(a)] for a in outer]
    blocks = []
    for idx in np.ndindex(*outer_shape) if outer else [()]:
        sel = [slice(None)] * data.ndim
        for a, i in zip(outer, idx):
            sel[axes.index(a)] = slice(i, i + 1)
        part = np.ascontiguousarray(data[tuple(sel)])
        dims = []
        for k in reversed(range(data.ndim)):             # file order: fastest axis first
            a = axes[k]
            start = (idx[outer.index(a)] if a in outer else 0) + starts.get(a, 0)
            dims.append((a.encode(), start, part.shape[k]))
        if with_mosaic:
            dims.append((b"M", len(blocks), 1))
        blocks.append((dims, part))
    pixel = PIXEL[data.dtype.name]
    body = bytearray()
    header_size = 32 + 512
    positions = []
    for dims, part in blocks:
        pos = header_size + len(body)
        positions.append(pos)
        entry = _entry(pixel, pos, dims, compression=2 if lzw else 0)
        raw = part.astype(part.dtype.newbyteorder("<")).tobytes()
        if lzw:
            raw = lzw(raw)
        payload = struct.pack("<iiq", len(metadata), 0, len(raw)) + entry
        payload += b"\0" * max(256 - len(payload), 0) + metadata + raw
        body += _segment(b"ZISRAWSUBBLOCK", payload)
    directory_position = header_size + len(body)
    payload = struct.pack("<i", len(blocks)) + b"\0" * 124
    for (dims, part), pos in zip(blocks, positions):
        payload += _entry(pixel, pos, dims, compression=2 if lzw else 0)
    body += _segment(b"ZISRAWDIRECTORY", payload)
    guid = bytes(range(16))
    head = struct.pack("<iiii16s16siqqiq", 1, 0, 0, 0, guid, guid, 0, directory_position, 0, 0, 0)
    with open(fName, "wb") as f:
        f.write(struct.pack("<16sqq", b"ZISRAWFILE", 512, 512) + head + b"\0" * (512 - len(head)))
        f.write(body)


def cases():
    """name -> (array, axes, block_axes, starts, with_mosaic)"""
    rng = np.random.default_rng(17)
    return {
        "zyx_planes_u16": (rng.integers(0, 60000, (5, 6, 7)).astype(np.uint16), "ZYX", "YX", None, False),
        "tzyx_planes_u8": (rng.integers(0, 256, (3, 4, 5, 6)).astype(np.uint8), "TZYX", "YX", {"T": 2, "Z": 10}, False),
        "tzyx_stacks_f32": (rng.normal(size=(2, 3, 4, 5)).astype(np.float32), "TZYX", "ZYX", None, False),
        "czyx_mosaic_u16": (rng.integers(0, 60000, (1, 4, 5, 6)).astype(np.uint16), "CZYX", "YX", {"C": 1}, True),
        # written with LZW sub-blocks (write_czi(..., lzw=encoder)): smooth, so that strings grow
        "zyx_planes_u16_lzw": ((1000 * np.exp(-np.linspace(-2, 2, 6 * 40 * 50).reshape(6, 40, 50) ** 2)).astype(np.uint16),
                               "ZYX", "YX", None, False),
    }


def lzw_encoder():
    """the TIFF 6.0 encoder of tests/test_tiff_codecs.py"""
    import os
    import sys
    tests = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, tests)
    try:
        import test_tiff_codecs
    finally:
        sys.path.remove(tests)
    return test_tiff_codecs._lzw_encode


def write_case(fName, key):
    data, axes, block_axes, starts, mosaic = cases()[key]
    write_czi(fName, data, axes, block_axes, starts, with_mosaic=mosaic, lzw=lzw_encoder() if key.endswith("_lzw") else None)
    return data
