"""Carl Zeiss CZI (ZISRAW) stacks without third-party decoders: the plain case light-sheet and confocal acquisitions
produce -- uncompressed (or LZW) greyscale sub-blocks listed in a sub-block directory.

The reference reads CZI through a vendored `czifile` (spimagine/lib/czifile.py; imgutils.py:44-47 readCziFile,
data_model.py:557-584 CZIData).  From the ZISRAW layout this module implements:

  * the file header segment (`ZISRAWFILE`: position of the sub-block directory);
  * the sub-block directory (`ZISRAWDIRECTORY`: `DV` entries -- pixel type, file position, compression, one
    (dimension, start, size, stored size) record per dimension, fastest dimension first);
  * sub-block segments (`ZISRAWSUBBLOCK`: sizes, a copy of the entry padded to 256 bytes, XML metadata, pixels).

The array is assembled as czifile does: its shape spans `min(start) .. max(start + size)` of every dimension except
the mosaic index M, each sub-block is pasted at `start - min(start)`.  `read_into` pastes straight into caller memory
(a page-locked buffer of frames.FrameSource); `time_point(t)` reads only the sub-blocks of one T index.

LZW sub-blocks (compression 2) are decoded by libspimtiff.so, as czifile decodes them with tifffile's LZW.
Refused with CziError naming the field: JPEG / JPEG-XR sub-blocks, colour and complex pixel
types, pyramid levels (stored size != size), files without a directory.
"""
import os
import struct

import numpy as np

__all__ = ["CziError", "CziFile", "readCziFile"]

_PIXEL = {0: "<u1", 1: "<u2", 2: "<f4", 12: "<i4", 13: "<i8"}
_PIXEL_NAMES = {3: "Bgr24", 4: "Bgr48", 8: "Bgr96Float", 9: "Bgra32", 10: "Gray64ComplexFloat", 11: "Bgr192ComplexFloat"}
_SEG = struct.Struct("<16sqq")
_ENTRY = struct.Struct("<2siqiiBB4si")
_DIM = struct.Struct("<4siifi")


class CziError(ValueError):
    pass


class _Block(object):
    __slots__ = ("dtype", "position", "dims", "start", "shape", "mosaic", "compression")


class CziFile(object):
    """`axes` (e.g. "TZYX"), `shape`, `start` (lowest index per axis), `dtype` of the assembled array."""

    def __init__(self, fName):
        self.fName = fName
        with open(fName, "rb") as f:
            sid, _, _ = self._segment(f, 0)
            if sid != b"ZISRAWFILE":
                raise CziError("%s: not a CZI file" % fName)
            head = f.read(80)
            if len(head) < 80:
                raise CziError("%s: truncated file header" % fName)
            directory = struct.unpack("<iiii16s16siqqiq", head)[7]
            if not directory:
                raise CziError("%s: no sub-block directory (directory_position = 0)" % fName)
            sid, _, _ = self._segment(f, directory)
            if sid != b"ZISRAWDIRECTORY":
                raise CziError("%s: no sub-block directory at %d" % (fName, directory))
            count = struct.unpack("<i", f.read(4))[0]
            f.seek(124, 1)
            blocks = [self._entry(f) for _ in range(count)]
        if not blocks:
            raise CziError("%s: no sub-blocks" % fName)
        with_index = [b for b in blocks if b.mosaic is not None]
        self.blocks = sorted(with_index, key=lambda b: b.mosaic) if with_index else blocks   # czifile.py:301-310
        b0 = self.blocks[0]
        self.axes = "".join(d for d, _, _ in b0.dims)
        for b in self.blocks:
            if "".join(d for d, _, _ in b.dims) != self.axes:
                raise CziError("%s: sub-blocks differ in their dimensions" % fName)
        self.dtype = np.dtype(np.result_type(*[b.dtype for b in self.blocks]))
        starts = np.array([b.start for b in self.blocks])
        ends = starts + np.array([b.shape for b in self.blocks])
        self.start = tuple(int(s) for s in starts.min(axis=0))
        self.shape = tuple(int(e - s) for e, s in zip(ends.max(axis=0), self.start))

    def _segment(self, f, at):
        f.seek(at)
        raw = f.read(_SEG.size)
        if len(raw) < _SEG.size:
            raise CziError("%s: no segment at %d" % (self.fName, at))
        sid, allocated, used = _SEG.unpack(raw)
        return sid.split(b"\0")[0], allocated, used

    def _entry(self, f):
        raw = f.read(_ENTRY.size)
        if len(raw) < _ENTRY.size:
            raise CziError("%s: truncated directory" % self.fName)
        schema, pixel, position, _, compression, pyramid, _, _, ndim = _ENTRY.unpack(raw)
        if schema != b"DV":
            raise CziError("%s: directory entry of schema %r" % (self.fName, schema))
        if pixel not in _PIXEL:
            raise CziError("%s: pixel type %s is not supported" % (self.fName, _PIXEL_NAMES.get(pixel, pixel)))
        if compression not in (0, 2):
            raise CziError("%s: this sub-block compression is not supported (compression = %d: %s)"
                           % (self.fName, compression, {1: "JPEG", 4: "JPEG XR"}.get(compression, "camera specific")))
        b = _Block()
        b.dtype, b.position, b.mosaic, b.compression = np.dtype(_PIXEL[pixel]), position, None, compression
        dims = []
        for _ in range(ndim):
            name, start, size, _, stored = _DIM.unpack(f.read(_DIM.size))
            name = name.split(b"\0")[0].decode("ascii")
            if name == "M":
                b.mosaic = start
                continue
            if stored and stored != size:
                raise CziError("%s: resampled sub-blocks (pyramid levels) are not supported (dimension %s: %d stored "
                               "for %d)" % (self.fName, name, stored, size))
            dims.append((name, start, size))
        b.dims = dims[::-1]                       # the file lists the fastest dimension first
        b.start = tuple(s for _, s, _ in b.dims)
        b.shape = tuple(n for _, _, n in b.dims)
        return b

    def _pixels(self, f, b):
        sid, _, _ = self._segment(f, b.position)
        if sid != b"ZISRAWSUBBLOCK":
            raise CziError("%s: no sub-block at %d" % (self.fName, b.position))
        metadata_size, _, data_size = struct.unpack("<iiq", f.read(16))
        ndim = _ENTRY.unpack(f.read(_ENTRY.size))[8]
        # the entry copy is padded so that sizes + entry take 256 bytes (16 + 240), then the XML, then the pixels
        f.seek(ndim * _DIM.size + max(240 - (_ENTRY.size + ndim * _DIM.size), 0) + metadata_size, 1)
        want = int(np.prod(b.shape)) * b.dtype.itemsize
        fsize = os.fstat(f.fileno()).st_size
        if data_size < 0 or data_size > fsize or want > (1 << 40):  # file-controlled sizes must not size reads / arrays
            raise CziError("%s: the sub-block at %d declares %d bytes for %d pixels bytes" % (
                self.fName, b.position, data_size, want))
        if b.compression == 2:
            # LZW as in TIFF (czifile decodes it with tifffile's decoder): libspimtiff.so, include/spimtiff.h
            import ctypes
            from .tiffio import load_codecs
            raw = f.read(data_size)
            a = np.empty(want // b.dtype.itemsize, b.dtype)
            n = ctypes.c_size_t(0)
            rc = load_codecs().spt_lzw_decode(raw, len(raw), a.ctypes.data, want, ctypes.byref(n))
            if rc not in (0, -3) or n.value < want:
                raise CziError("%s: damaged LZW sub-block at %d" % (self.fName, b.position))
            return a.reshape(b.shape)
        if data_size < want:
            raise CziError("%s: the sub-block at %d holds %d bytes, %d are needed" % (self.fName, b.position, data_size, want))
        a = np.fromfile(f, b.dtype, want // b.dtype.itemsize)
        if a.size * b.dtype.itemsize < want:
            raise CziError("%s: the sub-block at %d leaves the file" % (self.fName, b.position))
        return a.reshape(b.shape)

    def read_into(self, out, blocks=None, origin=None):
        """Paste sub-blocks (default: all) into `out`, whose element [0, 0, ...] is index `origin` (default: start)."""
        origin = self.start if origin is None else origin
        with open(self.fName, "rb") as f:
            for b in (self.blocks if blocks is None else blocks):
                index = tuple(slice(s - o, s - o + n) for s, o, n in zip(b.start, origin, b.shape))
                out[index] = self._pixels(f, b)
        return out

    def asarray(self):
        """the whole file, shaped `shape` (zeros where no sub-block lies), native byte order"""
        return self.read_into(np.zeros(self.shape, self.dtype.newbyteorder("=")))

    def time_point(self, t, out=None):
        """the array of T index `start[T] + t` alone (T axis dropped): only its sub-blocks are read"""
        if "T" not in self.axes:
            raise CziError("%s: no T dimension (axes %s)" % (self.fName, self.axes))
        k = self.axes.index("T")
        if not 0 <= t < self.shape[k]:
            raise IndexError("0 <= t < %d, but t = %d" % (self.shape[k], t))
        want = self.start[k] + t
        mine = [b for b in self.blocks if b.start[k] <= want < b.start[k] + b.shape[k]]
        if any(b.shape[k] != 1 for b in mine):
            arr = self.asarray().take(t, axis=k)          # sub-blocks that span several time points: read them all
            if out is None:
                return arr
            np.copyto(out.reshape(arr.shape), arr)        # the caller's buffer is what FrameSource uploads
            return out.reshape(arr.shape)
        shape = self.shape[:k] + (1,) + self.shape[k + 1:]
        full = np.zeros(shape, self.dtype.newbyteorder("=")) if out is None else out.reshape(shape)
        if out is not None:
            full[...] = 0
        origin = self.start[:k] + (want,) + self.start[k + 1:]
        self.read_into(full, mine, origin)
        return full.reshape(self.shape[:k] + self.shape[k + 1:])


def readCziFile(fName):
    """imgutils.py:44-47"""
    return np.squeeze(CziFile(fName).asarray())
