"""Baseline TIFF stacks without third-party decoders: the file format either side of the render path.

The reference reads and writes 2/3/4-d TIFF through `tifffile` (spimagine/utils/imgutils.py:18-27 read3dTiff /
write3dTiff, spimagine/models/data_model.py:178-218 TiffData), which is not part of this image.  Microscopy stacks
are almost always the plainest TIFF there is -- one uncompressed greyscale image per page, stored in strips -- and
that subset is what this module implements from the TIFF 6.0 / BigTIFF layout:

  * classic TIFF (magic 42, 32-bit offsets) and BigTIFF (magic 43, 64-bit offsets), little and big endian;
  * pages of one sample per pixel, 8 / 16 / 32 / 64-bit unsigned, signed or IEEE float;
  * strips (any RowsPerStrip); contiguous uncompressed pages are detected and read with ONE readinto per page;
  * tiles (TIFF 6.0 section 15: OME-TIFF writers), stored or compressed, edge tiles clipped;
  * compressed strips: LZW (5) and PackBits (32773) through libspimtiff.so (csrc/tiff_codecs.c, include/spimtiff.h),
    deflate (8 / 32946) through zlib, each with the horizontal-differencing predictor (Predictor = 2) for integers;
    every decoder runs with the GIL released, so FrameSource's reader thread overlaps it with rendering;
  * ImageJ hyperstacks: `ImageDescription = "ImageJ=...\nimages=N\nslices=Z\nframes=T"` gives the (T, Z, Y, X)
    shape, and ImageJ's "> 4 GB" layout (a single IFD followed by all N images back to back) is understood.

Anything else (JPEG / ZSTD / LZMA, the floating-point predictor, RGB, planar) raises TiffError naming the tag, so a caller can fall back
to a full decoder and wrap the result in frames.NumpyData.  `TiffFile.read_into` fills caller memory -- page-locked
buffers of frames.FrameSource -- straight from the file, which the tifffile path of the reference cannot do.
"""
import ctypes
import os
import struct
import zlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np

__all__ = ["TiffError", "TiffFile", "read3dTiff", "write3dTiff", "imread", "imsave"]


class TiffError(ValueError):
    pass


# tag ids (TIFF 6.0 section 8)
_WIDTH, _LENGTH, _BITS, _COMPRESSION, _PHOTOMETRIC, _DESCRIPTION = 256, 257, 258, 259, 262, 270
_STRIP_OFFSETS, _SAMPLES, _ROWS_PER_STRIP, _STRIP_COUNTS = 273, 277, 278, 279
_PLANAR, _PREDICTOR, _TILE_WIDTH, _TILE_LENGTH, _TILE_OFFSETS, _TILE_COUNTS, _SAMPLE_FORMAT = 284, 317, 322, 323, 324, 325, 339
_NONE, _LZW, _DEFLATE, _DEFLATE_OLD, _PACKBITS = 1, 5, 8, 32946, 32773
# field type -> (struct code, bytes)
_TYPES = {1: ("B", 1), 2: ("c", 1), 3: ("H", 2), 4: ("I", 4), 5: ("II", 8), 6: ("b", 1), 7: ("B", 1), 8: ("h", 2),
          9: ("i", 4), 10: ("ii", 8), 11: ("f", 4), 12: ("d", 8), 13: ("I", 4), 16: ("Q", 8), 17: ("q", 8), 18: ("Q", 8)}
_KIND = {1: "u", 2: "i", 3: "f"}


_CODEC_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "libspimtiff.so")
_codecs = None


def load_codecs():
    """ctypes binding of include/spimtiff.h (built by spimagine_b200.build.build_tiff)"""
    global _codecs
    if _codecs is None:
        if not os.path.exists(_CODEC_PATH):
            raise TiffError("%s is missing: build it with python -m spimagine_b200.build" % _CODEC_PATH)
        lib = ctypes.CDLL(_CODEC_PATH)
        decode = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]
        lib.spt_version.restype = ctypes.c_int
        lib.spt_version.argtypes = []
        for name in ("spt_lzw_decode", "spt_packbits_decode"):
            getattr(lib, name).restype = ctypes.c_int
            getattr(lib, name).argtypes = decode
        lib.spt_undo_differencing.restype = ctypes.c_int
        lib.spt_undo_differencing.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int,
                                              ctypes.c_int]
        _codecs = lib
    return _codecs


class _Page(object):
    __slots__ = ("width", "length", "dtype", "offsets", "counts", "rows_per_strip", "description", "compression",
                 "predictor", "tile")

    def contiguous(self):
        """-> file offset of the image if it is stored as it is and its strips follow one another without gaps,
        else None"""
        if self.compression != _NONE or self.tile:
            return None
        pos = self.offsets[0]
        for o, c in zip(self.offsets, self.counts):
            if o != pos:
                return None
            pos += c
        return self.offsets[0]

    @property
    def nbytes(self):
        return self.width * self.length * self.dtype.itemsize


class TiffFile(object):
    """Page directory of a TIFF stack.  `shape` is (pages, Y, X), or (T, Z, Y, X) for an ImageJ hyperstack with
    frames > 1; `dtype` carries the file's byte order."""

    decode_threads = 0  # workers for compressed pages; 0: one per core, at most 16

    def __init__(self, fName):
        self.fName = fName
        self.pages = []
        with open(fName, "rb") as f:
            self._size = os.fstat(f.fileno()).st_size
            self._parse(f)
        if not self.pages:
            raise TiffError("%s: no images" % fName)
        p0 = self.pages[0]
        self.dtype = p0.dtype
        self._ij = self._imagej(p0.description)
        n = len(self.pages)
        if n == 1 and self._ij.get("images", 1) > 1:
            # ImageJ's layout for stacks beyond 4 GB: one IFD, every image back to back behind the first
            n = self._ij["images"]
            start = p0.contiguous()
            if start is None or start + n * p0.nbytes > self._size:
                raise TiffError("%s: ImageJ stack of %d images does not fit the file" % (fName, n))
            self._flat = start
        else:
            for p in self.pages[1:]:
                if (p.width, p.length, p.dtype) != (p0.width, p0.length, p0.dtype):
                    raise TiffError("%s: pages differ in size or type" % fName)
            starts = [p.contiguous() for p in self.pages]
            self._flat = starts[0] if all(s is not None and s == starts[0] + i * p0.nbytes
                                          for i, s in enumerate(starts)) else None
        self.n_images = n
        frames, slices = self._ij.get("frames", 1), self._ij.get("slices", 1)
        if frames > 1 and frames * slices == n and self._ij.get("channels", 1) == 1:
            self.shape = (frames, slices, p0.length, p0.width)
        else:
            self.shape = (n, p0.length, p0.width)

    # ---- directory ----
    def _parse(self, f):
        head = f.read(16)
        if head[:2] == b"II":
            self._bo = "<"
        elif head[:2] == b"MM":
            self._bo = ">"
        else:
            raise TiffError("%s: not a TIFF file" % self.fName)
        magic = struct.unpack(self._bo + "H", head[2:4])[0]
        if magic == 42:
            self._big = False
            ifd = struct.unpack(self._bo + "I", head[4:8])[0]
        elif magic == 43:
            self._big = True
            if struct.unpack(self._bo + "HH", head[4:8]) != (8, 0):
                raise TiffError("%s: malformed BigTIFF header" % self.fName)
            ifd = struct.unpack(self._bo + "Q", head[8:16])[0]
        else:
            raise TiffError("%s: not a TIFF file (magic %d)" % (self.fName, magic))
        seen = set()
        while ifd:
            if ifd in seen or ifd >= self._size:
                raise TiffError("%s: broken image directory chain" % self.fName)
            seen.add(ifd)
            tags, ifd = self._read_ifd(f, ifd)
            self.pages.append(self._page(tags))

    def _read_ifd(self, f, at):
        bo = self._bo
        cnt_fmt, ent, off_fmt, inline = ("Q", 20, "Q", 8) if self._big else ("H", 12, "I", 4)
        f.seek(at)
        head = f.read(struct.calcsize(cnt_fmt))
        if len(head) < struct.calcsize(cnt_fmt):
            raise TiffError("%s: truncated image directory" % self.fName)
        n = struct.unpack(bo + cnt_fmt, head)[0]
        fsize = os.fstat(f.fileno()).st_size
        if n * ent > fsize:  # a file-controlled count must not size a read
            raise TiffError("%s: image directory with %d entries does not fit the file" % (self.fName, n))
        raw = f.read(n * ent + struct.calcsize(off_fmt))
        if len(raw) < n * ent + struct.calcsize(off_fmt):
            raise TiffError("%s: truncated image directory" % self.fName)
        tags = {}
        for i in range(n):
            e = raw[i * ent:(i + 1) * ent]
            tag, typ = struct.unpack(bo + "HH", e[:4])
            count = struct.unpack(bo + off_fmt, e[4:4 + inline])[0]
            if typ not in _TYPES:
                continue
            code, size = _TYPES[typ]
            nbytes = size * count
            value = e[4 + inline:]
            if nbytes > inline:
                if tag not in (_BITS, _DESCRIPTION, _STRIP_OFFSETS, _STRIP_COUNTS, _SAMPLE_FORMAT, _TILE_OFFSETS,
                               _TILE_COUNTS):  # all others fit inline
                    continue  # a big value of a tag this reader does not use (colour maps, ImageJ metadata, ...)
                where = struct.unpack(bo + off_fmt, value[:inline])[0]
                if nbytes > fsize or where > fsize - nbytes:
                    raise TiffError("%s: tag %d points beyond the file" % (self.fName, tag))
                here = f.tell()
                f.seek(where)
                value = f.read(nbytes)
                f.seek(here)
                if len(value) < nbytes:
                    raise TiffError("%s: tag %d points beyond the file" % (self.fName, tag))
            if typ == 2:
                tags[tag] = value[:nbytes].split(b"\0")[0].decode("latin-1")
            elif len(code) == 1:
                tags[tag] = np.frombuffer(value[:nbytes], dtype=bo + {"B": "u1", "b": "i1", "H": "u2", "h": "i2",
                                                                       "I": "u4", "i": "i4", "Q": "u8", "q": "i8",
                                                                       "f": "f4", "d": "f8"}[code]).tolist()
            # rationals are not needed by this reader
        nxt = struct.unpack(bo + off_fmt, raw[n * ent:])[0]
        return tags, nxt

    def _page(self, t):
        def one(tag, default=None):
            v = t.get(tag)
            if v is None:
                if default is None:
                    raise TiffError("%s: required tag %d is missing" % (self.fName, tag))
                return default
            return v[0] if isinstance(v, list) else v

        compression, predictor = one(_COMPRESSION, 1), one(_PREDICTOR, 1)
        if compression not in (_NONE, _LZW, _DEFLATE, _DEFLATE_OLD, _PACKBITS):
            raise TiffError("%s: this compression is not supported (tag 259 = %d)" % (self.fName, compression))
        if predictor not in (1, 2):
            raise TiffError("%s: this predictor is not supported (tag 317 = %d)" % (self.fName, predictor))
        if predictor == 2 and compression not in (_LZW, _DEFLATE, _DEFLATE_OLD):
            # TIFF 6.0 defines the predictor for LZW (deflate inherited it); libtiff writes the tag with other
            # schemes but does not difference the samples, other writers do: ambiguous, so refuse
            raise TiffError("%s: predictor with compression %d is ambiguous (tag 317 = 2)" % (self.fName, compression))
        if one(_SAMPLES, 1) != 1:
            raise TiffError("%s: %d samples per pixel are not supported (tag 277)" % (self.fName, one(_SAMPLES)))
        bits = one(_BITS, 1)
        kind = _KIND.get(one(_SAMPLE_FORMAT, 1), "u")
        if bits not in (8, 16, 32, 64) or (kind == "f" and bits < 32):
            raise TiffError("%s: %d-bit samples are not supported (tag 258)" % (self.fName, bits))
        if predictor == 2 and kind == "f":
            raise TiffError("%s: horizontal differencing of floating-point samples (tag 317 = 2)" % self.fName)
        p = _Page()
        p.compression, p.predictor = compression, predictor
        p.width, p.length = int(one(_WIDTH)), int(one(_LENGTH))
        p.dtype = np.dtype(self._bo + kind + str(bits // 8))
        if p.dtype.itemsize == 1:
            p.dtype = np.dtype(kind + "1")
        p.tile = None
        if _TILE_WIDTH in t:
            # TIFF 6.0 section 15: tiles of TileWidth x TileLength, left to right, top to bottom, the ones at the right
            # and bottom edges padded to full size
            tw, tl = int(one(_TILE_WIDTH)), int(one(_TILE_LENGTH))
            if tw <= 0 or tl <= 0:
                raise TiffError("%s: tiles of %d x %d (tags 322, 323)" % (self.fName, tw, tl))
            p.tile = (tw, tl)
            p.offsets = [int(o) for o in t.get(_TILE_OFFSETS, [])]
            p.counts = [int(c) for c in t.get(_TILE_COUNTS, [])]
            n = (-(-p.width // tw)) * (-(-p.length // tl))
            if len(p.offsets) != n or len(p.counts) != n:
                raise TiffError("%s: %d tile offsets and %d byte counts for %d tiles (tags 324, 325)"
                                % (self.fName, len(p.offsets), len(p.counts), n))
            p.rows_per_strip = tl
            p.description = t.get(_DESCRIPTION, "")
            return p
        p.offsets = [int(o) for o in t.get(_STRIP_OFFSETS, [])]
        if not p.offsets:
            raise TiffError("%s: no strips (tag 273)" % self.fName)
        p.rows_per_strip = min(int(one(_ROWS_PER_STRIP, p.length)), p.length)
        row = p.width * p.dtype.itemsize
        counts = t.get(_STRIP_COUNTS)
        if counts is None or len(counts) != len(p.offsets):  # old writers leave the counts out
            if compression != _NONE:
                raise TiffError("%s: compressed strips without byte counts (tag 279)" % self.fName)
            n = len(p.offsets)
            counts = [row * p.rows_per_strip] * (n - 1) + [p.nbytes - row * p.rows_per_strip * (n - 1)]
        p.counts = [int(c) for c in counts]
        if compression != _NONE:
            if len(p.offsets) != -(-p.length // p.rows_per_strip):
                raise TiffError("%s: %d strips for %d rows of %d per strip" % (self.fName, len(p.offsets), p.length,
                                                                              p.rows_per_strip))
        elif sum(p.counts) < p.nbytes:
            raise TiffError("%s: strips hold %d bytes, the image needs %d" % (self.fName, sum(p.counts), p.nbytes))
        p.description = t.get(_DESCRIPTION, "")
        return p

    @staticmethod
    def _imagej(text):
        if not isinstance(text, str) or not text.startswith("ImageJ="):
            return {}
        out = {}
        for line in text.splitlines():
            k, _, v = line.partition("=")
            if k in ("images", "slices", "frames", "channels"):
                try:
                    if int(v) > 0:  # a negative or zero count is a damaged description: ignore the key
                        out[k] = int(v)
                except ValueError:
                    pass
        return out

    # ---- pixels ----
    @staticmethod
    def _fill(f, view):
        got = 0
        while got < len(view):
            n = f.readinto(view[got:])
            if not n:
                raise TiffError("short read")
            got += n

    def read_into(self, out, first=0, count=None):
        """Fill `out` (C-contiguous, dtype == self.dtype, count * Y * X elements) with images first .. first+count-1
        in file order, reading straight into its memory."""
        count = self.n_images - first if count is None else count
        if first < 0 or count < 0 or first + count > self.n_images:
            raise IndexError("images %d..%d of %d" % (first, first + count - 1, self.n_images))
        p0 = self.pages[0]
        if out.dtype != self.dtype or not out.flags.c_contiguous or out.nbytes != count * p0.nbytes:
            raise ValueError("read_into needs a C-contiguous %s array of %d bytes" % (self.dtype, count * p0.nbytes))
        buf = memoryview(out.reshape(-1)).cast("B")
        with open(self.fName, "rb", buffering=0) as f:
            if self._flat is not None:
                f.seek(self._flat + first * p0.nbytes)
                self._fill(f, buf)
                return out
            packed = [i for i in range(count) if self.pages[first + i].contiguous() is None
                      and (self.pages[first + i].compression != _NONE or self.pages[first + i].tile)]
            for i in range(count):
                p = self.pages[first + i]
                dst = buf[i * p0.nbytes:(i + 1) * p0.nbytes]
                if p.compression != _NONE or p.tile:
                    continue
                start = p.contiguous()
                if start is not None:
                    f.seek(start)
                    self._fill(f, dst)
                    continue
                at = 0
                for o, c in zip(p.offsets, p.counts):
                    c = min(c, p0.nbytes - at)
                    f.seek(o)
                    self._fill(f, dst[at:at + c])
                    at += c
            if packed:
                # pages are independent and every decoder releases the GIL: one page per worker
                def work(i):
                    self._decode_page(f.fileno(), self.pages[first + i], buf[i * p0.nbytes:(i + 1) * p0.nbytes])

                workers = min(len(packed), self.decode_threads or min(16, os.cpu_count() or 1))
                if workers <= 1:
                    for i in packed:
                        work(i)
                else:
                    with ThreadPoolExecutor(workers) as pool:
                        list(pool.map(work, packed))
        return out

    def _decode_page(self, fd, p, dst):
        """compressed strips of one page -> dst (a writable byte view of the page), read with pread so that several
        pages can be in flight on one descriptor; a strip decodes to rows_per_strip rows (the last one to what is
        left), writers may pad it: the surplus is dropped"""
        size = p.dtype.itemsize
        swap = 0 if p.dtype.isnative else 1
        if p.tile:
            tw, tl = p.tile
            across = -(-p.width // tw)
            image = np.frombuffer(dst, np.uint8).reshape(p.length, p.width * size)
            tile = np.empty((tl, tw * size), np.uint8)
            view = memoryview(tile.reshape(-1))
            for k, (o, c) in enumerate(zip(p.offsets, p.counts)):
                self._decode_block(fd, p, o, c, view)
                if p.predictor == 2:
                    load_codecs().spt_undo_differencing(tile.ctypes.data, tl, tw, size, swap)
                y, x = (k // across) * tl, (k % across) * tw
                h, w = min(tl, p.length - y), min(tw, p.width - x)
                image[y:y + h, x * size:(x + w) * size] = tile[:h, :w * size]
            return
        row = p.width * size
        at = 0
        for o, c in zip(p.offsets, p.counts):
            want = min(row * p.rows_per_strip, p.nbytes - at)
            self._decode_block(fd, p, o, c, dst[at:at + want])
            at += want
        if p.predictor == 2:
            page = (ctypes.c_char * p.nbytes).from_buffer(dst)
            rc = load_codecs().spt_undo_differencing(ctypes.addressof(page), p.length, p.width, size, swap)
            if rc:
                raise TiffError("%s: predictor of %d-byte samples" % (self.fName, size))

    def _decode_block(self, fd, p, o, c, part):
        """one strip or tile: c bytes at file offset o -> exactly len(part) bytes"""
        want = len(part)
        if c < 0 or o < 0 or o + c > self._size:      # checked before allocating c bytes
            raise TiffError("%s: the block at %d (%d bytes) leaves the file" % (self.fName, o, c))
        raw = os.pread(fd, c, o)
        if len(raw) < c:
            raise TiffError("%s: the block at %d leaves the file" % (self.fName, o))
        if p.compression == _NONE:
            n = min(c, want)
            part[:n] = raw[:n]
        elif p.compression in (_DEFLATE, _DEFLATE_OLD):
            try:
                got = zlib.decompressobj().decompress(raw, want)
            except zlib.error as e:
                raise TiffError("%s: damaged deflate block at %d (%s)" % (self.fName, o, e))
            n = len(got)
            part[:n] = got
        else:
            lib = load_codecs()
            fn = lib.spt_lzw_decode if p.compression == _LZW else lib.spt_packbits_decode
            n = ctypes.c_size_t(0)
            rc = fn(raw, c, ctypes.addressof((ctypes.c_char * want).from_buffer(part)), want, ctypes.byref(n))
            if rc not in (0, -3):
                raise TiffError("%s: damaged %s block at %d" % (self.fName, "LZW" if p.compression == _LZW
                                                                else "PackBits", o))
            n = n.value
        if n < want:
            raise TiffError("%s: the block at %d holds %d bytes, %d are needed" % (self.fName, o, n, want))

    def asarray(self, native=True):
        """The whole stack, shaped `self.shape`; native=True returns the machine's byte order (what the renderer's
        upload wants), converting in place if the file's differs."""
        out = np.empty(self.shape, self.dtype)
        self.read_into(out)
        if native and not out.dtype.isnative:
            out = out.byteswap().view(out.dtype.newbyteorder("="))
        return out


def read3dTiff(fName):
    """imgutils.py:22-23"""
    return TiffFile(fName).asarray()


imread = read3dTiff


def write3dTiff(data, fName, bigtiff=None):
    """imgutils.py:26-27 (argument order of the reference: data first).  2/3/4-d arrays of the supported element
    types; one strip per page, pages back to back, an ImageJ description for (T, Z, Y, X) data.  BigTIFF once the
    file would pass 4 GB (or on request)."""
    data = np.asarray(data)
    if data.ndim not in (2, 3, 4):
        raise ValueError("data.ndim = %d (not 2, 3 or 4)" % data.ndim)
    if data.dtype.kind not in "uif" or data.dtype.itemsize not in (1, 2, 4, 8) or (
            data.dtype.kind == "f" and data.dtype.itemsize < 4):
        raise TiffError("element type %s cannot be stored" % data.dtype)
    dt = data.dtype.newbyteorder("<") if data.dtype.itemsize > 1 else data.dtype
    pages = np.ascontiguousarray(data, dtype=dt).reshape((-1,) + data.shape[-2:])
    n, ny, nx = pages.shape
    page_bytes = ny * nx * dt.itemsize
    desc = b""
    if data.ndim == 4:
        desc = ("ImageJ=1.52\nimages=%d\nslices=%d\nframes=%d\nhyperstack=true\n" % (
            n, data.shape[1], data.shape[0])).encode("latin-1") + b"\0"
    big = (n * (page_bytes + 512) + len(desc) > 0xFFFF0000) if bigtiff is None else bool(bigtiff)
    off_fmt, cnt_fmt, ent_fmt, long_t = ("Q", "Q", "<HHQQ", 16) if big else ("I", "H", "<HHII", 4)
    fmt_code = {"u": 1, "i": 2, "f": 3}[dt.kind]

    def entries(image_at, desc_at):
        e = [(_WIDTH, long_t, 1, nx), (_LENGTH, long_t, 1, ny), (_BITS, 3, 1, 8 * dt.itemsize), (_COMPRESSION, 3, 1, 1),
             (_PHOTOMETRIC, 3, 1, 1)]
        if desc_at:
            e.append((_DESCRIPTION, 2, len(desc), desc_at))
        e += [(_STRIP_OFFSETS, long_t, 1, image_at), (_SAMPLES, 3, 1, 1), (_ROWS_PER_STRIP, long_t, 1, ny),
              (_STRIP_COUNTS, long_t, 1, page_bytes), (_SAMPLE_FORMAT, 3, 1, fmt_code)]
        return e

    n_ent = len(entries(0, 1 if desc else 0))
    ifd_bytes = struct.calcsize("<" + cnt_fmt) + n_ent * struct.calcsize(ent_fmt) + struct.calcsize("<" + off_fmt)
    n_ent_rest = len(entries(0, 0))
    ifd_rest = struct.calcsize("<" + cnt_fmt) + n_ent_rest * struct.calcsize(ent_fmt) + struct.calcsize("<" + off_fmt)
    header = 16 if big else 8
    # layout: header | description | all images | all directories
    desc_at = header if desc else 0
    images_at = header + len(desc)
    images_at += (-images_at) % 16
    dirs_at = images_at + n * page_bytes
    dirs_at += dirs_at % 2
    with open(fName, "wb") as f:
        if big:
            f.write(struct.pack("<2sHHHQ", b"II", 43, 8, 0, dirs_at))
        else:
            f.write(struct.pack("<2sHI", b"II", 42, dirs_at))
        f.write(desc)
        f.write(b"\0" * (images_at - header - len(desc)))
        pages.tofile(f)
        f.write(b"\0" * (dirs_at - images_at - n * page_bytes))
        at = dirs_at
        for i in range(n):
            ents = entries(images_at + i * page_bytes, desc_at if i == 0 else 0)
            size = ifd_bytes if (i == 0 and desc) else ifd_rest
            nxt = at + size if i + 1 < n else 0
            f.write(struct.pack("<" + cnt_fmt, len(ents)))
            for tag, typ, count, value in ents:
                if typ == 3 and not big:  # a SHORT value sits left-justified in the 4-byte field
                    f.write(struct.pack("<HHIHH", tag, typ, count, value, 0))
                elif typ == 3:
                    f.write(struct.pack("<HHQHHI", tag, typ, count, value, 0, 0))
                elif typ == 2 and count <= (8 if big else 4):
                    raise TiffError("description too short to be stored by reference")
                else:
                    f.write(struct.pack(ent_fmt, tag, typ, count, value))
            f.write(struct.pack("<" + off_fmt, nxt))
            at += size


def imsave(fName, data, **kw):
    """tifffile's argument order (file first), as imgutils.py:18 imports it."""
    write3dTiff(data, fName, **kw)
