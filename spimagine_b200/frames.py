"""Frame sources for 3D+t playback: the data containers either side of VolumeRenderer.update_data.

Mirrors the part of the reference's data model that feeds the renderer during timelapse playback
(spimagine/models/data_model.py: GenericData :59-95, SpimData :97-148, TiffData :178-218, RawData :221-261,
RawMultipleFiles / TiffFolderData / TiffMultipleFiles :262-404, NumpyData :408-432, XwingData :475-515, DataModel :652-757, the
prefetching DataLoadThread / DataModel :600-757; spimagine/utils/imgutils.py: parseIndexFile :48-66, parseMetaFile
:69-86, fromSpimFolder :129-146, createSpimFolder :162-191) -- same class names, same `sizeT() / size() /
stackUnits / container[pos]` protocol -- with one change in where the bytes land: `FrameSource` reads time points
AHEAD of the playback position straight from the file into PAGE-LOCKED buffers (file.readinto, no intermediate
array), so that TimelapsePlayer.play(..., pinned=True) / VolumeRenderer.update_data(..., pinned=True) can move them
at PCIe rate without staging.  The reference reads every time point into fresh pageable memory (np.fromfile) on the
GUI thread or a Qt thread and re-uploads it synchronously (gui/glwidget.py:372-374).

Only containers whose bytes are laid out as the renderer wants them (C-order stacks of one element type) are
rebuilt here: raw files, SpimData folders and TIFF stacks (TiffData over utils/tiffio.py: stored as they are, or
LZW / deflate / PackBits strips or tiles decoded one page per worker thread).  CZI files of uncompressed greyscale sub-blocks are read by utils/cziio.py (CZIData).
JPEG-compressed TIFF and compressed CZI decode through third-party libraries in the reference (tifffile, czifile); their arrays can be wrapped in NumpyData or any
object with the same protocol.
"""
from __future__ import absolute_import, print_function

import glob
import logging
import os
import re
import threading

import numpy as np

from . import _lib

logger = logging.getLogger(__name__)


# ---------------------------------------------------------------------------------------------- containers
class GenericData(object):
    """abstract base class for 4d data: overwrite size() and __getitem__() (data_model.py:59-95)"""
    dataFileError = Exception("not a valid file")

    def __init__(self, name=""):
        self.stackSize = None
        self.stackUnits = None
        self.name = name

    def sizeT(self):
        return self.size()[0]

    def size(self):
        return self.stackSize

    def __len__(self):
        return int(self.sizeT())

    def __getitem__(self, pos):
        return None

    # additions used by FrameSource: element type of a time point and a way to read one into caller memory
    @property
    def dtype(self):
        return np.dtype(np.uint16)

    def read_into(self, pos, out):
        """Fill `out` (C-contiguous, shape size()[1:], this container's dtype) with time point `pos`."""
        np.copyto(out, self[pos], casting="no")


def parseIndexFile(fname):
    """returns (t,z,y,x) dimensions of a spim stack (imgutils.py:48-66)"""
    try:
        lines = open(fname).readlines()
    except IOError:
        print("could not open and read ", fname)
        return None
    items = lines[0].replace("\t", ",").split(",")
    try:
        stackSize = [int(i) for i in items[-4:-1]] + [len(lines)]
    except Exception as e:
        print(e)
        print("couldnt parse ", fname)
        return None
    stackSize.reverse()
    return stackSize


def parseMetaFile(fName):
    """returns pixelSizes (dx,dy,dz) (imgutils.py:69-86; np.float is gone from numpy: float())"""
    with open(fName) as f:
        s = f.read()
        try:
            z1 = float(re.findall("StartZ.*", s)[0].split("\t")[2])
            z2 = float(re.findall("StopZ.*", s)[0].split("\t")[2])
            zN = float(re.findall("NumberOfPlanes.*", s)[0].split("\t")[2])
            return (.162, .162, (1. * z2 - z1) / (zN - 1.))
        except Exception as e:
            print(e)
            print("coulndt parse ", fName)
            return (1., 1., 1.)


class SpimData(GenericData):
    """data class for spim data saved in folder fName (data_model.py:97-148)
    fname/
    |-- metadata.txt
    |-- data/
       |--data.bin      little-endian uint16, (t, z, y, x)
       |--index.txt
    """

    def __init__(self, fName=""):
        super(SpimData, self).__init__(fName)
        self.load(fName)

    def load(self, fName):
        if fName:
            try:
                self.stackSize = parseIndexFile(os.path.join(fName, "data/index.txt"))
                self.stackUnits = parseMetaFile(os.path.join(fName, "metadata.txt"))
                if self.stackSize is None:
                    raise IOError("no index file")
                self.fName = fName
            except Exception as e:
                print(e)
                self.fName = ""
                raise Exception("couldnt open %s as SpimData" % fName)

    def _offset(self, pos):
        if pos < 0 or pos >= self.stackSize[0]:
            raise IndexError("0 <= pos <= %i, but pos = %i" % (self.stackSize[0] - 1, pos))
        voxels = int(np.prod(self.stackSize[1:], dtype=np.int64))
        return 2 * int(pos) * voxels, voxels  # python ints: no overflow for big files

    def __getitem__(self, pos):
        if self.stackSize and self.fName:
            offset, voxels = self._offset(pos)
            with open(os.path.join(self.fName, "data/data.bin"), "rb") as f:
                f.seek(offset)
                return np.fromfile(f, dtype="<u2", count=voxels).reshape(self.stackSize[1:])
        return None

    def read_into(self, pos, out):
        offset, voxels = self._offset(pos)
        buf = memoryview(out.reshape(-1)).cast("B")
        assert out.dtype == np.dtype("<u2") and out.flags.c_contiguous and len(buf) == 2 * voxels
        with open(os.path.join(self.fName, "data/data.bin"), "rb", buffering=0) as f:
            f.seek(offset)
            got = 0
            while got < len(buf):
                n = f.readinto(buf[got:])
                if not n:
                    raise IOError("%s: short read of time point %d" % (self.fName, pos))
                got += n


class RawData(GenericData):
    """one raw file holding a (t,) z, y, x stack of `dtype` (data_model.py:221-261).  The reference loads the whole
    file with np.fromfile; here it is memory-mapped, so that a timelapse larger than host memory plays too.  A 3-d
    shape is one time point here; the reference's padding expression ((1,) * (len(shape) - 4), :236-237) is a no-op
    and leaves it 3-d, which makes every slice a time point of a 2-d image."""

    def __init__(self, fName="", shape=None, dtype=np.uint16):
        GenericData.__init__(self, fName)
        self.load(fName, shape, dtype)

    def load(self, fname, shape=None, dtype=np.uint16, stackUnits=[1., 1., 1.]):
        if fname:
            if shape is None or dtype is None:
                raise ValueError("RawData needs shape and dtype (the reference asks for them in a dialog)")
            shape = tuple(int(s) for s in shape)
            if len(shape) < 4:
                shape = (1,) * (4 - len(shape)) + shape
            elif len(shape) > 4:
                raise ValueError("shape should have length of 4!")
            try:
                self.data = np.memmap(fname, dtype=np.dtype(dtype), mode="r", shape=shape)
            except Exception as e:
                print(e)
                self.fName = ""
                raise Exception("couldnt open %s as RawData" % fname)
            self.stackSize = shape
            self.stackUnits = stackUnits
            self.fName = fname
            self._dtype = np.dtype(dtype)

    @property
    def dtype(self):
        return self._dtype

    def __getitem__(self, pos):
        if self.stackSize and self.fName:
            return self.data[pos]
        return None


class TiffData(GenericData):
    """2/3/4d tiff data (data_model.py:178-218).  The reference decodes the whole file into memory with tifffile;
    here only the page directory is parsed (utils/tiffio.py: strips as stored or LZW / deflate / PackBits, classic /
    BigTIFF, ImageJ hyperstacks) and a time point is read from the file when it is asked for -- by FrameSource straight into a
    page-locked buffer.  Big-endian files are byte-swapped after the read."""

    def __init__(self, fName=""):
        GenericData.__init__(self, fName)
        self.load(fName)

    def load(self, fName, stackUnits=[1., 1., 1.]):
        if fName:
            from .utils.tiffio import TiffFile
            try:
                tif = TiffFile(fName)
                # np.squeeze of the leading axes as in the reference: (1, Y, X) is one slice, (T, 1, Y, X) one volume
                shape = tuple(s for s in tif.shape[:-2] if s != 1) + tuple(tif.shape[-2:])
                self.stackSize = (1,) * (4 - len(shape)) + shape
            except Exception as e:
                print(e)
                self.fName = ""
                raise Exception("couldnt open %s as TiffData (%s)" % (fName, str(e)))
            self._tif = tif
            self.stackUnits = stackUnits
            self.fName = fName

    @property
    def dtype(self):
        return self._tif.dtype.newbyteorder("=")

    def read_into(self, pos, out):
        if pos < 0 or pos >= self.stackSize[0]:
            raise IndexError("0 <= pos <= %i, but pos = %i" % (self.stackSize[0] - 1, pos))
        nz = self.stackSize[1]
        raw = out.view(self._tif.dtype) if out.dtype.itemsize > 1 else out
        self._tif.read_into(raw, first=int(pos) * nz, count=nz)
        if not self._tif.dtype.isnative:
            out.byteswap(inplace=True)

    def __getitem__(self, pos):
        if self.stackSize and self.fName:
            out = np.empty(self.stackSize[1:], self.dtype)
            self.read_into(pos, out)
            return out
        return None


class _FilePerTimePoint(GenericData):
    """Time point t is file t of a sorted list (data_model.py:262-404: RawMultipleFiles, TiffFolderData,
    TiffMultipleFiles).  Sizes come from the first file; every read goes straight into the caller's buffer."""

    def _read_file_into(self, fname, out):
        raise NotImplementedError

    def read_into(self, pos, out):
        if pos < 0 or pos >= len(self.fNames):
            raise IndexError("0 <= pos <= %i, but pos = %i" % (len(self.fNames) - 1, pos))
        self._read_file_into(self.fNames[pos], out)

    def __getitem__(self, pos):
        if len(self.fNames) > 0 and pos < len(self.fNames):
            out = np.empty(tuple(self.stackSize[1:]), self.dtype)
            self.read_into(pos, out)
            return out
        return None


class RawMultipleFiles(_FilePerTimePoint):
    """2/3d raw data, one file per time point (data_model.py:262-308).  shape describes ONE file: (1, z, y, x) as
    the reference wants it (it drops the first entry, :290), or (z, y, x) / (y, x), which the reference cannot
    reshape."""

    def __init__(self, fnames=[], shape=None, dtype=None):
        GenericData.__init__(self, "[" + ", ".join(fnames) + "]")
        self.fNames = self.fnames = list(fnames)
        self.load(self.fNames, shape, dtype)

    def load(self, fnames, shape, dtype, stackUnits=[1., 1., 1.]):
        if fnames:
            if shape is None or dtype is None:
                raise ValueError("RawMultipleFiles needs shape and dtype (the reference asks for them in a dialog)")
            shape = tuple(int(s) for s in shape)
            if len(shape) == 4:
                if shape[0] != 1:
                    raise ValueError("a 4-d shape describes one file: its first entry must be 1")
                shape = shape[1:]
            elif len(shape) > 4:
                raise ValueError("shape should have length of at most 4!")
            shape = (1,) * (3 - len(shape)) + shape
            self._dtype = np.dtype(dtype)
            need = int(np.prod(shape, dtype=np.int64)) * self._dtype.itemsize
            if os.path.getsize(fnames[0]) < need:
                raise Exception("couldnt open %s as RawData" % fnames[0])
            self.stackSize = (len(fnames),) + shape
            self.stackUnits = stackUnits

    @property
    def dtype(self):
        return self._dtype

    def _read_file_into(self, fname, out):
        buf = memoryview(out.reshape(-1)).cast("B")
        with open(fname, "rb", buffering=0) as f:
            got = 0
            while got < len(buf):
                n = f.readinto(buf[got:])
                if not n:
                    raise IOError("%s: short read" % fname)
                got += n


def parse_index_xwing(fname):
    """(z, y, x) of an xwing stack: the last three numbers of the first line of default.index.txt, reversed
    (imgutils.py:89-107)"""
    try:
        with open(fname) as f:
            items = f.readline().replace("\t", ",").replace("\n", "").split(",")
        return [int(i) for i in items[-3:]][::-1]
    except Exception as e:
        print(e)
        print("couldnt parse ", fname)
        return None


def parse_meta_xwing(fName):
    """pixel sizes (dx, dy, dz) from the JSON on the first line of default.metadata.txt (imgutils.py:110-127)"""
    import json
    try:
        with open(fName) as f:
            s = json.loads(f.readline())
        return (float(s["VoxelDimX"]), float(s["VoxelDimY"]), float(s["VoxelDimZ"]))
    except Exception as e:
        print(e)
        print("coulndt parse ", fName)
        return (1., 1., 1.)


class XwingData(RawMultipleFiles):
    """xwing data saved in folder fName (data_model.py:475-515): default.index.txt, default.metadata.txt and one
    little-endian uint16 stack per time point under stacks/default/*.raw"""

    def __init__(self, dirname=""):
        GenericData.__init__(self, dirname)
        self.fNames = self.fnames = self._stack_names = []
        if dirname:
            try:
                names = sorted(glob.glob(os.path.join(dirname, "stacks", "default", "*.raw")))
                shape = parse_index_xwing(os.path.join(dirname, "default.index.txt"))
                if shape is None or not names:
                    raise IOError("no index file or no stacks")
                self.fNames = self.fnames = self._stack_names = names
                self._dtype = np.dtype("<u2")
                self.stackSize = [len(names)] + shape
                self.stackUnits = parse_meta_xwing(os.path.join(dirname, "default.metadata.txt"))
            except Exception as e:
                print(e)
                self.fNames = self.fnames = self._stack_names = []
                raise Exception("couldnt open %s as XwingData" % dirname)

    def __getitem__(self, pos):
        if self.stackSize and len(self._stack_names) > 0:
            if pos < 0 or pos >= self.stackSize[0]:
                raise IndexError("0 <= pos <= %i, but pos = %i" % (self.stackSize[0] - 1, pos))
            return super(XwingData, self).__getitem__(pos)
        return None


class TiffMultipleFiles(_FilePerTimePoint):
    """2/3d tiff data, one file per time point (data_model.py:364-404)"""

    def __init__(self, fName=[]):
        GenericData.__init__(self, "[" + ", ".join(fName) + "]")
        self.fNames = list(fName)
        self.load(self.fNames)

    def load(self, fNames, stackUnits=[1., 1., 1.]):
        if fNames:
            from .utils.tiffio import TiffFile
            try:
                first = TiffFile(fNames[0])
                single = tuple(first.shape)
                if len(single) != 3:
                    raise Exception("tiff stacks seem to be neither 2d nor 3d")
                self._file_dtype = first.dtype
                self.stackSize = (len(fNames),) + single
            except Exception as e:
                print(e)
                self.fName = ""
                raise Exception("couldnt open %s as TiffData" % fNames[0])
            self.stackUnits = stackUnits

    @property
    def dtype(self):
        return self._file_dtype.newbyteorder("=")

    def _read_file_into(self, fname, out):
        from .utils.tiffio import TiffFile
        tif = TiffFile(fname)
        if tuple(tif.shape) != tuple(self.stackSize[1:]) or tif.dtype.newbyteorder("=") != self.dtype:
            raise ValueError("%s: %s %s, expected %s %s" % (fname, tif.shape, tif.dtype, self.stackSize[1:], self.dtype))
        tif.read_into(out.view(tif.dtype) if out.dtype.itemsize > 1 else out)
        if not tif.dtype.isnative:
            out.byteswap(inplace=True)


class TiffFolderData(TiffMultipleFiles):
    """3d tiff data inside a folder: every *.tif / *.tiff, sorted by name (data_model.py:311-361)"""

    def __init__(self, fName=""):
        GenericData.__init__(self, fName)
        self.fNames = []
        self.fName = ""
        if fName:
            names = sorted(f for f in glob.glob(os.path.join(fName, "*")) if re.match(r".*\.(tif|tiff)", f))
            if len(names) == 0:
                raise Exception("folder %s seems to be empty" % fName)
            self.fNames = names
            self.load(names)
            self.fName = fName


class NumpyData(GenericData):
    """a (t,) z, y, x array already in memory (data_model.py:408-432)"""

    def __init__(self, data, stackUnits=[1., 1., 1.], copy=False):
        GenericData.__init__(self, "NumpyData")
        if data.ndim not in (2, 3, 4):
            raise TypeError("data should be 3 or 4 dimensional! shape = %s" % str(data.shape))
        self.data = (data.copy() if copy else data).reshape((1,) * (4 - data.ndim) + data.shape)
        self.stackSize = self.data.shape
        self.stackUnits = stackUnits

    @property
    def dtype(self):
        return self.data.dtype

    def __getitem__(self, pos):
        return self.data[pos]


def fromSpimFolder(fName, dataFileName="data/data.bin", indexFileName="data/index.txt", pos=0, count=1):
    """`count` time points of a SpimData folder from `pos` on as one (t, z, y, x) array (imgutils.py:132-148): pos is
    clamped to the folder, count to what is left behind pos, count <= 0 means everything from pos on."""
    stackSize = parseIndexFile(os.path.join(fName, indexFileName))
    if not stackSize:
        return None
    stackSize = list(stackSize)
    pos = max(min(pos, stackSize[0] - 1), 0)
    stackSize[0] = min(count, stackSize[0] - pos) if count > 0 else max(0, stackSize[0] - pos)
    voxels = int(np.prod(stackSize[1:], dtype=np.int64))
    with open(os.path.join(fName, dataFileName), "rb") as f:
        f.seek(2 * pos * voxels)
        return np.fromfile(f, dtype="<u2", count=stackSize[0] * voxels).reshape(stackSize)


class CZIData(GenericData):
    """czi files (data_model.py:557-584): the squeezed array of the file must be 3-d (one stack) or 4-d (t, z, y, x).
    The reference decodes the whole file in the constructor; here only the directory is parsed (utils/cziio.py:
    uncompressed greyscale sub-blocks) and a time point is read when it is asked for.  A file that cannot be opened
    raises (the reference prints the error and leaves a container without a size)."""

    def __init__(self, fName=None):
        GenericData.__init__(self, fName)
        self.load(fName)

    def load(self, fName, stackUnits=[1., 1., 1.]):
        if fName:
            from .utils.cziio import CziFile
            try:
                self._czi = CziFile(fName)
                squeezed = tuple(n for n in self._czi.shape if n != 1)
                if len(squeezed) not in (3, 4):
                    raise ValueError("in file %s: data.ndim = %s (not 3 or 4)" % (fName, len(squeezed)))
            except Exception as e:
                print(e)
                self.fName = ""
                raise Exception("couldnt open %s as CZIData" % fName)
            self._squeezed = squeezed
            # a 4-d file whose leading axis is T is read one time point at a time
            axes = [a for a, n in zip(self._czi.axes, self._czi.shape) if n != 1]
            self._by_time = len(squeezed) == 4 and axes[0] == "T"
            self._all = None
            self.stackSize = squeezed if len(squeezed) == 4 else (1,) + squeezed
            self.stackUnits = stackUnits
            self.fName = fName

    @property
    def dtype(self):
        return self._czi.dtype.newbyteorder("=")

    def __getitem__(self, pos):
        if self._by_time:
            return self._czi.time_point(pos).reshape(self._squeezed[1:])
        if self._all is None:
            self._all = self._czi.asarray().reshape(self._squeezed)
        return self._all if len(self._squeezed) == 3 else self._all[pos]

    def read_into(self, pos, out):
        """sub-blocks of time point `pos` pasted straight into `out` (FrameSource's page-locked buffers)"""
        if self._by_time and out.flags.c_contiguous and out.dtype == self.dtype:
            self._czi.time_point(pos, out=out)
        else:
            GenericData.read_into(self, pos, out)


class OverlayData(GenericData):
    """Two volumes of one shape wiped over each other along `axis` (models/overlay_volumes.py:9-53): time point i shows
    y in front of position i and x from i on, so data[0] is x and data[n] is y (n = the length of that axis,
    n + 1 time points).  The reference keeps one output array and repaints only the slab between the last position
    and the new one; so does this class (a step of the slider costs one slab, not one volume), and like there the
    array that is returned is reused by the next call."""

    def __init__(self, x, y, axis=-1):
        super(OverlayData, self).__init__()
        if x.shape != y.shape:
            raise ValueError("shapes of the two arrays have to be equal!")
        self.x, self.y, self.axis = x, y, axis
        self.out = x.copy()
        self._last_index = 0

    @property
    def dtype(self):
        return self.out.dtype

    def _slab(self, lo, hi):
        index = [slice(None)] * self.out.ndim
        index[self.axis] = slice(lo, hi)
        return tuple(index)

    def __getitem__(self, i):
        if i != self._last_index:
            lo, hi = min(i, self._last_index), max(i, self._last_index)
            source = self.y if i > self._last_index else self.x   # moving on uncovers y, moving back restores x
            self.out[self._slab(lo, hi)] = source[self._slab(lo, hi)]
            self._last_index = i
        return self.out

    def size(self):
        return (self.out.shape[self.axis] + 1,) + tuple(self.out.shape)


class Img2dData(GenericData):
    """one 2-d image as a (1, 1, Y, X) stack (data_model.py:150-175).  The reference decodes through
    `imgutils.openImageFile`, which its imgutils does not define, so there every file ends in "couldnt open ... as
    Img2dData"; here png / jpg / bmp are decoded by PIL where it is installed (colour images as their luminance,
    8-bit -> uint8, 16-bit -> uint16, anything else -> float32), with the same exception otherwise."""

    def __init__(self, fName=""):
        GenericData.__init__(self, fName)
        self.load(fName)

    def load(self, fName, stackUnits=[1., 1., 1.]):
        if fName:
            try:
                from PIL import Image
                with Image.open(fName) as im:
                    if im.mode in ("1", "P", "RGB", "RGBA", "LA", "CMYK", "YCbCr"):
                        im = im.convert("L")
                    a = np.asarray(im)
                if a.dtype not in (np.uint8, np.uint16):
                    a = a.astype(np.uint16 if a.dtype.kind in "ui" and a.min() >= 0 and a.max() < 65536 else np.float32)
                self.img = np.ascontiguousarray(a)[None]
                self.stackSize = (1,) + self.img.shape
            except Exception as e:
                print(e)
                self.fName = ""
                raise Exception("couldnt open %s as Img2dData" % fName)
            self.stackUnits = stackUnits
            self.fName = fName

    @property
    def dtype(self):
        return self.img.dtype

    def __getitem__(self, pos):
        if self.stackSize and self.fName:
            return self.img
        return None


class DemoData(GenericData):
    """The synthetic demo volume (data_model.py:434-472): a shell with ten meridian stripes plus an off-centre blob,
    float32, fading by exp(-0.3 t) over the time points.  DemoData(N) is N^3 with sizeT() == N, as in the reference;
    DemoData() there reads a logo stack shipped with the package (10 time points of 80^3) -- that asset is not part
    of this package, so the 80^3 synthetic volume stands in for it with the same size() and sizeT()."""

    def __init__(self, N=None):
        GenericData.__init__(self, "DemoData")
        self.load(N)

    def load(self, N=None):
        self.fName = ""
        self.stackUnits = (1, 1, 1)
        if N is None:
            N = 80
            self.stackSize = (10, N, N, N)
            self.nT = 10
        else:
            self.stackSize = (1, N, N, N)
            self.nT = N
        x = np.linspace(-1, 1, N)
        Z, Y, X = np.meshgrid(x, x, x, indexing="ij")
        R = np.sqrt(X ** 2 + Y ** 2 + Z ** 2)
        R2 = np.sqrt((X - .4) ** 2 + (Y + .2) ** 2 + Z ** 2)
        phi = np.arctan2(Z, Y)
        theta = np.arctan2(X, np.sqrt(Y ** 2 + Z ** 2))
        bend = np.exp(-np.sin(2 * (phi + np.pi / 2.)))
        # python's sum over the stripes, starting from the integer 0, in the reference's order
        stripes = sum(np.exp(-150 * (-theta - t + .1 * (t - np.pi / 2.) * bend) ** 2)
                      for t in np.linspace(-np.pi / 2., np.pi / 2., 10))
        u = np.exp(-500 * (R - 1.) ** 2) * stripes * (1 + Z)
        u2 = np.exp(-7 * R2 ** 2)
        self.data = (10000 * (u + 2 * u2)).astype(np.float32)

    @property
    def dtype(self):
        return np.dtype(np.float32)

    def sizeT(self):
        return self.nT

    def __getitem__(self, pos):
        return (self.data * np.exp(-.3 * pos)).astype(np.float32)


class EmptyData(GenericData):
    """one uint16 voxel of zeros: what the GUI shows before anything is loaded (data_model.py:518-531)"""

    def __init__(self):
        GenericData.__init__(self, "EmptyData")
        self.stackSize = (1, 1, 1, 1)
        self.fName = ""
        self.nT = 1
        self.stackUnits = (1, 1, 1)
        self.data = np.zeros((1, 1, 1)).astype(np.uint16)

    @property
    def dtype(self):
        return self.data.dtype

    def sizeT(self):
        return self.nT

    def __getitem__(self, pos):
        return self.data


def createSpimFolder(fName, data=None, stackSize=[10, 10, 32, 32], stackUnits=(.162, .162, .162)):
    """Write a SpimData folder (imgutils.py:162-191; the reference opens data.bin with the invalid mode "wa")."""
    os.makedirs(os.path.join(fName, "data"), exist_ok=True)
    if data is not None:
        stackSize = data.shape
        with open(os.path.join(fName, "data/data.bin"), "wb") as f:
            np.ascontiguousarray(data).astype("<u2").tofile(f)
    Nt, Nz, Ny, Nx = stackSize
    with open(os.path.join(fName, "data/index.txt"), "w") as f:
        for i in range(Nt):
            f.write("%i\t0.0000\t1,%i,%i,%i\t0\n" % (i, Nx, Ny, Nz))
    with open(os.path.join(fName, "metadata.txt"), "w") as f:
        f.write("timelapse.NumberOfPlanes\t=\t%i\t0\n" % Nz)
        f.write("timelapse.StartZ\t=\t0\t0\n")
        f.write("timelapse.StopZ\t=\t%.2f\t0\n" % (stackUnits[2] * (Nz - 1.)))


# ---------------------------------------------------------------------------------------------- prefetching source
class FrameSource(object):
    """Time points of a container, read ahead into a ring of page-locked buffers by a background thread.

        src = FrameSource(SpimData(folder), frames=player.my_frames(n), depth=3)
        for t, rend in player.play(src, frames=src.frames, pinned=True, max_val=...):
            ...

    source[t] blocks until time point t is in memory and returns a page-locked ndarray (shape size()[1:]) that stays
    valid until the frame after the next one has been requested -- long enough for the asynchronous upload and the
    render of frame t (TimelapsePlayer.play touches one frame at a time); the reader runs depth - 2 frames ahead
    (depth = 2: no read-ahead, every request waits for the disk).  Frames must be requested in
    the order given by `frames` (a playback order, possibly a rank's share t = rank, rank + world, ...; it wraps
    around for looping playback).  Replaces DataLoadThread / DataModel.prefetch (data_model.py:600-757): same idea --
    the neighbourhood ahead of the position is loaded in the background -- without the per-frame allocation and with
    the bytes landing where the DMA engine can read them.
    """

    def __init__(self, container, frames=None, depth=3, pinned=True):
        if depth < 2:
            raise ValueError("depth must be at least 2")
        self.container = container
        self.frames = list(range(len(container))) if frames is None else list(frames)
        self.depth = int(depth)
        shape = tuple(int(s) for s in container.size()[1:])
        alloc = _lib.pinned_empty if pinned else (lambda s, d: np.empty(s, d))
        self._ring = [alloc(shape, container.dtype) for _ in range(self.depth)]
        self._slot_frame = [None] * self.depth   # play-order index held by each slot
        self._next_load = 0                      # play-order index the reader loads next
        self._next_get = 0                       # play-order index the consumer asks for next
        self._cv = threading.Condition()
        self._stop = False
        self._error = None
        self.bytes_read = 0
        self._thread = threading.Thread(target=self._run, name="spimagine-frame-reader", daemon=True)
        self._thread.start()

    # container protocol
    def __len__(self):
        return len(self.container)

    def sizeT(self):
        return self.container.sizeT()

    def size(self):
        return self.container.size()

    @property
    def stackUnits(self):
        return self.container.stackUnits

    def _run(self):
        try:
            while True:
                with self._cv:
                    # The consumer keeps the frame it asked for last and the one before (G = frames asked for so
                    # far: G-1 and G-2 are in use), frame k lives in slot k % depth, so frame k may be loaded once
                    # k - depth <= G - 3: the reader runs depth - 2 frames ahead of the one being shown.
                    while not self._stop and (not self.frames or self._next_load > self._next_get + self.depth - 3):
                        self._cv.wait()
                    if self._stop:
                        return
                    k = self._next_load
                    slot = k % self.depth
                t = self.frames[k % len(self.frames)]
                self.container.read_into(t, self._ring[slot])   # file.readinto releases the GIL
                with self._cv:
                    self._slot_frame[slot] = k
                    self._next_load = k + 1
                    self.bytes_read += self._ring[slot].nbytes
                    self._cv.notify_all()
        except Exception as e:  # surfaced by the next __getitem__
            with self._cv:
                self._error = e
                self._cv.notify_all()

    def __getitem__(self, t):
        with self._cv:
            k = self._next_get
            if not self.frames or self.frames[k % len(self.frames)] != t:
                raise IndexError("FrameSource: time point %r requested out of play order (next is %r)" % (
                    t, self.frames[k % len(self.frames)] if self.frames else None))
            self._next_get = k + 1
            self._cv.notify_all()          # frame k - 2 is released: the reader may go one frame further
            while self._slot_frame[k % self.depth] != k:
                if self._error is not None:   # the reader died before it got to this frame
                    raise self._error
                self._cv.wait()
            return self._ring[k % self.depth]

    def close(self):
        with self._cv:
            self._stop = True
            self._cv.notify_all()
        self._thread.join()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---------------------------------------------------------------------------------------------- data model
class DataModel(object):
    """The reference's data model without Qt (data_model.py:652-757): a container plus a position, time points cached
    in `data` (pos -> array), the neighbourhood pos .. pos + prefetchSize (mod sizeT) kept loaded by a background
    thread (DataLoadThread, :600-649), container chosen from a path by loadFromPath.  Same method names; the Qt
    signals _dataSourceChanged / _dataPosChanged are lists of callables here (source_changed, pos_changed).
    For playback at PCIe rate use FrameSource (page-locked ring, fixed play order); this class is the drop-in for
    callers that hop around the time axis (the GUI's slider)."""

    def __init__(self, dataContainer=None, prefetchSize=0):
        assert prefetchSize >= 0
        self.source_changed, self.pos_changed = [], []
        self._lock = threading.Lock()
        self._wake = threading.Condition(self._lock)
        self._thread = None
        self._stopped = True
        self.dataContainer = None
        if dataContainer:
            self.setContainer(dataContainer, prefetchSize)

    @classmethod
    def fromPath(cls, fName, prefetchSize=0):
        d = cls()
        d.loadFromPath(fName, prefetchSize)
        return d

    def setContainer(self, dataContainer=None, prefetchSize=0):
        self.stopDataLoadThread()
        self.dataContainer = dataContainer
        self.prefetchSize = prefetchSize
        self.nset = [0]
        self.data = {}
        self.__dict__.pop("pos", None)
        if self.dataContainer:
            self._stopped = False
            self._thread = threading.Thread(target=self._run, name="spimagine-data-load", daemon=True)
            self._thread.start()
            for f in self.source_changed:
                f()
            self.setPos(0)

    def __repr__(self):
        return "DataModel: %s \t %s" % (self.dataContainer.name, self.size())

    def stopDataLoadThread(self):
        with self._wake:
            self._stopped = True
            self._wake.notify_all()
        if self._thread is not None:
            self._thread.join()
            self._thread = None

    close = stopDataLoadThread

    def __del__(self):
        try:
            self.stopDataLoadThread()
        except Exception:
            pass

    def _run(self):
        """keep exactly the neighbourhood loaded: drop what left it, load what entered it"""
        while True:
            with self._wake:
                while not self._stopped and set(self.data) == set(self.nset):
                    self._wake.wait()
                if self._stopped:
                    return
                want = list(self.nset)
                for k in set(self.data).difference(want):
                    del self.data[k]
                missing = [k for k in want if k not in self.data]
                container = self.dataContainer
            for k in missing:
                newdata = container[k]           # the read happens outside the lock
                with self._wake:
                    if self._stopped or container is not self.dataContainer:
                        return
                    if k in self.nset:
                        self.data[k] = newdata

    def prefetch(self, pos):
        with self._wake:
            self.nset[:] = [int(k) for k in self.neighborhood(pos)]
            self._wake.notify_all()

    def sizeT(self):
        if self.dataContainer:
            return self.dataContainer.sizeT()

    def size(self):
        if self.dataContainer:
            return self.dataContainer.size()

    def name(self):
        if self.dataContainer:
            return self.dataContainer.name

    def stackUnits(self):
        if self.dataContainer:
            return self.dataContainer.stackUnits

    def setPos(self, pos):
        if pos < 0 or pos >= self.sizeT():
            raise IndexError("setPos(pos): %i outside of [0,%i]!" % (pos, self.sizeT() - 1))
        if not hasattr(self, "pos") or self.pos != pos:
            self.pos = pos
            for f in self.pos_changed:
                f(pos)
            self.prefetch(self.pos)

    def __getitem__(self, pos):
        if not hasattr(self, "data"):
            print("something is wrong in datamodel as its lacking a 'data' atttribute!")
            return None
        with self._wake:
            newdata = self.data.get(pos)
        if newdata is None:
            newdata = self.dataContainer[pos]
            with self._wake:
                self.data[pos] = newdata
        self.prefetch(pos)
        return newdata

    def neighborhood(self, pos):
        return np.arange(pos, pos + self.prefetchSize + 1) % self.sizeT()

    def loadFromPath(self, fName, prefetchSize=0):
        """data_model.py:733-757: lists of tif / raw files, a tif / raw file, a SpimData / xwing / tiff folder.
        png / jpg / bmp images (decoded by PIL), czi files.  (raw files need a shape: give RawData /
        RawMultipleFiles to setContainer instead.)"""
        if isinstance(fName, (tuple, list)):
            if re.match(r".*\.(tif|tiff)", fName[0]):
                self.setContainer(TiffMultipleFiles(fName), prefetchSize)
            else:
                raise ValueError("a list of %s: only lists of tif files can be opened from their paths alone" % fName[0])
        elif re.match(r".*\.(tif|tiff)", fName):
            self.setContainer(TiffData(fName), prefetchSize=0)
        elif re.match(r".*\.(png|jpg|bmp)", fName):
            self.setContainer(Img2dData(fName), prefetchSize=0)
        elif re.match(r".*\.czi", fName):
            self.setContainer(CZIData(fName), prefetchSize=0)
        elif os.path.isdir(fName):
            if os.path.exists(os.path.join(fName, "metadata.txt")):
                self.setContainer(SpimData(fName), prefetchSize)
            elif os.path.exists(os.path.join(fName, "default.index.txt")):
                self.setContainer(XwingData(fName), prefetchSize)
            else:
                self.setContainer(TiffFolderData(fName), prefetchSize=prefetchSize)
        else:
            raise ValueError("%s: no container for this path (tif / png / jpg / bmp / czi file, SpimData / xwing / tiff folder)" % fName)
