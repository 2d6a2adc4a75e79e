"""Colour maps for the display hand-off (VolumeRenderer.set_lut / output_rgba, keyframes.record_keyframes).

The reference ships its maps as one-pixel-high PNG strips `colormaps/cmap_<name>.png` and reads them into
`__COLORMAPDICT__[name]`, an (N, 3) float array in [0, 1] (spimagine/config/loadcolormaps.py:29-63), which
GLWidget.set_colormap uploads as the LUT texture of gui/shaders/texture.frag.  Those image files are assets of the
reference and are not part of this package:

  * `loadcolormaps(folder)` reads any folder of such strips (the reference's own, if it is installed) with PIL;
  * "grays", "hot" and "jet" are generated from their published piecewise-linear definitions, so that the headless
    paths have the usual choices without any file.
"""
import os
import re

import numpy as np

__all__ = ["builtin", "get", "loadcolormaps", "array_from_image"]


def _piecewise(x, knots):
    xs, ys = zip(*knots)
    return np.interp(x, xs, ys)


def builtin(name, n=256):
    """(n, 3) float64 in [0, 1]: "grays", "hot" (black - red - yellow - white) or "jet" (blue - cyan - yellow - red)"""
    x = np.linspace(0., 1., n)
    if name in ("grays", "gray", "grey", "greys"):
        rgb = (x, x, x)
    elif name == "hot":
        rgb = (_piecewise(x, [(0, .0416), (.365079, 1), (1, 1)]),
               _piecewise(x, [(0, 0), (.365079, 0), (.746032, 1), (1, 1)]),
               _piecewise(x, [(0, 0), (.746032, 0), (1, 1)]))
    elif name == "jet":
        rgb = (_piecewise(x, [(0, 0), (.35, 0), (.66, 1), (.89, 1), (1, .5)]),
               _piecewise(x, [(0, 0), (.125, 0), (.375, 1), (.64, 1), (.91, 0), (1, 0)]),
               _piecewise(x, [(0, .5), (.11, 1), (.34, 1), (.65, 0), (1, 0)]))
    else:
        raise KeyError("colormap = '%s' not built in, valid: ['grays', 'hot', 'jet'] (or loadcolormaps(folder))" % name)
    return np.stack(rgb, axis=1)


def array_from_image(fName):
    """(h, w, 3) floats in [0, 1] of an RGB(A) image file (loadcolormaps.py:29-39)"""
    from PIL import Image
    with Image.open(fName) as im:
        img = np.asarray(im)
    if img.ndim < 3:
        raise TypeError("image %s appears not to be a 2d rgb image" % fName)
    return 1. / 255 * img[:, :, :3]


def loadcolormaps(basePath):
    """{name: (w, 3) array} of every cmap_<name>.png in basePath: the top row of each strip
    (loadcolormaps.py:42-63); files that do not decode are reported and left out"""
    cmaps = {}
    reg = re.compile(r"cmap_(.*)\.png")
    for fName in sorted(os.listdir(basePath)):
        match = reg.match(fName)
        if match:
            try:
                cmaps[match.group(1)] = array_from_image(os.path.join(basePath, fName))[0, :, :]
            except Exception as e:
                print(e)
                print("could not load %s" % fName)
    return cmaps


def get(name, folder=None):
    """the map called `name`: from `folder` (or $SPIMAGINE_COLORMAPS) when it holds cmap_<name>.png, else built in"""
    folder = folder or os.environ.get("SPIMAGINE_COLORMAPS")
    if folder and os.path.exists(os.path.join(folder, "cmap_%s.png" % name)):
        return array_from_image(os.path.join(folder, "cmap_%s.png" % name))[0, :, :]
    return builtin(name)
