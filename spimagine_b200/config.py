"""User configuration: the `~/.spimagine` file of the reference (spimagine/config/config.py:14-57,
spimagine/config/myconfigparser.py:19-60) as far as the render path reads it.

The file is a list of `key = value` lines without a section header.  Keys the render path uses:

    id_device       CUDA device of a VolumeRenderer made without `device=` (the reference: OpenCL device index)
    max_steps       samples per ray, the reference's -D maxSteps build option (default 200 -> 208 MIP samples)
    interpolation   "linear" / "nearest": the GUI's initial sampler (VolumeRenderer's own default stays "linear")
    texture_width   side of the square image the GUI / spim_render render at (default 800)
    colormap, spin_axis, window_width, window_height, box_linewidth: display settings, parsed and exposed only

id_platform / use_gpu / _qualifier_constant_to_global select and work around OpenCL devices; they are parsed (so an
existing file stays valid) and ignored: there is one platform, and no CPU path.

Differences to the reference, on purpose: importing this module never creates `~/.spimagine` (the reference
touches an empty file on first import), the path can be redirected with $SPIMAGINE_CONFIG, and the environment
variables SPIMAGINE_CUDA_DEVICE / SPIMAGINE_MAX_STEPS override the file for one process.
"""
import configparser
import logging
import os
from itertools import chain

logger = logging.getLogger(__name__)

__all__ = ["MyConfigParser", "defaults", "load", "get_param"]


class MyConfigParser(configparser.ConfigParser):
    """A section-less `key = value` file (myconfigparser.py:19-60): get(key, default) never raises."""

    def __init__(self, fName=None, defaults={}, create_file=True):
        configparser.ConfigParser.__init__(self, defaults)
        self.dummySection = "dummy"
        if fName:
            if create_file and not os.path.exists(fName):
                try:
                    open(fName, "w").close()
                except Exception as e:
                    logger.debug("failed to create %s (%s)", fName, e)
            self.read(fName)

    def read(self, fName):
        try:
            with open(fName) as f:
                self.read_file(chain(("[%s]" % self.dummySection,), f))
        except Exception as e:  # a missing or malformed file means "defaults"
            logger.debug(e)

    def get(self, key, defaultValue=None, **kwargs):
        try:
            return configparser.ConfigParser.get(self, self.dummySection, key, **kwargs)
        except Exception as e:
            logger.debug("%s (%s)", e, key)
            return defaultValue


# config.py:18-31
defaults = {
    "id_device": 0,
    "spin_axis": 1,
    "id_platform": 0,
    "use_gpu": 1,
    "colormap": "viridis",
    "texture_width": 800,
    "window_width": 900,
    "window_height": 800,
    "max_steps": 200,
    "box_linewidth": 1.,
    "interpolation": "linear",
    "_qualifier_constant_to_global": 0,
}
_TYPES = {"box_linewidth": float, "colormap": str, "interpolation": str, "_qualifier_constant_to_global": bool}


def config_file():
    return os.environ.get("SPIMAGINE_CONFIG") or os.path.expanduser("~/.spimagine")


def load(fName=None):
    """-> {key: typed value} of every key in `defaults`, from fName (default: $SPIMAGINE_CONFIG or ~/.spimagine).
    A value that does not convert (max_steps = many) raises ValueError, as the reference's import does."""
    parser = MyConfigParser(fName or config_file(), create_file=False)
    return {k: get_param(k, parser) for k in defaults}


def get_param(name, parser=None):
    """config.py:34-35 `_get_param`: type(config_parser.get(name, defaults[name]))"""
    parser = parser if parser is not None else MyConfigParser(config_file(), create_file=False)
    return _TYPES.get(name, int)(parser.get(name, defaults[name]))


def default_device():
    """CUDA device of a renderer made without `device=`: $SPIMAGINE_CUDA_DEVICE, else id_device of the file"""
    env = os.environ.get("SPIMAGINE_CUDA_DEVICE")
    return int(env) if env not in (None, "") else get_param("id_device")


def default_max_steps():
    """samples per ray of a renderer made without `max_steps=`: $SPIMAGINE_MAX_STEPS, else max_steps of the file
    (the reference compiles config.__DEFAULTMAXSTEPS__ into its kernels, volumerender.py:153-160)"""
    env = os.environ.get("SPIMAGINE_MAX_STEPS")
    return int(env) if env not in (None, "") else get_param("max_steps")


def __getattr__(name):
    # the reference's module constants (config.py:38-55), evaluated when asked for instead of at import
    keys = {"__ID_DEVICE__": "id_device", "__ID_PLATFORM__": "id_platform", "__USE_GPU__": "use_gpu",
            "__DEFAULTCOLORMAP__": "colormap", "__DEFAULT_TEXTURE_WIDTH__": "texture_width",
            "__DEFAULT_WIDTH__": "window_width", "__DEFAULT_HEIGHT__": "window_height",
            "__DEFAULT_SPIN_AXIS__": "spin_axis", "__DEFAULT_BOX_LINEWIDTH__": "box_linewidth",
            "__DEFAULTMAXSTEPS__": "max_steps", "__DEFAULT_INTERP__": "interpolation",
            "__QUALIFIER_CONSTANT_TO_GLOBAL__": "_qualifier_constant_to_global"}
    if name == "__CONFIGFILE__":
        return config_file()
    if name in keys:
        return get_param(keys[name])
    raise AttributeError(name)
