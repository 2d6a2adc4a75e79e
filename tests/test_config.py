"""spimagine_b200.config: the reference's ~/.spimagine file (spimagine/config/config.py, myconfigparser.py) as far as
the render path reads it.  The first test is the reference's own tests/test_config/test_config.py.  CPU only."""
import os

import pytest

from spimagine_b200 import config
from spimagine_b200.config import MyConfigParser


def test_config(tmp_path):
    """tests/test_config/test_config.py:15-39 of the reference"""
    vals = {"id_platform": 99, "id_device": 101, "colormap": "foo", "texture_width": 754, "window_width": 123,
            "window_height": 123, "max_steps": 400}
    fpath = str(tmp_path / "config_example.txt")
    with open(fpath, "w") as f:
        f.write("\n".join(["%s = %s " % (k, v) for k, v in vals.items()]))
    config_parser = MyConfigParser(fpath)
    for k, v in vals.items():
        assert v == type(v)(config_parser.get(k))


def test_defaults_missing_file_and_creation(tmp_path):
    missing = str(tmp_path / "nothing_here")
    p = MyConfigParser(missing, create_file=False)
    assert not os.path.exists(missing) and p.get("max_steps") is None and p.get("max_steps", 7) == 7
    got = config.load(missing)
    assert got == {"id_device": 0, "spin_axis": 1, "id_platform": 0, "use_gpu": 1, "colormap": "viridis",
                   "texture_width": 800, "window_width": 900, "window_height": 800, "max_steps": 200,
                   "box_linewidth": 1., "interpolation": "linear", "_qualifier_constant_to_global": False}
    assert not os.path.exists(missing)                 # loading never writes
    MyConfigParser(missing)                            # the reference's parser touches the file
    assert os.path.exists(missing) and os.path.getsize(missing) == 0
    garbage = tmp_path / "garbage"
    garbage.write_text("this is [not a config\n=\n")
    assert config.load(str(garbage))["max_steps"] == 200


def test_file_syntax(tmp_path):
    f = tmp_path / "c"
    f.write_text("# a comment\nMAX_STEPS: 300\ninterpolation=nearest\n\nbox_linewidth = 2.5\n; another\nid_device = 3\n")
    got = config.load(str(f))
    assert got["max_steps"] == 300 and got["interpolation"] == "nearest" and got["box_linewidth"] == 2.5
    assert got["id_device"] == 3 and got["texture_width"] == 800
    f.write_text("max_steps = many\n")
    with pytest.raises(ValueError):
        config.load(str(f))


def test_renderer_defaults_come_from_the_file_and_the_environment(tmp_path, monkeypatch):
    f = tmp_path / "c"
    f.write_text("max_steps = 120\nid_device = 2\ntexture_width = 640\n")
    monkeypatch.setenv("SPIMAGINE_CONFIG", str(f))
    monkeypatch.delenv("SPIMAGINE_MAX_STEPS", raising=False)
    monkeypatch.delenv("SPIMAGINE_CUDA_DEVICE", raising=False)
    assert config.default_max_steps() == 120 and config.default_device() == 2
    assert config.__DEFAULTMAXSTEPS__ == 120 and config.__ID_DEVICE__ == 2 and config.__DEFAULT_TEXTURE_WIDTH__ == 640
    assert config.__DEFAULT_INTERP__ == "linear" and config.__DEFAULTCOLORMAP__ == "viridis"
    assert config.__CONFIGFILE__ == str(f)
    monkeypatch.setenv("SPIMAGINE_MAX_STEPS", "64")
    monkeypatch.setenv("SPIMAGINE_CUDA_DEVICE", "5")
    assert config.default_max_steps() == 64 and config.default_device() == 5
    with pytest.raises(AttributeError):
        config.__NO_SUCH_THING__
    # the renderer asks these two functions (no GPU here: read the source instead of constructing one)
    import inspect
    from spimagine_b200 import volumerender
    src = inspect.getsource(volumerender.VolumeRenderer.__init__)
    assert "config.default_device()" in src and "config.default_max_steps()" in src
