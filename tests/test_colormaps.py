"""spimagine_b200.colormaps: LUTs for the display hand-off (spimagine/config/loadcolormaps.py:29-63).  CPU only."""
import numpy as np
import pytest

from spimagine_b200 import colormaps

PIL_Image = pytest.importorskip("PIL.Image")


def test_builtin_maps():
    g = colormaps.builtin("grays")
    assert g.shape == (256, 3) and np.array_equal(g[:, 0], np.linspace(0, 1, 256)) and np.array_equal(g[:, 0], g[:, 2])
    for name in ("hot", "jet"):
        m = colormaps.builtin(name, 64)
        assert m.shape == (64, 3) and m.min() >= 0 and m.max() <= 1
    hot = colormaps.builtin("hot")
    assert np.all(np.diff(hot, axis=0) >= 0)                       # every channel rises
    assert hot[0].tolist() == [.0416, 0, 0] and hot[-1].tolist() == [1, 1, 1]
    assert hot[128, 0] == 1 and 0 < hot[128, 1] < 1 and hot[128, 2] == 0      # orange in the middle
    jet = colormaps.builtin("jet")
    assert jet[0].argmax() == 2 and jet[-1].argmax() == 0 and jet[128].argmax() == 1   # blue ... green ... red
    with pytest.raises(KeyError):
        colormaps.builtin("nope")


def test_strips_are_read_like_the_reference_reads_them(tmp_path, monkeypatch):
    rng = np.random.default_rng(5)
    strip = rng.integers(0, 256, (3, 40, 3), dtype=np.uint8)       # three rows: only the first one counts
    rgba = np.concatenate([strip, np.full((3, 40, 1), 255, np.uint8)], axis=2)
    PIL_Image.fromarray(strip).save(str(tmp_path / "cmap_mine.png"))
    PIL_Image.fromarray(rgba).save(str(tmp_path / "cmap_with_alpha.png"))
    PIL_Image.fromarray(strip[:, :, 0]).save(str(tmp_path / "cmap_gray_file.png"))     # not RGB: reported, skipped
    PIL_Image.fromarray(strip).save(str(tmp_path / "other.png"))
    maps = colormaps.loadcolormaps(str(tmp_path))
    assert sorted(maps) == ["mine", "with_alpha"]
    want = 1. / 255 * strip[0]
    assert maps["mine"].shape == (40, 3) and np.array_equal(maps["mine"], want)
    assert np.array_equal(maps["with_alpha"], want)
    assert np.array_equal(colormaps.get("mine", str(tmp_path)), want)
    assert np.array_equal(colormaps.get("hot", str(tmp_path)), colormaps.builtin("hot"))   # not in the folder
    monkeypatch.setenv("SPIMAGINE_COLORMAPS", str(tmp_path))
    assert np.array_equal(colormaps.get("mine"), want)
