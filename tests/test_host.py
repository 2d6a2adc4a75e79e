"""CPU-side checks: the C-ABI library exists and exports what include/spimcuda.h declares, the host-side
mirror of the reference interface behaves like the reference, and the product refuses to run without CUDA."""
import ctypes
import os
import re

import numpy as np
import pytest

import scenes
from spimagine_b200 import _lib
from spimagine_b200.utils import transform_matrices as tm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "spimcuda.h")).read()
    return sorted(set(re.findall(r"SPV_API[^;(]*?\b(spv_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    assert os.path.exists(_lib.LIB_PATH), "build with python -m spimagine_b200.build"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "libspimcuda.so does not export %s" % n
        assert n in _lib.SIGNATURES, "no ctypes signature for %s" % n
    assert set(_lib.SIGNATURES) == set(names)
    assert _lib.load().spv_version() == 100


def test_param_struct_layout_matches_header():
    assert ctypes.sizeof(_lib.MipParams) == 6 * 4 + 4 * 4 + 4 * 4
    assert ctypes.sizeof(_lib.IsoParams) == 6 * 4 + 7 * 4


def test_no_cpu_fallback():
    """Without a CUDA device the product must fail loudly, not render on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    from spimagine_b200 import VolumeRenderer
    with pytest.raises(_lib.SpvError):
        VolumeRenderer((32, 32))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "spimagine_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower().replace("no cpu", ""), "%s mentions the oracle" % f


def test_perspective_matches_glu():
    P = tm.mat4_perspective(60, 1., .1, 10)
    f = 1. / np.tan(np.pi / 6)
    assert P.dtype == np.float32
    np.testing.assert_allclose(P, [[f, 0, 0, 0], [0, f, 0, 0], [0, 0, -10.1 / 9.9, -2. / 9.9], [0, 0, -1, 0]],
                               rtol=1e-6)


def test_rotation_is_orthonormal_and_about_axis():
    R = tm.mat4_rotation(.7, 0, 1, 0)
    np.testing.assert_allclose(np.dot(R[:3, :3], R[:3, :3].T), np.eye(3), atol=1e-6)
    np.testing.assert_allclose(np.dot(R, [0, 1, 0, 1]), [0, 1, 0, 1], atol=1e-7)
    np.testing.assert_allclose(R[0, 0], np.cos(.7), atol=1e-6)
    np.testing.assert_allclose(R[0, 2], np.sin(.7), atol=1e-6)
    E = tm.mat4_rotation_euler(.1, .2, .3)
    np.testing.assert_allclose(np.linalg.det(E), 1., atol=1e-6)


def test_translate_scale_ortho_lookat():
    np.testing.assert_array_equal(tm.mat4_translate(1, 2, 3)[:, 3], [1, 2, 3, 1])
    np.testing.assert_array_equal(np.diag(tm.mat4_scale(2, 3, 4)), [2, 3, 4, 1])
    O = tm.mat4_ortho(-2, 2, -1, 1, -10, 10)
    np.testing.assert_allclose(np.dot(O, [2, 1, -10, 1]), [1, 1, 1, 1], atol=1e-6)
    L = tm.mat4_lookat([0, 0, 10], [0, 0, 0], [0, 1, 0])
    np.testing.assert_allclose(np.dot(L, [0, 0, 0, 1]), [0, 0, -10, 1], atol=1e-6)
    F = tm.mat4_stereo_perspective(45, 1., .1, 10, 0)
    np.testing.assert_allclose(F, tm.mat4_perspective(45, 1., .1, 10), rtol=1e-5, atol=1e-7)


def test_vol_g_is_deterministic():
    a = scenes.vol_g(24, np.uint16, seed=3)
    b = scenes.vol_g(24, np.uint16, seed=3)
    assert a.dtype == np.uint16 and a.max() == 60000
    np.testing.assert_array_equal(a, b)
    assert scenes.vol_g(16, np.float32).max() == 1.0


def test_fast_4x4_inverse_is_scipy_inv_bit_for_bit():
    """update_matrices inverts with the LAPACK routines scipy.linalg.inv uses (the reference's call,
    volumerender.py:312-313) minus scipy's per-call validation; every matrix family a camera produces, and the ones
    scipy routes elsewhere (lower triangular, float32, singular, non-finite), must give the identical result."""
    from scipy.linalg import inv
    from spimagine_b200.volumerender import _inv4
    rng = np.random.default_rng(1)
    for i in range(200):
        s = tm.mat4_scale(*(rng.random(3) + .2))
        rot = tm.mat4_rotation(rng.uniform(0, 6), *rng.normal(size=3))
        for c in (np.dot(tm.mat4_translate(*rng.normal(size=3)), s), np.dot(tm.mat4_identity(), s),
                  tm.mat4_translate(0, 0, -4.), tm.mat4_perspective(rng.uniform(20, 90), rng.uniform(.5, 2), .1, 10),
                  tm.mat4_ortho(-1, 1, -1, 1, -1, 1), np.dot(np.dot(tm.mat4_translate(0, 0, -4), rot), s),
                  np.tril(rng.random((4, 4)) + np.eye(4)), np.triu(rng.random((4, 4)) + np.eye(4)),
                  rng.random((4, 4)).astype(np.float32)):
            want, got = inv(c), _inv4(c)
            assert want.dtype == got.dtype and np.array_equal(want, got)
    with pytest.raises(np.linalg.LinAlgError):
        _inv4(np.zeros((4, 4)))
    with pytest.raises(ValueError):
        _inv4(np.full((4, 4), np.nan))


def test_matrix_helpers_equal_the_references_own(tmp_path):
    """tests/golden/matrices_ref.json holds the reference's transform_matrices.py evaluated on fixed arguments
    (tests/golden/make_matrix_golden.py): same values, same dtypes."""
    import json
    with open(os.path.join(ROOT, "tests", "golden", "matrices_ref.json")) as f:
        calls = json.load(f)["calls"]
    assert len(calls) >= 20
    for c in calls:
        got = np.asarray(getattr(tm, c["fn"])(*c["args"]))
        assert str(got.dtype) == c["dtype"], (c["fn"], got.dtype, c["dtype"])
        assert np.array_equal(got.astype(np.float64), np.array(c["value"])), (c["fn"], c["args"])


def _host_only_renderer():
    """A VolumeRenderer without a device: the host-side methods only (the C calls are recorded, not made)."""
    from spimagine_b200 import VolumeRenderer

    class _NoLib(object):
        def spv_set_matrices(self, ctx, invP, invM):
            return 0

    r = VolumeRenderer.__new__(VolumeRenderer)
    r._lib, r._ctx = _NoLib(), None
    r._check = lambda rc: None
    r._invM = np.zeros(16, np.float32)
    r._invP = np.zeros(16, np.float32)
    r._invM_ptr, r._invP_ptr = _lib.fp(r._invM), _lib.fp(r._invP)
    r.projection = tm.mat4_perspective()
    r.modelView = tm.mat4_identity()      # the constructor's defaults (set before a volume exists)
    return r


def test_host_side_equals_the_references_volumerenderer():
    """tests/golden/renderer_host_ref.json: the reference's own VolumeRenderer.set_units / set_projection /
    set_modelView -> update_matrices, _stack_scale_mat and _get_downsampled_data_slices run without OpenCL
    (tests/golden/make_renderer_host_golden.py).  The float32 matrices handed to the kernels are bit-identical."""
    import json
    with open(os.path.join(ROOT, "tests", "golden", "renderer_host_ref.json")) as f:
        ref = json.load(f)
    assert len(ref["matrices"]) >= 18
    for c in ref["matrices"]:
        r = _host_only_renderer()
        r.dataImg = type("Img", (), {"shape": tuple(c["shape_xyz"]), "dtype": np.uint16})()
        r.set_units(c["units"])
        P = np.array(c["projection"]).astype(c["projection_dtype"])   # the helpers return float32 for some, float64
        M = np.array(c["modelView"]).astype(c["modelView_dtype"])     # for others: the inverse is taken in that type
        r.set_projection(P)
        r.set_modelView(M)
        assert np.array_equal(np.asarray(r._stack_scale_mat(), np.float64), np.array(c["mScale"]))
        assert r._invM.dtype == np.float32 == np.dtype(c["invM_dtype"])
        assert np.array_equal(r._invM.astype(np.float64), np.array(c["invM_f32"])), c["shape_xyz"]
        assert np.array_equal(r._invP.astype(np.float64), np.array(c["invP_f32"])), c["shape_xyz"]
        # the caches (scale matrix per shape / units, inverse per projection) must not go stale
        r.set_units([3., 1., 2.])
        r.set_units(c["units"])
        r.set_projection(tm.mat4_perspective(33, 1., .2, 7))
        r.set_projection(P)
        assert np.array_equal(r._invM.astype(np.float64), np.array(c["invM_f32"]))
        assert np.array_equal(r._invP.astype(np.float64), np.array(c["invP_f32"]))
    for c in ref["downsample"]:
        r = _host_only_renderer()
        r.memMax = c["memMax"]
        r.dtype = np.dtype(c["dtype"]).type
        s = r._get_downsampled_data_slices(np.zeros(c["shape"], c["dtype"]))
        got = None if s is None else [[x.start, x.stop, x.step] for x in s]
        assert got == c["slices"], c


def test_set_data_follows_the_references_rules():
    """tests/golden/setdata_ref.json: the reference's own set_data / set_dtype / set_shape / update_data run with
    recorders in place of gputools (tests/golden/make_setdata_golden.py) on seeded arrays of eleven element types:
    the element type the renderer settles on, the image shape, the strided downsampling under a memory budget, the
    texels that reach the device and the exceptions are the same here.  (Arrays of a type the device converts travel
    as they are: their texels are what astype gives, which the GPU tests check on the device.)"""
    import hashlib
    import json
    import sys
    golden = os.path.join(ROOT, "tests", "golden")
    sys.path.insert(0, golden)
    try:
        from make_setdata_golden import make
    finally:
        sys.path.remove(golden)
    with open(os.path.join(golden, "setdata_ref.json")) as f:
        cases = json.load(f)["cases"]
    assert len(cases) == 32

    class _RecLib(object):
        def __init__(self):
            self.calls = []

        def __getattr__(self, name):
            if not name.startswith("spv_"):
                raise AttributeError(name)
            return lambda *a: (self.calls.append((name,) + a[1:]), 0)[1]

    for c in cases:
        r = _host_only_renderer()
        r._lib = _RecLib()
        r.width = r.height = 8
        r.memMax = c["memMax"]
        r.stackUnits = np.ones(3)
        r.set_dtype(np.dtype(c["first"]).type)
        data = make(c["seed"], tuple(c["shape"]), c["dtype"])
        if c.get("raises"):
            with pytest.raises(NotImplementedError):
                r.set_data(data, autoConvert=c["autoConvert"], copyData=c["copyData"])
            continue
        r.set_data(data, autoConvert=c["autoConvert"], copyData=c["copyData"])
        tag = (c["first"], c["dtype"], c["memMax"])
        assert np.dtype(r.dtype).name == c["renderer_dtype"], tag
        assert list(r.dataImg.shape) == c["image_shape_xyz"] and np.dtype(r.dataImg.dtype).name == c["image_dtype"], tag
        got = None if r.dataSlices is None else [[s.start, s.stop, s.step] for s in r.dataSlices]
        assert got == c["slices"], tag
        texels = np.ascontiguousarray(r._data).astype(r.dtype)       # what the device holds after its conversion
        assert list(texels.shape) == c["uploaded_shape"] and texels.dtype.name == c["uploaded_dtype"], tag
        assert hashlib.sha1(texels.tobytes()).hexdigest() == c["uploaded_sha1"], tag
        if np.dtype(c["dtype"]).name == c["renderer_dtype"]:           # no conversion: same reference semantics
            assert (r._data is data) == c["keeps_callers_array"], tag
        uploads = [call for call in r._lib.calls if call[0].startswith(("spv_set_volume", "spv_update_volume"))]
        assert len(uploads) == 1 and uploads[0][0] in ("spv_set_volume", "spv_set_volume_from"), tag
        nx, ny, nz = uploads[0][-3:]
        assert [nx, ny, nz] == c["image_shape_xyz"], tag


class _RecordingLib(object):
    """Stands in for libspimcuda on the CPU: every spv_* call is recorded and answers 0; the two render-to-host
    calls hand back a zeroed staging buffer, so that the Python side of a frame runs to the end."""

    def __init__(self):
        self.calls = []
        self._bufs = []

    def _give_host(self, ref, floats):
        buf = (ctypes.c_float * floats)()
        self._bufs.append(buf)
        ptr = ctypes.cast(buf, _lib._FP)
        ctypes.memmove(ctypes.addressof(ref._obj), ctypes.addressof(ptr), ctypes.sizeof(ptr))

    def __getattr__(self, name):
        if not name.startswith("spv_"):
            raise AttributeError(name)

        def call(*a):
            rec = [name]
            for x in a[1:]:
                obj = getattr(x, "_obj", None)
                if isinstance(obj, (_lib.MipParams, _lib.IsoParams)):
                    rec.append({f: (list(getattr(obj, f)) if f == "box" else getattr(obj, f)) for f, _ in obj._fields_})
                elif isinstance(x, (int, float)):
                    rec.append(x)
            self.calls.append(rec)
            if name in ("spv_render_mip_to_host", "spv_render_iso_to_host") and getattr(a[-1], "_obj", None) is not None:
                self._give_host(a[-1], 7 * self.size[0] * self.size[1])
            return 0
        return call


def test_render_dispatch_follows_the_references(monkeypatch):
    """tests/golden/dispatch_ref.json: the reference's VolumeRenderer constructed and driven through render() with
    gputools replaced by a recorder (tests/golden/make_dispatch_golden.py): constructor defaults, interpolation
    defines, the kernel an element type selects and every scalar the kernels receive.  Here the same calls on the real
    Python class over a recording stand-in for the library must put the same numbers across the C ABI."""
    import json
    from spimagine_b200 import VolumeRenderer
    with open(os.path.join(ROOT, "tests", "golden", "dispatch_ref.json")) as f:
        rows = json.load(f)["rows"]
    f32 = lambda v: float(np.float32(v))  # noqa: E731

    def make(size=(48, 32), **kw):
        lib = _RecordingLib()
        lib.size = size
        monkeypatch.setattr(_lib, "load", lambda: lib)
        return VolumeRenderer(size, **kw), lib

    n_render = 0
    for row in rows:
        if row["what"] == "constructor":
            r, lib = make(interpolation=row["interpolation"])
            d = row["defaults"]
            assert (float(r.gamma), float(r.maxVal), float(r.minVal), float(r.alphaPow)) == (
                d["gamma"], d["maxVal"], d["minVal"], d["alphaPow"])
            assert (float(r.occ_strength), int(r.occ_radius), int(r.occ_n_points)) == (
                d["occ_strength"], d["occ_radius"], d["occ_n_points"])
            assert [float(b) for b in r.boxBounds] == d["boxBounds"] and [float(u) for u in r.stackUnits] == d["stackUnits"]
            assert np.dtype(r.dtype).name == d["dtype"] and (r.width, r.height) == (d["width"], d["height"])
            assert np.array_equal(np.asarray(r.projection, np.float64), np.array(d["projection"]))
            assert np.array_equal(np.asarray(r.modelView, np.float64), np.array(d["modelView"]))
            opts = row["log"][0]["build_options"]
            linear = "SAMPLER_FILTER=CLK_FILTER_LINEAR" in opts
            assert ["spv_set_interp", 1 if linear else 0] in lib.calls
            assert "maxSteps=%d" % r.max_steps in opts
            assert r.output.shape == (d["height"], d["width"]) and r.isGPU
        elif row["what"] == "bad interpolation":
            assert row["raises"] == "KeyError"
            with pytest.raises(KeyError):
                make(interpolation="cubic")
        elif row["what"] == "render without data":
            r, lib = make()
            del lib.calls[:]
            assert (r.render() is None) == row["returns_none"]
            assert not [c for c in lib.calls if c[0].startswith("spv_render")] and not row["log"]
        else:
            r, lib = make()
            r.set_data(np.zeros((5, 6, 7), row["dtype"]))
            for name, v in row["setters"]:
                getattr(r, name)(v)
            del lib.calls[:]
            assert (r.render(**row["kwargs"]) is None) == row["returns_none"]
            renders = [c for c in lib.calls if c[0].startswith("spv_render")]
            kernels = [e["kernel"] for e in row["log"]]
            assert list(r.output.shape) == row["output_shape"]
            if not kernels:
                assert not renders
                continue
            n_render += 1
            assert len(renders) == 1
            p = renders[0][1]
            s = [v for _, v in row["log"][0]["scalars"]]
            assert s[:2] == [r.width, r.height] and p["box"] == s[2:8]
            if kernels == ["max_project_short"] or kernels == ["max_project_float"]:
                assert renders[0][0] == "spv_render_mip_to_host"
                assert (kernels[0] == "max_project_short") == (np.dtype(r.dtype).name in ("uint16", "uint8"))
                assert [p["min_val"], p["max_val"], p["gamma"], p["alpha_pow"]] == s[8:12]
                assert [p["num_parts"], p["current_part"]] == s[12:14] and p["flags"] == 0 and p["max_steps"] == 200
            else:
                assert kernels == ["iso_surface", "conv_vec_x", "conv_vec_y", "occlusion", "conv_x", "conv_y", "shading"]
                assert renders[0][0] == "spv_render_iso_to_host"
                assert [p["iso_val"], p["gamma"]] == s[8:10] and p["flags"] == 0 and p["max_steps"] == 200
                assert s[10] == float(np.dtype(r.dtype).name in ("uint16", "uint8"))        # isShortType
                # the blur radii are constants of the library (spv_api.cu render_iso_impl: 7 and 5 taps)
                assert [e["scalars"][-1][1] for e in row["log"][1:3]] == [7., 7.]
                assert [e["scalars"][-1][1] for e in row["log"][4:6]] == [5., 5.]
                occ = [v for _, v in row["log"][3]["scalars"]]
                assert [p["occ_radius"], p["occ_n_points"]] == occ[2:4]
                shade = [v for _, v in row["log"][6]["scalars"]]
                assert p["occ_strength"] == shade[2] == f32(r.occ_strength)
    assert n_render == 8


def test_matrix_caches_notice_in_place_changes():
    """update_matrices caches the scale matrix and the projection inverse; a caller that modifies stackUnits or the
    projection array in place (or assigns a list) must still get the matrices of the current values."""
    from scipy.linalg import inv
    r = _host_only_renderer()
    r.dataImg = type("Img", (), {"shape": (40, 30, 20), "dtype": np.uint16})()
    M = np.dot(tm.mat4_translate(0, 0, -4), tm.mat4_rotation(.4, 0, 1, 0))
    P = tm.mat4_perspective(60, 1., .1, 10)
    r.set_units([1., 1., 2.])
    r.set_projection(P)
    r.set_modelView(M)

    def expect():
        want_m = inv(np.dot(r.modelView, r._stack_scale_mat())).flatten().astype(np.float32)
        want_p = inv(r.projection).flatten().astype(np.float32)
        assert np.array_equal(r._invM, want_m) and np.array_equal(r._invP, want_p)

    expect()
    r.stackUnits[2] = 5.          # in place
    r.update_matrices()
    expect()
    r.stackUnits = [2., 1., 1.]   # a plain list assigned by the caller
    r.update_matrices()
    expect()
    P[0, 0] *= 1.5                # the projection array modified in place
    r.update_matrices()
    expect()
    r.projection = P.astype(np.float64)   # same values, another element type: the inverse is taken in that type
    r.update_matrices()
    expect()


def test_bench_arguments(monkeypatch):
    """bench.py's contract with the driver: defaults that finish within minutes, the workload's own volume size,
    every host thread for the CPU arms even under torchrun's OMP_NUM_THREADS=1."""
    import sys
    sys.path.insert(0, ROOT)
    try:
        import bench
    finally:
        sys.path.remove(ROOT)
    monkeypatch.setattr(sys, "argv", ["bench.py"])
    a = bench.parse()
    assert (a.gpus, a.steps, a.warmup, a.impl, a.workload, a.vol, a.img) == (1, 720, 20, "ours", "sweep", 512, 1024)
    assert a.warmup >= 3
    monkeypatch.setattr(sys, "argv", ["bench.py", "--workload", "iso"])
    assert bench.parse().vol == 1024                      # BASELINE configs[2]
    monkeypatch.setattr(sys, "argv", ["bench.py", "--workload", "iso", "--vol", "256", "--impl", "reference"])
    a = bench.parse()
    assert a.vol == 256 and a.impl == "reference"
    monkeypatch.setenv("OMP_NUM_THREADS", "1")
    n = bench.use_all_host_threads()
    assert n == len(os.sched_getaffinity(0)) >= 1
    assert bench.SAMPLES_PER_RAY == 208 and bench.MAX_STEPS == 200
    cams = bench.sweep_cameras(4)
    assert len(cams) == 4 and np.asarray(cams[0][0]).shape == (4, 4)


def test_package_exports_what_spimagine_exports():
    """spimagine/__init__.py:20-27: the data containers, TransformData and the TIFF helpers are importable from the
    package root (`import spimagine_b200 as spimagine`); importing the package needs neither a GPU nor the library."""
    import spimagine_b200 as sp
    for name in ("VolumeRenderer", "DataModel", "SpimData", "TiffData", "TiffFolderData", "NumpyData", "RawData",
                 "RawMultipleFiles", "XwingData", "GenericData", "TransformData", "read3dTiff", "write3dTiff",
                 "Quaternion", "mat4_perspective", "mat4_translate", "mat4_rotation", "pinned_empty"):
        assert hasattr(sp, name), name
    assert sp.TransformData().zoom == 1 and sp.NumpyData(np.zeros((2, 3, 4))).size() == (1, 2, 3, 4)


def test_batched_sequence_launch_schedule():
    """render_sequence over a list of views, several frames per launch (volumerender._render_sequence_batched) with the
    device calls stubbed out: the first launch takes half a batch (from 4 frames per launch up), every later one a whole
    batch; two launches are in flight before the first frame is handed out; frames come back in order with the modelView
    they were rendered with; a generator is consumed lazily (never more than two launches ahead)."""
    from spimagine_b200 import VolumeRenderer

    class _Ctx(object):
        value = 1

    def run(n, batch, half=True):
        r = _host_only_renderer()
        r._ctx = _Ctx()
        r.width, r.height = 4, 3
        r.output_depth = None
        r.first_batch_half = half

        class _Lib(object):
            def spv_sync(self, ctx):
                return 0

            def spv_set_matrices(self, ctx, invP, invM):
                return 0
        r._lib = _Lib()
        launches, pulled, events = [], [0], []

        def render_batch(Ms, to_host=True):
            launches.append([float(M[0, 3]) for M in Ms])
            events.append(("launch", len(launches) - 1, pulled[0]))
            return len(launches) - 1

        def batch_frames_of(which, copy=None):
            return [(np.full((3, 4), v, np.float32), np.full((3, 4), -v, np.float32)) for v in launches[which]]
        r.render_batch, r.batch_frames_of = render_batch, batch_frames_of

        def views():
            for i in range(n):
                pulled[0] += 1
                M = tm.mat4_identity()
                M[0, 3] = float(i)
                yield M
        got = []
        for f in VolumeRenderer._render_sequence_batched(r, views(), None, batch):
            got.append((float(f.output[0, 0]), float(f.output_alpha[0, 0]), float(f.modelView[0, 3])))
            events.append(("frame", len(got) - 1, pulled[0]))
        return [len(l) for l in launches], got, events

    for n, batch, sizes in ((20, 10, [5, 10, 5]), (23, 7, [4, 7, 7, 5]), (23, 16, [8, 15]), (5, 2, [2, 2, 1]), (3, 10, [3]),
                            (720, 10, [5] + [10] * 71 + [5]), (0, 10, [])):
        got_sizes, got, events = run(n, batch)
        assert got_sizes == sizes, (n, batch, got_sizes)
        assert got == [(float(i), -float(i), float(i)) for i in range(n)]
        if len(sizes) >= 2:  # two launches are issued before the first frame is handed out, never more than two ahead
            first_frame = next(k for k, e in enumerate(events) if e[0] == "frame")
            assert [e[0] for e in events[:first_frame]] == ["launch", "launch"]
            for kind, idx, pulled in events:
                if kind == "frame":
                    done = sum(sizes[:1 + next(j for j in range(len(sizes)) if sum(sizes[:j + 1]) > idx)])
                    assert pulled <= done + sizes[min(len(sizes) - 1, 1 + next(j for j in range(len(sizes)) if sum(sizes[:j + 1]) > idx))]
    assert run(20, 10, half=False)[0] == [10, 10]


def test_iso_sequence_issue_order():
    """render_sequence(method="iso_surface") with the device calls recorded (volumerender._render_sequence_iso): with
    output + alpha only, renders run four frames ahead of the frame handed out and read-backs two ahead.  The invariants
    the device side relies on: a frame is rendered into slot k & 1 only after the read-back of the frame that used the
    slot before it has been enqueued; a slot's pinned planes are rewritten (read-back of frame k + 2) only after frame k
    has been handed out and the consumer has come back; frames are handed out in order."""
    from spimagine_b200 import VolumeRenderer

    class _Ctx(object):
        value = 1

    def run(n, planes):
        r = _host_only_renderer()
        r._ctx = _Ctx()
        r.width = r.height = 4
        r.maxVal, r.gamma, r.max_steps = 100., 1., 200
        r.occ_strength, r.occ_radius, r.occ_n_points = .1, 21, 30
        r.boxBounds = [-1, 1, -1, 1, -1, 1]
        r.stackUnits = np.ones(3)
        r.dataImg = type("D", (), {"shape": (8, 8, 8)})()
        log, state = [], {"slot": 0, "frame": -1}

        class _Lib(object):
            def spv_set_matrices(self, ctx, invP, invM):
                return 0

            def spv_select_slot(self, ctx, slot):
                state["slot"] = slot
                return 0

            def spv_render_iso(self, ctx, p):
                state["frame"] += 1
                assert state["slot"] == state["frame"] & 1
                log.append(("render", state["frame"]))
                return 0

            def spv_read_pinned_async(self, ctx, pl):
                assert pl == planes
                k = sum(1 for e in log if e[0] == "copy")
                assert state["slot"] == k & 1
                log.append(("copy", k))
                return 0

            def spv_set_tuning(self, ctx, knob, value):
                return 0

            def spv_sync(self, ctx):
                return 0
        r._lib = _Lib()
        r._adopt_slot = lambda slot, pl, clear=False: log.append(("adopt", slot))
        handed = 0
        for _ in VolumeRenderer._render_sequence_iso(r, (tm.mat4_identity() for _ in range(n)), planes, planes == 2):
            assert log[-1] == ("adopt", handed & 1)
            log.append(("frame", handed))
            handed += 1
        assert handed == n
        pos = {e: i for i, e in enumerate(log)}
        for k in range(n):
            assert pos[("render", k)] < pos[("copy", k)] < pos[("frame", k)]
            if k >= 2:
                assert pos[("copy", k - 2)] < pos[("render", k)]      # the slot's previous frame is on its way out
                assert pos[("frame", k - 2)] < pos[("copy", k)]       # the pinned planes it sat in have been seen
        ahead = max(sum(1 for e in log[:pos[("frame", k)]] if e[0] == "render") - k for k in range(n)) if n else 0
        return ahead

    assert run(12, 2) == 4      # frame k is handed out with frames k .. k + 3 rendered
    assert run(12, 7) == 3      # whole frames stay three ahead (bound by the host link either way)
    assert run(1, 2) == 1 and run(3, 2) == 3 and run(0, 2) == 0
