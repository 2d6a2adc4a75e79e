"""Generates tests/golden/matrices_ref.json: the camera-matrix helpers of the REFERENCE
(/root/reference/spimagine/utils/transform_matrices.py, loaded by path with `spimagine` entered as a bare namespace) evaluated on a fixed set
of arguments, plus VolumeRenderer._stack_scale_mat's expression (volumerender.py:299-308).  The tests compare
spimagine_b200.utils.transform_matrices element by element.

    python tests/golden/make_matrix_golden.py
"""
import importlib.util
import json
import os

import numpy as np

REF = "/root/reference/spimagine/utils/transform_matrices.py"
HERE = os.path.dirname(os.path.abspath(__file__))

CALLS = [
    ("mat4_identity", []),
    ("mat4_scale", [2., 3., .5]), ("mat4_scale", []),
    ("mat4_translate", [1., -2., 3.5]), ("mat4_translate", []),
    ("mat4_rotation", [.7, 0, 1, 0]), ("mat4_rotation", [2.1, 1., 2., -3.]), ("mat4_rotation", []),
    ("mat4_rotation", [1e-3, 0, 1, 0]), ("mat4_rotation", [6.2, .3, .3, .9]),
    ("mat4_rotation_euler", [.1, .2, .3]), ("mat4_rotation_euler", []),
    ("mat4_perspective", []), ("mat4_perspective", [60, 1., .1, 10]), ("mat4_perspective", [60, 1., 1, 10]),
    ("mat4_perspective", [35, 1.5, .01, 100]),
    ("mat4_frustrum", [-1., 1., -.5, .5, .1, 10.]),
    ("mat4_stereo_perspective", [45, 1., .1, 10, 0]), ("mat4_stereo_perspective", [60, 1.2, .1, 10, .05]),
    ("mat4_ortho", []), ("mat4_ortho", [-2., 2., -2., 2., -1.5, 1.5]), ("mat4_ortho", [-1, 1, -1, 1, -1, 1]),
    ("mat4_lookat", [[0, 0, 10], [0, 0, 0], [0, 1, 0]]), ("mat4_lookat", [[1, 2, 3], [.1, -.2, 0], [0, 0, 1]]),
]


def main():
    # the module imports spimagine.utils.quaternion: enter `spimagine` as a bare namespace so that its __init__
    # (pyopencl, Qt) does not run
    import sys
    import types
    for name in ("spimagine", "spimagine.utils"):
        m = types.ModuleType(name)
        m.__path__ = [os.path.join("/root/reference", *name.split("."))]
        sys.modules[name] = m
    spec = importlib.util.spec_from_file_location("spimagine.utils.transform_matrices", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    out = []
    for name, args in CALLS:
        if not hasattr(ref, name):
            continue
        m = np.asarray(getattr(ref, name)(*args))
        out.append({"fn": name, "args": args, "dtype": str(m.dtype), "value": m.astype(np.float64).tolist()})
    with open(os.path.join(HERE, "matrices_ref.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_matrix_golden.py", "calls": out}, f, indent=1)
    print("wrote %d matrices" % len(out), sorted(set(c["fn"] for c in out)))


if __name__ == "__main__":
    main()
