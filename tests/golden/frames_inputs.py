"""The synthetic inputs of the frame-container golden vectors, written with this repo's own writers from fixed seeds:
used by tests/golden/make_frames_golden.py (which reads them back with the REFERENCE's containers) and by
tests/test_frames.py (which reads them back with spimagine_b200.frames and compares)."""
import os

import numpy as np


def build(root):
    """-> dict name -> constructor arguments (paths under `root`)"""
    from spimagine_b200 import frames
    rng = np.random.default_rng(11)
    spim = rng.integers(0, 65536, size=(5, 6, 7, 9)).astype(np.uint16)
    frames.createSpimFolder(os.path.join(root, "spim"), spim, stackUnits=(.162, .162, .81))
    raw4 = rng.integers(0, 65536, size=(3, 4, 5, 6)).astype(np.uint16)
    raw4.tofile(os.path.join(root, "stack_u16.raw"))
    rawf = rng.normal(size=(2, 3, 4, 5)).astype(np.float32)
    rawf.tofile(os.path.join(root, "stack_f32.raw"))
    many = []
    for t in range(4):
        fn = os.path.join(root, "t%02d.raw" % t)
        rng.integers(0, 256, size=(3, 5, 4)).astype(np.uint8).tofile(fn)
        many.append(fn)
    xw = os.path.join(root, "xwing")
    os.makedirs(os.path.join(xw, "stacks", "default"), exist_ok=True)
    for t in range(3):
        rng.integers(0, 65536, size=(4, 6, 8)).astype("<u2").tofile(os.path.join(xw, "stacks", "default", "%06d.raw" % t))
    with open(os.path.join(xw, "default.index.txt"), "w") as f:
        f.write("0\t0.000\t8, 6, 4\n1\t0.100\t8, 6, 4\n")
    with open(os.path.join(xw, "default.metadata.txt"), "w") as f:
        f.write('{"VoxelDimX": 0.26, "VoxelDimY": 0.26, "VoxelDimZ": 1.5}\n')
    arr3 = rng.normal(size=(4, 5, 6)).astype(np.float32)
    arr2 = rng.integers(0, 255, size=(7, 8)).astype(np.uint8)
    return {
        "SpimData": ("SpimData", [os.path.join(root, "spim")], {}),
        "RawData_u16": ("RawData", [os.path.join(root, "stack_u16.raw")], {"shape": (3, 4, 5, 6), "dtype": np.uint16}),
        "RawData_f32": ("RawData", [os.path.join(root, "stack_f32.raw")], {"shape": (2, 3, 4, 5), "dtype": np.float32}),
        "RawMultipleFiles": ("RawMultipleFiles", [many], {"shape": (1, 3, 5, 4), "dtype": np.uint8}),
        "XwingData": ("XwingData", [xw], {}),
        "NumpyData_3d": ("NumpyData", [arr3], {}),
        "NumpyData_2d": ("NumpyData", [arr2], {"stackUnits": [.5, .5, 2.]}),
    }
