#!/usr/bin/env python
"""Generate tests/golden/*.npz from the REFERENCE's own kernel text.

Run in the dev container (needs /root/reference): oracle/build.py compiles the reference's
spimagine/volumerender/kernels/*.cl for the host (oracle/_ref/libspim_ref.so) and every case of
tests/golden_cases.py is rendered with it.  The .npz files are committed; the tests compare the C restatement
(CPU) and libspimcuda's exact sampler (GPU) against them bit for bit.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import build, oracle  # noqa: E402
import golden_cases  # noqa: E402


def main():
    build.build_oracle()
    if build.build_ref() is None or not oracle.available("reference"):
        raise SystemExit("the reference tree is not available: golden vectors can only be made where it is")
    for name, case in sorted(golden_cases.CASES.items()):
        rend = oracle.OracleRenderer(golden_cases.SIZE, interpolation=case.get("interpolation", "linear"),
                                     kind="reference")
        res = golden_cases.run_case(rend, name)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **res)
        print(name, {k: (v.shape, float(np.nanmax(np.where(np.isfinite(v), v, 0)))) for k, v in res.items()})
    for name, case in sorted(golden_cases.CONFIG_CASES.items()):
        rend = oracle.OracleRenderer(case["size"], kind="reference")
        res = golden_cases.run_config_case(rend, name)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **res)
        print(name, {k: (v.shape, float(v.max())) for k, v in res.items()})


if __name__ == "__main__":
    main()
