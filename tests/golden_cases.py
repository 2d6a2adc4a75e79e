"""The golden-vector cases: name -> how to set a renderer up.  Shared by tests/golden/make_golden.py (which runs
them through oracle/_ref, the reference's own kernel text built for the host) and by the tests (which run them
through the C restatement on CPU and through libspimcuda on the GPU)."""
import numpy as np

import scenes
from spimagine_b200.utils.transform_matrices import mat4_perspective, mat4_rotation, mat4_translate

SIZE = (44, 36)  # (width, height): not square, not a multiple of the 8x4 tile in y... 36 = 9*4, 44 = 5.5*8


def _tilted():
    return scenes.tilted_camera()


def _iso_cam():
    return np.dot(mat4_translate(0, 0, -5), mat4_rotation(.3, 1, 1, 0)), mat4_perspective()


CASES = {
    # max_project_float, tests/test_rendering/test_simple_rendering.py:17-36
    "mip_f32_two_blobs": dict(data=lambda: scenes.two_blobs(32), cam=_tilted, render=dict(maxVal=255.)),
    # max_project_short on the same data
    "mip_u16_two_blobs": dict(data=lambda: scenes.two_blobs(32).astype(np.uint16), cam=_tilted,
                              render=dict(maxVal=255.)),
    "mip_u8_blob": dict(data=lambda: scenes.gaussian(24, 250.).astype(np.uint8), cam=_tilted,
                        render=dict(maxVal=200., minVal=10.)),
    # tests/test_volumerender/test_volumerender.py:173-188 (opacity)
    "mip_f32_alpha": dict(data=lambda: scenes.gaussian(32, 200.), cam=_tilted, alpha_pow=.6,
                          render=dict(maxVal=200.)),
    "mip_u16_alpha": dict(data=lambda: scenes.gaussian(32, 200.).astype(np.uint16), cam=_tilted, alpha_pow=1.5,
                          render=dict(maxVal=150.)),
    # tests/test_volumerender/test_volumerender.py:48-57 (nearest), ragged shape
    "mip_u16_nearest_ragged": dict(data=lambda: scenes.random_vol((20, 24, 28), np.uint16, 1), cam=_tilted,
                                   interpolation="nearest", render=dict(maxVal=65535.)),
    "mip_f32_random_ragged": dict(data=lambda: scenes.random_vol((19, 33, 27), np.float32, 2), cam=_tilted,
                                  units=(1., 1., 2.5), render=dict(maxVal=1.)),
    "mip_f32_box": dict(data=lambda: scenes.linspace_vol(24), cam=_tilted, box=[-.5, .7, -1, .4, -.3, .9],
                        render=dict(maxVal=1.)),
    "mip_f32_gamma": dict(data=lambda: scenes.two_blobs(32), cam=_tilted, render=dict(maxVal=255., gamma=.6)),
    "mip_f32_parts": dict(data=lambda: scenes.two_blobs(32), cam=_tilted, render=dict(maxVal=255.), parts=3),
    # tests/test_rendering/test_simple_rendering.py:85-95: constant volume -> 123/200
    "mip_f32_const": dict(data=lambda: np.full((8, 8, 8), 123., np.float32), cam=_tilted, render=dict(maxVal=200.)),
    # tests/test_rendering/test_simple_rendering.py:55-68 (iso surface of a uint16 sphere)
    "iso_u16_sphere": dict(data=lambda: scenes.iso_sphere(32), cam=_iso_cam, render=dict(maxVal=20.),
                           method="iso_surface"),
    "iso_f32_blobs": dict(data=lambda: scenes.two_blobs(32), cam=_tilted, render=dict(maxVal=120., gamma=1.3),
                          method="iso_surface", occ=(.4, 9, 12)),
}


# BASELINE.json configs at their own sizes; the golden file keeps every `rows`-th image row
CONFIG_CASES = {
    # configs[0]: max_project of a 128^3 float32 synthetic Gaussian-blob volume to 512x512
    "c1_mip_f32_volg128_512": dict(size=(512, 512), rows=8, data=lambda: scenes.vol_g(128, np.float32, seed=0),
                                   cam=lambda: scenes.gui_camera(2 * np.pi * 40 / 360, 4.0), render=dict(maxVal=1.)),
}
CASES_ALL = dict(CASES)
CASES_ALL.update(CONFIG_CASES)


def run_config_case(rend, name):
    """-> dict(output, alpha), every `rows`-th row of the frame."""
    c = CONFIG_CASES[name]
    rend.set_data(c["data"]())
    M, P = c["cam"]()
    rend.set_modelView(M)
    rend.set_projection(P)
    rend.render(**c["render"])
    r = c["rows"]
    return {"output": np.array(rend.output[::r]), "alpha": np.array(rend.output_alpha[::r])}


def run_case(rend, name):
    """Drive `rend` (OracleRenderer or VolumeRenderer: same calls) through case `name`.
    -> dict of result arrays."""
    c = CASES[name]
    rend.set_data(c["data"]())
    if "units" in c:
        rend.set_units(c["units"])
    M, P = c["cam"]()
    rend.set_modelView(M)
    rend.set_projection(P)
    rend.set_alpha_pow(c.get("alpha_pow", 0.))
    rend.set_box_boundaries(c.get("box", [-1, 1, -1, 1, -1, 1]))
    if "occ" in c:
        rend.set_occ_strength(c["occ"][0])
        rend.set_occ_radius(c["occ"][1])
        rend.set_occ_n_points(c["occ"][2])
    method = c.get("method", "max_project")
    out = {}
    if "parts" in c:
        n = c["parts"]
        for part in range(n):
            rend.render(method=method, numParts=n, currentPart=part, **c["render"])
            out["output_part%d" % part] = np.array(rend.output)
            out["alpha_part%d" % part] = np.array(rend.output_alpha)
        return out
    rend.render(method=method, **c["render"])
    out["output"] = np.array(rend.output)
    out["alpha"] = np.array(rend.output_alpha)
    if method == "iso_surface":
        out["depth"] = np.array(rend.output_depth)
        out["normals"] = np.array(rend.output_normals)
        out["occlusion"] = np.array(rend.output_occlusion)
    return out
