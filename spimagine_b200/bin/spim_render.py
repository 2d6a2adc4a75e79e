#!/usr/bin/env python
"""Headless rendering from the command line: one image of a volume, or every frame of a keyframe path.

Mirror of the reference's `spim_render` (spimagine/bin/spim_render.py:35-147): same options, defaults and camera
(scale . rotation . translate, perspective(60, 1, 1, 10) or ortho(-1, 1, ...)), rendering through this package's
VolumeRenderer.  Differences, all at the edges:
  * the reference saves `out = rend.render()`, which is None (volumerender.py:508-547 returns nothing); this one
    saves `rend.output`;
  * images are written with PIL: 8-bit as round(255 * value) like imageio's float conversion, `--16bit` with the
    linear map of scipy.misc.toimage(mode="I") ((v - min) * (high - low) / (max - min) + low), 16 bits per pixel;
  * the window's upper end defaults to the data maximum (found on the device) instead of the reference's 0, which
    shows integer data as a white box; `--maxval` sets it;
  * additions: `-f raw` (with --shape / --dtype), `--iso` (iso_surface instead of max_project),
    and `--keyframes path.json --frames N`: the GUI's record loop (gui/keyframe_view.py:644-653) into
    <output directory>/output_NNN.png through keyframes.record_keyframes.

    python -m spimagine_b200.bin.spim_render -i mydata.tif -o myoutput.png -t 0 0 -4 -u 1 1 4
"""
from __future__ import absolute_import, print_function

import argparse
import os
import sys

import numpy as np


def build_parser():
    parser = argparse.ArgumentParser(formatter_class=argparse.RawTextHelpFormatter,
                                     description="""renders max projections of 3d data

    example usage:

    Tif data: \t \tspim_render  -i mydata.tif -o myoutput.png -t 0 0 -4 -u 1 1 4
    Bscope data:  \tspim_render -f bscope -i mydataFolder -o myoutput.png -t 0 0 -4 -u 1 1 4
    """)
    parser.add_argument("-f", "--format", dest="format", metavar="format",
                        help="format currently supported:\n    tif (default)\n    bscope\n    raw (needs --shape)",
                        type=str, default="tif", required=False)
    parser.add_argument("-i", "--input", dest="input", metavar="infile",
                        help="name of the input file to render", type=str, default=None, required=True)
    parser.add_argument("-o", "--output", dest="output", metavar="outfile",
                        help="name of the output file,  png extension is recommended", type=str, default="out.png")
    parser.add_argument("-p", "--pos", dest="pos", metavar="timepoint position",
                        help="timepoint to render if format=='bscope' ", type=int, default=0)
    parser.add_argument("-w", "--width", dest="width", metavar="width",
                        help="pixelwidth of the rendered output ", type=int, default=400)
    parser.add_argument("-s", "--scale", dest="scale", metavar="scale", type=float, nargs=1, default=[1.])
    parser.add_argument("-u", "--units", dest="units", metavar="units", type=float, nargs=3, default=[1., 1., 5.])
    parser.add_argument("-t", "--translate", dest="translate", type=float, nargs=3, default=[0, 0, -4],
                        metavar=("x", "y", "z"))
    parser.add_argument("-r", "--rotation", dest="rotation", type=float, nargs=4, default=[0, 1, 0, 0],
                        metavar=("w", "x", "y", "z"))
    parser.add_argument("-R", "--range", dest="range", type=float, nargs=2, default=None,
                        help="if --16bit is set, the range of the data values to consider, defaults to [min,max]",
                        metavar=("min", "max"))
    parser.add_argument("-O", "--Orthoview", help="use parallel projection (default: perspective)",
                        dest="ortho", action="store_true")
    parser.add_argument("--16bit", help="render into 16 bit png", dest="is16Bit", action="store_true")
    # additions
    parser.add_argument("--shape", type=int, nargs="+", default=None, help="raw format: (t) z y x")
    parser.add_argument("--dtype", type=str, default="uint16", help="raw format: element type")
    parser.add_argument("--iso", action="store_true", help="iso_surface at maxval / 2 instead of max_project")
    parser.add_argument("--maxval", type=float, default=None, help="upper end of the window (default: data maximum)")
    parser.add_argument("--keyframes", type=str, default=None, help="keyframe file saved by the spimagine GUI")
    parser.add_argument("--frames", type=int, default=100, help="number of frames of the keyframe path")
    parser.add_argument("--device", type=int, default=None, help="CUDA device")
    parser.add_argument("--colormap", type=str, default="grays",
                        help="colour map of the keyframe frames: grays, hot, jet, or cmap_<name>.png in --colormap-folder")
    parser.add_argument("--colormap-folder", dest="colormap_folder", type=str, default=None,
                        help="folder of cmap_<name>.png strips (default: $SPIMAGINE_COLORMAPS)")
    return parser


def open_container(args):
    """-> a frames.GenericData holding every time point of the input"""
    from spimagine_b200 import frames
    if args.format == "tif":
        return frames.TiffData(args.input)
    if args.format == "bscope":
        return frames.SpimData(args.input)
    if args.format == "raw":
        return frames.RawData(args.input, shape=args.shape, dtype=np.dtype(args.dtype))
    raise ValueError("format %s not supported (should be tif/bscope/raw)" % args.format)


def model_view(args):
    from spimagine_b200.utils.transform_matrices import mat4_rotation, mat4_translate, mat4_scale
    M = mat4_scale(*(args.scale * 3))
    M = np.dot(mat4_rotation(*args.rotation), M)
    return np.dot(mat4_translate(*args.translate), M)


def to_uint8(out):
    # in float64 like imageio's conversion of float images in [0, 1] (x * 255 + 0.499999999, truncated)
    return (np.clip(np.asarray(out, np.float64), 0, 1) * 255 + 0.499999999).astype(np.uint8)


def to_uint16(out, low, high):
    cmin, cmax = float(np.amin(out)), float(np.amax(out))
    scale = (high - low) / (cmax - cmin) if cmax > cmin else 0.
    return np.clip((np.asarray(out, np.float64) - cmin) * scale + low, 0, 65535).astype(np.uint16)


def save_image(fName, out, is16Bit=False, rng=None):
    from PIL import Image
    if not is16Bit:
        Image.fromarray(to_uint8(out)).save(fName)
        return
    if not rng:
        print("min/max: ", np.amin(out), np.amax(out))
        rng = (np.amin(out), np.amax(out))
    Image.fromarray(to_uint16(out, float(rng[0]), float(rng[1]))).save(fName)


def main(argv=None):
    parser = build_parser()
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) == 0:
        parser.print_help()
        return 0
    args = parser.parse_args(argv)
    for k, v in vars(args).items():
        print(k, v)

    from spimagine_b200 import VolumeRenderer
    from spimagine_b200.utils.transform_matrices import mat4_ortho, mat4_perspective

    container = open_container(args)
    pos = min(max(args.pos, 0), len(container) - 1)  # fromSpimFolder clamps (imgutils.py:133-135)
    data = container[pos]

    device = args.device if args.device is not None else (
        int(os.environ["LOCAL_RANK"]) if "LOCAL_RANK" in os.environ else None)
    rend = VolumeRenderer((args.width, args.width), device=device)
    try:
        rend.set_data(data)
        rend.set_units(args.units)
        # the reference never sets a window (maxVal stays 0 = raw values clamped to [0, 1], white for integer data)
        rend.set_max_val(args.maxval if args.maxval is not None else rend.data_min_max[1])

        if args.keyframes:
            from spimagine_b200 import keyframes
            keyList = keyframes.KeyFrameList.load_from_JSON(args.keyframes)
            outdir = args.output if os.path.isdir(args.output) or not os.path.splitext(args.output)[1] \
                else (os.path.dirname(args.output) or ".")
            from spimagine_b200 import colormaps
            lut = colormaps.get(args.colormap, args.colormap_folder)
            # under torchrun (one process per GPU) every rank records its share of the frames
            names = keyframes.record_keyframes(rend, keyList, args.frames, outdir, lut=lut,
                                               source=container if len(container) > 1 else None,
                                               isPerspective=not args.ortho,
                                               use_source_units=False,  # -u wins, as for a single time point
                                               rank=int(os.environ.get("RANK", "0")),
                                               world=int(os.environ.get("WORLD_SIZE", "1")))
            print("%d frames written to %s" % (len(names), outdir))
            return 0

        rend.set_modelView(model_view(args))
        if args.ortho:
            rend.set_projection(mat4_ortho(-1, 1, -1, 1, -1, 1))
        else:
            rend.set_projection(mat4_perspective(60, 1., 1, 10))
        rend.render(method="iso_surface" if args.iso else "max_project")
        save_image(args.output, rend.output, args.is16Bit, args.range)
    finally:
        rend.close()
    return 0


if __name__ == '__main__':
    sys.exit(main())
