#!/usr/bin/env python
"""bench.py -- the headline benchmark: max-intensity projection of a 512^3 uint16 volume to 1024x1024 over a
360-degree modelView sweep (BASELINE.json configs[1]), frames/s and Gsamples/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A step is one frame of the sweep (one launch of the max-projection kernel).  With N > 1 the frames of the sweep
are sharded over the ranks (frame f -> rank f mod N, the 3D+t playback decomposition: no data-path collective),
so per-GPU work is fixed: weak scaling; `value` is all frames rendered by all ranks / max-over-ranks device time.

  value      device-resident: volume in HBM, camera matrices passed as kernel arguments, nothing read back
  e2e        through the public API a spimagine caller uses: rend.set_modelView(M); rend.render(); rend.output --
             host matrix inversion, launch, and the device->host read of output + alpha into pinned memory
  roofline   HBM: algorithmic bytes per frame (every voxel once + both output planes) / average launch time,
             against the measured copy bandwidth of MEASURED_PEAKS.json; plus the texture-sample view
  cpu_baseline / --impl reference   the reference's own kernel text compiled for the host (oracle/_ref) when
             that build is present, else the C restatement, on all host cores
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

VOL_N = 512
IMG = 1024
MAX_STEPS = 200
SAMPLES_PER_RAY = (MAX_STEPS // 16 + 1) * 16  # 208, volume_kernel.cl:100-122
PEAK_VALUE = 60000.
SWEEP = 360
METRIC = "MIP frames/s, 512^3 uint16 -> 1024^2, 360-degree modelView sweep"
# ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch of the max-projection kernel on this
# workload: profiles/r01_mip_ncu_summary.json "prof_mip_zpair_session3" (407.9 MB read + 9.3 MB written; the z-paired
# volume is 512 MiB, of which a frame touches the part its rays cross)
NCU_TRAFFIC_BYTES = 417157120


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=720)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--vol", type=int, default=None,
                    help="edge of the synthetic volume (default: %d; 1024 for the iso workload = configs[2])" % VOL_N)
    ap.add_argument("--img", type=int, default=IMG)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c4", action="store_true",
                    help="sweep workload: leave out the `c4` record (BASELINE configs[3], the 2048^3 sort-last composite; "
                         "at --gpus 1 its single-GPU baselines)")
    ap.add_argument("--c4-vol", type=int, default=2048)
    ap.add_argument("--c4-img", type=int, default=2048)
    ap.add_argument("--c4-steps", type=int, default=24)
    ap.add_argument("--no-mip-overlap", action="store_true",
                    help="sweep workload: every frame's kernel on the render stream, one after the other (as round 1)")
    ap.add_argument("--batch", type=int, default=10,
                    help="sweep workload: frames per launch of the device-resident loop and of render_sequence "
                         "(spv_render_mip_batch: the frames of a launch share the volume in L2; at most 16; 1 = one launch "
                         "per frame through spv_render_mip, as before)")
    ap.add_argument("--no-axis", action="store_true",
                    help="sweep workload: mip_fast_kernel on the z-paired array for every frame (tuning knob 16 = 0, as "
                         "round 1) instead of the view-aligned layered copies of spv_mip_axis.cu; implies --batch 1")
    ap.add_argument("--no-occ-table", action="store_true",
                    help="iso workload: hash every ambient-occlusion tap in every frame (tuning knob 17 = 0) instead of "
                         "reading its pixel offsets from the per-image table")
    ap.add_argument("--no-iso-overlap", action="store_true",
                    help="iso workload, one GPU: every frame's screen-space passes on the render stream (as round 1)")
    ap.add_argument("--dtype", default="u16", choices=["u16", "f32"],
                    help="sweep workload: element type of the volume (f32 with --vol 128 --img 512 is BASELINE configs[0])")
    ap.add_argument("--alpha-pow", type=float, default=0.,
                    help="sweep workload with front-to-back attenuation (set_alpha_pow; volume_kernel.cl:300-318): "
                         "mip_alpha_kernel instead of mip_fast_kernel; no c4 record, no texture-sample roofline")
    ap.add_argument("--mip-path", default=None, choices=[None, "tmu", "smem"],
                    help="max-projection kernel family (default: the library's choice)")
    ap.add_argument("--workload", default="sweep", choices=["sweep", "slab", "timelapse", "iso", "blur", "keyframes"],
                    help="sweep: BASELINE configs[1], frames sharded over the GPUs (default). slab: configs[3], one "
                         "--vol^3 uint16 volume split into z-slabs over the GPUs, sort-last max composite. timelapse: "
                         "configs[4], --frames time points of --tl-shape uint16, time point t on GPU t mod N. iso: "
                         "configs[2], --vol^3 uint16 iso_surface with ambient occlusion and shading; N > 1: sort-last. "
                         "blur: SURVEY 8f-4, BlurProcessor(sigma=4) on a --vol^3 uint16 volume into the renderer's resident "
                         "array (one GPU). keyframes: SURVEY 8f-3, the GUI's record loop over a keyframe path (max projection "
                         "and iso-surface stretches) on a --vol^3 uint16 volume, pipelined against the reference-shaped "
                         "synchronous loop, plus the TIFF time-point reader (one GPU)")
    ap.add_argument("--frames", type=int, default=100, help="timelapse workload: time points in the whole series")
    ap.add_argument("--tl-shape", default="512,1024,1024", help="timelapse workload: (Nz,Ny,Nx) of one time point")
    ap.add_argument("--skip", action="store_true", help="enable empty-space skipping on the min/max brick grid")
    ap.add_argument("--slabs-per-rank", type=int, default=2,
                    help="slab workload: slabs per GPU, dealt in serpentine order (front/back slabs pair up: balanced "
                         "sample counts for any view), at most 4, all marched by one kernel launch")
    ap.add_argument("--bricks", type=int, default=0,
                    help="slab workload, --gpus 1 only: the single-GPU-brick baseline -- the volume as this many z-slabs, "
                         "each rendered by its own launch one after the other, max-merged, then windowed")
    ap.add_argument("--composite", default="peer", choices=["peer", "nccl"],
                    help="slab workload: peer = partials stored straight into the band owners' memory over NVLink "
                         "(spv_render_mip_composite); nccl = all-reduce(MAX) of the raw plane")
    args = ap.parse_args()
    if args.vol is None:
        args.vol = 1024 if args.workload == "iso" else VOL_N
    return args


def sweep_cameras(n=SWEEP):
    import scenes
    return [scenes.gui_camera(2 * math.pi * f / n, 4.0) for f in range(n)]


def sweep_thetas(steps, rank, world):
    """View angles (radians) of the K = steps timed frames of rank r of N: 2 pi (j / K + r / (K N)), j < K.  Every
    rank's frames cover the 360-degree sweep with the same spacing, whatever K and N are, so `value` measures the sweep
    the metric names and the work per GPU does not change with N (the frame time depends on the angle:
    profiles/r01_exp_mip_time_vs_angle.txt)."""
    return [2 * math.pi * (j / float(steps) + rank / float(steps * world)) for j in range(steps)]


def sweep_metric(args):
    """`metric` of the sweep workload -- the same string in both arms."""
    f32 = getattr(args, "dtype", "u16") == "f32"
    m = METRIC
    if (args.vol, args.img, f32) != (VOL_N, IMG, False):
        m = "MIP frames/s, %d^3 %s -> %d^2, 360-degree modelView sweep" % (args.vol, "float32" if f32 else "uint16", args.img)
    if getattr(args, "alpha_pow", 0.):
        m += ", alpha_pow = %g" % args.alpha_pow
    return m


def sweep_config(args):
    """`config` of the headline workload -- the same dict in both arms (--impl ours / reference)."""
    return {
        "workload": "Vol-G(%d, %s, seed 0) max_project -> %dx%d, max_steps=200 (208 samples per hit ray); step j "
                    "of rank r of N renders the sweep angle 360 (j / K + r / (K N)) degrees, K = steps: every rank's "
                    "frames cover the 360-degree sweep evenly" % (
                        args.vol, "float32" if getattr(args, "dtype", "u16") == "f32" else "uint16", args.img, args.img),
        "camera": "perspective(60,1,.1,10), translate(0,0,-4) . rotation(theta + 1e-3, y)",
        "window": "minVal 0, maxVal %s, gamma 1, alpha_pow 0, box +-1, units 1" % (
            "1" if getattr(args, "dtype", "u16") == "f32" else "60000"),
    }


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled while the timed regions run."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def bind_to_gpu_numa_node(local_rank):
    """One process per GPU on a two-socket host: run on the cores next to this rank's GPU, so that the page-locked
    staging the results are copied into (first touched by this process) lies in that socket's memory.  Multi-rank runs
    only -- a single rank keeps every core for its cpu_baseline leg.  Plumbing; failures are ignored."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        n = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n)
        cpus = [64 * i + b for i, w in enumerate(mask) for b in range(64) if (w >> b) & 1]
        allowed = set(os.sched_getaffinity(0))
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return 0


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arms are meant to use every host core (the other
    ranks of the reference arm exit at once, and the GPU arm's host work is negligible).  libgomp is process-global:
    setting the count through it reaches the oracle libraries, which link the same runtime."""
    import ctypes
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(int(n))
    except OSError:
        pass
    return n


def cpu_reference(vol, cams, img, steps, warmup, budget_s, alpha_pow=0., peak=PEAK_VALUE):
    """Time the reference kernels on the host cores.  -> dict(fps, gsamples, kind, cores, sample, ms_per_step)"""
    from oracle import oracle
    use_all_host_threads()
    kind = "reference" if oracle.available("reference") else "port"
    # timing: the reference's text with the host equivalents of its OpenCL build options (-cl-fast-relaxed-math,
    # -cl-mad-enable: -O3 -ffast-math -mavx2 -mfma) where that build exists and the CPU can run it
    build = "reference_fast" if kind == "reference" and oracle.available("reference_fast") and oracle.cpu_has_avx2() else kind
    r = oracle.OracleRenderer((img, img), kind=build, max_steps=MAX_STEPS)
    r.set_data(vol)
    r.set_projection(cams[0][1])
    r.set_max_val(peak)
    r.set_alpha_pow(alpha_pow)
    lib = r.lib
    cores = int(lib.so_num_threads())
    # one full frame to size the sample
    r.set_modelView(cams[0][0])
    t0 = time.perf_counter()
    r.render()
    t_full = time.perf_counter() - t0
    rowstep = int(min(64, max(1, math.ceil(steps * t_full / budget_s))))
    lib.so_set_row_sampling(0, rowstep)
    for i in range(warmup):
        r.set_modelView(cams[i % len(cams)][0])
        r.render()
    hits = 0
    t0 = time.perf_counter()
    for i in range(steps):
        lib.so_set_row_sampling(i % rowstep, rowstep)
        r.set_modelView(cams[i % len(cams)][0])
        r.render()
    dt = time.perf_counter() - t0
    lib.so_set_row_sampling(0, 1)
    for i in range(min(steps, len(cams))):
        r.set_modelView(cams[i % len(cams)][0])
        hits += r.count_hit_rays()
    mean_hits = hits / float(min(steps, len(cams)))
    frames = steps / float(rowstep)  # each step rendered 1/rowstep of the rows of its frame
    fps = frames / dt
    sample = "%d steps, each every %d-th row of one %dx%d frame of the sweep (1/%d of its rays); %s on %d threads" % (
        steps, rowstep, img, img, rowstep,
        ("reference kernel text built for the host with the reference's fast-math options (oracle/_ref, -O3 -ffast-math "
         "-mavx2 -mfma)" if build == "reference_fast" else "reference kernel text built for the host (oracle/_ref, -O2)")
        if kind == "reference" else "C restatement of the reference kernels (oracle/)", cores)
    return {"fps": fps, "gsamples": fps * mean_hits * SAMPLES_PER_RAY / 1e9, "kind": kind, "cores": cores,
            "sample": sample, "ms_per_step": 1e3 * dt / steps, "seconds": dt}


def run_reference(args, rank):
    """--impl reference: the reference's own kernel text compiled for the host (oracle/_ref; the C restatement where
    that build is missing) on all host cores, same config / steps / warm-up as the GPU arm, frames of rank 0."""
    if rank != 0:
        return
    import scenes
    f32 = args.dtype == "f32"
    vol = scenes.vol_g(args.vol, np.float32 if f32 else np.uint16, seed=0)
    cams = [scenes.gui_camera(th, 4.0) for th in sweep_thetas(args.steps, 0, max(1, args.gpus))]
    res = cpu_reference(vol, cams, args.img, args.steps, args.warmup, 100.0, alpha_pow=args.alpha_pow,
                        peak=1. if f32 else PEAK_VALUE)
    line = {
        "impl": "reference", "metric": sweep_metric(args), "value": res["fps"], "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32" if f32 else "u16->f32",
        "data": "synthetic",
        "config": dict(sweep_config(args), window=sweep_config(args)["window"].replace(
            "alpha_pow 0", "alpha_pow %g" % args.alpha_pow)) if args.alpha_pow else sweep_config(args),
        "gsamples_per_s": res["gsamples"],
        "cpu_baseline": {"value": res["fps"], "unit": "frames/s", "cores": res["cores"], "kind": res["kind"],
                         "sample": res["sample"]},
        "e2e": {"value": res["fps"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def vol_g_slab_device(N, z_lo, z_hi, seed, device):
    """Slices [z_lo, z_hi) of Vol-G(N, uint16, seed) generated on the GPU with torch (plumbing: bench input only).
    Same recipe as tests/scenes.py::vol_g; every rank evaluates the same expressions for a given global slice,
    so slabs and halos agree bit for bit across ranks.  Peak scaled to 60000 by the analytic maximum bound."""
    import torch
    rng = np.random.default_rng(seed)
    c = rng.uniform(-.6, .6, (8, 3))
    sg = rng.uniform(.08, .25, 8)
    a = rng.uniform(.3, 1., 8)
    lin = torch.linspace(-1, 1, N, device=device, dtype=torch.float32)
    out = torch.empty((z_hi - z_lo, N, N), dtype=torch.uint16, device=device)
    # normalisation: maximum of the blob sum on a coarse 64^3 grid (identical on every rank)
    g = torch.linspace(-1, 1, 64, device=device, dtype=torch.float32)
    coarse = torch.zeros((64, 64, 64), device=device)
    for i in range(8):
        k = float(1. / (2 * sg[i] ** 2))
        coarse += float(a[i]) * (torch.exp(-k * (g - float(c[i, 0])) ** 2)[:, None, None] *
                                 torch.exp(-k * (g - float(c[i, 1])) ** 2)[None, :, None] *
                                 torch.exp(-k * (g - float(c[i, 2])) ** 2)[None, None, :])
    scale = 58000. / float(coarse.max())
    step = 16
    for zb in range(z_lo, z_hi, step):
        ze = min(zb + step, z_hi)
        v = torch.zeros((ze - zb, N, N), device=device)
        for i in range(8):
            k = float(1. / (2 * sg[i] ** 2))
            gz = float(a[i]) * torch.exp(-k * (lin[zb:ze] - float(c[i, 0])) ** 2)
            gy = torch.exp(-k * (lin - float(c[i, 1])) ** 2)
            gx = torch.exp(-k * (lin - float(c[i, 2])) ** 2)
            v += gz[:, None, None] * (gy[:, None] * gx[None, :])[None]
        idx = torch.arange(zb * N * N, ze * N * N, device=device, dtype=torch.int64).reshape(ze - zb, N, N)
        h = (idx * 2654435761 + seed * 40503) & 0xffffffff
        h = ((h ^ (h >> 15)) * 2246822519) & 0xffffffff
        h = (h ^ (h >> 13)) & 0xffffff
        v = v * scale + (0.01 * 58000. / float(1 << 24)) * h.to(torch.float32)
        out[zb - z_lo:ze - z_lo] = v.clamp_(0, 65535).round_().to(torch.int32).to(torch.uint16)
    return out


def vol_g_device(shape, seed, t, device):
    """One Vol-G time point (blob centres drifting 0.01 t) of shape (nz, ny, nx), uint16, generated on
    the GPU with torch (plumbing: bench input only)."""
    import torch
    rng = np.random.default_rng(seed)
    c = rng.uniform(-.6, .6, (8, 3)) + 0.01 * t
    sg = rng.uniform(.08, .25, 8)
    a = rng.uniform(.3, 1., 8)
    nz, ny, nx = shape
    lz = torch.linspace(-1, 1, nz, device=device, dtype=torch.float32)
    ly = torch.linspace(-1, 1, ny, device=device, dtype=torch.float32)
    lx = torch.linspace(-1, 1, nx, device=device, dtype=torch.float32)
    out = torch.empty(shape, dtype=torch.uint16, device=device)
    step = 16
    for zb in range(0, nz, step):
        ze = min(zb + step, nz)
        v = torch.zeros((ze - zb, ny, nx), device=device)
        for i in range(8):
            k = float(1. / (2 * sg[i] ** 2))
            gz = float(a[i]) * torch.exp(-k * (lz[zb:ze] - float(c[i, 0])) ** 2)
            gy = torch.exp(-k * (ly - float(c[i, 1])) ** 2)
            gx = torch.exp(-k * (lx - float(c[i, 2])) ** 2)
            v += gz[:, None, None] * (gy[:, None] * gx[None, :])[None]
        idx = torch.arange(zb * ny * nx, ze * ny * nx, device=device, dtype=torch.int64).reshape(ze - zb, ny, nx)
        h = (idx * 2654435761 + (seed + t) * 40503) & 0xffffffff
        h = ((h ^ (h >> 15)) * 2246822519) & 0xffffffff
        h = (h ^ (h >> 13)) & 0xffffff
        v = v * 30000. + (0.01 * 58000. / float(1 << 24)) * h.to(torch.float32)
        out[zb:ze] = v.clamp_(0, 65535).round_().to(torch.int32).to(torch.uint16)
    return out


def run_timelapse(args, rank, local_rank, world):
    """BASELINE configs[4]: 3D+t playback, time point t on GPU t mod N, no data-path collective (weak scaling in the
    number of time points a GPU holds; `value` counts time points played by all GPUs per second).
      value : the owned time points are RESIDENT in HBM (one texture array each); a step renders the next one
      e2e   : STREAMED through the public API -- per step one time point is uploaded from page-locked host memory
              (TimelapsePlayer.play -> update_data(pinned=True)), rendered, and output + alpha read back"""
    import torch
    import torch.distributed as dist
    import scenes
    from spimagine_b200 import pinned_empty
    from spimagine_b200.multigpu import TimelapsePlayer, frames_for_rank

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    shape = tuple(int(x) for x in args.tl_shape.split(","))
    W = args.img
    mine = frames_for_rank(args.frames, rank, world)
    player = TimelapsePlayer((W, W), rank=rank, world=world, device=local_rank, max_steps=MAX_STEPS, pinned_outputs=True)
    ring = [pinned_empty(shape, np.uint16) for _ in range(min(4, len(mine)))]  # streamed source: distinct time points
    t0 = time.perf_counter()
    for j, t in enumerate(mine):
        d = vol_g_device(shape, 100, t, dev)
        if j < len(ring):
            ring[j][...] = d.cpu().numpy()
        player.preload({t: (d.data_ptr(), shape, np.uint16)}, frames=[t], device_ptrs=True)
        del d
    torch.cuda.synchronize()
    t_prep = time.perf_counter() - t0
    P = scenes.gui_camera(0, 4.0)[1]
    cams = [scenes.gui_camera(2 * math.pi * f / 720, 4.0)[0] for f in range(720)]  # slow spin
    for r in player.resident.values():
        r.set_projection(P)
        r.set_max_val(PEAK_VALUE)
        r.set_units([1., 1., 2.])  # 512 slices of twice the pixel pitch: a cubic field of view

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def resident_step(i):
        r = player.resident[mine[i % len(mine)]]
        r.set_modelView(cams[i % 720])
        r.render_device_only()

    for i in range(args.warmup):
        resident_step(i)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    for i in range(args.steps):
        resident_step(i)
    for r in player.resident.values():
        r.sync()
    t_res = time.perf_counter() - t0  # contexts run on their own streams: host clock around a full drain
    barrier()

    # streamed, end to end
    src = {i: ring[i % len(ring)] for i in range(args.steps + 3)}
    settings = dict(projection=P, max_val=PEAK_VALUE, units=[1., 1., 2.])
    for _ in player.play(src, {i: cams[i % 720] for i in src}, frames=range(3), pinned=True, **settings):
        pass
    barrier()
    chk = 0.
    t0 = time.perf_counter()
    for t, r in player.play(src, {i: cams[i % 720] for i in src}, frames=range(3, args.steps + 3), pinned=True, **settings):
        chk += float(r.output[W // 2, W // 2])
    torch.cuda.synchronize()
    t_str = time.perf_counter() - t0
    barrier()
    if world > 1:
        tt = torch.tensor([t_res, t_str], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_res, t_str = float(tt[0]), float(tt[1])
    if rank == 0:
        nbytes = int(np.prod(shape)) * 2
        print(json.dumps({
            "metric": "3D+t max_project playback, time points/s, %d x %s uint16 -> %d^2" % (args.frames, "x".join(map(str, shape)), W),
            "value": args.steps * world / t_res, "unit": "time points/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_res / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u16->f32", "data": "synthetic",
            "config": {"workload": "Vol-G time points (seed 100, centres drifting), time point t on GPU t mod %d, %d resident "
                                   "per GPU (%d MiB each as z-paired texels), fixed projection + slow spin" % (
                                       world, len(mine), 2 * nbytes >> 20)},
            "e2e": {"value": args.steps * world / t_str, "unit": "time points/s", "h2d_bytes_per_step": nbytes + 128,
                    "d2h_bytes_per_step": 2 * W * W * 4, "checksum": chk,
                    "note": "streamed: TimelapsePlayer.play(pinned=True) uploads every time point from page-locked host "
                            "memory (%.1f GB/s per GPU), renders it and reads output + alpha back" % (nbytes / (t_str / args.steps) / 1e9)},
            "gpu_launches": args.steps, "prepare_s": t_prep}))
    player.close()
    if world > 1:
        dist.destroy_process_group()


def run_iso(args, rank, local_rank, world):
    """BASELINE configs[2]: --vol^3 uint16 iso_surface at maxVal/2 with the AO defaults (.1, 21, 30) -> --img^2, 36-frame
    sweep.  One GPU: VolumeRenderer.  N GPUs: z-slabs with an iso halo, sort-last (MIN-reduce of the candidate
    sample indices, owner resolves, SUM-reduce, post passes) -- strong scaling, same image on every rank."""
    import hashlib
    import torch
    import torch.distributed as dist
    import scenes
    import ctypes as C
    from spimagine_b200 import VolumeRenderer, _lib
    from spimagine_b200.multigpu import SlabMaxProjector, iso_halo, partition_slabs, slab_with_halo

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    N, W = args.vol, args.img
    dev = torch.device("cuda", local_rank)
    iso_max = 30000.
    if world == 1:
        rend = VolumeRenderer((W, W), device=local_rank, max_steps=MAX_STEPS, pinned_outputs=True)
        stream = torch.cuda.Stream(device=local_rank)
        torch.cuda.set_stream(stream)
        rend.use_stream(stream.cuda_stream)
        vol = vol_g_slab_device(N, 0, N, 1, dev)
        rend.set_data_device(vol.data_ptr(), (N, N, N), np.uint16)
        halo = 0
    else:
        halo = iso_halo(N, MAX_STEPS)
        z0, z1 = partition_slabs(N, world)[rank]
        lo, hi = slab_with_halo(z0, z1, N, halo)
        rend = SlabMaxProjector((W, W), rank=rank, world=world, device=local_rank, max_steps=MAX_STEPS,
                                pinned_outputs=True, halo=halo, composite=args.composite)
        torch.cuda.set_stream(rend._stream)
        vol = vol_g_slab_device(N, lo, hi, 1, dev)
        rend.set_slab((np.uint16, N, N), N, z0, z1, device_ptr=vol.data_ptr())
        if args.composite == "peer":
            rend.connect()
    rend.sync()
    del vol
    torch.cuda.empty_cache()
    rend.set_max_val(iso_max)
    if args.no_occ_table:
        _lib.check(rend._lib.spv_set_tuning(rend._ctx, 17, 0), rend._ctx)
    NF = 36
    cams = [scenes.gui_camera(2 * math.pi * f / NF, 4.0) for f in range(NF)]
    rend.set_projection(cams[0][1])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if world == 1 and not args.no_iso_overlap:
        # frames alternate between the two output slots; a frame's screen-space passes run beside the next frame's search
        _lib.check(rend._lib.spv_set_tuning(rend._ctx, 14, 1), rend._ctx)

    def device_step(i):
        if world == 1:
            _lib.check(rend._lib.spv_select_slot(rend._ctx, i & 1), rend._ctx)
        rend.set_modelView(cams[i % NF][0])
        if world == 1:
            p = _lib.IsoParams(rend._box(), iso_max / 2, 1., MAX_STEPS, .1, 21, 30, 0)
            _lib.check(rend._lib.spv_render_iso(rend._ctx, C.byref(p)), rend._ctx)
        elif args.composite == "peer":
            rend.enqueue_iso_composite()
        else:
            rend.iso_search()
            dist.all_reduce(rend.iso_k_tensor(), op=dist.ReduceOp.MIN)
            rend.iso_resolve()
            dist.all_reduce(rend.iso_planes_tensor(), op=dist.ReduceOp.SUM)
            _lib.check(rend._lib.spv_iso_slab_post(rend._ctx, C.byref(rend._iso_params())), rend._ctx)

    for i in range(args.warmup):
        device_step(i)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = rend.launch_count()
    e0.record()
    for i in range(args.steps):
        device_step(i)
    if world == 1:
        _lib.check(rend._lib.spv_stream_join(rend._ctx), rend._ctx)  # the last frames' passes, before the closing event
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches_counted = rend.launch_count() - launches0  # as the library counts its own launches
    if world == 1:
        _lib.check(rend._lib.spv_set_tuning(rend._ctx, 14, 0), rend._ctx)
        _lib.check(rend._lib.spv_select_slot(rend._ctx, 0), rend._ctx)
    digest = hashlib.sha1()
    hit_px = 0
    for i in range(8):  # untimed: hash of everything the first frames produce (compared across GPU counts)
        rend.set_modelView(cams[i][0])
        rend.render(method="iso_surface")
        for a in (rend.output, rend.output_depth, rend.output_normals, rend.output_occlusion):
            digest.update(np.ascontiguousarray(a).tobytes())
        hit_px = int(np.isfinite(rend.output_depth).sum())
    barrier()
    d2h0 = rend.d2h_bytes()
    t0 = time.perf_counter()
    for i in range(args.steps):
        rend.set_modelView(cams[i % NF][0])
        rend.render(method="iso_surface")
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    d2h_sync = (rend.d2h_bytes() - d2h0) / float(args.steps)
    barrier()
    # one GPU: the same frames through render_sequence (output + alpha of every frame reach pinned host memory; frame
    # i+1's search runs beside frame i's screen-space passes and read-back)
    t_seq = None
    if world == 1:
        for r_ in rend.render_sequence((cams[i % NF][0] for i in range(6)), method="iso_surface", iso_planes=2):
            pass
        chk = 0.
        b0 = rend.d2h_bytes()
        t0 = time.perf_counter()
        for r_ in rend.render_sequence((cams[i % NF][0] for i in range(args.steps)), method="iso_surface", iso_planes=2):
            chk += float(r_.output[W // 2, W // 2])
        torch.cuda.synchronize()
        t_seq = time.perf_counter() - t0
        d2h_seq = (rend.d2h_bytes() - b0) / float(args.steps)
    # where the time of a sort-last frame goes (peer composite): device time per phase, statistics on, untimed
    phases = None
    if world > 1 and args.composite == "peer":
        rend.time_phases(True)
        acc = {}
        for i in range(12):
            device_step(i)
            for k_, v_ in rend.last_phases_us().items():
                acc.setdefault(k_, []).append(v_)
        rend.time_phases(False)
        names = [k_ for k_ in rend.PHASES if len(acc.get(k_, [])) > 2]
        pv = torch.tensor([float(np.mean(acc[k_][2:])) for k_ in names], device="cuda", dtype=torch.float64)
        pmax = pv.clone()
        dist.all_reduce(pmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(pv, op=dist.ReduceOp.SUM)
        phases = dict((k_, {"mean_over_ranks_us": float(pv[j]) / world, "max_over_ranks_us": float(pmax[j])})
                      for j, k_ in enumerate(names))
    barrier()
    if world > 1:
        t = torch.tensor([ms, t_e2e * 1e3], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, t_e2e = float(t[0]), float(t[1]) / 1e3
    if rank == 0:
        peaks, peak_src = measured_peaks()
        alg = float(N) ** 3 * 2 + 6 * W * W * 4   # every voxel once + (3 + 3) result planes
        achieved = alg / (ms * 1e-3 / args.steps) / 1e9 / world
        traffic, traffic_src = ncu_traffic("iso_%d_%d" % (N, W), "iso") if world == 1 else (None, "single-GPU captures only")
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            # the oracle's whole pipeline (search, blur, occlusion, blur, shading) on all host cores, a few frames of the sweep
            from oracle import oracle
            use_all_host_threads()
            kind = "reference" if oracle.available("reference") else "port"
            build = "reference_fast" if kind == "reference" and oracle.available("reference_fast") and oracle.cpu_has_avx2() else kind
            host_vol = vol_g_slab_device(N, 0, N, 1, dev).cpu().numpy()
            o = oracle.OracleRenderer((W, W), kind=build, max_steps=MAX_STEPS)
            o.set_data(host_vol)
            o.set_projection(cams[0][1])
            o.set_max_val(iso_max)
            o.set_modelView(cams[0][0])
            o.render(method="iso_surface")
            nf = 6
            t0 = time.perf_counter()
            for i in range(nf):
                o.set_modelView(cams[(i * 6) % NF][0])
                o.render(method="iso_surface")
            t_cpu = (time.perf_counter() - t0) / nf
            cpu = {"value": 1. / t_cpu, "unit": "frames/s", "cores": int(o.lib.so_num_threads()), "kind": kind,
                   "sample": "%d whole frames of the sweep (every 6th), iso_surface + blur + occlusion + blur + shading; %s" % (
                       nf, "reference kernel text, -O3 -ffast-math -mavx2 -mfma (the reference's fast-math options)"
                       if build == "reference_fast" else "-O2")}
            del host_vol, o
        line_extra = {
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s (per GPU)",
                         "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": peak_src, "algorithmic_bytes_per_frame": alg,
                         "kernel": "spv::iso_fast_kernel<u16, linear, skip>" if world == 1 else "spv::iso_slab_search_kernel<u16, linear, skip>",
                         "overlap": ("frames alternate between two output slots; a frame's screen-space passes run on a second "
                                     "stream beside the next frame's search (tuning knob 14)") if world == 1 and not args.no_iso_overlap else None,
                         "note": "SURVEY 8d's algorithmic bytes count every voxel once; the search leaves cells that cannot hold "
                                 "a crossing in one step (exact hierarchical traversal), so the DRAM traffic is a few per cent of "
                                 "that and the fraction can exceed 1: this kernel is bound by the latency of its longest warps "
                                 "(DESIGN 4.3), not by HBM"}}
        if cpu:
            line_extra["cpu_baseline"] = cpu
        if phases:
            line_extra["phases_us"] = phases
        print(json.dumps(dict({
            "metric": "iso_surface frames/s, %d^3 uint16 -> %d^2, AO + shading, 36-frame sweep" % (N, W),
            "value": args.steps / (ms * 1e-3), "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u16->f32", "data": "synthetic",
            "config": {"workload": "Vol-G(%d, uint16, seed 1) generated on device, iso at %g, AO (.1, 21, 30), max_steps=200%s" % (
                N, iso_max / 2, "" if world == 1 else (", %d z-slabs with %d halo slices, sort-last over " % (world, halo) + (
                    "peer memory (candidates pushed to the band owners, MIN + redistribution by the owners, finished "
                    "pixels stored into every rank by the crossing's owner; arrival counters, no NCCL on the data path)"
                    if args.composite == "peer" else "NCCL (MIN int32 2 planes, SUM float32 7 planes)")))},
            "e2e": ({"value": args.steps / t_seq, "unit": "frames/s", "h2d_bytes_per_step": 128, "d2h_bytes_per_step": int(round(d2h_seq)),
                     "note": "VolumeRenderer.render_sequence(modelViews, method='iso_surface', iso_planes=2): output + alpha of "
                             "every frame reach pinned host memory (the rectangle the projected box can touch: the rest holds "
                             "no surface and reads out 0 / alpha 0 already); renders issued three frames ahead of the frame "
                             "handed out, read-backs two ahead"}
                    if t_seq else
                    {"value": args.steps / t_e2e, "unit": "frames/s", "h2d_bytes_per_step": 128, "d2h_bytes_per_step": 2 * W * W * 4,
                     "note": "SlabMaxProjector.set_modelView + render(method='iso_surface') on every rank: output + alpha read "
                             "back per frame (8 MiB)"}),
            "e2e_synchronous": {"value": args.steps / t_e2e, "unit": "frames/s", "h2d_bytes_per_step": 128,
                                "d2h_bytes_per_step": int(round(d2h_sync)),
                                "note": "set_modelView + render(method='iso_surface'), blocking per frame: output + alpha read "
                                        "back (the rectangle the projected box can touch); depth, normals and occlusion stay "
                                        "on the device until they are looked at (lazy attributes)"},
            "surface_pixels_last": hit_px, "image_sha1_first8": digest.hexdigest(),
            # one GPU (counted by the library): iso_fast, normal blur, occlusion (tap table: one launch), occlusion blur
            # + shading; sort-last: search, resolve, fix-up, 2+2 blur launches, occlusion list + queue, shading (the two
            # NCCL reductions are not counted)
            "gpu_launches": int(launches_counted) if world == 1 else args.steps * (13 if args.composite == "peer" else 10)},
                          **line_extra)))
    rend.close()
    if world > 1:
        dist.destroy_process_group()


def run_blur(args, rank, local_rank, world):
    """the reference's default image processor, BlurProcessor(sigma=4) = gputools.convolve_sep3 with 19
    taps per axis (models/imageprocessor.py:47-56), on a --vol^3 uint16 volume whose result goes to
    renderer.update_data (gui/mainwidget.py:455-465).  value: volumes/s of the three device passes on a resident
    volume; e2e: host volume -> processor chain -> renderer's resident array (apply_chain), against the
    reference-shaped host path timed beside it (apply() to a host array, then update_data)."""
    import torch
    import scenes
    from oracle import filters as forc
    from spimagine_b200 import VolumeRenderer, imageprocessor as ip

    if world > 1:
        if rank == 0:
            print(json.dumps({"metric": "blur", "unavailable": "the blur workload is single-GPU (replicas only)"}))
        return
    torch.cuda.set_device(local_rank)
    N = args.vol
    steps, warmup = min(args.steps, 50), max(3, min(args.warmup, 5))
    vol = scenes.vol_g(N, np.uint16, seed=0)
    proc = ip.BlurProcessor(sigma=4.)
    taps = proc._taps()
    vf = ip.VolumeFilter(local_rank)
    dvol = torch.from_numpy(vol.view(np.int16)).to("cuda:%d" % local_rank)
    torch.cuda.synchronize()
    clocks = ClockSampler(local_rank)
    clocks.start()
    by_mode = {}
    for fuse in (1, 0):  # x + y as one kernel (opt-in) / three passes (the default, timed last): same result
        vf.set_tuning(0, fuse)
        ms, pass_ms = [], []
        for i in range(warmup + steps):
            vf.load_device(dvol.data_ptr(), vol.shape, np.uint16)  # read in place: 512 MiB result > L2 between steps
            vf.convolve_sep3(*taps)
            vf.sync()
            if i >= warmup:
                ms.append(vf.last_ms())
                pass_ms.append(vf.last_pass_ms())
        by_mode[fuse] = float(np.mean(ms))
    dev_ms = by_mode[0]
    pass_ms = [float(x) for x in np.mean(np.array(pass_ms), axis=0)]  # x, y, z pass of the three-pass mode
    got = vf.result()
    # the spectrum processor (FFTProcessor, libspimfft.so) on the same resident volume, timed beside the blur
    plan = ip.SpectrumPlan(local_rank)
    fft_ms = []
    for i in range(3 + 10):
        plan.spectrum_device(dvol.data_ptr(), vol.shape, np.uint16)
        if i >= 3:
            fft_ms.append(plan.last_ms())
    fft_ms = float(np.mean(fft_ms))
    plan.close()
    # parity on the spot: a corner block against the CPU restatement (outputs within 9 voxels of the block's cut
    # faces would see voxels the block does not have)
    b = min(N, 96)
    want = forc.convolve_sep3(vol[:b, :b, :b], *taps)
    k = b if b == N else b - 9
    parity = bool(np.array_equal(got[:k, :k, :k], want[:k, :k, :k]))
    rend = VolumeRenderer((args.img, args.img), device=local_rank)
    rend.set_data(vol)
    t_chain, t_host = [], []
    for i in range(3 + 5):
        t0 = time.perf_counter()
        ip.apply_chain(rend, vol, [proc])
        rend.sync()
        t1 = time.perf_counter()
        rend.update_data(proc.apply(vol))  # the reference's shape: result to the host, cast, upload again
        rend.sync()
        t2 = time.perf_counter()
        if i >= 3:
            t_chain.append(t1 - t0)
            t_host.append(t2 - t1)
    clk = clocks.stop()
    t_chain, t_host = float(np.mean(t_chain)), float(np.mean(t_host))
    # CPU restatement on all host cores, a bounded sample of the same workload
    cb = min(N, 256)
    sample = np.ascontiguousarray(vol[:cb, :cb, :cb])
    use_all_host_threads()
    forc.convolve_sep3(sample[:32], *taps)
    t0 = time.perf_counter()
    forc.convolve_sep3(sample, *taps)
    t_cpu = (time.perf_counter() - t0) * (N / cb) ** 3
    peaks, peak_src = measured_peaks()
    nvox = float(N) ** 3
    alg = nvox * (2 + 4)  # every voxel read once (uint16) and the float32 result written once
    moved = nvox * (2 + 4 + 4 + 4 + 4 + 4)  # what the three passes move: x u16 -> f32, y f32 -> f32, z f32 -> f32
    achieved = alg / (dev_ms * 1e-3) / 1e9
    print(json.dumps({
        "metric": "BlurProcessor(sigma=4) volumes/s, %d^3 uint16 -> float32 (separable 19-tap convolution)" % N,
        "value": 1e3 / dev_ms, "unit": "volumes/s", "n_gpus": 1, "steps": steps, "warmup": warmup, "ms_per_step": dev_ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u16->f32", "data": "synthetic",
        "config": {"workload": "Vol-G(%d, uint16, seed 0), BlurProcessor(sigma=4): gputools.convolve_sep3 with 19 taps "
                               "per axis, zero boundary" % N,
                   "l2": "the %d MiB float32 result and the work volumes exceed the 126 MB L2" % (nvox * 4 / 2 ** 20)},
        "gvoxels_per_s": nvox / (dev_ms * 1e-3) / 1e9, "parity_subblock_bitwise": parity,
        "ms_three_passes": by_mode[0], "ms_fused_xy_plus_z": by_mode[1],
        "e2e": {"value": 1. / t_chain, "unit": "volumes/s", "h2d_bytes_per_step": int(nvox * 2), "d2h_bytes_per_step": 0,
                "note": "apply_chain(renderer, host volume, [BlurProcessor(4)]): upload (pageable host memory), three "
                        "passes, conversion to the renderer's uint16 texels and the z-pair array build on the device; "
                        "the call returns when the renderer's volume is replaced",
                "host_round_trip_value": 1. / t_host,
                "host_round_trip_note": "the reference's shape with the same kernels: proc.apply(data) -> float32 host "
                                        "array -> renderer.update_data(result)"},
        "gpu_launches": steps * 3, "clocks": clk,
        "fft_processor": {"ms": fft_ms, "volumes_per_s": 1e3 / fft_ms,
                          "note": "FFTProcessor on the same resident volume: pad + convert pass, real-to-complex cuFFT "
                                  "(library), fused magnitude / shift / scale / crop pass; device time of the three"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": achieved / peaks["hbm_gbs"], "traffic": None, "peak_source": peak_src,
                     "algorithmic_bytes_per_step": alg, "bytes_moved_by_the_three_passes": moved,
                     "frac_of_moved_bytes": moved / (dev_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                     "kernel": "spv::conv_x2_kernel<u16,19> (x, row pairs, pipelined) + 2 x spv::conv_axisw_kernel<19,16,4> "
                               "(y, z; four columns per thread), packed fma.rn.f32x2; timed together",
                     "passes": [{"pass": name, "ms": ms_k, "bytes": nvox * by, "gbytes_per_s": nvox * by / (ms_k * 1e-3) / 1e9,
                                 "frac_of_peak": nvox * by / (ms_k * 1e-3) / 1e9 / peaks["hbm_gbs"]}
                                for name, ms_k, by in zip(("x: uint16 -> float32", "y: float32 -> float32", "z: float32 -> float32"),
                                                          pass_ms, (6, 8, 8))],
                     "note": "each pass against the bytes it moves itself; the fraction above is against the algorithmic "
                             "bytes of the whole filter (every voxel read once, the result written once), which three "
                             "separable passes through memory cannot reach: at the copy rate they take %.3f ms, and the "
                             "57 fused multiply-adds per voxel alone take %.3f ms of the FMA pipes" % (
                                 moved / peaks["hbm_gbs"] / 1e6, 57 * nvox / (148 * 128 * 1.965e9) * 1e3)},
        "cpu_baseline": {"value": 1. / t_cpu, "unit": "volumes/s", "cores": use_all_host_threads(), "kind": "port",
                         "sample": "one %d^3 corner block of the volume through oracle/filter_oracle.c (OpenMP, all host "
                                   "cores), time scaled by (%d/%d)^3" % (cb, N, cb)}}))
    rend.close()
    vf.close()


def run_keyframes(args, rank, local_rank, world):
    """the record loop of the keyframe panel (gui/keyframe_view.py:644-653 -> GLWidget.render,
    gui/glwidget.py:610-636) over a path with max-projection and iso-surface stretches on a --vol^3 uint16 volume.
    value: frames/s of keyframes.render_keyframes (every frame lands in pinned host memory; frame i+1 renders while
    frame i is read back); beside it the reference-shaped loop (apply the transform, render(), blocking read-back)
    with the same kernels.  Also times frames.TiffData.read_into (a time point from an uncompressed TIFF straight
    into page-locked memory) against np.fromfile of the same bytes."""
    import tempfile
    import torch
    import scenes
    from spimagine_b200 import VolumeRenderer, frames, keyframes as kf, pinned_empty
    from spimagine_b200.utils import tiffio
    from spimagine_b200.utils.quaternion import Quaternion

    if world > 1:
        if rank == 0:
            print(json.dumps({"metric": "keyframes", "unavailable": "the keyframes workload is single-GPU (replicas only)"}))
        return
    torch.cuda.set_device(local_rank)
    N, W = args.vol, args.img
    n_frames = max(8, min(args.steps, 720))
    vol = scenes.vol_g(N, np.uint16, seed=0)
    keys = kf.KeyFrameList()
    keys.addItem(kf.KeyFrame(0., kf.TransformData(quatRot=Quaternion(1, 0, 0, 0), zoom=1., maxVal=PEAK_VALUE)))
    keys.addItem(kf.KeyFrame(.35, kf.TransformData(quatRot=Quaternion(.6, .2, .7, .1), zoom=1.25, maxVal=PEAK_VALUE * .7,
                                                  gamma=.8), 2.))
    keys.addItem(kf.KeyFrame(.6, kf.TransformData(quatRot=Quaternion(.1, .9, -.3, .2), zoom=.9, maxVal=PEAK_VALUE,
                                                 isIso=True)))
    keys.addItem(kf.KeyFrame(.8, kf.TransformData(quatRot=Quaternion(-.3, .5, .6, .4), zoom=1.1, maxVal=PEAK_VALUE,
                                                 bounds=[-.8, .8, -1, 1, -1, 1])))
    keys.addItem(kf.KeyFrame(1., kf.TransformData(quatRot=Quaternion(0, 0, 1, 0), zoom=1., maxVal=PEAK_VALUE)))
    rend = VolumeRenderer((W, W), device=local_rank, max_steps=MAX_STEPS, pinned_outputs=True)
    rend.set_data(vol)
    clocks = ClockSampler(local_rank)

    def pipelined():
        acc = 0.
        for pos, td, r in kf.render_keyframes(rend, keys, n_frames, iso_planes=2):
            acc += float(r.output[W // 2, W // 2])
        return acc

    def synchronous():
        acc = 0.
        for pos, t in kf.keyframe_times(n_frames):
            _, method = kf.apply_transform(rend, keys.getTransform(t))
            rend.render(method=method)
            acc += float(rend.output[W // 2, W // 2])
        return acc

    n_iso = sum(1 for _, t in kf.keyframe_times(n_frames) if keys.getTransform(t).isIso)
    for _ in range(3):
        pipelined()
    synchronous()
    clocks.start()
    l0 = rend.launch_count()
    t0 = time.perf_counter()
    chk_p = pipelined()
    t_pipe = time.perf_counter() - t0
    launches = rend.launch_count() - l0
    t0 = time.perf_counter()
    chk_s = synchronous()
    t_sync = time.perf_counter() - t0
    clk = clocks.stop()
    rend.close()

    # the TIFF time-point reader (page cache warm: the rate is the reader's, not the disk's)
    tmp = tempfile.mkdtemp(prefix="spv_bench_")
    fn = os.path.join(tmp, "vol.tif")
    tiffio.write3dTiff(vol, fn)
    d = frames.TiffData(fn)
    dst = pinned_empty(vol.shape, np.uint16)
    d.read_into(0, dst)
    t0 = time.perf_counter()
    for _ in range(3):
        d.read_into(0, dst)
    t_tif = (time.perf_counter() - t0) / 3
    same = bool(np.array_equal(dst, vol))
    vol.tofile(os.path.join(tmp, "vol.raw"))
    np.fromfile(os.path.join(tmp, "vol.raw"), dtype=np.uint16)
    t0 = time.perf_counter()
    for _ in range(3):
        np.fromfile(os.path.join(tmp, "vol.raw"), dtype=np.uint16)
    t_raw = (time.perf_counter() - t0) / 3
    for f in ("vol.tif", "vol.raw"):
        os.remove(os.path.join(tmp, f))
    os.rmdir(tmp)
    fps = n_frames / t_pipe
    print(json.dumps({
        "metric": "keyframe record loop frames/s, %d^3 uint16 -> %d^2 (max_project + iso_surface stretches)" % (N, W),
        "value": fps, "unit": "frames/s", "n_gpus": 1, "steps": n_frames, "warmup": 3 * n_frames,
        "ms_per_step": 1e3 * t_pipe / n_frames, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u16->f32", "data": "synthetic",
        "config": {"workload": "Vol-G(%d, uint16, seed 0), 5 keyframes (slerp, eased stretch, window / gamma / box "
                               "changes, one iso_surface stretch of %d frames), %d frames at recordPos / nFrames" % (
                                   N, n_iso, n_frames),
                   "l2": "the %d MiB z-paired volume exceeds the 126 MB L2 and the view changes every frame" % (
                       float(N) ** 3 * 4 / 2 ** 20)},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 128,
                "d2h_bytes_per_step": 2 * W * W * 4,
                "note": "keyframes.render_keyframes: host interpolation of the path, setters, one render per frame, "
                        "output + alpha of every frame read back into pinned host memory (the planes a recorded frame "
                        "shows; iso_planes=2), two frames in flight; wall clock",
                "synchronous_value": n_frames / t_sync,
                "synchronous_note": "the reference's loop with the same kernels: apply the transform, render(), "
                                    "blocking read-back per frame",
                "checksums_agree": bool(chk_p == chk_s)},
        "gpu_launches": int(launches), "clocks": clk,
        "tiff_reader": {"gbytes_per_s": vol.nbytes / t_tif / 1e9, "np_fromfile_raw_gbytes_per_s": vol.nbytes / t_raw / 1e9,
                        "bytes": int(vol.nbytes), "identical": same,
                        "note": "frames.TiffData.read_into: one readinto per time point into page-locked memory "
                                "(contiguous pages), page cache warm; beside it np.fromfile of the raw bytes into "
                                "fresh pageable memory (the reference's SpimData path)"}}))


def run_bricks(args, local_rank=0):
    """The single-GPU-brick baseline of BASELINE configs[3]: the same slab kernels on ONE GPU, the volume cut into
    --bricks z-slabs, each rendered by its own launch one after the other (what a single GPU does when the volume
    has to be processed brick by brick), the raw partials max-merged (torch.maximum: plumbing), then windowed."""
    import hashlib
    import torch
    import ctypes as C
    from spimagine_b200 import _lib
    from spimagine_b200.multigpu import SlabMaxProjector, partition_slabs, slab_with_halo

    torch.cuda.set_device(local_rank)
    N, W, B = args.vol, args.img, args.bricks
    stream = torch.cuda.Stream(device=local_rank)
    torch.cuda.set_stream(stream)
    parts = []
    for i, (z0, z1) in enumerate(partition_slabs(N, B)):
        lo, hi = slab_with_halo(z0, z1, N)
        slab = vol_g_slab_device(N, lo, hi, 2, torch.device("cuda", local_rank))
        r = SlabMaxProjector((W, W), rank=i, world=B, device=local_rank, max_steps=MAX_STEPS, pinned_outputs=True)
        r.use_stream(stream.cuda_stream)
        r.set_slab((np.uint16, N, N), N, z0, z1, device_ptr=slab.data_ptr())
        r.sync()
        del slab
        torch.cuda.empty_cache()
        parts.append(r)
    cams = sweep_cameras()
    last = parts[-1]
    last.set_max_val(PEAK_VALUE)
    last.set_projection(cams[0][1])
    mats = []
    for M, P in cams:
        last.set_modelView(M)
        mats.append((last._invP.copy(), last._invM.copy()))
    raws = [r._raw_tensor() for r in parts]
    raw_params = _lib.MipParams(last._box(), 0., PEAK_VALUE, 1., 0., 1, 0, MAX_STEPS, _lib.MIP_RAW_ONLY)
    lib = last._lib

    def step(i):
        invP, invM = mats[(i * 7) % SWEEP]
        for r in parts:
            lib.spv_set_matrices(r._ctx, _lib.fp(invP), _lib.fp(invM))
            rc = lib.spv_render_mip(r._ctx, C.byref(raw_params))
            if rc:
                _lib.check(rc, r._ctx)
        for t in raws[:-1]:
            torch.maximum(raws[-1], t, out=raws[-1])
        lib.spv_mip_finish(last._ctx, C.byref(raw_params))

    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    digest = hashlib.sha1()
    for i in range(min(8, args.steps)):
        step(i)
        out = np.empty((W, W), np.float32)
        _lib.check(lib.spv_read(last._ctx, _lib.BUF_OUT, _lib.fp(out), out.size), last._ctx)
        digest.update(out.tobytes())
    line = {
        "metric": "MIP frames/s, %d^3 uint16 -> %d^2, single-GPU-brick baseline" % (N, W),
        "value": args.steps / (ms * 1e-3), "unit": "frames/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u16->f32", "data": "synthetic",
        "config": {"workload": "Vol-G(%d, uint16, seed 2) generated on device, max_project -> %dx%d, max_steps=200, "
                               "%d z-slabs rendered one after the other on one GPU, max-merged, windowed" % (N, W, W, B)},
        "gpu_launches": args.steps * (B + 1), "image_sha1_first8": digest.hexdigest()}
    for r in parts:
        r.close()
    del parts, raws
    torch.cuda.empty_cache()
    return line


def run_slab(args, rank, local_rank, world, own_pg=True):
    """BASELINE configs[3]: --vol^3 uint16 split into `world` z-slabs, every frame = raw slab render on each GPU,
    all-reduce(MAX) over NCCL, window.  Strong scaling: the total work per frame is fixed."""
    import hashlib
    import torch
    import torch.distributed as dist
    import scenes
    import ctypes as C
    from spimagine_b200 import _lib
    from spimagine_b200.multigpu import SlabMaxProjector, partition_slabs_multi, slab_with_halo

    torch.cuda.set_device(local_rank)
    if world > 1 and own_pg:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    N, W = args.vol, args.img
    K = max(1, args.slabs_per_rank)
    mine = partition_slabs_multi(N, world, K)[rank]
    rend = SlabMaxProjector((W, W), rank=rank, world=world, device=local_rank, max_steps=MAX_STEPS,
                            pinned_outputs=True, composite=args.composite, slabs_per_rank=K)
    if rend._stream is None:
        rend._stream = torch.cuda.Stream(device=local_rank)
        rend.use_stream(rend._stream.cuda_stream)
    torch.cuda.set_stream(rend._stream)
    t_gen = t_upload = 0.
    slab_bytes = 0
    for i, (z0, z1) in enumerate(mine):
        lo, hi = slab_with_halo(z0, z1, N)
        slab_bytes = max(slab_bytes, (hi - lo) * N * N * 4)
        t0 = time.perf_counter()
        slab = vol_g_slab_device(N, lo, hi, 2, torch.device("cuda", local_rank))
        torch.cuda.synchronize()
        t_gen += time.perf_counter() - t0
        t0 = time.perf_counter()
        if i + 1 < len(mine):
            rend.add_slab((np.uint16, N, N), N, z0, z1, device_ptr=slab.data_ptr())
        else:
            rend.set_slab((np.uint16, N, N), N, z0, z1, device_ptr=slab.data_ptr())
        rend.sync()
        t_upload += time.perf_counter() - t0
        del slab
        torch.cuda.empty_cache()
    rend.set_max_val(PEAK_VALUE)
    cams = sweep_cameras()
    rend.set_projection(cams[0][1])
    mats = []
    for M, P in cams:
        rend.set_modelView(M)
        mats.append((rend._invP.copy(), rend._invM.copy()))
    lib, ctx = rend._lib, rend._ctx
    params = _lib.MipParams(rend._box(), 0., PEAK_VALUE, 1., 0., 1, 0, MAX_STEPS, _lib.MIP_RAW_ONLY)
    raw = rend._raw_tensor()
    peer = args.composite == "peer"
    if peer:
        rend.connect()
        params = _lib.MipParams(rend._box(), 0., PEAK_VALUE, 1., 0., 1, 0, MAX_STEPS, 0)
    launches_per_step = 4 if peer else 2

    def step(i):
        invP, invM = mats[(i * 7) % SWEEP]
        lib.spv_set_matrices(ctx, _lib.fp(invP), _lib.fp(invM))
        if peer:  # render + push over NVLink, counters, owner's max + window + redistribution: 4 launches, no NCCL
            rc = lib.spv_render_mip_composite(ctx, C.byref(params))
            if rc:
                _lib.check(rc, ctx)
            return
        rc = lib.spv_render_mip(ctx, C.byref(params))
        if rc:
            _lib.check(rc, ctx)
        if world > 1:
            dist.all_reduce(raw, op=dist.ReduceOp.MAX)
        lib.spv_mip_finish(ctx, C.byref(params))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    rend.enable_stats(True)
    step(0)
    hits0, issued0 = rend.last_stats()
    rend.enable_stats(False)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    def n_launches():
        return rend.launch_count()

    l0 = n_launches()
    for i in range(args.warmup):
        step(i)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = n_launches() - l0 - launches_per_step * args.warmup
    if peer:
        _lib.check(lib.spv_comp_check(ctx), ctx)

    # end to end: the public render() call; the composited image is complete on every GPU and read back to pinned
    # host memory on the display rank (rank 0)
    if peer:
        rend.readback_ranks = {0}
    digest = hashlib.sha1()
    for i in range(8):  # untimed: image hash of the first frames (compared across GPU counts and composites)
        rend.set_modelView(cams[(i * 7) % SWEEP][0])
        rend.render()
        if rend.readback_ranks is None or rank in rend.readback_ranks:
            digest.update(rend.output.tobytes())
    barrier()
    d2h0 = rend.d2h_bytes()
    t0 = time.perf_counter()
    for i in range(args.steps):
        rend.set_modelView(cams[(i * 7) % SWEEP][0])
        rend.render()
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    d2h_e2e = (rend.d2h_bytes() - d2h0) / float(args.steps)  # this rank's (rank 0: the display rank's) bytes per frame
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms, t_e2e * 1e3], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, t_e2e = float(t[0]), float(t[1]) / 1e3
        cnt = torch.tensor([float(issued0)], device="cuda", dtype=torch.float64)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        issued_total = float(cnt[0])
    else:
        issued_total = float(issued0)
    line = None
    if rank == 0:
        fps = args.steps / (ms * 1e-3)
        peaks, peak_src = measured_peaks()
        alg_bytes = float(N) ** 3 * 2 + 2 * W * W * 4
        line = {
            "metric": "MIP frames/s, %d^3 uint16 -> %d^2, sort-last z-slabs + %s max composite" % (
                N, W, "peer-memory" if peer else "NCCL"),
            "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u16->f32", "data": "synthetic",
            "config": {"workload": "Vol-G(%d, uint16, seed 2) generated on device, max_project -> %dx%d, "
                                   "max_steps=200, %d z-slab(s) with one halo slice (%d per GPU, serpentine), %s" % (
                                       N, W, W, world * K, K,
                                       "partials stored into the band owners' memory over NVLink, owner max + window, "
                                       "redistribution (spv_render_mip_composite, no NCCL on the data path)" if peer else
                                       "all-reduce(MAX) of the %d MiB raw plane over NCCL, then window" % (W * W * 4 >> 20)),
                       "composite": args.composite,
                       "l2": "each slab (%d MB as z-paired texels) exceeds the 126 MB L2" % (slab_bytes >> 20),
                       "parallelism": "sort-last, %d slab(s) on %d GPU(s)" % (world * K, world)},
            "gsamples_per_s": fps * hits0 * SAMPLES_PER_RAY / 1e9,
            "hit_rays_per_frame": hits0, "issued_samples_per_frame_all_ranks": issued_total,
            "e2e": {"value": args.steps / t_e2e, "unit": "frames/s", "h2d_bytes_per_step": 128,
                    "d2h_bytes_per_step": int(round(d2h_e2e)),
                    "note": "SlabMaxProjector.set_modelView + render() on every rank, composited output + alpha read "
                            "back into pinned host memory%s (the rectangle the projected box can touch; the rest of the "
                            "planes holds the miss values already)" % (" on rank 0 (the display rank)" if peer else " on every rank")},
            "image_sha1_first8": digest.hexdigest(),
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": alg_bytes / (ms * 1e-3 / args.steps) / 1e9 / world,
                         "peak": peaks["hbm_gbs"], "unit": "GB/s (per GPU)",
                         "frac": alg_bytes / (ms * 1e-3 / args.steps) / 1e9 / world / peaks["hbm_gbs"],
                         "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_frame_all_gpus": alg_bytes},
            "gen_s": t_gen, "upload_s": t_upload,
        }
    rend.close()
    del rend, raw
    torch.cuda.empty_cache()
    if world > 1 and own_pg:
        dist.destroy_process_group()
    return line if rank == 0 else None


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.workload == "timelapse":
        run_timelapse(args, rank, local_rank, world)
        return
    if args.workload == "blur":
        return run_blur(args, rank, local_rank, world)
    if args.workload == "iso":
        run_iso(args, rank, local_rank, world)
        return
    if args.workload == "keyframes":
        return run_keyframes(args, rank, local_rank, world)
    if args.workload == "slab" and args.bricks > 0:
        if world != 1:
            raise SystemExit("--bricks is the single-GPU baseline: run it with --gpus 1")
        print(json.dumps(run_bricks(args)))
        return
    if args.workload == "slab":
        line = run_slab(args, rank, local_rank, world)
        if rank == 0:
            print(json.dumps(line))
        return

    line = run_sweep(args, rank, local_rank, world)
    if rank == 0:
        print(json.dumps(line))


C4_BASELINE_TMP = "/tmp/spv_c4_baseline_%d_%d.json"


def c4_record(args, rank, local_rank, world):
    """BASELINE configs[3] under the same clock as the headline: a --c4-vol^3 uint16 volume -> --c4-img^2, sort-last.
      --gpus 1 : the two single-GPU denominators -- the volume as ONE resident array rendered by one launch
                 ("monolithic"), and cut into 8 z-slabs rendered one after the other and max-merged (the
                 single-GPU-brick baseline north_star names).  Written to /tmp as well, for a later N > 1 run on this box.
      --gpus N : z-slabs over the ranks, 4 per rank in serpentine order, partials pushed into the band owners' memory
                 over NVLink inside the render kernel (peer composite), and the same with an NCCL all-reduce(MAX);
                 speed-up against the brick baseline of this box (/tmp) or else of profiles/r02_c4_baseline_n1.json.
    The process group of the sweep is reused."""
    import copy
    a = copy.copy(args)
    a.vol, a.img = args.c4_vol, args.c4_img
    a.steps, a.warmup = max(4, args.c4_steps), max(3, min(args.warmup, 5))
    keep = ("value", "ms_per_step", "steps", "warmup", "image_sha1_first8", "gsamples_per_s", "hit_rays_per_frame",
            "issued_samples_per_frame_all_ranks", "gpu_launches", "gen_s", "upload_s")

    def brief(line):
        d = dict((k, line[k]) for k in keep if k in line)
        d["workload"] = line["config"]["workload"]
        if "e2e" in line:
            d["e2e"] = {"value": line["e2e"]["value"], "d2h_bytes_per_step": line["e2e"]["d2h_bytes_per_step"]}
        if "roofline" in line:
            d["roofline"] = line["roofline"]
        return d

    rec = {"metric": "MIP frames/s, %d^3 uint16 -> %d^2, sort-last (BASELINE configs[3])" % (a.vol, a.img),
           "n_gpus": world, "scaling": "strong", "unit": "frames/s"}
    tmp = C4_BASELINE_TMP % (a.vol, a.img)
    if world == 1:
        a.composite, a.slabs_per_rank, a.bricks = "nccl", 1, 0
        mono = run_slab(a, 0, local_rank, 1, own_pg=False)
        a.bricks = 8
        bricks = run_bricks(a, local_rank)
        rec["monolithic"] = brief(mono)
        rec["single_gpu_brick"] = brief(bricks)
        rec["value"] = mono["value"]
        rec["ms_per_step"] = mono["ms_per_step"]
        rec["same_image"] = mono["image_sha1_first8"] == bricks["image_sha1_first8"]
        try:
            with open(tmp, "w") as f:
                json.dump({"monolithic_ms": mono["ms_per_step"], "single_gpu_brick_ms": bricks["ms_per_step"],
                           "image_sha1_first8": bricks["image_sha1_first8"], "source_sha1": source_sha1()}, f)
        except OSError:
            pass
        return rec
    a.bricks = 0
    a.composite, a.slabs_per_rank = "peer", 4
    peer = run_slab(a, rank, local_rank, world, own_pg=False)
    a.composite = "nccl"
    nccl = run_slab(a, rank, local_rank, world, own_pg=False)
    if rank != 0:
        return None
    rec["peer_composite"] = brief(peer)
    rec["nccl_composite"] = brief(nccl)
    rec["value"] = peer["value"]
    rec["ms_per_step"] = peer["ms_per_step"]
    rec["image_sha1_first8"] = peer["image_sha1_first8"]
    rec["same_image_peer_nccl"] = peer["image_sha1_first8"] == nccl["image_sha1_first8"]
    W = a.img
    # per rank and frame: partials of the bands it does not own go to their owners; its own finished band goes to everyone
    rec["nvlink_bytes_per_frame_per_rank"] = {"partials_pushed": W * W * 4 * (world - 1) // world,
                                              "finished_band_stored_to_peers": W * W * 4 * (world - 1) // world}
    base, src = None, None
    if os.path.exists(tmp):
        with open(tmp) as f:
            base, src = json.load(f), "this box, the --gpus 1 run before this one (%s)" % tmp
    else:
        path = os.path.join(ROOT, "profiles", "r02_c4_baseline_n1.json")
        if os.path.exists(path) and (a.vol, a.img) == (2048, 2048):
            with open(path) as f:
                base, src = json.load(f), "profiles/r02_c4_baseline_n1.json (another B200 of this pool, `bench.py --gpus 1`)"
    if base is not None:
        rec["baseline"] = dict(base, source=src)
        rec["speedup_vs_single_gpu_brick"] = base["single_gpu_brick_ms"] / peer["ms_per_step"]
        rec["speedup_vs_monolithic"] = base["monolithic_ms"] / peer["ms_per_step"]
        rec["same_image_as_baseline"] = base.get("image_sha1_first8") == peer["image_sha1_first8"]
    return rec


KERNEL_SOURCES = {"mip": ("spv_mip.cu", "spv_mip_axis.cu", "spv_common.cuh", "spv_kernels.h"), "iso": ("spv_iso.cu", "spv_common.cuh", "spv_kernels.h")}


def source_sha1(family="mip"):
    """SHA-1 over the sources of a kernel family: ties a committed ncu traffic figure to the code it was measured on
    (profiles/r02_mip_traffic.json is written by scripts/ncu_traffic.py)."""
    import hashlib
    h = hashlib.sha1()
    for f in KERNEL_SOURCES[family]:
        with open(os.path.join(ROOT, "spimagine_b200", "csrc", f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def ncu_traffic(workload_key, family="mip"):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from the committed ncu capture of
    this workload, or (None, why) when there is none for the current kernel sources."""
    path = os.path.join(ROOT, "profiles", "r02_mip_traffic.json")
    if not os.path.exists(path):
        return None, "no ncu capture committed for this round (scripts/ncu_traffic.py writes %s)" % os.path.basename(path)
    with open(path) as f:
        rec = json.load(f).get(workload_key)
    if rec is None:
        return None, "no ncu capture of this workload in profiles/r02_mip_traffic.json"
    if rec.get("source_sha1") != source_sha1(family):
        return None, "the committed ncu capture (profiles/r02_mip_traffic.json) predates the current kernel sources"
    return int(rec["dram_bytes_read"] + rec["dram_bytes_write"]), \
        "profiles/r02_mip_traffic.json: ncu dram__bytes_read.sum + dram__bytes_write.sum, mean over %d launches of %s" % (
            rec.get("launches", 0), rec.get("kernel", "?"))


def run_sweep(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    import scenes
    import ctypes as C
    from spimagine_b200 import VolumeRenderer, _lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU path in spimagine_b200)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        bind_to_gpu_numa_node(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    f32 = args.dtype == "f32"
    peak = 1. if f32 else PEAK_VALUE
    vol = scenes.vol_g(args.vol, np.float32 if f32 else np.uint16, seed=0)
    K = args.steps
    thetas = sweep_thetas(K, rank, world)
    cams = [scenes.gui_camera(th, 4.0) for th in thetas]
    W = H = args.img
    rend = VolumeRenderer((W, H), device=local_rank, max_steps=MAX_STEPS, pinned_outputs=True)
    # a dedicated (non-default) stream shared by torch and the renderer, so torch's events bracket the kernels
    stream = torch.cuda.Stream(device=local_rank)
    torch.cuda.set_stream(stream)
    rend.use_stream(stream.cuda_stream)
    t0 = time.perf_counter()
    rend.set_data(vol)
    torch.cuda.synchronize()
    t_upload = time.perf_counter() - t0
    rend.set_max_val(peak)
    rend.set_skipping(args.skip)
    rend.set_projection(cams[0][1])
    if args.mip_path is not None:
        rend.set_mip_path(args.mip_path)  # raises where the library has no such path

    # camera matrices of this rank's frames as the kernels take them (volumerender.py:310-316)
    mats = []
    for M, P in cams:
        rend.set_modelView(M)
        mats.append((rend._invP.copy(), rend._invM.copy()))
    lib, ctx = rend._lib, rend._ctx
    rend.set_alpha_pow(args.alpha_pow)
    params = _lib.MipParams(rend._box(), 0., peak, 1., float(args.alpha_pow), 1, 0, MAX_STEPS, 0)

    overlap = [False]
    if args.no_axis:
        _lib.check(lib.spv_set_tuning(ctx, 16, 0), ctx)
        args.batch = 1
    B = max(1, min(int(args.batch), _lib.MAX_BATCH))
    if B > 1 and lib.spv_mip_batch_possible(ctx, C.byref(params)) != 1:
        B = 1  # float volumes, attenuated renders, skipping, the software-sampled path: one launch per frame
    invM_all = np.ascontiguousarray(np.stack([m[1] for m in mats]), dtype=np.float32)  # [K][16]

    def device_batch(i0, n):
        """frames i0 .. i0 + n - 1 (mod K) in ONE launch; results stay on the device"""
        idx = [(i0 + j) % K for j in range(n)]
        inv = invM_all[idx] if idx != list(range(idx[0], idx[0] + n)) else invM_all[idx[0]:idx[0] + n]
        inv = np.ascontiguousarray(inv)
        used = C.c_int()
        rc = lib.spv_render_mip_batch(ctx, C.byref(params), _lib.fp(inv), n, 0, C.byref(used))
        if rc:
            _lib.check(rc, ctx)

    def device_step(i):
        if overlap[0]:  # frames alternate between the two output slots; slot 1's kernel runs on a second stream
            lib.spv_select_slot(ctx, i & 1)
        invP, invM = mats[i % K]
        lib.spv_set_matrices(ctx, _lib.fp(invP), _lib.fp(invM))
        rc = lib.spv_render_mip(ctx, C.byref(params))
        if rc:
            _lib.check(rc, ctx)

    # algorithmic samples: hit rays x 208, counted on the device in an untimed pass over exactly the timed frames
    rend.enable_stats(True)
    hits, issued = [], []
    for i in range(K):
        device_step(i)
        h, s = rend.last_stats()
        hits.append(h)
        issued.append(s)
    rend.enable_stats(False)
    mean_hits = float(np.mean(hits))
    mean_issued = float(np.mean(issued))

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ----
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if B > 1:
        # K frames in ceil(K / B) launches of (up to) B frames each; the warm-up steps the same way
        for i in range(0, args.warmup, B):
            device_batch(i, min(B, args.warmup - i))
        barrier()
        launches0 = rend.launch_count()
        e0.record()
        for i in range(0, K, B):
            device_batch(i, min(B, K - i))
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = rend.launch_count() - launches0
    else:
        if not args.no_mip_overlap and not args.alpha_pow and args.mip_path != "smem":
            _lib.check(lib.spv_set_tuning(ctx, 15, 1), ctx)
            overlap[0] = True
        launches0 = rend.launch_count()
        for i in range(args.warmup):
            device_step(i)
        barrier()
        e0.record()
        for i in range(K):
            device_step(i)
        if overlap[0]:
            _lib.check(lib.spv_stream_join(ctx), ctx)  # the last slot-1 frame, before the closing event
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if overlap[0]:
            _lib.check(lib.spv_set_tuning(ctx, 15, 0), ctx)
            _lib.check(lib.spv_select_slot(ctx, 0), ctx)
            overlap[0] = False
        launches = rend.launch_count() - launches0 - args.warmup
    kernels_per_frame = launches / float(K)
    axis_used = rend.mip_axis_used()

    # ---- end to end through the public API: render_sequence() over the same frames.  Every frame's output and alpha
    # reach pinned host memory inside the timed region; frame i+1 renders while frame i is in flight
    for r in rend.render_sequence((cams[i % K][0] for i in range(min(args.warmup, 5))), batch=B):
        pass
    barrier()
    checksum_seq = 0.0
    b0 = rend.d2h_bytes()
    t0 = time.perf_counter()
    for r in rend.render_sequence((cams[i][0] for i in range(K)), batch=B):
        checksum_seq += float(r.output[H // 2, W // 2])
    torch.cuda.synchronize()
    t_seq = time.perf_counter() - t0
    d2h_seq = (rend.d2h_bytes() - b0) / float(K)
    barrier()

    # ---- end to end, one blocking call per frame (the reference's frame loop as it stands) ----
    def api_step(i):
        rend.set_modelView(cams[i % K][0])
        rend.render()
        return rend.output

    for i in range(min(args.warmup, 5)):
        api_step(i)
    barrier()
    checksum = 0.0
    b0 = rend.d2h_bytes()
    t0 = time.perf_counter()
    for i in range(K):
        out = api_step(i)
        checksum += float(out[H // 2, W // 2])
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    d2h_sync = (rend.d2h_bytes() - b0) / float(K)
    barrier()

    # what the host link gives this rank while every rank copies at once: the same bytes per frame as render_sequence
    # moves, as plain device -> pinned-host copies back to back (torch: plumbing) -- the ceiling of `e2e` on this box
    nprobe = max(1, int(d2h_seq) // 4)
    dev_buf = torch.empty(nprobe, dtype=torch.float32, device="cuda")
    host_buf = torch.empty(nprobe, dtype=torch.float32, pin_memory=True)
    for _ in range(3):
        host_buf.copy_(dev_buf, non_blocking=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(50):
        host_buf.copy_(dev_buf, non_blocking=True)
    torch.cuda.synchronize()
    d2h_probe = 50 * nprobe * 4 / (time.perf_counter() - t0) / 1e9
    barrier()
    del dev_buf, host_buf

    clocks = sampler.stop() if rank == 0 else None

    if world > 1:
        pr = torch.tensor([d2h_probe], device="cuda", dtype=torch.float64)
        dist.all_reduce(pr, op=dist.ReduceOp.MIN)
        d2h_probe = float(pr[0])
        t = torch.tensor([ms, t_e2e * 1e3, t_seq * 1e3], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, t_e2e, t_seq = float(t[0]), float(t[1]) / 1e3, float(t[2]) / 1e3
        cnt = torch.tensor([mean_hits, mean_issued, d2h_seq, d2h_sync], device="cuda", dtype=torch.float64)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        mean_hits, mean_issued, d2h_seq, d2h_sync = (float(x) / world for x in cnt)

    line = None
    if rank == 0:
        total_frames = K * world
        fps = total_frames / (ms * 1e-3)
        peaks, peak_src = measured_peaks()
        launch_s = ms * 1e-3 / K           # per frame (the timed region / frames)
        alg_bytes = vol.nbytes + 2 * W * H * 4  # per frame: every voxel once + both result planes
        achieved = alg_bytes / launch_s / 1e9
        frames_per_launch = K / float(launches) if launches else 1.
        axis_path = axis_used[0] >= 0
        tex_peak = rend.texrate_probe(4000)
        # the same probe with the benchmark camera's footprint: neighbouring rays one pixel apart at the volume's
        # centre (2 d tan(fovy/2) / W box units, d = 4), samples L/192 apart along the view direction (mean in-box
        # path L = 1.23 box units), at (up to 24 of) the angles that were timed, each weighted by the
        # samples its frame issues; every fetch hits L1
        tpu = args.vol / 2.                       # texels per box unit
        pitch = 2. * 4. * np.tan(np.radians(30.)) / W * tpu
        step = 1.23 / (MAX_STEPS // 16 * 16) * tpu
        pick = sorted(set(int(round(x)) for x in np.linspace(0, K - 1, min(K, 24))))
        if axis_path:
            pick = []  # the probe lays lanes out as 2x2-pixel quads on the z-paired array: not this path's footprint
        foot = []
        for j in pick:
            th = thetas[j] + 1e-3
            c, s_ = np.cos(th), np.sin(th)
            foot.append(rend.texrate_probe(2000, footprint=[[pitch * c, 0., -pitch * s_], [0., pitch, 0.],
                                                            [-step * s_, 0., -step * c]]))
        wts = np.array([issued[j] for j in pick], float)
        if wts.sum() == 0:  # attenuated renders do not count samples
            wts[:] = 1.
        tex_foot = wts.sum() / sum(w / r for w, r in zip(wts, foot)) if pick else None  # samples / sum(samples / rate)
        traffic, traffic_src = ncu_traffic("sweep_%d_%d%s" % (args.vol, W, "_b%d" % B if axis_path else ""))
        kernel_name = rend.mip_kernel_name() if hasattr(rend, "mip_kernel_name") else "spv::mip_fast_kernel<u16, linear>"
        if axis_path:
            kernel_name = "spv::mip_axis_kernel<u16> (layer axis %s, %s)" % (
                "xyz"[axis_used[0]], ("2x2-pixel quads", "4x1 row quads", "1x4 column quads")[axis_used[1]])
        metric = sweep_metric(args)
        if f32:
            kernel_name = ("spv::mip_axis_kernel<f32> (layer axis %s, %s)" % (
                "xyz"[axis_used[0]], ("2x2-pixel quads", "4x1 row quads", "1x4 column quads")[axis_used[1]])) if axis_path else \
                "spv::mip_fast_kernel<f32, linear>"
        if args.alpha_pow:
            kernel_name = ("spv::mip_axis_kernel<u16, attenuated> (layer axis %s)" % "xyz"[axis_used[0]]) if axis_path else \
                "spv::mip_alpha_kernel<%s, linear>" % args.dtype
        line = {
            "metric": metric, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": K,
            "warmup": args.warmup, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if f32 else "u16->f32", "data": "synthetic",
            "config": sweep_config(args),
            "notes": {
                "l2": ("the %d MB volume (%d MB as the paired layered copy the kernel samples) exceeds the 126 MB L2 and the "
                       "view changes every step%s" % (vol.nbytes >> 20, vol.nbytes >> 19, (
                           "; the %d frames of one launch (distinct views, each computed in full) are scheduled tile row by "
                           "tile row so that they share the slab of the volume that row crosses while it is in L2 -- that "
                           "schedule is the kernel's design, nothing is reused between launches or steps" % B) if B > 1 else "")),
                "frames_per_launch": B,
                "skipping": bool(args.skip), "parallelism": "frames sharded over %d GPU(s), volume replicated" % world,
                "angles_deg_rank0": [round(math.degrees(t), 3) for t in thetas[:4]] + ["..."] if K > 4 else
                                    [round(math.degrees(t), 3) for t in thetas]},
            "gsamples_per_s": fps * mean_hits * SAMPLES_PER_RAY / 1e9,
            "hit_rays_per_frame": mean_hits,
            "issued_samples_per_frame": mean_issued,
            "e2e": {"value": total_frames / t_seq, "unit": "frames/s", "h2d_bytes_per_step": 128,
                    "d2h_bytes_per_step": int(round(d2h_seq)),
                    "d2h_gbytes_per_s_per_rank": d2h_seq * K / t_seq / 1e9,
                    "d2h_ceiling_gbytes_per_s_per_rank": d2h_probe,
                    "d2h_ceiling_frames_per_s": world * d2h_probe * 1e9 / max(1., d2h_seq),
                    "d2h_ceiling_note": "plain device -> pinned host copies of the same size, all ranks at once, slowest "
                                        "rank: what the host link of this box leaves for `e2e`",
                    "note": ("VolumeRenderer.render_sequence(modelViews): per frame the host inverts the 4x4 modelView; %d "
                             "frames go into one launch (spv_render_mip_batch) and the launch's output + alpha planes are "
                             "copied to pinned host memory while the next launch renders (only the rectangle the projected "
                             "box can touch travels: %.0f%% of 2*W*H*4 bytes; the rest already holds the miss values); a pixel "
                             "of every frame is read on the host" % (B, 100. * d2h_seq / (2 * W * H * 4))) if B > 1 else
                            ("VolumeRenderer.render_sequence(modelViews): per frame the host inverts the 4x4 modelView, "
                             "launches, and output + alpha are copied to pinned host memory (only the rows the projected "
                             "box can touch travel: %.0f%% of 2*W*H*4 bytes; the others already hold the miss values); "
                             "frame i+1 renders while frame i is in flight; a pixel of every frame is read on the host" % (
                                 100. * d2h_seq / (2 * W * H * 4))),
                    "checksum": checksum_seq},
            "e2e_synchronous": {"value": total_frames / t_e2e, "unit": "frames/s", "h2d_bytes_per_step": 128,
                                "d2h_bytes_per_step": int(round(d2h_sync)),
                                "note": "the drop-in call of the reference's frame loop, blocking per frame: "
                                        "VolumeRenderer.set_modelView + render() = host 4x4 inversion, one launch, output "
                                        "+ alpha copied in 12 bands into pinned host memory by two copy streams that wait "
                                        "on per-band completion counters (cuStreamWaitValue32), wait",
                                "checksum": checksum},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes * frames_per_launch,
                         "avg_launch_us": launch_s * 1e6 * frames_per_launch,
                         "frames_per_launch": frames_per_launch,
                         "algorithmic_bytes_per_frame": alg_bytes, "us_per_frame": launch_s * 1e6,
                         "kernels_per_frame": kernels_per_frame, "kernel": kernel_name,
                         "note": ("one launch renders %d frames whose CTAs run tile row by tile row, so the frames share the "
                                  "volume in L2: `traffic` (DRAM bytes per launch) is BELOW the algorithmic bytes (every voxel "
                                  "once PER FRAME + the result planes); the path is bound by the texture unit, see "
                                  "roofline_tex" % B) if B > 1 else None,
                         "overlap": None if (B > 1 or args.no_mip_overlap or args.alpha_pow or args.mip_path == "smem") else
                         "frames alternate between two output slots and two streams: a frame's kernel starts in the tail of "
                         "the one before (tuning knob 15); avg_launch_us is the timed region / frames"},
            "roofline_tex": {"bound": "texture samples", "issued_gsamples_per_s": mean_issued / launch_s / 1e9,
                             "algorithmic_gsamples_per_s": mean_hits * SAMPLES_PER_RAY / launch_s / 1e9,
                             "peak_gsamples_per_s": tex_peak / 1e9,
                             "frac_issued": mean_issued / launch_s / tex_peak,
                             "frac_algorithmic": mean_hits * SAMPLES_PER_RAY / launch_s / tex_peak,
                             "peak_source": "spv_texrate_probe on this GPU: cache-resident trilinear uint16 fetches",
                             "peak_at_render_footprint_gsamples_per_s": tex_foot / 1e9 if tex_foot else None,
                             "frac_issued_at_render_footprint": mean_issued / launch_s / tex_foot if tex_foot else None,
                             "render_footprint": None if not tex_foot else {"ray_spacing_texels": pitch, "sample_spacing_texels": step,
                                                  "gsamples_per_s_by_angle_deg": dict(
                                                      ("%.1f" % math.degrees(thetas[j]), r / 1e9) for j, r in zip(pick, foot)),
                                                  "note": "spv_texrate_probe_footprint at the timed angles (builder-defined "
                                                          "denominator, not SURVEY 8d's): the same independent fetches "
                                                          "laid out like this camera's rays (a warp's 8x4 tile one "
                                                          "pixel apart, 16 samples in flight along the view "
                                                          "direction), all warps on one L1-resident region: what the "
                                                          "texture unit delivers for this footprint without misses"}},
            "upload_s": t_upload,
        }
        if args.alpha_pow:
            line["config"] = dict(line["config"], window=line["config"]["window"].replace("alpha_pow 0", "alpha_pow %g" % args.alpha_pow))
            del line["roofline_tex"]   # executed samples depend on where a ray goes dark: not counted
            line["gsamples_per_s"] = None
            line["issued_samples_per_frame"] = None
            line["hit_rays_per_frame"] = None
        if world == 1 and not args.no_cpu_baseline:
            sub = cams[::max(1, K // 24)][:24]
            res = cpu_reference(vol, sub, args.img, len(sub), 1, 20.0, alpha_pow=args.alpha_pow, peak=peak)
            line["cpu_baseline"] = {"value": res["fps"], "unit": "frames/s", "cores": res["cores"],
                                    "kind": res["kind"], "sample": res["sample"],
                                    "gsamples_per_s": res["gsamples"]}
    rend.close()
    del rend
    torch.cuda.empty_cache()
    if not args.no_c4 and not args.alpha_pow and (args.vol, args.img, args.dtype) == (VOL_N, IMG, "u16"):
        rec = c4_record(args, rank, local_rank, world)
        if rank == 0:
            line["c4"] = rec
    if world > 1:
        dist.destroy_process_group()
    return line


if __name__ == "__main__":
    main()
