"""Multi-GPU rendering: one process per GPU (torch.distributed), new relative to the reference, which is
single-device.

  partition_slabs / SlabMaxProjector   sort-last max projection of a volume split into z-slabs.  Every rank
      marches the SAME rays (global camera, tnear/tfar, dt, sample indices) but evaluates only the samples
      whose trilinear footprint starts in its own slices; the un-windowed partial maxima are merged with one
      all-reduce(MAX) over NCCL / NVLink and windowed afterwards.  max is associative, commutative and
      idempotent, so the composite equals the single-GPU render bit for bit.
  frames_for_rank / TimelapsePlayer    3D+t playback: frame t is rendered by rank t mod world, no data-path
      collective at all.

torch is plumbing here (process group, the all-reduce on a tensor that aliases the renderer's device
buffer); all rendering is libspimcuda.
"""
from __future__ import absolute_import, print_function

import ctypes as C

import numpy as np

from . import _lib
from .volumerender import VolumeRenderer


def partition_slabs(nz, world):
    """Split slices [0, nz) into `world` contiguous, near-equal, non-empty slabs -> list of (z0, z1)."""
    if world < 1 or nz < world:
        raise ValueError("cannot split %d slices over %d ranks" % (nz, world))
    base, rem = divmod(nz, world)
    out, z = [], 0
    for r in range(world):
        n = base + (1 if r < rem else 0)
        out.append((z, z + n))
        z += n
    return out


def partition_slabs_multi(nz, world, k):
    """world * k contiguous slabs dealt to the ranks in serpentine order: rank r gets slabs r, 2*world-1-r,
    2*world+r, ...  Front and back slabs pair up, so the per-rank sample counts stay balanced whatever the view
    direction (a perspective camera puts up to 3x more samples into the slabs nearest to it).
    -> list over ranks of lists of (z0, z1)."""
    parts = partition_slabs(nz, world * k)
    out = [[] for _ in range(world)]
    for i, p in enumerate(parts):
        rnd, j = divmod(i, world)
        out[j if rnd % 2 == 0 else world - 1 - j].append(p)
    return out


def slab_with_halo(z0, z1, nz, halo=1):
    """Slices a rank must hold to evaluate the samples it owns: `halo` slices either side (1 for max projection;
    iso surfaces need iso_halo(...) because the gradient taps reach +-2h along z)."""
    return max(z0 - halo, 0), min(z1 + halo, nz)


def iso_halo(nz, max_steps=200, gamma=1., path=2 * 3 ** .5):
    """Halo slices a slab needs for sort-last iso surfaces: the 12-tap gradient reaches +-2h with h = dt * gamma^2 in
    normalised texture units (iso_kernel.cl:163-192), dt = path / (max_steps - 1), plus one ray step for the
    bracket refinement.  `path` = longest in-box ray length in eye-ray units (box diagonal for an unscaled
    modelView).  The kernels check the actual requirement per pixel and the render raises if this was too small."""
    dt = path / (max_steps - 1.)
    return int(np.ceil(2 * dt * gamma * gamma * nz + 0.5 * dt * nz)) + 3


def frames_for_rank(n_frames, rank, world):
    """Frame sharding for timelapse playback: frame t -> rank t mod world."""
    return list(range(rank, n_frames, world))


def composite_max(tensor, group=None):
    """In-place element-wise maximum over all ranks (the sort-last composite).  Bit-exact for any rank count."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.MAX, group=group)
    return tensor


class _DevBuffer(object):
    """Exposes a raw device pointer through __cuda_array_interface__ so torch can alias it without a copy."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class SlabMaxProjector(VolumeRenderer):
    """Sort-last max projection across the ranks of a torch.distributed process group.

    Same setters as VolumeRenderer.  `set_data(volume)` takes the WHOLE (Nz, Ny, Nx) volume (or anything
    sliceable along z, e.g. a memmap) and uploads only this rank's slab + halo; `set_slab(slab, gnz, z0, z1)`
    takes the slab directly.  `render()` leaves the composited image in `output` on every rank.
    """

    def __init__(self, size=None, interpolation='linear', group=None, rank=None, world=None, composite="nccl",
                 slabs_per_rank=1, **kw):
        """composite = "nccl": all-reduce(MAX) of the raw plane (torch.distributed / NCCL), then window;
        "peer": the render kernel stores its partials straight into the band owners' memory over NVLink and the
        owners window + redistribute (spv_render_mip_composite) -- call connect() (one process per GPU) or
        connect_local([...]) (several renderers in one process) once after construction / resize."""
        import torch
        import torch.distributed as dist
        self._torch = torch
        self._dist = dist
        self.group = group
        if composite not in ("nccl", "peer"):
            raise KeyError("composite = '%s' not defined, valid: ['nccl', 'peer']" % composite)
        self.composite = composite
        self._connected = False
        self.slabs_per_rank = int(slabs_per_rank)
        self.halo = int(kw.pop("halo", 1))  # slices either side of the owned slab (iso surfaces: iso_halo(nz))
        self.readback_ranks = None  # None: every rank reads the composited image back; or a set of ranks (display rank)
        self._parts = []  # helper renderers holding this rank's other slabs (slabs_per_rank > 1)
        self._ctor = (interpolation, dict(kw))
        if rank is None or world is None:
            if dist.is_available() and dist.is_initialized():
                rank, world = dist.get_rank(group), dist.get_world_size(group)
            else:
                rank, world = 0, 1
        self.rank, self.world = rank, world
        super(SlabMaxProjector, self).__init__(size, interpolation, **kw)
        self._stream = None
        if world > 1 and torch.cuda.is_available():
            # one non-default stream for both the render kernels and the collective: ordered without host syncs
            self._stream = torch.cuda.Stream(device=self.device)
            self.use_stream(self._stream.cuda_stream)

    def set_data(self, data, autoConvert=True, copyData=False):
        if not autoConvert and not data.dtype in self.dtypes:
            raise NotImplementedError("data type should be either %s not %s" % (self.dtypes, data.dtype))
        nz = data.shape[0]
        mine = partition_slabs_multi(nz, self.world, self.slabs_per_rank)[self.rank]
        self.clear_parts()
        for i, (z0, z1) in enumerate(mine):
            lo, hi = slab_with_halo(z0, z1, nz, self.halo)
            slab = np.asarray(data[lo:hi])
            if slab.dtype.type not in self.dtypes:
                slab = slab.astype(self.dtype, copy=False)
            if i + 1 < len(mine):
                self.add_slab(slab, nz, z0, z1)
            else:
                self.set_slab(slab, nz, z0, z1)

    # ---- several slabs per rank: helper contexts hold the data, this context's kernel marches all of them ----
    def _register_parts(self):
        arr = (_lib._CTX * max(1, len(self._parts)))(*[h._ctx for h in self._parts])
        self._check(self._lib.spv_set_extra_slabs(self._ctx, arr, len(self._parts)))

    def clear_parts(self):
        self._check(self._lib.spv_set_extra_slabs(self._ctx, None, 0))
        for h in self._parts:
            h.close()
        self._parts = []

    def add_slab(self, slab, gnz, z0, z1, device_ptr=None):
        """One more slab (at most 3) for this rank to render besides the one given to set_slab(): every ray marches
        the samples owned by each resident slab in the same kernel launch."""
        interpolation, kw = self._ctor
        h = SlabMaxProjector((16, 16), interpolation, rank=self.rank, world=self.world, composite="nccl",
                             halo=self.halo, **kw)
        h.set_layout(getattr(self, "layout", "zpair"))
        h._check(h._lib.spv_share_stream(h._ctx, self._ctx))
        h.set_slab(slab, gnz, z0, z1, device_ptr=device_ptr)
        self._parts.append(h)
        self._register_parts()

    def use_stream(self, cuda_stream=None):
        super(SlabMaxProjector, self).use_stream(cuda_stream)
        for h in getattr(self, "_parts", []):
            h._check(h._lib.spv_share_stream(h._ctx, self._ctx))

    def close(self):
        if getattr(self, "_parts", None) and getattr(self, "_ctx", None) is not None and self._ctx.value:
            self._lib.spv_set_extra_slabs(self._ctx, None, 0)
        for h in getattr(self, "_parts", []):
            h.close()
        self._parts = []
        super(SlabMaxProjector, self).close()

    def set_slab(self, slab, gnz, z0, z1, device_ptr=None):
        """slab: ndarray holding global slices [max(z0-halo,0), min(z1+halo,gnz)) with halo = self.halo; or pass
        device_ptr (int) to a C-order device copy of it together with slab=(dtype, ny, nx)."""
        lo, hi = slab_with_halo(z0, z1, gnz, self.halo)
        if device_ptr is None:
            slab = np.ascontiguousarray(slab)
            if slab.shape[0] != hi - lo:
                raise ValueError("slab holds %d slices, expected %d (with halo)" % (slab.shape[0], hi - lo))
            dtype, ny, nx = slab.dtype, slab.shape[1], slab.shape[2]
            ptr, on_dev = slab.ctypes.data, 0
        else:
            dtype, ny, nx = slab
            ptr, on_dev = int(device_ptr), 1
        dtype = np.dtype(dtype)
        self.set_dtype(dtype.type)
        self.dataSlices = None
        self.set_shape((nx, ny, gnz))
        self.slab = (z0, z1)
        self._check(self._lib.spv_set_volume_slab_halo(self._ctx, ptr, on_dev, _lib.DTYPE_CODES[dtype], nx, ny, gnz,
                                                       z0, z1, self.halo))
        self._need_alloc = True
        self.update_matrices()

    # ---- peer-memory composite ----
    def connect(self):
        """Exchange the IPC handles of the composite staging with every other rank of the process group."""
        self._check(self._lib.spv_comp_init(self._ctx, self.rank, self.world))
        if self.world > 1:
            buf = C.create_string_buffer(192)
            self._check(self._lib.spv_comp_export(self._ctx, buf, 192))
            handles = [None] * self.world
            self._dist.all_gather_object(handles, bytes(buf.raw), group=self.group)
            for r, h in enumerate(handles):
                if r != self.rank:
                    self._check(self._lib.spv_comp_import(self._ctx, r, h, len(h)))
            self._dist.barrier(group=self.group)
        self._connected = True

    @staticmethod
    def connect_local(renderers):
        """Several SlabMaxProjector(composite="peer") objects living in ONE process (one per rank, on one or more
        GPUs): wire their staging buffers together directly."""
        for r in renderers:
            r._check(r._lib.spv_comp_init(r._ctx, r.rank, r.world))
        for a in renderers:
            for b in renderers:
                if a is not b:
                    a._check(a._lib.spv_comp_import_local(a._ctx, b.rank, b._ctx))
            a._connected = True

    def enqueue_composite(self):
        """Enqueue one composited max projection (no host synchronisation): peers must enqueue theirs as well."""
        if self.alphaPow != 0:
            raise NotImplementedError("sort-last compositing needs alpha_pow == 0")
        if not self._connected:
            raise RuntimeError("SlabMaxProjector(composite='peer'): call connect() / connect_local() first")
        p = _lib.MipParams(self._box(), float(self.minVal), float(self.maxVal), float(self.gamma), 0.,
                           1, 0, int(self.max_steps), 0)
        self._check(self._lib.spv_render_mip_composite(self._ctx, C.byref(p)))

    def collect(self):
        """Wait for the enqueued composite and read output / output_alpha back."""
        self._check(self._lib.spv_comp_check(self._ctx))
        if self.readback_ranks is not None and self.rank not in self.readback_ranks:
            return  # the image is complete in this rank's device buffer (spv_device_ptr) but stays there
        flat, n = self._fetch(2)
        shape = (self.height, self.width)
        self.output = flat[:n].reshape(shape)
        self.output_alpha = flat[n:2 * n].reshape(shape)

    def resize(self, size):
        super(SlabMaxProjector, self).resize(size)
        self._connected = False  # the staging moved: connect() again


    # The pipelined / device-only entry points of VolumeRenderer call the single-context kernels directly; on a slab
    # context they would return this rank's windowed partial alone.  Every frame of a slab renderer goes through the
    # composite instead (the peers must make the same calls, as with render()).
    def render_sequence(self, modelViews, method="max_project", depth=2, iso_planes=7):
        if not hasattr(self, 'dataImg'):
            print("no data provided, set_data(data) before")
            return
        if method not in ("max_project", "iso_surface"):
            raise KeyError("method = '%s' not defined, valid: ['max_project', 'iso_surface']" % method)
        for M in modelViews:
            self.set_modelView(M)
            self.render(method=method)
            yield self

    def render_device_only(self, numParts=1, currentPart=0):
        """Enqueue one composited max projection without reading anything back (the peers must do the same)."""
        if numParts != 1 or self.alphaPow != 0:
            raise NotImplementedError("sort-last compositing needs alpha_pow == 0 and numParts == 1")
        if self.composite == "peer":
            self.enqueue_composite()
            return
        p = _lib.MipParams(self._box(), float(self.minVal), float(self.maxVal), float(self.gamma), 0.,
                           1, 0, int(self.max_steps), _lib.MIP_RAW_ONLY)
        self._check(self._lib.spv_render_mip(self._ctx, C.byref(p)))
        if self.world > 1 and self._dist.is_initialized():
            with self._torch.cuda.stream(self._stream):
                composite_max(self._raw_tensor(), self.group)
        self._check(self._lib.spv_mip_finish(self._ctx, C.byref(p)))

    def _raw_tensor(self):
        p = C.c_void_p()
        self._check(self._lib.spv_device_ptr(self._ctx, _lib.BUF_RAW, C.byref(p)))
        return self._torch.as_tensor(_DevBuffer(p.value, (self.height, self.width)),
                                     device="cuda:%d" % self.device)

    def _render_max_project(self, dtype=np.float32, numParts=1, currentPart=0):
        if self.alphaPow != 0 or numParts != 1:
            raise NotImplementedError("sort-last compositing needs alpha_pow == 0 and numParts == 1 "
                                      "(front-to-back attenuation is order dependent)")
        if self.composite == "peer":
            self.enqueue_composite()
            self.collect()
            return
        p = _lib.MipParams(self._box(), float(self.minVal), float(self.maxVal), float(self.gamma), 0.,
                           1, 0, int(self.max_steps), _lib.MIP_RAW_ONLY)
        torch = self._torch
        self._check(self._lib.spv_render_mip(self._ctx, C.byref(p)))
        if self.world > 1 and self._dist.is_initialized():
            with torch.cuda.stream(self._stream):
                composite_max(self._raw_tensor(), self.group)
        self._check(self._lib.spv_mip_finish(self._ctx, C.byref(p)))
        flat, n = self._fetch(2)
        shape = (self.height, self.width)
        self.output = flat[:n].reshape(shape)
        self.output_alpha = flat[n:2 * n].reshape(shape)

    # ---- sort-last iso surface ----
    def _iso_params(self, raw_only=False):
        return _lib.IsoParams(self._box(), float(self.maxVal / 2), float(self.gamma), int(self.max_steps),
                              float(self.occ_strength), int(self.occ_radius), int(self.occ_n_points),
                              _lib.ISO_RAW_ONLY if raw_only else 0)

    def _dev_tensor(self, which, shape, typestr):
        p = C.c_void_p()
        self._check(self._lib.spv_device_ptr(self._ctx, which, C.byref(p)))
        buf = _DevBuffer(p.value, shape)
        buf.__cuda_array_interface__["typestr"] = typestr
        return self._torch.as_tensor(buf, device="cuda:%d" % self.device)

    def iso_k_tensor(self):
        """int32 (2, H, W) view of the crossing-candidate planes (element-wise MIN over the ranks)."""
        return self._dev_tensor(_lib.BUF_KPLANES, (2, self.height, self.width), "<i4")

    def iso_planes_tensor(self):
        """float32 (7, H, W) view of [out | alpha | depth | occ | normals(3)] (element-wise SUM over the ranks)."""
        return self._dev_tensor(_lib.BUF_OUT, (7, self.height, self.width), "<f4")

    def iso_search(self):
        if self._parts:
            raise NotImplementedError("sort-last iso_surface renders one slab per rank")
        self._check(self._lib.spv_iso_slab_search(self._ctx, C.byref(self._iso_params())))

    def iso_resolve(self):
        self._check(self._lib.spv_iso_slab_resolve(self._ctx, C.byref(self._iso_params())))

    def iso_finish(self, raw_only=False):
        self._check(self._lib.spv_iso_slab_post(self._ctx, C.byref(self._iso_params(raw_only))))
        flat, n = self._fetch(7)
        self._check(self._lib.spv_iso_slab_check(self._ctx))
        shape = (self.height, self.width)
        self.output = flat[:n].reshape(shape)
        self.output_alpha = flat[n:2 * n].reshape(shape)
        self.output_depth = flat[2 * n:3 * n].reshape(shape)
        self.output_occlusion = flat[3 * n:4 * n].reshape(shape)
        self.output_normals = flat[4 * n:7 * n].reshape(shape + (3,))

    def enqueue_iso_composite(self, raw_only=False):
        """Enqueue one sort-last iso surface whose exchanges all run over peer memory (spv_render_iso_composite); no
        host synchronisation: the peers must enqueue theirs as well."""
        if not self._connected:
            raise RuntimeError("SlabMaxProjector(composite='peer'): call connect() / connect_local() first")
        if self._parts:
            raise NotImplementedError("sort-last iso_surface renders one slab per rank")
        self._check(self._lib.spv_render_iso_composite(self._ctx, C.byref(self._iso_params(raw_only))))

    PHASES = ("search", "wait_candidates", "min_redistribute", "wait_min", "resolve", "wait_resolve", "screen_passes",
              "band_gather_wait")

    def time_phases(self, on=True):
        self._check(self._lib.spv_set_tuning(self._ctx, 13, int(bool(on))))

    def last_phases_us(self):
        """{phase: microseconds} of the last enqueue_iso_composite that ran with time_phases(True)."""
        ms = (C.c_float * 8)()
        n = C.c_int()
        self._check(self._lib.spv_last_phases_ms(self._ctx, ms, 8, C.byref(n)))
        return dict((self.PHASES[i], 1e3 * ms[i]) for i in range(n.value))

    def collect_iso(self):
        """Wait for the enqueued iso composite and read all planes back."""
        self._check(self._lib.spv_comp_check(self._ctx))
        self._check(self._lib.spv_iso_slab_check(self._ctx))
        if self.readback_ranks is not None and self.rank not in self.readback_ranks:
            return
        flat, n = self._fetch(2)  # depth / normals / occlusion follow when they are first looked at, as on one GPU
        shape = (self.height, self.width)
        self.output = flat[:n].reshape(shape)
        self.output_alpha = flat[n:2 * n].reshape(shape)
        self._iso_pending = True

    def _render_isosurface(self, raw_only=False):
        """composite="peer": spv_render_iso_composite.  composite="nccl": search on every slab -> all-reduce(MIN) of the candidate sample indices -> the owner of each crossing
        resolves it -> all-reduce(SUM) assembles the planes -> post passes on every rank."""
        if self.composite == "peer":
            self.enqueue_iso_composite(raw_only)
            self.collect_iso()
            return
        torch, dist = self._torch, self._dist
        multi = self.world > 1 and dist.is_initialized()
        self.iso_search()
        if multi:
            with torch.cuda.stream(self._stream):
                dist.all_reduce(self.iso_k_tensor(), op=dist.ReduceOp.MIN, group=self.group)
        self.iso_resolve()
        if multi:
            with torch.cuda.stream(self._stream):
                dist.all_reduce(self.iso_planes_tensor(), op=dist.ReduceOp.SUM, group=self.group)
        self.iso_finish(raw_only)


class TimelapsePlayer(object):
    """Frame-sharded 3D+t playback: rank r plays frames r, r+world, ... of a (T, Nz, Ny, Nx) source;
    no data-path collective.  Replaces the reference's per-time-step full re-upload from pageable memory
    (DataModel -> GLWidget.dataModel_changed -> renderer.update_data, gui/glwidget.py:372-374).

    Two ways to hold the frames this rank owns:
      preload(source)   every owned frame becomes resident in HBM (its own texture array): playing a frame is
                        one render, nothing is uploaded -- 2 GiB per 512x1024x1024 uint16 frame in the z-paired layout
      play(source)      streamed: frame t+1 is uploaded while frame t is being read back; with a page-locked source
                        (spimagine_b200.pinned_empty) the upload runs at PCIe rate and never blocks the host
    """

    def __init__(self, size, rank=0, world=1, **kw):
        self.rank, self.world = rank, world
        self.size = size
        self._kw = dict(kw)
        self.rend = VolumeRenderer(size, **kw)
        # time points are rendered through the primary (z) copy, streamed or resident: a streamed one is rendered once
        # (no second layered copy per upload), and a time point looks the same whichever way it is played
        self.rend.set_view_copies("primary")
        self.resident = {}  # t -> VolumeRenderer holding frame t

    def my_frames(self, n_frames):
        return frames_for_rank(n_frames, self.rank, self.world)

    def close(self):
        for r in self.resident.values():
            r.close()
        self.resident = {}
        self.rend.close()

    # ---- resident playback ----
    def preload(self, source, frames=None, device_ptrs=False):
        """Make the owned frames resident.  source[t] -> (Nz, Ny, Nx) ndarray, or with device_ptrs=True a tuple
        (device pointer, shape, dtype) of a C-order device array on this GPU."""
        frames = self.my_frames(len(source)) if frames is None else frames
        for t in frames:
            r = VolumeRenderer(self.size, **self._kw)
            r.set_view_copies("primary")
            if device_ptrs:
                ptr, shape, dtype = source[t]
                r.set_data_device(ptr, shape, dtype)
                r.sync()
            else:
                r.set_data(source[t])
            self.resident[t] = r
        return frames

    def _configure(self, r, modelView, settings):
        for k, v in settings.items():
            getattr(r, "set_" + k)(v)
        if modelView is not None:
            r.set_modelView(modelView)

    def render_resident(self, t, modelView=None, method="max_project", **settings):
        """Render resident frame t.  settings: projection=, max_val=, min_val=, gamma=, units=, ... (set_* names)."""
        r = self.resident[t]
        self._configure(r, modelView, settings)
        r.render(method=method)
        return r

    # ---- streamed playback ----
    def play(self, source, modelViews=None, frames=None, pinned=False, method="max_project", **settings):
        """Generator over (t, renderer) for the owned frames of `source`; renderer.output* hold frame t.
        pinned=True: source[t] are page-locked arrays (spimagine_b200.pinned_empty) that stay untouched while the
        generator runs -- uploads are asynchronous at PCIe rate."""
        frames = self.my_frames(len(source)) if frames is None else frames
        r = self.rend
        first = not hasattr(r, "dataImg")
        for t in frames:
            vol = source[t]
            if first or tuple(vol.shape[::-1]) != r.dataImg.shape or vol.dtype != r.dataImg.dtype:
                r.set_data(vol)
                first = False
            else:
                r.update_data(vol, pinned=pinned)
            self._configure(r, None if modelViews is None else modelViews[t], settings)
            r.render(method=method)
            yield t, r

    def render_frames(self, source, modelViews=None, **render_kw):
        """source[t] -> (Nz, Ny, Nx) ndarray.  Returns {t: image} for the frames this rank owns."""
        out = {}
        for t, r in self.play(source, modelViews):
            if render_kw:
                r.render(**render_kw)
            out[t] = np.array(r.output)
        return out
