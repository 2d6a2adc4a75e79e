/*
 * spimcuda.h -- C ABI of libspimcuda.so, the B200 (sm_100a) replacement for the
 * device side of spimagine's volume renderer.
 *
 * Every entry point below stands in for one group of gputools / pyopencl calls
 * that the reference's spimagine/volumerender/volumerender.py makes (file:line
 * cited per function, relative to the reference tree).  Plain pointers and
 * sizes only; all functions return 0 on success or a negative SPV_E* /
 * positive cudaError_t code, never throw, never exit.  A context is not
 * thread-safe; distinct contexts are independent.
 */
#ifndef SPIMCUDA_H_
#define SPIMCUDA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPV_VERSION 100

#if defined(__GNUC__)
#define SPV_API __attribute__((visibility("default")))
#else
#define SPV_API
#endif

typedef struct spv_ctx spv_ctx;

/* volume texel types: VolumeRenderer.dtypes, volumerender.py:67 */
enum { SPV_F32 = 0, SPV_U16 = 1, SPV_U8 = 2 };

/* result buffers: VolumeRenderer.buf / buf_alpha / buf_depth / buf_normals /
 * buf_occlusion, volumerender.py:185-193.  SPV_BUF_RAW is new: the un-windowed
 * ray maximum that the sort-last composite reduces across GPUs. */
enum { SPV_BUF_OUT = 0, SPV_BUF_ALPHA = 1, SPV_BUF_DEPTH = 2, SPV_BUF_NORMALS = 3, SPV_BUF_OCC = 4, SPV_BUF_RAW = 5,
       SPV_BUF_KPLANES = 6 /* int32 [2][h][w]: first sample > iso / first sample <= iso (sort-last iso surface) */ };

/* sampler implementations */
enum {
  SPV_SAMPLER_TMU = 0,   /* hardware-filtered tex3D (9-bit weights), fma sample positions, brick skipping */
  SPV_SAMPLER_EXACT = 1  /* fp32 software trilinear from 8 point fetches, positions accumulated like the
                            reference loop: bit-comparable with the reference kernel text built for the host */
};

enum {
  SPV_EINVAL = -22,   /* bad argument */
  SPV_ENODATA = -61,  /* render before set_volume */
  SPV_ENOMEM = -12
};

/* ---- lifetime: VolumeRenderer.__init__/resize/reset_buffer, volumerender.py:72-134, 181-197 ---- */
SPV_API int spv_create(int device, int width, int height, spv_ctx **out);
SPV_API int spv_destroy(spv_ctx *ctx);
SPV_API int spv_resize(spv_ctx *ctx, int width, int height);
/* run on a caller-owned CUDA stream (cudaStream_t as void*); NULL restores the context's own stream */
SPV_API int spv_set_stream(spv_ctx *ctx, void *cuda_stream);
/* run on the stream `other` currently uses (several slab contexts of one GPU render in order on one stream) */
SPV_API int spv_share_stream(spv_ctx *ctx, spv_ctx *other);
SPV_API int spv_sync(spv_ctx *ctx);

/* ---- volume: set_shape / update_data, volumerender.py:260-294 (OCLImage.empty + write_array) ----
 * host: C-order (z,y,x) array of nz*ny*nx texels of `dtype`.  Allocates the 3-D array, uploads,
 * builds the min/max brick grid and the global min/max in the same pass. */
SPV_API int spv_set_volume(spv_ctx *ctx, const void *host, int dtype, int nx, int ny, int nz);
/* same shape and dtype as the current volume: re-upload only (the timelapse path, glwidget.py:372-374) */
SPV_API int spv_update_volume(spv_ctx *ctx, const void *host);
/* same, from PAGE-LOCKED host memory, without waiting: the transfer runs at PCIe rate on the context's stream and the
 * call returns at once; the buffer must stay untouched until spv_sync (or any synchronising call) returns.  This is
 * the streamed-timelapse path: the frame source fills pinned buffers (spv_host_alloc) ahead of the playback. */
SPV_API int spv_update_volume_async(spv_ctx *ctx, const void *pinned_host);
SPV_API int spv_host_alloc(size_t nbytes, void **host);  /* page-locked, usable from every device */
SPV_API int spv_host_free(void *host);
/* Host arrays of ANOTHER element type (replaces the host-side `data.astype(self.dtype)` of set_data / update_data,
 * volumerender.py:245-246, 290-291): the source bytes travel over PCIe as they are and are converted to `dtype`
 * texels on the device, chunk by chunk inside the ingest pipeline, with the semantics of the C cast numpy's astype
 * performs (float targets: round to nearest; integer targets: integers wrap, floats truncate towards zero and wrap
 * within the int32 range). */
enum { SPV_SRC_I8 = 0, SPV_SRC_U8 = 1, SPV_SRC_I16 = 2, SPV_SRC_U16 = 3, SPV_SRC_I32 = 4, SPV_SRC_U32 = 5,
       SPV_SRC_I64 = 6, SPV_SRC_U64 = 7, SPV_SRC_F16 = 8, SPV_SRC_F32 = 9, SPV_SRC_F64 = 10, SPV_SRC_BOOL = 11 };
SPV_API int spv_set_volume_from(spv_ctx *ctx, const void *host, int src_type, int dtype, int nx, int ny, int nz);
SPV_API int spv_update_volume_from(spv_ctx *ctx, const void *host, int src_type);
/* as above from a DEVICE pointer (C-order linear); used by frame sources that keep timepoints in HBM */
SPV_API int spv_set_volume_device(spv_ctx *ctx, const void *dev, int dtype, int nx, int ny, int nz);
/* update_data (volumerender.py:279-294) from a DEVICE array of the current shape and of element type src_type
 * (SPV_SRC_*): converted to the volume's texel type on the device with astype semantics where the types differ, then
 * stored into the resident array.  Enqueue-only on the context's stream; whatever produced `dev` must have finished.
 * This is how a filtered volume (spv_filter_*) reaches the renderer without crossing PCIe. */
SPV_API int spv_update_volume_device_from(spv_ctx *ctx, const void *dev, int src_type);
/* One z-slab of a larger volume for sort-last rendering (new).  The host/dev pointer
 * holds slices [z_lo, z_hi) of a global volume of gnz slices, where z_lo = max(z0-1,0) and
 * z_hi = min(z1+1,gnz) (one halo slice either side); the context then renders only the ray samples
 * whose trilinear footprint starts in slices [z0, z1) (the last slab also owns everything beyond). */
SPV_API int spv_set_volume_slab(spv_ctx *ctx, const void *host, int on_device, int dtype, int nx, int ny, int gnz, int z0,
                        int z1);
/* as spv_set_volume_slab with `halo` >= 1 slices either side: the pointer holds slices [max(z0-halo,0),
 * min(z1+halo,gnz)).  Sort-last iso surfaces need 2 * dt * gamma^2 * gnz + one ray step of halo (the 12-tap gradient
 * reaches +-2h along z, iso_kernel.cl:163-192); max projections need 1. */
SPV_API int spv_set_volume_slab_halo(spv_ctx *ctx, const void *src, int on_device, int dtype, int nx, int ny, int gnz,
                                     int z0, int z1, int halo);
/* global min/max of the resident volume (replaces GLWidget._get_min_max, gui/glwidget.py:328-344) */
SPV_API int spv_volume_minmax(spv_ctx *ctx, float *vmin, float *vmax);

/* -D SAMPLER_FILTER=..., volumerender.py:69-70, 146-147: 1 = linear, 0 = nearest */
SPV_API int spv_set_interp(spv_ctx *ctx, int linear);
SPV_API int spv_set_sampler(spv_ctx *ctx, int sampler);
/* what read_imageui + a LINEAR sampler means for integer volumes (undefined by OpenCL):
 * 1 = interpolate like a float image (default), 0 = nearest */
SPV_API int spv_set_int_filter(spv_ctx *ctx, int linear);
/* Storage layout of INTEGER volumes, effective at the next spv_set_volume*:
 *   SPV_LAYOUT_3D    3-D array, one hardware trilinear fetch per sample
 *   SPV_LAYOUT_ZPAIR (default) 2-D layered array of {v[z], v[z+1]} texel pairs: one hardware bilinear fetch plus an
 *                    fp32 lerp along z per sample -- half the texture-unit work, twice the memory, exact z weight.
 * float32 volumes always use SPV_LAYOUT_3D. */
enum { SPV_LAYOUT_3D = 0, SPV_LAYOUT_ZPAIR = 1 };
SPV_API int spv_set_layout(spv_ctx *ctx, int layout);
/* empty-space skipping on the min/max brick grids (texture-unit sampler; results are identical either way):
 * -1 = auto (default: on for iso surfaces, off for max projection, where it only pays on sparse volumes),
 * 0 = off, 1 = on for both */
SPV_API int spv_set_skipping(spv_ctx *ctx, int on);

/* invPBuf / invMBuf.write_array, volumerender.py:310-316: row-major float[16] each */
SPV_API int spv_set_matrices(spv_ctx *ctx, const float *invP, const float *invM);

/* ---- max projection: _render_max_project, volumerender.py:327-386 -> max_project_float/short ---- */
typedef struct {
  float box[6];      /* boxMin_x, boxMax_x, boxMin_y, boxMax_y, boxMin_z, boxMax_z */
  float min_val, max_val, gamma, alpha_pow;
  int num_parts, current_part;
  int max_steps;     /* -D maxSteps (config.__DEFAULTMAXSTEPS__ = 200) */
  int flags;         /* SPV_MIP_* */
} spv_mip_params;
enum {
  SPV_MIP_RAW_ONLY = 1, /* write only SPV_BUF_RAW (+ alpha): the per-slab partial of a sort-last render;
                           needs alpha_pow == 0 and num_parts == 1 */
  SPV_MIP_PUSH = 2      /* internal to spv_render_mip_composite: raw partials go to the band owners' staging */
};
SPV_API int spv_render_mip(spv_ctx *ctx, const spv_mip_params *p);
/* Render + read-back in one call (replaces run_kernel followed by the blocking buf.get() / buf_alpha.get() of
 * _render_max_project, volumerender.py:366-390).  The frame is cut into `bands` horizontal bands (<= 0: the
 * context's default, tuning knob 2); the rows of a finished band are copied into the selected slot's pinned staging
 * ([out | alpha]) while the rest renders.  Where the driver offers stream memory operations the whole frame is ONE
 * launch: its CTAs are dealt from the top and bottom rows inwards and count themselves into per-band counters that the
 * copy streams wait on (cuStreamWaitValue32); otherwise one launch per band.  wait != 0: returns when the frame is in host memory; wait == 0:
 * returns at once, collect with spv_wait_slot.  *host (may be NULL) = the slot's staging. */
SPV_API int spv_render_mip_to_host(spv_ctx *ctx, const spv_mip_params *p, int bands, int wait, float **host);
/* ---- several frames per launch (new; the record loop of a keyframe / rotation sequence:
 *      spimagine/gui/mainwidget.py renders frame after frame through _render_max_project, volumerender.py:327-390).
 *      n <= SPV_MAX_BATCH projections (plain or attenuated, one part) of the resident volume (integer in the paired layout, or float32) that share the
 *      projection, box, window and step count and differ in their model view: invM = n row-major float[16], the invP
 *      of spv_set_matrices.  ONE launch; its CTAs are dealt (tile row, frame, tile column), so the frames' CTAs of a
 *      tile row run together and share the volume in L2; every frame picks the layered copy of the volume (pairs along
 *      x, y or z; built on the device when first wanted, 4 bytes per voxel each) and the lane-to-pixel map under which
 *      a texture request stays inside one layer.  Results go to one of two sets of planes, [n][out | alpha], that
 *      alternate from call to call (*set = the one used); to_host != 0: the rows the projected box can touch also
 *      travel to the set's pinned planes behind the launch.  Enqueue-only: spv_batch_wait waits for a set and returns
 *      its planes (frame f: + f * 2 * width * height floats).  Pixel values equal spv_render_mip's with the same copy. */
#define SPV_MAX_BATCH 16
SPV_API int spv_render_mip_batch(spv_ctx *ctx, const spv_mip_params *p, const float *invM, int n, int to_host, int *set);
SPV_API int spv_batch_wait(spv_ctx *ctx, int set, float **host, float **dev, int *n_frames);
/* 1 if spv_render_mip_batch accepts these parameters on the resident volume and the context's settings, else 0 */
SPV_API int spv_mip_batch_possible(spv_ctx *ctx, const spv_mip_params *p);
/* layer axis (0 x, 1 y, 2 z) and lane map (0: 2x2-pixel quads, 1: 4x1 row quads, 2: 1x4 column quads) the last plain
 * projection used; -1 / -1 when it ran mip_fast_kernel (tuning knob 16 = 0, or a path the layered copies do not cover) */
SPV_API int spv_mip_axis_used(spv_ctx *ctx, int *axis, int *quad);
/* window + gamma of SPV_BUF_RAW into SPV_BUF_OUT after the cross-GPU max composite */
SPV_API int spv_mip_finish(spv_ctx *ctx, const spv_mip_params *p);

/* ---- sort-last composite over peer memory (new).  One context per GPU, one process per GPU (handles
 *      exchanged by the host program) or several contexts in one process.  The image is cut into `world` bands of
 *      rows, rank o owns band o.  spv_render_mip_composite = slab render whose raw partial maxima are stored
 *      straight into the band owners' staging (NVLink peer stores), arrival counters, the owner's max + window over
 *      its band stored into every rank's SPV_BUF_OUT, second counter round.  Enqueue-only: follow with spv_sync /
 *      spv_read*; the result equals the single-GPU render bit for bit on every rank. ---- */
SPV_API int spv_comp_init(spv_ctx *ctx, int rank, int world);              /* after spv_create / spv_resize */
SPV_API int spv_comp_export(spv_ctx *ctx, void *handles, size_t nbytes);   /* 192 bytes: 3 cudaIpcMemHandle_t */
SPV_API int spv_comp_import(spv_ctx *ctx, int peer, const void *handles, size_t nbytes); /* another process' export */
SPV_API int spv_comp_import_local(spv_ctx *ctx, int peer, spv_ctx *peer_ctx);            /* same process */
SPV_API int spv_render_mip_composite(spv_ctx *ctx, const spv_mip_params *p);
SPV_API int spv_comp_check(spv_ctx *ctx);  /* synchronises; -110 if a wait on a peer timed out (4 s) */
/* Several slabs of one volume on one GPU (each uploaded into its own context with spv_set_volume_slab): slab renders
 * of `ctx` also march the slabs of `others` (at most 3; same extent, dtype, layout and filter) -- one ray setup, one
 * launch, one partial.  The other contexts only hold data; they must outlive the renders.  n = 0 switches it off. */
SPV_API int spv_set_extra_slabs(spv_ctx *ctx, spv_ctx **others, int n);

/* ---- iso surface: _render_isosurface, volumerender.py:446-506
 *      iso_surface -> conv_vec_x/y(7) -> occlusion -> conv_x/y(5) -> shading ---- */
typedef struct {
  float box[6];
  float iso_val;     /* maxVal / 2, volumerender.py:463 */
  float gamma;
  int max_steps;
  float occ_strength;
  int occ_radius, occ_n_points;
  int flags;         /* SPV_ISO_* */
} spv_iso_params;
enum {
  SPV_ISO_RAW_ONLY = 1  /* the iso_surface kernel alone, no blur / occlusion / shading passes */
};
SPV_API int spv_render_iso(spv_ctx *ctx, const spv_iso_params *p);
/* Render + read-back of [out | alpha] in one call (replaces the run_kernel chain followed by the blocking buf.get() /
 * buf_alpha.get() of _render_isosurface, volumerender.py:499-506): the alpha plane is final once the march has run and
 * travels to the selected slot's pinned staging while the blur / occlusion / shading passes run; the value plane
 * follows.  depth / normals / occlusion stay on the device (spv_read_pinned fetches them).  wait / host as for
 * spv_render_mip_to_host. */
SPV_API int spv_render_iso_to_host(spv_ctx *ctx, const spv_iso_params *p, int wait, float **host);

/* ---- sort-last iso surface (new): every context holds one z-slab (+ halo) and the same camera.
 *   1. spv_iso_slab_search   per pixel over the samples the slab owns: SPV_BUF_KPLANES = {first k with s_k > iso,
 *                            first k with s_k <= iso} (INT_MAX: none)
 *   2. the caller reduces SPV_BUF_KPLANES element-wise with MIN over all ranks (ncclMin on int32)
 *   3. spv_iso_slab_resolve  the slab owning a pixel's crossing sample refines, takes the gradient and shades it into
 *                            the 7 planes starting at SPV_BUF_OUT ([out|alpha|depth|occ = 0|normals(3)]); all other
 *                            ranks write zeros there
 *   4. the caller reduces those 7*w*h floats element-wise with SUM over all ranks (x + 0 = x: bit-exact)
 *   5. spv_iso_slab_post     depth = INFINITY where nothing was hit, then the blur / occlusion / shading passes
 * The result equals spv_render_iso on the whole volume bit for bit.  spv_iso_slab_check synchronises and reports a
 * halo that was too small for the gradient taps. ---- */
SPV_API int spv_iso_slab_search(spv_ctx *ctx, const spv_iso_params *p);
SPV_API int spv_iso_slab_resolve(spv_ctx *ctx, const spv_iso_params *p);
SPV_API int spv_iso_slab_post(spv_ctx *ctx, const spv_iso_params *p);
SPV_API int spv_iso_slab_check(spv_ctx *ctx);
/* The same result with every exchange over peer memory instead of the two caller-side reductions (needs spv_comp_init +
 * the handle exchange, slot 0, one slab per context): the search stores the candidates of image band o into owner o's
 * staging, the owner takes the MIN and stores the band's final candidates into every rank, the rank owning a pixel's
 * crossing stores the finished pixel into every rank's planes; arrival counters in between; then the post passes.
 * Enqueue-only: follow with spv_comp_check + spv_iso_slab_check, then read. */
SPV_API int spv_render_iso_composite(spv_ctx *ctx, const spv_iso_params *p);

/* ---- results: buf.get(), volumerender.py:388-390, 499-506 ---- */
/* copies n floats (n = w*h, or 3*w*h for normals) to host memory; synchronises */
SPV_API int spv_read(spv_ctx *ctx, int which, float *host_dst, size_t n);
/* all MIP results (out, alpha) or iso results in ONE device->host transfer into pinned staging, then
 * scattered to the given host pointers (any may be NULL) */
SPV_API int spv_read_many(spv_ctx *ctx, float *out, float *alpha, float *depth, float *normals, float *occ);
/* zero-copy variant: one device->host transfer of the first `planes` planes of [out | alpha | depth | occ |
 * normals(3)] into the context's pinned staging buffer; *host points at it (valid until the next read/resize).
 * With planes <= 2 and a finished frame in the slot (max projection, composite, iso surface) only the rectangle the
 * projected box can touch is transferred (tuning knob 9): every pixel outside it is a miss, and the staging holds the
 * miss values there already (out 0; alpha 0, or -1 for float32 max projections) -- the staging is read-only for callers. */
SPV_API int spv_read_pinned(spv_ctx *ctx, int planes, float **host);
SPV_API int spv_device_ptr(spv_ctx *ctx, int which, void **dev_ptr);

/* ---- display hand-off (replaces buf.get() + the glTexImage2D re-upload + the LUT look-up of the fragment shader:
 *      volumerender.py:388-390, gui/gui_utils.py:121-162, gui/glwidget.py:412-444, gui/shaders/texture.frag:8-38).
 *      spv_set_lut: the colour map, n RGB triples in [0,1] (what GLWidget.set_colormap uploads as texture_LUT).
 *      spv_read_rgba8: one device pass turns the current value plane into packed RGBA8 exactly as texture.frag would
 *      shade it (rgb = LUT(v) in black mode / LUT(1-v) otherwise, linear LUT filtering, a = v, all-zero where the
 *      alpha plane is negative) and copies the h*w*4 bytes to host_dst: a quarter of the bytes of the float planes. */
SPV_API int spv_set_lut(spv_ctx *ctx, const float *rgb, int n);
SPV_API int spv_read_rgba8(spv_ctx *ctx, int mode_black, unsigned char *host_dst, size_t nbytes);

/* ---- pipelined sequences (new; replaces the blocking buf.get() per frame of the GUI spin / keyframe loops,
 *      gui/glwidget.py:636-692, volumerender.py:388-390): two output slots, each a full set of device result
 *      buffers plus pinned staging.  Frame i renders into slot i&1 while frame i-1 is still on its way to the host. ---- */
/* result buffers that subsequent renders write and reads fetch (0 or 1; slot 1 is allocated on first use).  A render
 * into a slot waits on the device for an asynchronous read of that slot that is still in flight. */
SPV_API int spv_select_slot(spv_ctx *ctx, int slot);
/* enqueue one device->host transfer of the first `planes` planes of the selected slot on the context's copy stream,
 * ordered after everything enqueued so far on the render stream; returns immediately (clipped like spv_read_pinned) */
SPV_API int spv_read_pinned_async(spv_ctx *ctx, int planes);
/* block until the last asynchronous read of `slot` has landed; *host = its pinned staging (valid until the next
 * read of that slot or a resize) */
SPV_API int spv_wait_slot(spv_ctx *ctx, int slot, float **host);

/* ---- volume filters: the image-processor chain between the data model and update_data
 *      (gui/mainwidget.py:455-465).  spv_filter_convolve_sep3 replaces gputools.convolve_sep3(data, hx, hy, hz) as
 *      BlurProcessor.apply / BlurXYZProcessor.apply call it (models/imageprocessor.py:47-71): three float32 passes
 *      (x, then y, then z), each out[i] = sum_ht h[ht] * in[i + Nh/2 - ht] over the taps that stay inside the volume,
 *      accumulated in ascending tap order with one fused multiply-add per tap.  gputools is not vendored in the
 *      reference tree: the algorithm is restated from its published kernel (parity unpinned).
 *      A filter object owns a stream and two float32 work volumes on `device`; it is not thread-safe. ---- */
typedef struct spv_filter spv_filter;
SPV_API int spv_filter_create(int device, spv_filter **out);
SPV_API int spv_filter_destroy(spv_filter *f);
/* the volume the next convolution reads: C-order (z,y,x), element type SPV_SRC_*.  Host sources are copied (the
 * pointer is borrowed for the call only); device sources of type float32 / uint16 / uint8 are read in place by the
 * next convolution and must stay valid until it has run.  Other types become float32 first (gputools: astype). */
SPV_API int spv_filter_load(spv_filter *f, const void *src, int on_device, int src_type, int nx, int ny, int nz);
/* convolves the loaded volume -- or, when called again, the previous result (a processor chain) */
SPV_API int spv_filter_convolve_sep3(spv_filter *f, const float *hx, int nhx, const float *hy, int nhy, const float *hz,
                                     int nhz);
SPV_API int spv_filter_sync(spv_filter *f);
/* the float32 result: in device memory (valid until the next load / convolution of this filter) or copied to the host */
SPV_API int spv_filter_result_device(spv_filter *f, float **dev);
SPV_API int spv_filter_read(spv_filter *f, float *host_dst, size_t n);
SPV_API int spv_filter_last_ms(spv_filter *f, float *ms);  /* device time of the last convolution (three passes) */
/* device time of each kernel of the last convolution: ms[0..2] = x, y, z pass (passes = 3), or fused x + y, z, 0
 * (passes = 2); `passes` may be NULL */
SPV_API int spv_filter_last_pass_ms(spv_filter *f, float ms[3], int *passes);
SPV_API const char *spv_filter_last_error(spv_filter *f);  /* f may be NULL: last create error */
/* knob 0: 1 = the x and y pass run as one kernel where both tap counts fall into the same size class of at most 31
 * taps (one float32 round trip of the volume less; measured slower than the three passes so far), 0 = three passes
 * (default); knob 1: variant of the y / z passes (process-wide): 1 = automatic (four columns per thread where rows
 * are multiples of 16 bytes, two where of 8 bytes, else one), 16 / 32 = one column, that many outputs per thread,
 * 1602 / 1604 = two / four columns per thread; knob 2 (process-wide): the x pass works on row pairs through a
 * software-pipelined tile loop (2, default), on row pairs with loads at the top of every tile (1), on single rows (0).
 * The variants with several columns / row pairs update two outputs per instruction (packed fma.rn.f32x2 of sm_100,
 * each half an IEEE fused multiply-add); results are identical bit for bit under every knob */
SPV_API int spv_filter_set_tuning(spv_filter *f, int knob, int value);
SPV_API int spv_filter_launch_count(spv_filter *f, unsigned long long *n);  /* kernels launched by this filter so far */

/* ---- diagnostics ---- */
SPV_API int spv_last_timing_ms(spv_ctx *ctx, float *ms);            /* device time of the last render call */
SPV_API int spv_last_stats(spv_ctx *ctx, unsigned long long *v, int n); /* [hit rays, texture samples issued] of the
                                                                last render when stats were enabled */
SPV_API int spv_enable_stats(spv_ctx *ctx, int on);
/* performance knobs that never change results; knob 0 = CTA shape / occupancy target of the max-projection kernel,
 * knob 1 = persistent CTAs pulling tiles from a counter (1) or one CTA per tile (0), knob 2 = default band count of
 * spv_render_mip_to_host, knob 3 = spv_render_mip_to_host stores straight into pinned host memory, knob 4 = warps that share a
 * ray of the iso-surface search, each taking a segment of its samples (1, 2 or 4), knob 5 = its CTAs are dealt from the
 * image centre outwards (1, default) or row by row (0), knob 6 = resident CTAs per SM of the occlusion queue kernel,
 * knob 7 = copy streams the band copies of spv_render_mip_to_host alternate between (1 or 2), knob 8 = order in which
 * the one-launch path of spv_render_mip_to_host deals its tile rows: 0 = from the top and bottom edges inwards, 1 = the rows
 * the projected box cannot touch first, then the box's rows top to bottom (default), knob 9 = with knob 8 = 1, rows
 * outside the hull of the projected box corners (+ 9 pixels) are not copied by spv_render_mip_to_host: every ray there
 * misses, and the pinned staging rows already hold the miss values (out 0, alpha 0 / -1), which the library keeps
 * track of per output slot (1, default; 0 = copy every row).  The staging memory must be treated as read-only.
 * knob 10 = tiles of every 8 (fixed pattern) that the software-sampled max projection (spv_set_mip_path) hands to the
 * texture unit (hybrid; changes which sampler a pixel gets, deterministically), knob 11 = its box / ring geometry (0..4),
 * knob 12 = sort-last iso frames run the screen-space passes on the rank's own band of rows and exchange the finished
 * bands (1) or on the whole image on every rank (0, default: the exchange costs more than the passes save), knob 13 = record CUDA events at the phase boundaries of
 * sort-last iso frames (spv_last_phases_ms; off by default), knob 14 = device-only iso-surface renders (spv_render_iso) put
 * their screen-space passes on a second stream, so that the search of the next frame -- rendered into the other output
 * slot, spv_select_slot -- runs beside them; reads through this library wait for the passes by themselves, a caller that
 * takes spv_device_ptr must call spv_stream_join or spv_sync first (off by default; render_sequence switches it on),
 * knob 15 = the same for plain max projections into output slot 1 (a frame starts in the tail of the one before),
 * knob 16 = layered copies max projections sample (never changes a result by more than the
 * texture unit's weight rounding; the hit mask and alpha plane never change): 1 (default) = per frame the copy with pairs
 * along x, y or z and the lane-to-pixel map under which a texture request stays inside one layer, chosen from the camera
 * alone (the x / y copies -- for float32 volumes, whose array is 3-D, the z copy as well -- are built on the device when
 * first wanted, 4 bytes per voxel each, 8 for float32), 2 = the primary z copy only (float32: mip_fast_kernel),
 * 0 = mip_fast_kernel on the z copy (round 1's path), 10 + 3 * axis + map = forced (tests),
 * knob 17 = the ambient-occlusion pass reads the pixel offsets of its taps from a table built once per (image size,
 * radius, tap count) -- 64 bytes per pixel and 32 taps -- and gathers depths from a shared-memory tile (1, default)
 * instead of hashing every tap in every frame (0); same result bit for bit,
 * knob 18 = spv_read_pinned_async of a clipped rectangle (output, alpha) first moves it into a device staging buffer
 * (one small kernel), after which the slot's planes are free for the next render while the rectangle crosses the host
 * link (1, default), or copies straight out of the slot, which then stays busy for the length of the copy (0),
 * knob 20 = iso-surface frames compute the shading in the epilogue of the occlusion blur (1, default: one launch and one
 * pass over the occlusion plane less) or in a launch of its own (0); same result bit for bit. */
SPV_API int spv_set_tuning(spv_ctx *ctx, int knob, int value);
/* Which kernel family renders plain (alpha_pow == 0, num_parts == 1) max projections of uint16 volumes
 * (max_project_short, volume_kernel.cl:270-345):
 *   SPV_MIP_PATH_TMU   one hardware-filtered texture fetch per sample (default)
 *   SPV_MIP_PATH_SMEM  ray-segment slabs of a linear uint16 copy of the volume staged in shared memory by TMA box loads,
 *                      fp32 software trilinear sampling with exact weights (spv_mip_smem.cu); samples outside the staged
 *                      boxes, other element types, slab contexts, multi-pass and raw renders use the TMU path.  Costs three
 *                      permuted linear copies of the volume (6 bytes per voxel), built on first use after an upload.
 * Both are within north_star's 1e-3 of the OpenCL sampler; they differ from each other by the texture unit's 8-bit weight
 * quantisation.  spv_mip_path_used reports what the last max projection ran on. */
enum { SPV_MIP_PATH_TMU = 0, SPV_MIP_PATH_SMEM = 1 };
/* the render stream waits for screen-space passes running beside it (tuning knob 14); asynchronous */
SPV_API int spv_stream_join(spv_ctx *ctx);
SPV_API int spv_set_mip_path(spv_ctx *ctx, int path);
SPV_API int spv_mip_path_used(spv_ctx *ctx, int *path);
SPV_API const char *spv_last_error(spv_ctx *ctx);                   /* ctx may be NULL: last create error */
SPV_API int spv_version(void);
/* out[i] = the current sampler's value at normalised position pos[3i..3i+2] (what read_imagef(volume, sampler,
 * pos).x is to the reference kernels): lets tests check the sampler in isolation */
SPV_API int spv_sample_points(spv_ctx *ctx, const float *host_pos, int n, float *host_out);
/* roofline calibration: measured rate (samples/s) of independent, cache-resident filtered fetches of the
 * resident volume's format with the current interpolation mode */
SPV_API int spv_texrate_probe(spv_ctx *ctx, int iters, double *samples_per_s);
/* the same probe with a given footprint: lane (lx, ly) of a warp's 8x4 tile fetches at base + lx*a + ly*b + j*m,
 * j = 0..15, with vec9 = {a, b, m} in texels (x, y, z).  All warps walk the same small region (every fetch hits L1):
 * the rate the texture unit can deliver for the ray and sample spacing of a given camera */
SPV_API int spv_texrate_probe_footprint(spv_ctx *ctx, int iters, const float *vec9, double *samples_per_s);
/* Device time of each phase of the last spv_render_iso_composite that ran with spv_set_tuning(ctx, 13, 1), in order: search,
 * wait for the peers' candidates, MIN + redistribution, wait, resolve, wait, screen-space passes, band gather + wait
 * (the last one only with more than one rank).  *count = phases written. */
SPV_API int spv_last_phases_ms(spv_ctx *ctx, float *ms, int n, int *count);
SPV_API int spv_launch_count(spv_ctx *ctx, unsigned long long *n);  /* kernels launched by this context so far */
/* Result bytes this context has enqueued for device -> host copies so far (what `.get()` of the result buffers moves in
 * the reference, volumerender.py:388-390, 499-506): rows that cannot hold a hit are not copied, so a frame
 * moves fewer than 2 * W * H * 4 bytes. */
SPV_API int spv_d2h_bytes(spv_ctx *ctx, unsigned long long *n);

#ifdef __cplusplus
}
#endif
#endif /* SPIMCUDA_H_ */
