// spv_api.cu -- the C ABI of libspimcuda.so (include/spimcuda.h): context, volume residency, render dispatch,
// result read-back.  Device side of spimagine/volumerender/volumerender.py; see the header for the reference
// call each entry point replaces.
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <thread>

#include <cuda.h>

#include "spv_kernels.h"

using namespace spv;

struct spv_ctx {
  int device = 0;
  int width = 0, height = 0;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool timed = false;
  // volume
  cudaArray_t arr = nullptr;
  cudaTextureObject_t tex_lin = 0, tex_near = 0, tex_pt = 0;
  int dtype = -1, nx = 0, ny = 0, gnz = 0, local_nz = 0, z_lo = 0, z0 = 0, z1 = 0;
  bool slab = false;
  int layout = LAYOUT_3D;        // of the resident array
  int want_layout = LAYOUT_ZPAIR; // for integer volumes (spv_set_layout)
  void *d_stage = nullptr;       // ingest staging on the device: [paired texels | linear chunk x 2]
  size_t stage_bytes = 0;
  void *d_conv = nullptr;        // texels converted from a device array of another element type (spv_update_volume_device_from)
  size_t conv_bytes = 0;
  size_t stage_sig[3] = {0, 0, 0};  // partition of d_stage used by the last upload
  char *h_ring = nullptr;        // page-locked ring (2 chunks) for pageable sources
  size_t ring_bytes = 0;
  cudaEvent_t ev_up_begin = nullptr, ev_h2d_done[2] = {nullptr, nullptr}, ev_consumed[2] = {nullptr, nullptr};
  bool ring_used[2] = {false, false}, lin_used[2] = {false, false};
  float2 *bricks = nullptr, *coarse = nullptr, *top = nullptr;
  int gx = 0, gy = 0, gz = 0, cgx = 0, cgy = 0, cgz = 0;
  float *d_minmax = nullptr;
  float h_minmax[2] = {0.f, 0.f};
  bool minmax_valid = false;
  bool bricks_valid = false;  // the min/max grids are built on first use after an upload (iso, skipping, min/max query)
  // settings
  int linear = 1, sampler = SPV_SAMPLER_TMU, int_filter = 1, skipping = -1, stats_on = 0, tile_variant = 0, persistent = 0;
  int bands = 0;  // default band count of spv_render_mip_to_host (knob 2); 0 = 12 where the copy stream can wait on the
                  // band counters (one launch per frame), else 2 (one launch per band) -- profiles/r01_exp_e2e.txt
  int iso_segments = 1;   // tuning knob 4 (measured on configs[2]: 1 -> 97 us, 2 -> 110 us, 4 -> 178 us)
  int iso_centre_out = 1; // tuning knob 5
  int row_mode = 1;     // spv_render_mip_to_host, one-launch path: order of the tile rows (tuning knob 8, MipArgs::row_mode)
  int clip_copies = 1;  // spv_render_mip_to_host, one-launch path: rows the box cannot project to are not copied
                        // (tuning knob 9); their staging rows hold the miss values already
  // per output slot: rows [dirty_lo, dirty_hi) of the pinned out / alpha staging may differ from the miss values
  // (out 0, alpha clean_alpha); every other row holds them
  int dirty_lo[2] = {0, 0}, dirty_hi[2] = {0, 0};
  int dirty_x0[2] = {0, 0}, dirty_x1[2] = {0, 0};  // ... and, of those rows, the columns [dirty_x0, dirty_x1)
  float clean_alpha[2] = {0.f, 0.f};
  int direct_host = 0;  // spv_render_mip_to_host: the kernel stores straight into the pinned staging (tuning knob 3)
  unsigned *d_tile_counter = nullptr;
  unsigned *d_band_done = nullptr;   // [MAX_BANDS] CTAs finished per band, counting up across frames (never reset)
  unsigned band_expect[64] = {0};    // value band b's counter reaches when the current frame's band is complete
  unsigned char *d_tile_hit = nullptr;  // per 8x4 warp tile: holds an iso-surface pixel
  unsigned *d_occ_queue = nullptr;      // occlusion work queue (launch_occlusion), per image size
  unsigned occ_frame = 0;
  int sms = 0;
  float *d_lut = nullptr;               // colour map of the display pass, n_lut RGB triples
  int n_lut = 0;
  unsigned char *d_rgba = nullptr, *h_rgba = nullptr;  // packed display image (device, pinned host), per image size
  float4 *d_taps = nullptr;             // occlusion tap table (launch_occ_taps), valid for taps_n taps
  int taps_n = 0;
  void *d_occ_table = nullptr;          // pixel offsets of every tap of every pixel (launch_occ_table), for occ_key
  int occ_key[4] = {0, 0, -1, 0};       // width, height, radius, n_points the table was built for
  int occ_table_on = 1;                 // tuning knob 17
  int stage_reads = 1;                  // tuning knob 18: spv_read_pinned_async frees the slot through a device staging copy
  int fuse_shading = 1;                 // tuning knob 20: iso frames shade in the epilogue of the occlusion blur
  Camera cam;
  // result buffers: one allocation  [out | alpha | depth | occ | normals(3) | raw | tmp | tmp_vec(3)]
  float *dbuf = nullptr;
  float *hpin = nullptr;  // pinned staging for [out | alpha | depth | occ | normals(3)]
  // two output slots for pipelined sequences (spv_select_slot): dbuf / hpin alias the selected one; slot 1 is
  // allocated on first use.  Asynchronous reads run on copy_stream, ordered against the renders by events.
  float *dbuf_s[2] = {nullptr, nullptr};
  float *hpin_s[2] = {nullptr, nullptr};
  int slot = 0;
  cudaStream_t copy_stream = nullptr, copy_stream2 = nullptr;  // band copies alternate between the two
  cudaEvent_t ev_copy2 = nullptr;
  int copy_streams = 2;  // tuning knob 7 (measured: 267 us per synchronous C2 frame with 2, 280 with 1)
  cudaEvent_t ev_rendered[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
  bool copy_pending[2] = {false, false};
  // spv_read_pinned_async of a clipped rectangle first moves it into a device staging buffer (a few us), so that the
  // slot's planes are free for the next render long before the rectangle has crossed the host link (~65 us)
  float *dstage_s[2] = {nullptr, nullptr};  // [out | alpha] per slot, allocated when first used
  cudaEvent_t ev_freed[2] = {nullptr, nullptr};
  bool freed_by_stage[2] = {false, false};  // the slot's last asynchronous read went through dstage: renders wait for ev_freed
  // sort-last composite over peer memory (spv_comp_*): I own image band comp_rank
  int comp_rank = -1, comp_world = 0, comp_band_rows = 0;
  float *comp_part = nullptr;      // [2 parities][world][band_rows * width] floats (max-projection partials), then
                                   // [2 parities][world][2][band_rows * width] ints (iso-surface k1 / k0 candidates)
  unsigned *comp_flags = nullptr;  // [COMP_PHASES][MAX_WORLD] arrival counters, written by the peers
  unsigned *comp_err = nullptr;
  float *peer_part[MAX_WORLD] = {nullptr};
  unsigned *peer_flags[MAX_WORLD] = {nullptr};
  float *peer_out[MAX_WORLD] = {nullptr};
  void *ipc_opened[MAX_WORLD][3] = {{nullptr}};
  unsigned comp_frame = 0;
  unsigned *d_iso_err = nullptr;  // sort-last iso surface: raised when a slab's halo does not cover the gradient taps
  spv_ctx *extra[MAX_EXTRA_SLABS] = {nullptr};  // spv_set_extra_slabs: contexts whose slabs my slab renders march too
  int n_extra = 0;
  int last_method = 0;    // 0 = mip, 1 = iso
  // per output slot: camera, box and alpha miss value of the finished frame it holds (valid), for the clipped read-back
  // of out + alpha (spv_read_pinned, spv_read_pinned_async): pixels outside the projected box are misses
  // full: every pixel of the slot's device planes outside the rectangle holds the miss values as well (iso frames with
  // their screen-space passes), so a copy may cover more than the rectangle
  struct SlotClip { bool valid = false; Camera cam; float box[6]; float miss_alpha = 0.f; bool full = false; };
  SlotClip clip_s[2];
  unsigned long long *d_stats = nullptr;
  unsigned long long h_stats[40] = {0};  // hit rays, samples fetched, longest warp / sum over warps (cycles; iso search)
  unsigned long long launches = 0;
  unsigned long long d2h_bytes = 0;
  // software-sampled max projection (spv_set_mip_path, spv_mip_smem.cu): permuted linear uint16 copies of the resident
  // volume (slowest axis x, y, z) and the tensor maps the kernel's TMA box loads go through
  int mip_path = 0, last_mip_path = 0;
  cudaEvent_t ev_ph[9] = {nullptr};  // phase boundaries of the last sort-last iso frame (recorded while statistics are on)
  int n_ph = 0;
  // iso surface: the screen-space passes of the frame in one output slot may run on post_stream while the search of
  // the next frame (other slot) runs on `stream` (tuning knob 14); tile flags per slot
  int iso_overlap = 0;
  int mip_overlap = 0;   // tuning knob 15: plain max projections into output slot 1 run on post_stream, beside slot 0's
  cudaEvent_t ev_uploaded = nullptr;  // fork point: what the render stream had enqueued when post_stream last caught up
  unsigned long long upload_seq = 1, side_seq = 0;  // uploads on the render stream / the last one post_stream waited for
  cudaStream_t post_stream = nullptr;
  cudaEvent_t ev_searched[2] = {nullptr, nullptr}, ev_posted[2] = {nullptr, nullptr};
  bool post_pending[2] = {false, false};
  unsigned char *d_tile_hit_s[2] = {nullptr, nullptr};
  int time_phases = 0;       // tuning knob 13: record the phase boundaries of sort-last iso frames (spv_last_phases_ms)
  int iso_post_sharded = 0;  // tuning knob 12: sort-last iso frames run the screen-space passes on the rank's own band
                             // (measured slower than every rank doing the whole image: profiles/r02_exp_iso_sortlast.txt)
  int smem_tex_of8 = 0;  // tuning knob 10: tiles (of every 8) the software-sampled kernel hands to the texture unit
  void *d_lin[3] = {nullptr, nullptr, nullptr};
  bool lin_valid = false;
  CUtensorMap tmaps[8][3];  // per box geometry (tuning knob 11)
  int smem_cfg = 0;  // result bytes enqueued for device -> host copies so far (spv_d2h_bytes)
  // view-aligned layered copies (spv_mip_axis.cu): pairs along x (0) and y (1); the z copy is `arr` itself.  Built on the
  // render stream when a frame first wants them, rebuilt after an upload (axis_seq is the upload they were built from).
  // (float32 volumes live in a 3-D array: for them the z copy [2] is a layered pair copy of its own as well)
  cudaArray_t axis_arr[3] = {nullptr, nullptr, nullptr};
  cudaTextureObject_t axis_tex[3] = {0, 0, 0};
  cudaSurfaceObject_t axis_surf[3] = {0, 0, 0};
  unsigned long long axis_seq[3] = {0, 0, 0};
  bool axis_failed[3] = {false, false, false};  // the copy could not be allocated: not tried again for this volume
  float batch_miss_alpha = 0.f;  // what the alpha planes of the batch sets hold outside the tracked rectangles
  int axis_mode = 1;       // tuning knob 16: 0 = off (mip_fast_kernel), 1 = per-frame choice among the three copies (the x / y
                           // copies are built when a frame first wants them), 2 = per-frame choice of the lane map on the
                           // primary z copy only (no second copy: streamed time points), 10 + 3 * axis + quad = forced
  int last_axis = -1, last_quad = -1;  // what the last plain projection used (-1: mip_fast_kernel)
  // multi-frame launches (spv_render_mip_batch): two sets of [cap][out | alpha] planes, device and pinned
  float *d_batch[2] = {nullptr, nullptr}, *h_batch[2] = {nullptr, nullptr};
  int batch_cap = 0, batch_set = 1, batch_n[2] = {0, 0};
  cudaEvent_t ev_batch_rendered[2] = {nullptr, nullptr}, ev_batch_copied[2] = {nullptr, nullptr};
  bool batch_copy_pending[2] = {false, false}, batch_render_pending[2] = {false, false};
  // pixel rectangle [x0, x1) x [y0, y1) of a frame's planes that may hold hits, per set and frame: pinned (h) and device (d)
  struct Rect { int x0, x1, y0, y1; };
  Rect batch_h_dirty[2][MAX_BATCH], batch_d_dirty[2][MAX_BATCH];
  std::string err;

  size_t n() const { return (size_t)width * height; }
  float *out() const { return dbuf; }
  float *alpha() const { return dbuf + n(); }
  float *depth() const { return dbuf + 2 * n(); }
  float *occ() const { return dbuf + 3 * n(); }
  float *normals() const { return dbuf + 4 * n(); }
  float *raw() const { return dbuf + 7 * n(); }
  float *tmp() const { return dbuf + 8 * n(); }
  float *tmp_vec() const { return dbuf + 9 * n(); }
};

// arrival-counter phases of the peer composites: 0, 1 max projection; 2, 3, 4 iso surface
constexpr int COMP_PHASES = 6;  // + 5: the finished bands of a sort-last iso frame have reached every rank

static thread_local std::string g_create_err;

static int fail(spv_ctx *c, int code, const char *what) {
  char b[512];
  snprintf(b, sizeof b, "%s (code %d)", what, code);
  if (c) c->err = b; else g_create_err = b;
  return code;
}
static int cufail(spv_ctx *c, cudaError_t e, const char *where) {
  char b[512];
  snprintf(b, sizeof b, "%s: %s", where, cudaGetErrorString(e));
  if (c) c->err = b; else g_create_err = b;
  return (int)e;
}
#define CU(call)                                              \
  do {                                                        \
    cudaError_t e_ = (call);                                  \
    if (e_ != cudaSuccess) return cufail(ctx, e_, #call);     \
  } while (0)
#define BIND()                                                \
  do {                                                        \
    if (!ctx) return fail(nullptr, SPV_EINVAL, "null ctx");   \
    CU(cudaSetDevice(ctx->device));                           \
  } while (0)

static void free_comp(spv_ctx *c) {
  for (int r = 0; r < MAX_WORLD; ++r) {
    for (int k = 0; k < 3; ++k) {
      if (c->ipc_opened[r][k]) cudaIpcCloseMemHandle(c->ipc_opened[r][k]);
      c->ipc_opened[r][k] = nullptr;
    }
    c->peer_part[r] = nullptr;
    c->peer_flags[r] = nullptr;
    c->peer_out[r] = nullptr;
  }
  if (c->comp_part) cudaFree(c->comp_part);
  if (c->comp_flags) cudaFree(c->comp_flags);
  if (c->comp_err) cudaFree(c->comp_err);
  c->comp_part = nullptr;
  c->comp_flags = nullptr;
  c->comp_err = nullptr;
  c->comp_rank = -1;
  c->comp_world = 0;
  c->comp_frame = 0;
}

static void free_axis(spv_ctx *c) {
  for (int d = 0; d < 3; ++d) {
    if (c->axis_tex[d]) cudaDestroyTextureObject(c->axis_tex[d]);
    if (c->axis_surf[d]) cudaDestroySurfaceObject(c->axis_surf[d]);
    if (c->axis_arr[d]) cudaFreeArray(c->axis_arr[d]);
    c->axis_tex[d] = 0;
    c->axis_surf[d] = 0;
    c->axis_arr[d] = nullptr;
    c->axis_seq[d] = 0;
    c->axis_failed[d] = false;
  }
}
static void free_batch(spv_ctx *c) {
  for (int s = 0; s < 2; ++s) {
    if (c->d_batch[s]) cudaFree(c->d_batch[s]);
    if (c->h_batch[s]) cudaFreeHost(c->h_batch[s]);
    c->d_batch[s] = c->h_batch[s] = nullptr;
    c->batch_copy_pending[s] = c->batch_render_pending[s] = false;
    c->batch_n[s] = 0;
  }
  c->batch_cap = 0;
}
static void free_buffers(spv_ctx *c) {
  free_comp(c);  // the peers' pointers into these buffers die with them: re-run spv_comp_init + exchange after a resize
  if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
  if (c->copy_stream2) cudaStreamSynchronize(c->copy_stream2);
  free_batch(c);
  for (int s = 0; s < 2; ++s) {
    if (c->dbuf_s[s]) cudaFree(c->dbuf_s[s]);
    if (c->hpin_s[s]) cudaFreeHost(c->hpin_s[s]);
    c->dbuf_s[s] = nullptr;
    c->hpin_s[s] = nullptr;
    c->copy_pending[s] = false;
    if (c->dstage_s[s]) cudaFree(c->dstage_s[s]);
    c->dstage_s[s] = nullptr;
    c->freed_by_stage[s] = false;
  }
  if (c->post_stream) cudaStreamSynchronize(c->post_stream);
  c->post_pending[0] = c->post_pending[1] = false;
  if (c->d_tile_hit) cudaFree(c->d_tile_hit);
  c->d_tile_hit = nullptr;
  if (c->d_tile_hit_s[1]) cudaFree(c->d_tile_hit_s[1]);
  c->d_tile_hit_s[0] = c->d_tile_hit_s[1] = nullptr;
  if (c->d_occ_queue) cudaFree(c->d_occ_queue);
  c->d_occ_queue = nullptr;
  if (c->d_occ_table) cudaFree(c->d_occ_table);
  c->d_occ_table = nullptr;
  c->occ_key[2] = -1;
  if (c->d_rgba) cudaFree(c->d_rgba);
  if (c->h_rgba) cudaFreeHost(c->h_rgba);
  c->d_rgba = c->h_rgba = nullptr;
  c->dbuf = nullptr;
  c->hpin = nullptr;
  c->slot = 0;
}
static void free_volume(spv_ctx *c) {
  free_axis(c);
  if (c->tex_lin) cudaDestroyTextureObject(c->tex_lin);
  if (c->tex_near) cudaDestroyTextureObject(c->tex_near);
  if (c->tex_pt) cudaDestroyTextureObject(c->tex_pt);
  if (c->arr) cudaFreeArray(c->arr);
  if (c->bricks) cudaFree(c->bricks);
  if (c->coarse) cudaFree(c->coarse);
  if (c->top) cudaFree(c->top);
  if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
  if (c->d_stage) cudaFree(c->d_stage);
  if (c->d_conv) cudaFree(c->d_conv);
  for (int d = 0; d < 3; ++d) {
    if (c->d_lin[d]) cudaFree(c->d_lin[d]);
    c->d_lin[d] = nullptr;
  }
  c->lin_valid = false;
  if (c->h_ring) cudaFreeHost(c->h_ring);
  c->d_stage = nullptr;
  c->d_conv = nullptr;
  c->conv_bytes = 0;
  c->stage_bytes = 0;
  c->h_ring = nullptr;
  c->ring_bytes = 0;
  c->ring_used[0] = c->ring_used[1] = c->lin_used[0] = c->lin_used[1] = false;
  c->tex_lin = c->tex_near = c->tex_pt = 0;
  c->arr = nullptr;
  c->bricks = c->coarse = c->top = nullptr;
  c->dtype = -1;
}

static int alloc_slot(spv_ctx *ctx, int s) {
  const size_t n = ctx->n();
  CU(cudaMalloc(&ctx->dbuf_s[s], 12 * n * sizeof(float)));
  CU(cudaMemsetAsync(ctx->dbuf_s[s], 0, 12 * n * sizeof(float), ctx->stream));
  CU(cudaMallocHost(&ctx->hpin_s[s], 7 * n * sizeof(float)));
  memset(ctx->hpin_s[s], 0, 7 * n * sizeof(float));
  ctx->dirty_lo[s] = ctx->dirty_hi[s] = 0;
  ctx->dirty_x0[s] = ctx->dirty_x1[s] = 0;
  ctx->clean_alpha[s] = 0.f;
  ctx->clip_s[s].valid = false;  // no finished frame in the new planes
  return 0;
}

// another path is about to write slot s's pinned staging: nothing is known about its rows any more
static void staging_dirty(spv_ctx *ctx, int s) {
  ctx->dirty_lo[s] = 0;
  ctx->dirty_hi[s] = ctx->height;
  ctx->dirty_x0[s] = 0;
  ctx->dirty_x1[s] = ctx->width;
}
// make everything outside the rectangle [xa, xb) x [ya, yb) of slot s's out / alpha staging hold the miss values (the
// staging is quiescent): what the tracked dirty rectangle covers outside the new one is refilled
static void staging_clean_outside_rect(spv_ctx *ctx, int s, int xa, int xb, int ya, int yb, float miss_alpha) {
  const int H = ctx->height, Wd = ctx->width;
  const size_t W = (size_t)ctx->width, n = ctx->n();
  if (ctx->clean_alpha[s] != miss_alpha) staging_dirty(ctx, s);
  const int dx0 = ctx->dirty_x0[s] < 0 ? 0 : ctx->dirty_x0[s], dx1 = ctx->dirty_x1[s] > Wd ? Wd : ctx->dirty_x1[s];
  const int dy0 = ctx->dirty_lo[s] < 0 ? 0 : ctx->dirty_lo[s], dy1 = ctx->dirty_hi[s] > H ? H : ctx->dirty_hi[s];
  const bool keep = xa < xb && ya < yb;
  auto fill = [&](int y, int x0, int x1) {
    if (x0 >= x1) return;
    float *o = ctx->hpin_s[s] + (size_t)y * W, *al = ctx->hpin_s[s] + n + (size_t)y * W;
    memset(o + x0, 0, (size_t)(x1 - x0) * sizeof(float));
    if (miss_alpha == 0.f) memset(al + x0, 0, (size_t)(x1 - x0) * sizeof(float));
    else std::fill(al + x0, al + x1, miss_alpha);
  };
  if (dx0 < dx1)
    for (int y = dy0; y < dy1; ++y) {
      if (!keep || y < ya || y >= yb) { fill(y, dx0, dx1); continue; }
      fill(y, dx0, dx1 < xa ? dx1 : xa);
      fill(y, dx0 > xb ? dx0 : xb, dx1);
    }
  ctx->clean_alpha[s] = miss_alpha;
  ctx->dirty_lo[s] = keep ? ya : 0;
  ctx->dirty_hi[s] = keep ? yb : 0;
  ctx->dirty_x0[s] = keep ? xa : 0;
  ctx->dirty_x1[s] = keep ? xb : 0;
}
// the event after which the device planes of slot s may be rendered into again (its last asynchronous read-back)
static cudaEvent_t slot_free_event(spv_ctx *ctx, int s) { return ctx->freed_by_stage[s] ? ctx->ev_freed[s] : ctx->ev_copied[s]; }
static void slot_clip_set(spv_ctx *ctx, int s, const Camera &cam, const float *box, float miss_alpha, bool full = false) {
  ctx->clip_s[s].valid = true;
  ctx->clip_s[s].full = full;
  ctx->clip_s[s].cam = cam;
  memcpy(ctx->clip_s[s].box, box, sizeof ctx->clip_s[s].box);
  ctx->clip_s[s].miss_alpha = miss_alpha;
}

static int alloc_buffers(spv_ctx *ctx, int w, int h) {
  if (w <= 0 || h <= 0) return fail(ctx, SPV_EINVAL, "spv_resize: width and height must be positive");
  free_buffers(ctx);
  ctx->width = w;
  ctx->height = h;
  int rc = alloc_slot(ctx, 0);
  if (rc) return rc;
  ctx->dbuf = ctx->dbuf_s[0];
  ctx->hpin = ctx->hpin_s[0];
  CU(cudaMalloc(&ctx->d_tile_hit, (size_t)((w + 7) / 8) * ((h + 3) / 4)));  // one flag per 8x4 warp tile
  CU(cudaMalloc(&ctx->d_tile_hit_s[1], (size_t)((w + 7) / 8) * ((h + 3) / 4)));  // slot 1's flags (iso overlap)
  ctx->d_tile_hit_s[0] = ctx->d_tile_hit;
  CU(cudaMalloc(&ctx->d_occ_queue, occ_queue_bytes(w, h)));
  CU(cudaMemsetAsync(ctx->d_occ_queue, 0, occ_queue_bytes(w, h), ctx->stream));
  return 0;
}

extern "C" {

SPV_API int spv_version(void) { return SPV_VERSION; }

SPV_API const char *spv_last_error(spv_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

SPV_API int spv_create(int device, int width, int height, spv_ctx **out) {
  if (!out) return fail(nullptr, SPV_EINVAL, "spv_create: null out");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess) return cufail(nullptr, e, "spv_create: cudaGetDeviceCount (no CUDA device: this library has no CPU path)");
  if (device < 0 || device >= ndev) return fail(nullptr, SPV_EINVAL, "spv_create: no such CUDA device");
  spv_ctx *ctx = new spv_ctx();
  ctx->device = device;
  for (int i = 0; i < 16; ++i) ctx->cam.invP[i] = ctx->cam.invM[i] = (i % 5 == 0) ? 1.f : 0.f;
#define CC(call)                                   \
  do {                                             \
    cudaError_t e_ = (call);                       \
    if (e_ != cudaSuccess) {                       \
      cufail(nullptr, e_, #call);                  \
      spv_destroy(ctx);                            \
      return (int)e_;                              \
    }                                              \
  } while (0)
  CC(cudaSetDevice(device));
  CC(cudaDeviceGetAttribute(&ctx->sms, cudaDevAttrMultiProcessorCount, device));
  CC(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
  ctx->stream = ctx->own_stream;
  CC(cudaEventCreate(&ctx->ev0));
  CC(cudaEventCreate(&ctx->ev1));
  CC(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  CC(cudaStreamCreateWithFlags(&ctx->copy_stream2, cudaStreamNonBlocking));
  CC(cudaStreamCreateWithFlags(&ctx->post_stream, cudaStreamNonBlocking));
  CC(cudaEventCreateWithFlags(&ctx->ev_uploaded, cudaEventDisableTiming));
  CC(cudaEventCreateWithFlags(&ctx->ev_copy2, cudaEventDisableTiming));
  CC(cudaEventCreateWithFlags(&ctx->ev_up_begin, cudaEventDisableTiming));
  for (int s = 0; s < 2; ++s) {
    CC(cudaEventCreateWithFlags(&ctx->ev_rendered[s], cudaEventDisableTiming));
    CC(cudaEventCreateWithFlags(&ctx->ev_copied[s], cudaEventDisableTiming));
    CC(cudaEventCreateWithFlags(&ctx->ev_freed[s], cudaEventDisableTiming));
    CC(cudaEventCreateWithFlags(&ctx->ev_h2d_done[s], cudaEventDisableTiming));
    CC(cudaEventCreateWithFlags(&ctx->ev_consumed[s], cudaEventDisableTiming));
    CC(cudaEventCreateWithFlags(&ctx->ev_searched[s], cudaEventDisableTiming));
    CC(cudaEventCreateWithFlags(&ctx->ev_posted[s], cudaEventDisableTiming));
    CC(cudaEventCreateWithFlags(&ctx->ev_batch_rendered[s], cudaEventDisableTiming));
    CC(cudaEventCreateWithFlags(&ctx->ev_batch_copied[s], cudaEventDisableTiming));
  }
  CC(cudaMalloc(&ctx->d_minmax, 2 * sizeof(float)));
  CC(cudaMalloc(&ctx->d_stats, 40 * sizeof(unsigned long long)));
  CC(cudaMalloc(&ctx->d_tile_counter, sizeof(unsigned)));
  CC(cudaMalloc(&ctx->d_band_done, 64 * sizeof(unsigned)));
  CC(cudaMemset(ctx->d_band_done, 0, 64 * sizeof(unsigned)));
  CC(cudaMalloc(&ctx->d_iso_err, sizeof(unsigned)));
  CC(cudaMemset(ctx->d_iso_err, 0, sizeof(unsigned)));
#undef CC
  int rc = alloc_buffers(ctx, width, height);
  if (rc) {
    g_create_err = ctx->err;
    spv_destroy(ctx);
    return rc;
  }
  *out = ctx;
  return 0;
}

SPV_API int spv_destroy(spv_ctx *ctx) {
  if (!ctx) return 0;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  if (ctx->post_stream) cudaStreamSynchronize(ctx->post_stream);
  free_volume(ctx);
  free_buffers(ctx);
  if (ctx->d_minmax) cudaFree(ctx->d_minmax);
  if (ctx->d_stats) cudaFree(ctx->d_stats);
  if (ctx->d_tile_counter) cudaFree(ctx->d_tile_counter);
  for (int i = 0; i < 9; ++i)
    if (ctx->ev_ph[i]) cudaEventDestroy(ctx->ev_ph[i]);
  if (ctx->d_band_done) cudaFree(ctx->d_band_done);
  if (ctx->d_iso_err) cudaFree(ctx->d_iso_err);
  if (ctx->d_taps) cudaFree(ctx->d_taps);
  if (ctx->d_occ_table) cudaFree(ctx->d_occ_table);
  if (ctx->d_lut) cudaFree(ctx->d_lut);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  if (ctx->ev_up_begin) cudaEventDestroy(ctx->ev_up_begin);
  for (int s = 0; s < 2; ++s) {
    if (ctx->ev_rendered[s]) cudaEventDestroy(ctx->ev_rendered[s]);
    if (ctx->ev_copied[s]) cudaEventDestroy(ctx->ev_copied[s]);
    if (ctx->ev_freed[s]) cudaEventDestroy(ctx->ev_freed[s]);
    if (ctx->ev_h2d_done[s]) cudaEventDestroy(ctx->ev_h2d_done[s]);
    if (ctx->ev_consumed[s]) cudaEventDestroy(ctx->ev_consumed[s]);
    if (ctx->ev_searched[s]) cudaEventDestroy(ctx->ev_searched[s]);
    if (ctx->ev_posted[s]) cudaEventDestroy(ctx->ev_posted[s]);
    if (ctx->ev_batch_rendered[s]) cudaEventDestroy(ctx->ev_batch_rendered[s]);
    if (ctx->ev_batch_copied[s]) cudaEventDestroy(ctx->ev_batch_copied[s]);
  }
  if (ctx->post_stream) cudaStreamDestroy(ctx->post_stream);
  if (ctx->ev_uploaded) cudaEventDestroy(ctx->ev_uploaded);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->copy_stream2) cudaStreamDestroy(ctx->copy_stream2);
  if (ctx->ev_copy2) cudaEventDestroy(ctx->ev_copy2);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx;
  return 0;
}

SPV_API int spv_resize(spv_ctx *ctx, int width, int height) {
  BIND();
  CU(cudaStreamSynchronize(ctx->stream));
  return alloc_buffers(ctx, width, height);
}

SPV_API int spv_set_stream(spv_ctx *ctx, void *cuda_stream) {
  BIND();
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
  return 0;
}

SPV_API int spv_share_stream(spv_ctx *ctx, spv_ctx *other) {
  BIND();
  if (!other || other->device != ctx->device) return fail(ctx, SPV_EINVAL, "spv_share_stream: need a context on the same device");
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->stream = other->stream;
  return 0;
}

SPV_API int spv_sync(spv_ctx *ctx) {
  BIND();
  CU(cudaStreamSynchronize(ctx->stream));
  CU(cudaStreamSynchronize(ctx->post_stream));
  CU(cudaStreamSynchronize(ctx->copy_stream));
  CU(cudaStreamSynchronize(ctx->copy_stream2));
  ctx->post_pending[0] = ctx->post_pending[1] = false;
  return 0;
}

// the render stream waits for the screen-space passes that run beside it (iso overlap): whatever is enqueued on the
// render stream next sees finished planes in slot s (s < 0: both slots)
static int join_post(spv_ctx *ctx, int s) {
  for (int k = 0; k < 2; ++k)
    if ((s < 0 || s == k) && ctx->post_pending[k]) {
      CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_posted[k], 0));
      ctx->post_pending[k] = false;
    }
  return 0;
}


SPV_API int spv_stream_join(spv_ctx *ctx) {
  BIND();
  return join_post(ctx, -1);
}

static size_t elem_size(int dtype) { return dtype == SPV_F32 ? 4 : (dtype == SPV_U16 ? 2 : 1); }

static int make_textures(spv_ctx *ctx) {
  cudaResourceDesc rd;
  memset(&rd, 0, sizeof rd);
  rd.resType = cudaResourceTypeArray;
  rd.res.array.array = ctx->arr;
  cudaTextureDesc td;
  memset(&td, 0, sizeof td);
  td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
  td.normalizedCoords = 0;
  // filtered: integer texels can only be filtered when read as normalised floats
  td.readMode = ctx->dtype == SPV_F32 ? cudaReadModeElementType : cudaReadModeNormalizedFloat;
  td.filterMode = cudaFilterModeLinear;
  CU(cudaCreateTextureObject(&ctx->tex_lin, &rd, &td, nullptr));
  td.filterMode = cudaFilterModePoint;
  CU(cudaCreateTextureObject(&ctx->tex_near, &rd, &td, nullptr));
  td.readMode = cudaReadModeElementType;
  CU(cudaCreateTextureObject(&ctx->tex_pt, &rd, &td, nullptr));
  return 0;
}

static int fmt_of(const spv_ctx *c) { return c->dtype + 3 * c->layout; }

static Volume volume_of(const spv_ctx *c) {
  Volume V;
  const bool lin = c->linear && (c->dtype == SPV_F32 || c->int_filter);
  V.filt = lin ? c->tex_lin : c->tex_near;
  V.pt = c->tex_pt;
  V.nx = c->nx; V.ny = c->ny; V.nz = c->gnz;
  V.local_nz = c->local_nz;
  V.fnx = (float)c->nx; V.fny = (float)c->ny; V.fnz = (float)c->gnz;
  V.scale = c->dtype == SPV_F32 ? 1.f : (c->dtype == SPV_U16 ? 65535.f : 255.f);
  V.z_lo = c->z_lo; V.z0 = c->z0; V.z1 = c->z1;
  V.bricks = c->bricks;
  V.gx = c->gx; V.gy = c->gy; V.gz = c->gz;
  return V;
}

// Copy n bytes with a few host threads (one thread saturates neither the memory system nor a PCIe 5 link).
enum { SRC_DEVICE = 0, SRC_PINNED = 1, SRC_PAGEABLE = 2 };

// Fill the resident array from `src` (C-order, local_nz slices) -- the ingest path (replaces OCLImage.write_array,
// volumerender.py:294).  Host sources move in chunks of ~32 MiB through a two-deep pipeline:
//   pageable memory  : host threads copy chunk i+1 into a page-locked ring while chunk i is on the PCIe link
//   page-locked      : the DMA engine reads the caller's buffer directly
//   copy_stream      : H2D of chunk i+1 overlaps the device-side work on chunk i
//   LAYOUT_ZPAIR     : pair_kernel builds {v[z], v[z+1]} texels from the linear chunk, cudaMemcpy3D moves them into
//                      the layers (stream order on the render stream)
//   LAYOUT_3D        : the DMA writes the array slices directly
//   src_type >= 0    : the host array has another element type (SPV_SRC_*): its bytes travel as they are, convert_kernel
//                      turns each chunk into texels on the device (instead of a host-side astype), then as above
static int upload(spv_ctx *ctx, const void *src, bool on_device, bool no_wait = false, int src_type = -1) {
  // a max projection running beside the render stream (tuning knob 15) may still be reading the array
  {
    int rcj = join_post(ctx, -1);
    if (rcj) return rcj;
  }
  ctx->upload_seq++;
  const size_t es = elem_size(ctx->dtype);
  const size_t slice = (size_t)ctx->nx * ctx->ny;
  const size_t slice_bytes = slice * es;
  const bool conv = src_type >= 0;
  const size_t in_slice_bytes = conv ? slice * src_elem_size(src_type) : slice_bytes;  // as the chunk travels
  if (conv && (on_device || in_slice_bytes == 0)) return fail(ctx, SPV_EINVAL, "upload: bad source element type");
  const int nz = ctx->local_nz;
  const bool zpair = ctx->layout == LAYOUT_ZPAIR;
  int kind = SRC_DEVICE;
  if (!on_device) {
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, src);
    if (e != cudaSuccess) cudaGetLastError();
    kind = (e == cudaSuccess && at.type == cudaMemoryTypeHost) ? SRC_PINNED : SRC_PAGEABLE;
  }
  ctx->bricks_valid = false;  // rebuilt by ensure_bricks() when something needs them
  ctx->minmax_valid = false;
  ctx->lin_valid = false;     // rebuilt by ensure_linear() when the software-sampled path renders

  // slices per chunk
  const size_t budget = kind == SRC_DEVICE ? ((size_t)128 << 20) : ((size_t)32 << 20);
  int zc = (int)(budget / (slice_bytes > in_slice_bytes ? slice_bytes : in_slice_bytes));
  if (zc < 1) zc = 1;
  if (zc > nz) zc = nz;
  const size_t lin_bytes = slice_bytes * (size_t)(zc + (zpair ? 1 : 0));  // + the upper partner of the last slice
  const size_t in_bytes = in_slice_bytes * (size_t)(zc + (zpair ? 1 : 0));
  const size_t pair_bytes = zpair ? slice_bytes * 2 * (size_t)zc : 0;
  const size_t need_dev = pair_bytes + (((zpair || conv) && kind != SRC_DEVICE) ? 2 * lin_bytes : 0) + (conv ? 2 * in_bytes : 0);
  if (need_dev > ctx->stage_bytes) {
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaStreamSynchronize(ctx->copy_stream));
    if (ctx->d_stage) cudaFree(ctx->d_stage);
    ctx->d_stage = nullptr;
    ctx->stage_bytes = 0;
    CU(cudaMalloc(&ctx->d_stage, need_dev));
    ctx->stage_bytes = need_dev;
  }
  if (kind == SRC_PAGEABLE && 2 * in_bytes > ctx->ring_bytes) {
    CU(cudaStreamSynchronize(ctx->copy_stream));
    if (ctx->h_ring) cudaFreeHost(ctx->h_ring);
    ctx->h_ring = nullptr;
    ctx->ring_bytes = 0;
    CU(cudaMallocHost(&ctx->h_ring, 2 * in_bytes));
    ctx->ring_bytes = 2 * in_bytes;
  }
  if (ctx->stage_sig[0] != pair_bytes || ctx->stage_sig[1] != lin_bytes || ctx->stage_sig[2] != in_bytes) {
    // the staging is cut differently than by the previous upload (which may still be in flight after an asynchronous
    // call): the per-half events no longer describe these regions
    CU(cudaStreamSynchronize(ctx->copy_stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->stage_sig[0] = pair_bytes; ctx->stage_sig[1] = lin_bytes; ctx->stage_sig[2] = in_bytes;
  }
  char *d_pair = (char *)ctx->d_stage;
  char *d_lin[2] = {d_pair + pair_bytes, d_pair + pair_bytes + lin_bytes};
  char *d_in[2] = {d_pair + pair_bytes + 2 * lin_bytes, d_pair + pair_bytes + 2 * lin_bytes + in_bytes};  // conv only
  char *h_ring[2] = {ctx->h_ring, ctx->h_ring ? ctx->h_ring + in_bytes : nullptr};

  auto to_layers = [&](const void *lin, int lin_nz, int zfirst, int zb, int ze) -> int {
    // lin holds slices [zfirst, zfirst + lin_nz) of the volume; build the paired texels of [zb, ze) and store them
    CU(launch_pair(lin, d_pair, ctx->dtype, slice, lin_nz, zb - zfirst, ze - zfirst, ctx->stream));
    cudaMemcpy3DParms p;
    memset(&p, 0, sizeof p);
    p.srcPtr = make_cudaPitchedPtr(d_pair, (size_t)ctx->nx * 2 * es, ctx->nx, ctx->ny);
    p.dstArray = ctx->arr;
    p.dstPos = make_cudaPos(0, 0, zb);
    p.extent = make_cudaExtent(ctx->nx, ctx->ny, ze - zb);
    p.kind = cudaMemcpyDeviceToDevice;
    CU(cudaMemcpy3DAsync(&p, ctx->stream));
    ctx->launches += 1;
    return 0;
  };

  if (kind == SRC_DEVICE) {
    if (!zpair) {
      cudaMemcpy3DParms p;
      memset(&p, 0, sizeof p);
      p.srcPtr = make_cudaPitchedPtr(const_cast<void *>(src), (size_t)ctx->nx * es, ctx->nx, ctx->ny);
      p.dstArray = ctx->arr;
      p.extent = make_cudaExtent(ctx->nx, ctx->ny, nz);
      p.kind = cudaMemcpyDeviceToDevice;
      CU(cudaMemcpy3DAsync(&p, ctx->stream));
    } else {
      for (int zb = 0; zb < nz; zb += zc) {
        int rc = to_layers(src, nz, 0, zb, zb + zc < nz ? zb + zc : nz);
        if (rc) return rc;
      }
    }
    return 0;
  }

  // host sources: the DMA runs on copy_stream; it must not overtake renders that still read the array
  CU(cudaEventRecord(ctx->ev_up_begin, ctx->stream));
  CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_up_begin, 0));
  int i = 0;
  for (int zb = 0; zb < nz; zb += zc, ++i) {
    const int ze = zb + zc < nz ? zb + zc : nz;
    const int zs = zpair ? (ze < nz ? ze + 1 : nz) : ze;  // slices [zb, zs) travel
    const size_t bytes = (size_t)(zs - zb) * in_slice_bytes;
    const int s = i & 1;
    const char *from = (const char *)src + (size_t)zb * in_slice_bytes;
    if (kind == SRC_PAGEABLE) {
      if (ctx->ring_used[s]) CU(cudaEventSynchronize(ctx->ev_h2d_done[s]));  // the DMA that last read this half
      parallel_memcpy(h_ring[s], from, bytes);
      ctx->ring_used[s] = true;
      from = h_ring[s];
    }
    if (zpair || conv) {
      if (ctx->lin_used[s]) CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_consumed[s], 0));
      CU(cudaMemcpyAsync(conv ? d_in[s] : d_lin[s], from, bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
      CU(cudaEventRecord(ctx->ev_h2d_done[s], ctx->copy_stream));
      CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_h2d_done[s], 0));
      if (conv) {
        CU(launch_convert(d_in[s], d_lin[s], src_type, ctx->dtype, (size_t)(zs - zb) * slice, ctx->stream));
        ctx->launches += 1;
      }
      if (zpair) {
        int rc = to_layers(d_lin[s], zs - zb, zb, zb, ze);
        if (rc) return rc;
      } else {
        cudaMemcpy3DParms p;
        memset(&p, 0, sizeof p);
        p.srcPtr = make_cudaPitchedPtr(d_lin[s], (size_t)ctx->nx * es, ctx->nx, ctx->ny);
        p.dstArray = ctx->arr;
        p.dstPos = make_cudaPos(0, 0, zb);
        p.extent = make_cudaExtent(ctx->nx, ctx->ny, ze - zb);
        p.kind = cudaMemcpyDeviceToDevice;
        CU(cudaMemcpy3DAsync(&p, ctx->stream));
      }
      CU(cudaEventRecord(ctx->ev_consumed[s], ctx->stream));
      ctx->lin_used[s] = true;
    } else {
      cudaMemcpy3DParms p;
      memset(&p, 0, sizeof p);
      p.srcPtr = make_cudaPitchedPtr(const_cast<char *>(from), (size_t)ctx->nx * es, ctx->nx, ctx->ny);
      p.dstArray = ctx->arr;
      p.dstPos = make_cudaPos(0, 0, zb);
      p.extent = make_cudaExtent(ctx->nx, ctx->ny, ze - zb);
      p.kind = cudaMemcpyHostToDevice;
      CU(cudaMemcpy3DAsync(&p, ctx->copy_stream));
      CU(cudaEventRecord(ctx->ev_h2d_done[s], ctx->copy_stream));
    }
  }
  if (!zpair && !conv && i > 0) CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_h2d_done[(i - 1) & 1], 0));  // renders wait for the data
  if (!no_wait) {  // the host pointer is only borrowed for this call
    CU(cudaStreamSynchronize(ctx->copy_stream));
    CU(cudaStreamSynchronize(ctx->stream));
  }
  return 0;
}

// min/max brick grids + global min/max of the resident volume: needed by iso-surface skipping, max-projection
// skipping and spv_volume_minmax, not by the brute-force max projection -- a timelapse that only plays back
// max projections never pays for them
static int ensure_bricks(spv_ctx *ctx) {
  if (ctx->bricks_valid) return 0;
  Volume V = volume_of(ctx);
  CU(launch_build_bricks(V, fmt_of(ctx), ctx->local_nz, ctx->bricks, ctx->coarse, ctx->cgx, ctx->cgy, ctx->cgz,
                         ctx->top, ctx->d_minmax, ctx->stream));
  ctx->launches += 4;
  ctx->bricks_valid = true;
  ctx->minmax_valid = false;
  return 0;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point lookup (no link-time dependency on libcuda)
typedef CUresult (*encode_tiled_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                   const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static encode_tiled_t encode_tiled_fn() {
  static encode_tiled_t fn = nullptr;
  static bool looked = false;
  if (!looked) {
    looked = true;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (encode_tiled_t)p;
    else
      cudaGetLastError();
  }
  return fn;
}

// can the software-sampled path render this context's volume at all?
static bool smem_path_possible(const spv_ctx *c) {
  const int m = c->nx < c->ny ? (c->nx < c->gnz ? c->nx : c->gnz) : (c->ny < c->gnz ? c->ny : c->gnz);
  return c->dtype == SPV_U16 && !c->slab && m >= 48 && encode_tiled_fn() != nullptr;
}

// The three permuted linear copies (slowest axis x, y, z; contiguous axis z, x, x) of the resident uint16 volume and
// their tensor maps.  Built from the array on the render stream on first use after an upload.
static int ensure_linear(spv_ctx *ctx) {
  if (ctx->lin_valid) return 0;
  const Volume V = volume_of(ctx);
  for (int D = 0; D < 3; ++D) {
    const int NA = D == 0 ? ctx->gnz : ctx->nx, NB = D == 1 ? ctx->gnz : ctx->ny,
              ND = D == 0 ? ctx->nx : (D == 1 ? ctx->ny : ctx->gnz);
    const size_t pitchA = ((size_t)NA + 7) / 8 * 8;  // rows start on 16-byte boundaries (tensor map stride rule)
    if (!ctx->d_lin[D]) CU(cudaMalloc(&ctx->d_lin[D], pitchA * NB * ND * sizeof(unsigned short)));
    CU(launch_permute(V, fmt_of(ctx), D, NA, NB, ND, pitchA, ctx->d_lin[D], ctx->stream));
    ctx->launches += 1;
    const cuuint64_t dims[3] = {(cuuint64_t)NA, (cuuint64_t)NB, (cuuint64_t)ND};
    const cuuint64_t strides[2] = {pitchA * 2, pitchA * 2 * (cuuint64_t)NB};
    const cuuint32_t es[3] = {1, 1, 1};
    for (int cfg = 0; cfg < mip_smem_configs(); ++cfg) {
      int box[3];
      mip_smem_box(cfg, box);
      const cuuint32_t bx[3] = {(cuuint32_t)box[0], (cuuint32_t)box[1], (cuuint32_t)box[2]};
      CUresult r = encode_tiled_fn()(&ctx->tmaps[cfg][D], CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, ctx->d_lin[D], dims, strides, bx,
                                     es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return fail(ctx, (int)r, "cuTensorMapEncodeTiled failed");
    }
  }
  ctx->lin_valid = true;
  return 0;
}

static int set_volume_impl(spv_ctx *ctx, const void *src, bool on_device, int dtype, int nx, int ny, int gnz, int z0,
                           int z1, bool slab, int halo = 1, int src_type = -1) {
  BIND();
  if (!src) return fail(ctx, SPV_EINVAL, "spv_set_volume: null data");
  if (dtype < 0 || dtype > 2) return fail(ctx, SPV_EINVAL, "spv_set_volume: dtype must be 0 (f32), 1 (u16) or 2 (u8)");
  if (nx <= 0 || ny <= 0 || gnz <= 0 || z0 < 0 || z1 > gnz || z0 >= z1)
    return fail(ctx, SPV_EINVAL, "spv_set_volume: bad extent");
  if (halo < 1) return fail(ctx, SPV_EINVAL, "spv_set_volume_slab: the halo must be at least one slice");
  const int z_lo = slab ? (z0 - halo > 0 ? z0 - halo : 0) : 0;
  const int z_hi = slab ? (z1 + halo < gnz ? z1 + halo : gnz) : gnz;
  const int local_nz = z_hi - z_lo;
  // z-paired layered layout for integer volumes when asked for and the layer count allows it
  const int layout = (ctx->want_layout == LAYOUT_ZPAIR && dtype != SPV_F32 && local_nz <= 2048 && nx <= 32768 &&
                      ny <= 32768) ? LAYOUT_ZPAIR : LAYOUT_3D;
  CU(cudaStreamSynchronize(ctx->stream));
  const bool same = ctx->arr && ctx->dtype == dtype && ctx->nx == nx && ctx->ny == ny && ctx->local_nz == local_nz &&
                    ctx->layout == layout;
  if (!same) {
    free_volume(ctx);
    ctx->dtype = dtype;
    ctx->layout = layout;
    ctx->nx = nx; ctx->ny = ny; ctx->local_nz = local_nz;
    const int bits = dtype == SPV_F32 ? 32 : (dtype == SPV_U16 ? 16 : 8);
    const cudaChannelFormatKind kind = dtype == SPV_F32 ? cudaChannelFormatKindFloat : cudaChannelFormatKindUnsigned;
    cudaChannelFormatDesc cd = cudaCreateChannelDesc(bits, layout == LAYOUT_ZPAIR ? bits : 0, 0, 0, kind);
    CU(cudaMalloc3DArray(&ctx->arr, &cd, make_cudaExtent(nx, ny, local_nz),
                         layout == LAYOUT_ZPAIR ? cudaArrayLayered : cudaArrayDefault));
    int rc = make_textures(ctx);
    if (rc) return rc;
    ctx->gx = (nx + BRICK - 1) / BRICK; ctx->gy = (ny + BRICK - 1) / BRICK; ctx->gz = (local_nz + BRICK - 1) / BRICK;
    ctx->cgx = (ctx->gx + 3) / 4; ctx->cgy = (ctx->gy + 3) / 4; ctx->cgz = (ctx->gz + 3) / 4;
    CU(cudaMalloc(&ctx->bricks, (size_t)ctx->gx * ctx->gy * ctx->gz * sizeof(float2)));
    CU(cudaMalloc(&ctx->coarse, (size_t)ctx->cgx * ctx->cgy * ctx->cgz * sizeof(float2)));
    CU(cudaMalloc(&ctx->top, (size_t)((ctx->cgx + 3) / 4) * ((ctx->cgy + 3) / 4) * ((ctx->cgz + 3) / 4) * sizeof(float2)));
  }
  ctx->gnz = gnz; ctx->z_lo = z_lo; ctx->z0 = z0; ctx->z1 = z1; ctx->slab = slab;
  return upload(ctx, src, on_device, false, src_type);
}

SPV_API int spv_set_volume(spv_ctx *ctx, const void *host, int dtype, int nx, int ny, int nz) {
  return set_volume_impl(ctx, host, false, dtype, nx, ny, nz, 0, nz, false);
}
SPV_API int spv_set_volume_device(spv_ctx *ctx, const void *dev, int dtype, int nx, int ny, int nz) {
  return set_volume_impl(ctx, dev, true, dtype, nx, ny, nz, 0, nz, false);
}
SPV_API int spv_set_volume_slab(spv_ctx *ctx, const void *src, int on_device, int dtype, int nx, int ny, int gnz, int z0,
                        int z1) {
  return set_volume_impl(ctx, src, on_device != 0, dtype, nx, ny, gnz, z0, z1, true);
}
SPV_API int spv_set_volume_slab_halo(spv_ctx *ctx, const void *src, int on_device, int dtype, int nx, int ny, int gnz,
                                     int z0, int z1, int halo) {
  return set_volume_impl(ctx, src, on_device != 0, dtype, nx, ny, gnz, z0, z1, true, halo);
}
SPV_API int spv_update_volume(spv_ctx *ctx, const void *host) {
  BIND();
  if (!ctx->arr) return fail(ctx, SPV_ENODATA, "spv_update_volume: no volume set");
  if (!host) return fail(ctx, SPV_EINVAL, "spv_update_volume: null data");
  return upload(ctx, host, false);
}

// a source type that already is the texel type needs no conversion pass
static int native_src(int src_type, int dtype) {
  return (src_type == SPV_SRC_F32 && dtype == SPV_F32) || (src_type == SPV_SRC_U16 && dtype == SPV_U16) ||
         (src_type == SPV_SRC_U8 && dtype == SPV_U8);
}
SPV_API int spv_set_volume_from(spv_ctx *ctx, const void *host, int src_type, int dtype, int nx, int ny, int nz) {
  if (src_elem_size(src_type) == 0) return fail(ctx, SPV_EINVAL, "spv_set_volume_from: unknown source element type");
  return set_volume_impl(ctx, host, false, dtype, nx, ny, nz, 0, nz, false, 1, native_src(src_type, dtype) ? -1 : src_type);
}
SPV_API int spv_update_volume_from(spv_ctx *ctx, const void *host, int src_type) {
  BIND();
  if (!ctx->arr) return fail(ctx, SPV_ENODATA, "spv_update_volume_from: no volume set");
  if (!host) return fail(ctx, SPV_EINVAL, "spv_update_volume_from: null data");
  if (src_elem_size(src_type) == 0) return fail(ctx, SPV_EINVAL, "spv_update_volume_from: unknown source element type");
  return upload(ctx, host, false, false, native_src(src_type, ctx->dtype) ? -1 : src_type);
}

SPV_API int spv_update_volume_device_from(spv_ctx *ctx, const void *dev, int src_type) {
  BIND();
  if (!ctx->arr) return fail(ctx, SPV_ENODATA, "spv_update_volume_device_from: no volume set");
  if (!dev) return fail(ctx, SPV_EINVAL, "spv_update_volume_device_from: null data");
  if (src_elem_size(src_type) == 0) return fail(ctx, SPV_EINVAL, "spv_update_volume_device_from: unknown source element type");
  if (native_src(src_type, ctx->dtype)) return upload(ctx, dev, true);
  const size_t n = (size_t)ctx->nx * ctx->ny * ctx->local_nz, need = n * elem_size(ctx->dtype);
  if (need > ctx->conv_bytes) {
    CU(cudaStreamSynchronize(ctx->stream));
    if (ctx->d_conv) cudaFree(ctx->d_conv);
    ctx->d_conv = nullptr;
    ctx->conv_bytes = 0;
    CU(cudaMalloc(&ctx->d_conv, need));
    ctx->conv_bytes = need;
  }
  CU(launch_convert(dev, ctx->d_conv, src_type, ctx->dtype, n, ctx->stream));
  ctx->launches += 1;
  return upload(ctx, ctx->d_conv, true);
}

SPV_API int spv_update_volume_async(spv_ctx *ctx, const void *pinned_host) {
  BIND();
  if (!ctx->arr) return fail(ctx, SPV_ENODATA, "spv_update_volume_async: no volume set");
  if (!pinned_host) return fail(ctx, SPV_EINVAL, "spv_update_volume_async: null data");
  cudaPointerAttributes at;
  cudaError_t e = cudaPointerGetAttributes(&at, pinned_host);
  if (e != cudaSuccess || at.type != cudaMemoryTypeHost) {
    cudaGetLastError();
    return fail(ctx, SPV_EINVAL, "spv_update_volume_async: the source must be page-locked host memory (spv_host_alloc / cudaHostRegister)");
  }
  return upload(ctx, pinned_host, false, true);
}

SPV_API int spv_host_alloc(size_t nbytes, void **host) {
  if (!host || nbytes == 0) return SPV_EINVAL;
  cudaError_t e = cudaHostAlloc(host, nbytes, cudaHostAllocPortable);
  if (e != cudaSuccess) return cufail(nullptr, e, "cudaHostAlloc");
  return 0;
}
SPV_API int spv_host_free(void *host) {
  if (host) cudaFreeHost(host);
  return 0;
}

SPV_API int spv_volume_minmax(spv_ctx *ctx, float *vmin, float *vmax) {
  BIND();
  if (!ctx->arr) return fail(ctx, SPV_ENODATA, "spv_volume_minmax: no volume set");
  {
    int rc = ensure_bricks(ctx);
    if (rc) return rc;
  }
  if (!ctx->minmax_valid) {
    CU(cudaMemcpyAsync(ctx->h_minmax, ctx->d_minmax, 2 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->minmax_valid = true;
  }
  if (vmin) *vmin = ctx->h_minmax[0];
  if (vmax) *vmax = ctx->h_minmax[1];
  return 0;
}

SPV_API int spv_set_interp(spv_ctx *ctx, int linear) {
  if (!ctx) return SPV_EINVAL;
  ctx->linear = linear != 0;
  return 0;
}
SPV_API int spv_set_sampler(spv_ctx *ctx, int sampler) {
  if (!ctx) return SPV_EINVAL;
  if (sampler != SPV_SAMPLER_TMU && sampler != SPV_SAMPLER_EXACT) return fail(ctx, SPV_EINVAL, "spv_set_sampler: unknown sampler");
  ctx->sampler = sampler;
  return 0;
}
SPV_API int spv_set_int_filter(spv_ctx *ctx, int linear) {
  if (!ctx) return SPV_EINVAL;
  ctx->int_filter = linear != 0;
  return 0;
}
SPV_API int spv_set_layout(spv_ctx *ctx, int layout) {
  if (!ctx) return SPV_EINVAL;
  if (layout != LAYOUT_3D && layout != LAYOUT_ZPAIR) return fail(ctx, SPV_EINVAL, "spv_set_layout: unknown layout");
  ctx->want_layout = layout;  // takes effect at the next spv_set_volume*
  return 0;
}
SPV_API int spv_set_skipping(spv_ctx *ctx, int on) {
  if (!ctx) return SPV_EINVAL;
  ctx->skipping = on < 0 ? -1 : (on != 0);
  return 0;
}
SPV_API int spv_set_mip_path(spv_ctx *ctx, int path) {
  if (!ctx) return SPV_EINVAL;
  if (path != SPV_MIP_PATH_TMU && path != SPV_MIP_PATH_SMEM) return fail(ctx, SPV_EINVAL, "spv_set_mip_path: unknown path");
  if (path == SPV_MIP_PATH_SMEM && !encode_tiled_fn())
    return fail(ctx, SPV_EINVAL, "spv_set_mip_path: this driver has no cuTensorMapEncodeTiled");
  ctx->mip_path = path;
  return 0;
}
SPV_API int spv_mip_path_used(spv_ctx *ctx, int *path) {
  if (!ctx || !path) return SPV_EINVAL;
  *path = ctx->last_mip_path;
  return 0;
}
SPV_API int spv_set_tuning(spv_ctx *ctx, int knob, int value) {
  if (!ctx) return SPV_EINVAL;
  if (knob == 0) ctx->tile_variant = value;
  else if (knob == 1) ctx->persistent = value != 0;
  else if (knob == 2) ctx->bands = value < 1 ? 1 : (value > 64 ? 64 : value);
  else if (knob == 3) ctx->direct_host = value != 0;
  else if (knob == 4) ctx->iso_segments = value == 4 ? 4 : (value == 2 ? 2 : 1);
  else if (knob == 5) ctx->iso_centre_out = value != 0;
  else if (knob == 7) ctx->copy_streams = value > 1 ? 2 : 1;
  else if (knob == 8) ctx->row_mode = value == 1 ? 1 : 0;
  else if (knob == 9) ctx->clip_copies = value != 0;
  else if (knob == 10) ctx->smem_tex_of8 = value < 0 ? 0 : (value > 8 ? 8 : value);
  else if (knob == 12) ctx->iso_post_sharded = value != 0;
  else if (knob == 13) ctx->time_phases = value != 0;
  else if (knob == 15) {
    if (!value && ctx->mip_overlap) {
      cudaSetDevice(ctx->device);
      join_post(ctx, -1);
    }
    ctx->mip_overlap = value != 0;
  }
  else if (knob == 14) {
    if (!value && ctx->iso_overlap) {  // switching off joins what is in flight
      cudaSetDevice(ctx->device);
      join_post(ctx, -1);
    }
    ctx->iso_overlap = value != 0;
  }
  else if (knob == 16) ctx->axis_mode = value;
  else if (knob == 11) ctx->smem_cfg = value < 0 || value >= mip_smem_configs() ? 0 : value;
  else if (knob == 6) occ_ctas_per_sm = value < 1 ? 1 : (value > 16 ? 16 : value);  // process-wide
  else if (knob == 17) ctx->occ_table_on = value != 0;
  else if (knob == 18) ctx->stage_reads = value != 0;
  else if (knob == 20) ctx->fuse_shading = value != 0;
  else return fail(ctx, SPV_EINVAL, "spv_set_tuning: unknown knob");
  return 0;
}
SPV_API int spv_enable_stats(spv_ctx *ctx, int on) {
  if (!ctx) return SPV_EINVAL;
  ctx->stats_on = on != 0;
  return 0;
}

SPV_API int spv_set_matrices(spv_ctx *ctx, const float *invP, const float *invM) {
  if (!ctx || !invP || !invM) return fail(ctx, SPV_EINVAL, "spv_set_matrices: null argument");
  memcpy(ctx->cam.invP, invP, 16 * sizeof(float));
  memcpy(ctx->cam.invM, invM, 16 * sizeof(float));
  return 0;
}

static int begin_render(spv_ctx *ctx) {
  if (ctx->stats_on) CU(cudaMemsetAsync(ctx->d_stats, 0, 40 * sizeof(unsigned long long), ctx->stream));
  CU(cudaEventRecord(ctx->ev0, ctx->stream));
  return 0;
}
static int end_render(spv_ctx *ctx) {
  CU(cudaEventRecord(ctx->ev1, ctx->stream));
  ctx->timed = true;
  if (ctx->stats_on)
    CU(cudaMemcpyAsync(ctx->h_stats, ctx->d_stats, 40 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
  return 0;
}

static bool bad_float(float f) { return !(f == f); }

// cuStreamWaitValue32 through the runtime's driver entry point lookup (no link-time dependency on libcuda);
// nullptr where the driver does not offer it
typedef CUresult (*wait_value_t)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
static wait_value_t wait_value_fn() {
  static wait_value_t fn = nullptr;
  static bool looked = false;
  if (!looked) {
    looked = true;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (wait_value_t)p;
    else
      cudaGetLastError();
  }
  return fn;
}

// 4x4 inverse (Gauss-Jordan with partial pivoting, double); false if singular
static bool invert4(const float *m, double *inv) {
  double a[4][8];
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) {
      a[r][c] = m[4 * r + c];
      a[r][4 + c] = r == c ? 1. : 0.;
    }
  for (int c = 0; c < 4; ++c) {
    int piv = c;
    for (int r = c + 1; r < 4; ++r)
      if (fabs(a[r][c]) > fabs(a[piv][c])) piv = r;
    if (fabs(a[piv][c]) < 1e-300) return false;
    if (piv != c)
      for (int k = 0; k < 8; ++k) { const double t = a[c][k]; a[c][k] = a[piv][k]; a[piv][k] = t; }
    const double d = 1. / a[c][c];
    for (int k = 0; k < 8; ++k) a[c][k] *= d;
    for (int r = 0; r < 4; ++r)
      if (r != c) {
        const double f = a[r][c];
        if (f != 0.)
          for (int k = 0; k < 8; ++k) a[r][k] -= f * a[c][k];
      }
  }
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) inv[4 * r + c] = a[r][4 + c];
  return true;
}

// Pixel rectangle [x0, x1) x [y0, y1) of a W x H image that the box can project to: the hull of its eight corners
// through projection . modelView, one pixel of slack.  The whole image when a corner lies behind the eye (or anything
// is degenerate); empty (all zeros) when the box is off screen.  Scheduling only where pixels are rendered anyway; where
// pixels outside it are NOT rendered or copied the callers add one more tile on every side (the kernel's fp32 slab test
// can differ from the exact geometry by ~1e-4 pixel at most).
static void hit_rect(const Camera &cam, const float *box, int W, int H, double &x0, double &x1, double &y0, double &y1) {
  x0 = 0.; x1 = W; y0 = 0.; y1 = H;
  double P[16], M[16];
  if (!invert4(cam.invP, P) || !invert4(cam.invM, M)) return;
  double xmin = 1e300, xmax = -1e300, ymin = 1e300, ymax = -1e300;
  for (int i = 0; i < 8; ++i) {
    const double c[4] = {box[i & 1], box[2 + ((i >> 1) & 1)], box[4 + ((i >> 2) & 1)], 1.};
    double e[4], q[4];
    for (int r = 0; r < 4; ++r) e[r] = M[4 * r] * c[0] + M[4 * r + 1] * c[1] + M[4 * r + 2] * c[2] + M[4 * r + 3] * c[3];
    for (int r = 0; r < 4; ++r) q[r] = P[4 * r] * e[0] + P[4 * r + 1] * e[1] + P[4 * r + 2] * e[2] + P[4 * r + 3] * e[3];
    if (!(q[3] > 1e-9)) return;
    const double x = (q[0] / q[3] + 1.) * 0.5 * W;  // u = (x / Nx) * 2 - 1
    const double y = (q[1] / q[3] + 1.) * 0.5 * H;  // v = (y / Ny) * 2 - 1
    if (!(x == x) || !(y == y)) return;
    xmin = x < xmin ? x : xmin;
    xmax = x > xmax ? x : xmax;
    ymin = y < ymin ? y : ymin;
    ymax = y > ymax ? y : ymax;
  }
  xmin -= 1.; ymin -= 1.; xmax += 1.; ymax += 1.;
  if (xmin >= (double)W || xmax <= 0. || ymin >= (double)H || ymax <= 0.) { x0 = x1 = y0 = y1 = 0.; return; }
  x0 = xmin > 0. ? xmin : 0.;
  x1 = xmax < (double)W ? xmax : (double)W;
  y0 = ymin > 0. ? ymin : 0.;
  y1 = ymax < (double)H ? ymax : (double)H;
}

// Tile rows [ta, tb) (8 pixels each) of an image of H rows that the box can project to (conservative; any answer is
// correct for scheduling).
static void hit_tile_rows(const Camera &cam, const float *box, int H, unsigned tiles_y, unsigned &ta, unsigned &tb) {
  double x0, x1, y0, y1;
  hit_rect(cam, box, 16, H, x0, x1, y0, y1);
  if (!(y1 > y0)) { ta = tb = 0; return; }
  const double lo = floor(y0 / 8.), hi = ceil(y1 / 8.) + 1.;
  ta = lo > 0. ? (unsigned)lo : 0u;
  tb = hi < (double)tiles_y ? (unsigned)hi : tiles_y;
  if (tb < ta) tb = ta;
}

// Pixel rectangle, in whole CTA tiles (16 x 8 pixels) with one tile of slack on every side, outside of which every ray
// of the frame misses the box; empty when the box is off screen.
static void miss_free_rect(const Camera &cam, const float *box, int W, int H, int &xa, int &xb, int &ya, int &yb) {
  double x0, x1, y0, y1;
  hit_rect(cam, box, W, H, x0, x1, y0, y1);
  if (!(x1 > x0) || !(y1 > y0)) { xa = xb = ya = yb = 0; return; }
  xa = ((int)floor(x0 / 16.) - 1) * 16;
  xb = ((int)ceil(x1 / 16.) + 1) * 16;
  ya = ((int)floor(y0 / 8.) - 1) * 8;
  yb = ((int)ceil(y1 / 8.) + 1) * 8;
  xa = xa > 0 ? xa : 0;
  ya = ya > 0 ? ya : 0;
  xb = xb < W ? xb : W;
  yb = yb < H ? yb : H;
}

// ---- view-aligned layered copies (spv_mip_axis.cu) ------------------------------------------------------------------
// The copy of the resident integer volume with pairs along axis lax (0 x, 1 y), built from the primary z copy on the
// render stream; post_stream is made to wait for it (frames in output slot 1 may render there).
static int ensure_axis(spv_ctx *ctx, int lax) {
  const bool f32 = ctx->dtype == SPV_F32;
  if (lax == 2 && !f32) return 0;  // integer volumes: the primary array is the z copy
  if (ctx->axis_arr[lax] && ctx->axis_seq[lax] == ctx->upload_seq) return 0;
  if (!ctx->axis_arr[lax]) {
    const int bits = f32 ? 32 : (ctx->dtype == SPV_U16 ? 16 : 8);
    cudaChannelFormatDesc cd = cudaCreateChannelDesc(bits, bits, 0, 0, f32 ? cudaChannelFormatKindFloat : cudaChannelFormatKindUnsigned);
    const cudaExtent ext = lax == 0 ? make_cudaExtent(ctx->gnz, ctx->ny, ctx->nx)
                                    : (lax == 1 ? make_cudaExtent(ctx->nx, ctx->gnz, ctx->ny) : make_cudaExtent(ctx->nx, ctx->ny, ctx->gnz));
    cudaError_t e = cudaMalloc3DArray(&ctx->axis_arr[lax], &cd, ext, cudaArrayLayered | cudaArraySurfaceLoadStore);
    if (e != cudaSuccess) {  // no room for another copy: this volume renders from what it has
      cudaGetLastError();
      ctx->axis_arr[lax] = nullptr;
      ctx->axis_failed[lax] = true;
      return 1;
    }
    cudaResourceDesc rd;
    memset(&rd, 0, sizeof rd);
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = ctx->axis_arr[lax];
    cudaTextureDesc td;
    memset(&td, 0, sizeof td);
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
    td.normalizedCoords = 0;
    td.readMode = f32 ? cudaReadModeElementType : cudaReadModeNormalizedFloat;
    td.filterMode = cudaFilterModeLinear;
    CU(cudaCreateTextureObject(&ctx->axis_tex[lax], &rd, &td, nullptr));
    CU(cudaCreateSurfaceObject(&ctx->axis_surf[lax], &rd));
  }
  CU(launch_axis_pair(volume_of(ctx), ctx->dtype, lax, ctx->axis_surf[lax], ctx->stream));
  ctx->launches += 1;
  ctx->axis_seq[lax] = ctx->upload_seq;
  CU(cudaEventRecord(ctx->ev_uploaded, ctx->stream));
  CU(cudaStreamWaitEvent(ctx->post_stream, ctx->ev_uploaded, 0));
  return 0;
}

// can plain projections of the resident volume go through mip_axis_kernel at all?
static bool axis_path_possible(const spv_ctx *c) {
  // integer volumes in the z-paired layout (the z copy exists already), or float32 volumes (3-D array: every copy is extra)
  const bool ints = c->dtype != SPV_F32 && c->layout == LAYOUT_ZPAIR && c->int_filter;
  const bool f32 = c->dtype == SPV_F32 && c->layout == LAYOUT_3D && c->axis_mode != 2 &&
                   !(c->axis_failed[0] && c->axis_failed[1] && c->axis_failed[2]);
  return c->axis_mode != 0 && c->arr && (ints || f32) && !c->slab && c->sampler == SPV_SAMPLER_TMU && c->linear &&
         !(c->skipping > 0) && !c->persistent && !c->stats_on && c->tile_variant == 0 &&
         c->mip_path != SPV_MIP_PATH_SMEM && c->gnz == c->local_nz;
}

// Layer axis and lane-to-pixel map of one frame.  At the middle of the central ray's path through the box: gx / gy =
// texels between horizontally / vertically adjacent pixels' samples, gk = texels between a ray's consecutive samples.
// A quad request is cheapest when its four lanes share a layer and the ray stays in it; the weights are a fit to
// profiles/r02_exp_multiframe_v3.txt (per-frame time ~ 1 + 0.36 * layers spanned by a quad + 0.09 * layers per step,
// + 0.10 for column quads).  other_copies: the x / y copies may be used (built if they are not there yet).
static void choose_axis(spv_ctx *ctx, const float *invP, const float *invM, const float *box, int max_steps,
                        bool other_copies, int &lax_out, int &quad_out) {
  lax_out = 2;
  quad_out = 0;
  if (ctx->axis_mode >= 10) {  // forced (tests, experiments)
    const int v = ctx->axis_mode - 10;
    lax_out = (v / 3) % 3;
    quad_out = v % 3;
    if (ctx->axis_failed[lax_out]) {  // fall back to a copy that exists or can exist
      lax_out = 2;
      if (ctx->dtype == SPV_F32 && ctx->axis_failed[2]) lax_out = ctx->axis_failed[1] ? 0 : 1;
    }
    return;
  }
  const int W = ctx->width, H = ctx->height;
  auto ray = [&](double px, double py, double *o, double *d) {
    const double u = px / W * 2. - 1., v = py / H * 2. - 1.;
    const double front[4] = {u, v, -1., 1.}, back[4] = {u, v, 1., 1.};
    double a[4], b[4], t[4];
    for (int r = 0; r < 4; ++r) {
      a[r] = b[r] = 0.;
      for (int c = 0; c < 4; ++c) { a[r] += invP[4 * r + c] * front[c]; b[r] += invP[4 * r + c] * back[c]; }
    }
    for (int r = 0; r < 4; ++r) { a[r] /= a[3] != 0. ? a[3] : 1.; }
    const double bw = b[3] != 0. ? b[3] : 1.;
    double len = 0.;
    for (int r = 0; r < 4; ++r) { t[r] = b[r] / bw - a[r]; len += t[r] * t[r]; }
    len = len > 0. ? sqrt(len) : 1.;
    double ow = 0.;
    for (int c = 0; c < 4; ++c) ow += invM[12 + c] * a[c];
    if (ow == 0.) ow = 1.;
    for (int r = 0; r < 3; ++r) {
      o[r] = d[r] = 0.;
      for (int c = 0; c < 4; ++c) { o[r] += invM[4 * r + c] * a[c]; d[r] += invM[4 * r + c] * t[c] / len; }
      o[r] /= ow;
    }
  };
  double o[3], d[3], ox[3], dx[3], oy[3], dy[3];
  ray(W * .5, H * .5, o, d);
  ray(W * .5 + 1., H * .5, ox, dx);
  ray(W * .5, H * .5 + 1., oy, dy);
  double tn = -1e300, tf = 1e300, dd = 0., od = 0.;
  for (int r = 0; r < 3; ++r) {
    dd += d[r] * d[r];
    od += o[r] * d[r];
    if (d[r] != 0.) {
      const double t0 = (box[2 * r] - o[r]) / d[r], t1 = (box[2 * r + 1] - o[r]) / d[r];
      tn = fmax(tn, fmin(t0, t1));
      tf = fmin(tf, fmax(t0, t1));
    }
  }
  if (!(dd > 0.)) return;
  const bool hit = tf > tn && tf > 0.;
  if (tn < 0.) tn = 0.;
  const double tm = hit ? .5 * (tn + tf) : -od / dd;  // else: closest approach to the box centre
  const double dt = hit ? (tf - tn) / (double)((max_steps / 16) * 16) : 2. / max_steps;
  const double n[3] = {(double)ctx->nx, (double)ctx->ny, (double)ctx->gnz};
  double best = 1e300;
  const bool f32 = ctx->dtype == SPV_F32;
  if (f32) lax_out = ctx->axis_failed[2] ? (ctx->axis_failed[1] ? 0 : 1) : 2;
  for (int lax = 2; lax >= 0; --lax) {
    if (ctx->axis_failed[lax] || (lax != 2 && !other_copies && !f32)) continue;
    if ((lax == 0 ? ctx->nx : (lax == 1 ? ctx->ny : ctx->gnz)) > 2048) continue;  // layer count of a layered array
    const double gx = fabs((ox[lax] + tm * dx[lax]) - (o[lax] + tm * d[lax])) * .5 * n[lax];
    const double gy = fabs((oy[lax] + tm * dy[lax]) - (o[lax] + tm * d[lax])) * .5 * n[lax];
    const double gk = fabs(.5 * dt * d[lax]) * n[lax];
    const double own = (lax == 2 && !f32) ? 0. : .02;  // a copy that may have to be built must earn it
    const double cost[3] = {.36 * (gx + gy) + .09 * gk + own, .36 * 3. * gx + .09 * gk + own, .36 * 3. * gy + .09 * gk + .10 + own};
    for (int q = 0; q < 3; ++q)
      if (cost[q] < best) { best = cost[q]; lax_out = lax; quad_out = q; }
  }
}

// One max projection.  bands > 1 (fast kernel only): the frame is rendered as `bands` horizontal bands launched back to
// back, and the rows of a finished band travel to the pinned staging on the copy stream while the next band renders.
static int render_mip_impl(spv_ctx *ctx, const spv_mip_params *p, int bands, bool to_host, const PushArgs *push = nullptr) {
  if (!p) return fail(ctx, SPV_EINVAL, "spv_render_mip: null params");
  if (!ctx->arr) return fail(ctx, SPV_ENODATA, "spv_render_mip: no volume set");
  if (p->num_parts < 1 || p->current_part < 0 || p->max_steps / p->num_parts < 16)
    return fail(ctx, SPV_EINVAL, "spv_render_mip: need num_parts >= 1 and max_steps/num_parts >= 16");
  if (bad_float(p->alpha_pow) || bad_float(p->gamma)) return fail(ctx, SPV_EINVAL, "spv_render_mip: NaN parameter");
  const bool raw_only = (p->flags & SPV_MIP_RAW_ONLY) != 0;
  const bool exact = ctx->sampler == SPV_SAMPLER_EXACT;
  const bool fast = !exact;                // texture-unit kernels: mip_fast_kernel, or mip_alpha_kernel when attenuated
  const bool plain = fast && p->alpha_pow == 0.f;
  if ((raw_only || ctx->slab) && !plain)
    return fail(ctx, SPV_EINVAL, "spv_render_mip: slab / raw renders need the TMU sampler and alpha_pow == 0");
  if (raw_only && p->num_parts != 1) return fail(ctx, SPV_EINVAL, "spv_render_mip: raw renders need num_parts == 1");
  MipArgs a;
  a.cam = ctx->cam;
  a.vol = volume_of(ctx);
  a.coarse = ctx->coarse;
  a.cgx = ctx->cgx; a.cgy = ctx->cgy; a.cgz = ctx->cgz;
  memcpy(a.box, p->box, sizeof a.box);
  a.min_val = p->min_val; a.max_val = p->max_val; a.gamma = p->gamma; a.alpha_pow = p->alpha_pow;
  a.num_parts = p->num_parts; a.current_part = p->current_part; a.max_steps = p->max_steps; a.flags = p->flags;
  a.tile_variant = ctx->tile_variant;
  a.width = ctx->width; a.height = ctx->height;
  a.out = ctx->out(); a.alpha = ctx->alpha(); a.raw = ctx->raw();
  a.stats = ctx->stats_on ? ctx->d_stats : nullptr;
  a.tile_counter = ctx->persistent ? ctx->d_tile_counter : nullptr;
  a.band_done = nullptr;
  a.band_rows = 0;
  a.row_mode = 0;
  a.hit_tile_a = a.hit_tile_b = 0;
  a.n_extra = 0;
  if (ctx->slab && ctx->n_extra > 0) {
    if (ctx->skipping > 0) return fail(ctx, SPV_EINVAL, "spv_render_mip: extra slabs need empty-space skipping off");
    for (int i = 0; i < ctx->n_extra; ++i) {
      const spv_ctx *o = ctx->extra[i];
      if (!o->arr || !o->slab || o->dtype != ctx->dtype || o->layout != ctx->layout || o->nx != ctx->nx || o->ny != ctx->ny ||
          o->gnz != ctx->gnz || o->linear != ctx->linear || o->int_filter != ctx->int_filter)
        return fail(ctx, SPV_EINVAL, "spv_render_mip: an extra slab does not match this context's volume (extent, dtype, layout, filter)");
      a.extra[a.n_extra++] = volume_of(o);
    }
  }
  memset(&a.push, 0, sizeof a.push);
  if (push) {
    a.push = *push;
    a.flags |= SPV_MIP_RAW_ONLY | SPV_MIP_PUSH;
  } else {
    a.flags &= ~SPV_MIP_PUSH;
  }
  const bool linear = ctx->linear && (ctx->dtype == SPV_F32 || ctx->int_filter);
  int rc = (plain && ctx->skipping > 0) ? ensure_bricks(ctx) : 0;
  if (rc) return rc;
  // software-sampled path (spv_set_mip_path): whole-frame launches of plain uint16 max projections
  const bool smem = ctx->mip_path == SPV_MIP_PATH_SMEM && plain && linear && !raw_only && !push && !(ctx->skipping > 0) &&
                    p->num_parts == 1 && p->current_part == 0 && !ctx->persistent && smem_path_possible(ctx);
  ctx->last_mip_path = smem ? SPV_MIP_PATH_SMEM : SPV_MIP_PATH_TMU;
  if (smem) {
    rc = ensure_linear(ctx);
    if (rc) return rc;
    bands = 1;
  }
  // view-aligned layered copy + lane map of this frame (spv_mip_axis.cu): a function of the camera alone, so a view
  // renders to the same bits whatever was rendered before it
  const bool axis = fast && linear && !raw_only && !push && !smem && p->num_parts == 1 && p->current_part == 0 &&
                    axis_path_possible(ctx);  // plain and attenuated projections alike
  int lax = 2, quad = 0;
  if (axis) {
    for (int tries = 0; tries < 3; ++tries) {  // a copy that cannot be allocated drops out of the choice
      choose_axis(ctx, ctx->cam.invP, ctx->cam.invM, p->box, p->max_steps, ctx->axis_mode != 2, lax, quad);
      rc = ensure_axis(ctx, lax);
      if (rc && !ctx->axis_failed[lax]) return rc;
      if (!ctx->axis_failed[lax]) break;
    }
  }
  const bool axis_ok = axis && !ctx->axis_failed[lax];  // (float32 volumes: all three copies may have failed)
  ctx->clip_s[ctx->slot].valid = false;
  ctx->last_axis = axis_ok ? lax : -1;
  ctx->last_quad = axis_ok ? quad : -1;
  MipAxisArgs ax;
  if (axis_ok) {
    memset(&ax, 0, sizeof ax);
    memcpy(ax.invP, ctx->cam.invP, sizeof ax.invP);
    memcpy(ax.invM[0], ctx->cam.invM, sizeof ax.invM[0]);
    ax.tex[0] = ctx->axis_tex[0]; ax.tex[1] = ctx->axis_tex[1];
    ax.tex[2] = ctx->dtype == SPV_F32 ? ctx->axis_tex[2] : ctx->tex_lin;
    ax.lax[0] = (unsigned char)lax; ax.quad[0] = (unsigned char)quad;
    ax.nx = ctx->nx; ax.ny = ctx->ny; ax.nz = ctx->gnz;
    ax.scale = a.vol.scale;
    memcpy(ax.box, p->box, sizeof ax.box);
    ax.min_val = p->min_val; ax.max_val = p->max_val; ax.gamma = p->gamma; ax.max_steps = p->max_steps;
    ax.alpha_pow = p->alpha_pow;
    ax.width = ctx->width; ax.height = ctx->height; ax.n_frames = 1;
  }
  // the kernel of this frame (or band of it): a carries what varies between the launches below
  auto launch_frame = [&](const MipArgs &m, cudaStream_t st) -> cudaError_t {
    if (!axis_ok) return launch_mip(m, fmt_of(ctx), linear, fast, exact, ctx->skipping > 0, ctx->slab, ctx->stats_on != 0, st);
    ax.out[0] = m.out; ax.alpha[0] = m.alpha;
    ax.y_begin = m.y_begin; ax.y_end = m.y_end;
    ax.band_done = m.band_done; ax.band_rows = m.band_rows; ax.row_mode = m.row_mode;
    ax.hit_tile_a = m.hit_tile_a; ax.hit_tile_b = m.hit_tile_b;
    ax.tile_x0[0] = ax.tile_y0[0] = 0;  // every tile of the band: callers of single frames own the planes' contents
    ax.tile_nx[0] = (unsigned short)((ctx->width + 15) / 16);
    ax.tile_ny[0] = (unsigned short)((m.y_end - m.y_begin + 7) / 8);
    ax.rend_x0[0] = ax.rend_y0[0] = 0;
    ax.rend_x1[0] = ax.rend_y1[0] = 0xffff;
    return launch_mip_axis(ax, ctx->dtype, st);
  };
  rc = begin_render(ctx);
  if (rc) return rc;
  const int s = ctx->slot;
  rc = join_post(ctx, s);  // an iso frame's screen-space passes may still be writing this slot beside the render stream
  if (rc) return rc;
  // overlap (tuning knob 15): frames that alternate between the two output slots -- render_sequence, bench.py -- put slot
  // 1's kernel on post_stream, so that it starts while slot 0's kernel is in its tail (SMs are busy 88-93 % of a launch)
  // and the other way round.  Readers of slot 1 wait for ev_posted[1] exactly as for an iso frame's passes.
  const bool side = ctx->mip_overlap && s == 1 && plain && !smem && !ctx->slab && !raw_only && !push && p->num_parts == 1 &&
                    !(ctx->skipping > 0) && !ctx->persistent && !ctx->stats_on && !ctx->direct_host;
  cudaStream_t kst = ctx->stream;  // the stream the kernel(s) of this frame run on
  if (side) {
    kst = ctx->post_stream;
    if (ctx->side_seq != ctx->upload_seq) {  // an upload on the render stream that post_stream has not waited for yet
      CU(cudaEventRecord(ctx->ev_uploaded, ctx->stream));
      CU(cudaStreamWaitEvent(kst, ctx->ev_uploaded, 0));
      ctx->side_seq = ctx->upload_seq;
    }
    if (ctx->copy_pending[s]) CU(cudaStreamWaitEvent(kst, slot_free_event(ctx, s), 0));  // an asynchronous read of this slot
  }
  if (to_host && ctx->copy_pending[s]) CU(cudaEventSynchronize(ctx->ev_copied[s]));  // staging about to be rewritten
  const int H = ctx->height;
  // read-back: columns [clip_xa, clip_xb) of the copied rows (the projected box's rectangle where rows are clipped too)
  int clip_xa = 0, clip_xb = ctx->width;
  auto clip_columns = [&]() {
    int xa, xb, ya, yb;
    miss_free_rect(a.cam, a.box, ctx->width, H, xa, xb, ya, yb);
    clip_xa = xa;
    clip_xb = xb;
  };
  // rows [c0, c1) x those columns of the value and alpha planes in one 3-D copy (depth 2 = the two planes)
  auto copy_rows = [&](int c0, int c1, cudaStream_t cs) -> int {
    if (c0 >= c1 || clip_xa >= clip_xb) return 0;
    const size_t Wd = (size_t)ctx->width;
    cudaMemcpy3DParms cp;
    memset(&cp, 0, sizeof cp);
    cp.srcPtr = make_cudaPitchedPtr(ctx->dbuf_s[ctx->slot], Wd * sizeof(float), Wd, (size_t)H);
    cp.dstPtr = make_cudaPitchedPtr(ctx->hpin_s[ctx->slot], Wd * sizeof(float), Wd, (size_t)H);
    cp.srcPos = make_cudaPos((size_t)clip_xa * sizeof(float), (size_t)c0, 0);
    cp.dstPos = cp.srcPos;
    cp.extent = make_cudaExtent((size_t)(clip_xb - clip_xa) * sizeof(float), (size_t)(c1 - c0), 2);
    cp.kind = cudaMemcpyDeviceToHost;
    CU(cudaMemcpy3DAsync(&cp, cs));
    ctx->d2h_bytes += 2 * (size_t)(clip_xb - clip_xa) * (size_t)(c1 - c0) * sizeof(float);
    return 0;
  };
  const bool direct = to_host && ctx->direct_host && plain && p->num_parts == 1 && !raw_only;
  if (direct) {  // zero-copy: the result planes are written over PCIe by the kernel's own 128-bit stores
    staging_dirty(ctx, s);
    a.out = ctx->hpin_s[s];
    a.alpha = ctx->hpin_s[s] + ctx->n();
    bands = 1;
  }
  if (!plain || bands < 1) bands = 1;
  if (bands > 64) bands = 64;
  int rows = ((H + bands - 1) / bands + 15) / 16 * 16;  // band height: a multiple of every CTA tile height
  if (rows < 16) rows = 16;
  // One launch for the whole frame where the driver offers stream memory operations: the CTAs count themselves into
  // their band's counter and the copy stream waits for each counter to reach the band's CTA total (cuStreamWaitValue32)
  // before it moves the band -- no per-band launch, no per-band tail, the bands can be small.
  if (to_host && !direct && bands > 1 && ctx->tile_variant == 0 && !ctx->persistent && !ctx->slab && !raw_only &&
      p->num_parts == 1 && !(ctx->skipping > 0) && wait_value_fn()) {
    a.y_begin = 0;
    a.y_end = H;
    a.band_done = ctx->d_band_done;
    a.band_rows = rows;
    const int nb = (H + rows - 1) / rows;
    int order[64];
    int clip_a = 0, clip_b = H;  // rows that are copied; the others are known to be misses and are not
    bool clean_after_launch = false;
    if (ctx->row_mode == 1) {
      // rows the box cannot project to are dealt first, then its rows top to bottom: bands without any of those rows
      // complete at once, the others in ascending order
      a.row_mode = 1;
      hit_tile_rows(a.cam, a.box, H, (unsigned)((H + 7) / 8), a.hit_tile_a, a.hit_tile_b);
      const int ya = (int)a.hit_tile_a * 8, yb = (int)a.hit_tile_b * 8;
      if (ctx->clip_copies) {
        // A ray of a row outside the hull of the projected corners runs along a line that does not meet the box, so
        // it is a miss: out 0, alpha 0 (integer volumes) / -1 (float32), which the staging rows hold already.  The
        // hull is taken in double precision with one pixel of slack; one more tile row on either side here (the
        // kernel's fp32 slab test can differ from the exact geometry by ~1e-4 pixel at most).
        clip_a = ya - 8 > 0 ? ya - 8 : 0;
        clip_b = yb + 8 < H ? yb + 8 : H;
        if (clip_a >= clip_b) clip_a = clip_b = 0;
        clip_columns();
        clean_after_launch = true;  // (outside the rectangle the copies write: the host refills it beside the kernel)
      } else {
        staging_dirty(ctx, s);
      }
      int k = 0;
      for (int b = 0; b < nb; ++b)
        if (b * rows + rows <= ya || b * rows >= yb) order[k++] = b;
      for (int b = 0; b < nb; ++b)
        if (!(b * rows + rows <= ya || b * rows >= yb)) order[k++] = b;
    } else {
      // the kernel deals tile rows from the top and bottom edges inwards: the bands complete in the order
      // 0, nb-1, 1, nb-2, ...
      for (int i = 0; i < nb; ++i) order[i] = (i & 1) ? nb - 1 - (i >> 1) : (i >> 1);
      staging_dirty(ctx, s);
    }
    CU(launch_frame(a, kst));
    ctx->launches += 1;
    if (side) {
      CU(cudaEventRecord(ctx->ev_posted[s], kst));
      ctx->post_pending[s] = true;
    }
    if (clean_after_launch)
      staging_clean_outside_rect(ctx, s, clip_xa, clip_xb, clip_a, clip_b, ctx->dtype == SPV_F32 ? -1.f : 0.f);
    const unsigned ctas_x = (unsigned)(ctx->width + 15) / 16;
    for (int i = 0; i < nb; ++i) {
      // enqueue the copies in the order the bands complete, alternating between the copy streams
      const int b = order[i];
      const int y0 = b * rows, y1 = y0 + rows < H ? y0 + rows : H;
      ctx->band_expect[b] += ctas_x * (unsigned)((y1 - y0 + 7) / 8);
      const int c0 = y0 > clip_a ? y0 : clip_a, c1 = y1 < clip_b ? y1 : clip_b;  // the band's rows that can hold hits
      if (c0 >= c1) continue;
      cudaStream_t cs = (ctx->copy_streams > 1 && (i & 1)) ? ctx->copy_stream2 : ctx->copy_stream;
      CUresult wr = wait_value_fn()(cs, (CUdeviceptr)(uintptr_t)(ctx->d_band_done + b), ctx->band_expect[b],
                                    CU_STREAM_WAIT_VALUE_GEQ);
      if (wr != CUDA_SUCCESS) return fail(ctx, (int)wr, "cuStreamWaitValue32 failed");
      rc = copy_rows(c0, c1, cs);
      if (rc) return rc;
    }
    ctx->last_method = 0;
    if (!raw_only) slot_clip_set(ctx, s, a.cam, a.box, (ctx->dtype == SPV_F32 ? -1.f : 0.f));
    rc = end_render(ctx);
    if (rc) return rc;
    if (ctx->copy_streams > 1) {
      CU(cudaEventRecord(ctx->ev_copy2, ctx->copy_stream2));
      CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_copy2, 0));
    }
    CU(cudaEventRecord(ctx->ev_copied[s], ctx->copy_stream));
    ctx->copy_pending[s] = true;
    return 0;
  }
  // band-per-launch path (and the single band of pipelined sequences): the same row clipping as above
  int clip_a = 0, clip_b = H;
  if (to_host && !direct) {
    if (ctx->clip_copies && ctx->row_mode == 1 && !ctx->slab && !raw_only && p->num_parts == 1) {
      unsigned ta, tb;
      hit_tile_rows(a.cam, a.box, H, (unsigned)((H + 7) / 8), ta, tb);
      clip_a = (int)ta * 8 - 8 > 0 ? (int)ta * 8 - 8 : 0;
      clip_b = (int)tb * 8 + 8 < H ? (int)tb * 8 + 8 : H;
      if (clip_a >= clip_b) clip_a = clip_b = 0;
      clip_columns();
      staging_clean_outside_rect(ctx, s, clip_xa, clip_xb, clip_a, clip_b, ctx->dtype == SPV_F32 ? -1.f : 0.f);
    } else {
      staging_dirty(ctx, s);
    }
  }
  for (int y0 = 0; y0 < H; y0 += rows) {
    const int y1 = y0 + rows < H ? y0 + rows : H;
    a.y_begin = y0;
    a.y_end = y1;
    if (smem) {
      a.tile_counter = ctx->d_tile_counter;
      CU(launch_mip_smem(a, fmt_of(ctx), ctx->smem_cfg, ctx->tmaps[ctx->smem_cfg], ctx->smem_tex_of8, ctx->stream));
    } else
      CU(launch_frame(a, kst));
    ctx->launches += 1;
    if (side) {
      CU(cudaEventRecord(ctx->ev_posted[s], kst));
      ctx->post_pending[s] = true;
    }
    const int c0 = y0 > clip_a ? y0 : clip_a, c1 = y1 < clip_b ? y1 : clip_b;
    if (to_host && !direct && c0 < c1) {
      CU(cudaEventRecord(ctx->ev_rendered[s], kst));
      CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_rendered[s], 0));
      rc = copy_rows(c0, c1, ctx->copy_stream);  // the band's part of the value and alpha planes in one copy
      if (rc) return rc;
    }
  }
  ctx->last_method = 0;
  if (!raw_only && !push) slot_clip_set(ctx, s, a.cam, a.box, (ctx->dtype == SPV_F32 ? -1.f : 0.f));
  rc = end_render(ctx);
  if (rc) return rc;
  if (to_host) {
    CU(cudaEventRecord(ctx->ev_copied[s], direct ? ctx->stream : ctx->copy_stream));
    ctx->copy_pending[s] = true;
  }
  return 0;
}

SPV_API int spv_render_mip(spv_ctx *ctx, const spv_mip_params *p) {
  BIND();
  return render_mip_impl(ctx, p, 1, false);
}

SPV_API int spv_render_mip_to_host(spv_ctx *ctx, const spv_mip_params *p, int bands, int wait, float **host) {
  BIND();
  if (p && (p->flags & SPV_MIP_RAW_ONLY)) return fail(ctx, SPV_EINVAL, "spv_render_mip_to_host: not for raw renders");
  if (bands <= 0) bands = ctx->bands > 0 ? ctx->bands : (wait_value_fn() ? 12 : 2);
  int rc = render_mip_impl(ctx, p, bands, true);
  if (rc) return rc;
  if (wait) {
    CU(cudaEventSynchronize(ctx->ev_copied[ctx->slot]));
    CU(cudaStreamSynchronize(ctx->stream));  // statistics / timing events of this frame
  }
  if (host) *host = ctx->hpin_s[ctx->slot];
  return 0;
}

// ---- several frames per launch --------------------------------------------------------------------------------------
static int ensure_batch(spv_ctx *ctx, int n) {
  if (n <= ctx->batch_cap) return 0;
  // room for the largest launch at once (a sequence's first launch may be a short one: no reallocation -- which waits
  // for everything in flight -- in the middle of it), unless the planes are huge
  if ((size_t)MAX_BATCH * 2 * ctx->n() * sizeof(float) <= ((size_t)256 << 20)) n = MAX_BATCH;
  CU(cudaStreamSynchronize(ctx->stream));
  CU(cudaStreamSynchronize(ctx->copy_stream));
  CU(cudaStreamSynchronize(ctx->copy_stream2));
  free_batch(ctx);
  const size_t bytes = (size_t)n * 2 * ctx->n() * sizeof(float);
  for (int s = 0; s < 2; ++s) {
    CU(cudaMalloc(&ctx->d_batch[s], bytes));
    CU(cudaMallocHost(&ctx->h_batch[s], bytes));
    memset(ctx->h_batch[s], 0, bytes);  // every pixel holds the miss values of integer volumes (out 0, alpha 0)
    CU(cudaMemsetAsync(ctx->d_batch[s], 0, bytes, ctx->stream));
    for (int f = 0; f < MAX_BATCH; ++f) ctx->batch_h_dirty[s][f] = ctx->batch_d_dirty[s][f] = spv_ctx::Rect{0, 0, 0, 0};
  }
  ctx->batch_miss_alpha = 0.f;
  ctx->batch_cap = n;
  return 0;
}

// the alpha planes of the batch sets hold `miss` outside the tracked rectangles: when the element type of the volume
// changes the miss value (0 for integer volumes, -1 for float32), everything counts as dirty once
static void batch_miss_value(spv_ctx *ctx, float miss) {
  if (ctx->batch_miss_alpha == miss) return;
  const size_t np = ctx->n();
  for (int s = 0; s < 2; ++s)
    for (int f = 0; f < ctx->batch_cap && f < MAX_BATCH; ++f) {
      // pinned planes: refilled here, once (the copies out of them are quiescent: the caller has synchronised);
      // device planes: everything counts as dirty, the next launch into a slot covers the whole image
      float *hf = ctx->h_batch[s] + (size_t)f * 2 * np;
      memset(hf, 0, np * sizeof(float));
      std::fill_n(hf + np, np, miss);
      ctx->batch_h_dirty[s][f] = spv_ctx::Rect{0, 0, 0, 0};
      ctx->batch_d_dirty[s][f] = spv_ctx::Rect{0, ctx->width, 0, ctx->height};
    }
  ctx->batch_miss_alpha = miss;
}

// n (<= SPV_MAX_BATCH) max projections (plain or attenuated) of the resident integer volume that differ in their model view only, in ONE
// launch (mip_axis_kernel): the frames' CTAs of one tile row run together, so the frames share the volume in L2.  Results
// go to one of two sets of planes, [n][out | alpha] floats, alternating from call to call; with to_host the rows the
// projected box can touch travel to the set's pinned planes behind the launch (the other rows hold the miss values), and
// the next call's launch overlaps that copy.  Enqueue-only; spv_batch_wait gives the pointers of a set.
SPV_API int spv_render_mip_batch(spv_ctx *ctx, const spv_mip_params *p, const float *invM, int n, int to_host, int *set_out) {
  BIND();
  if (!p || !invM || !set_out) return fail(ctx, SPV_EINVAL, "spv_render_mip_batch: null argument");
  if (n < 1 || n > MAX_BATCH) return fail(ctx, SPV_EINVAL, "spv_render_mip_batch: 1 <= n <= SPV_MAX_BATCH frames per launch");
  if (!ctx->arr) return fail(ctx, SPV_ENODATA, "spv_render_mip_batch: no volume set");
  if (p->num_parts != 1 || p->current_part != 0 || p->max_steps < 16 || p->flags != 0)
    return fail(ctx, SPV_EINVAL, "spv_render_mip_batch: one part, no flags, max_steps >= 16");
  if (bad_float(p->gamma) || bad_float(p->alpha_pow)) return fail(ctx, SPV_EINVAL, "spv_render_mip_batch: NaN parameter");
  if (!axis_path_possible(ctx))
    return fail(ctx, SPV_EINVAL, "spv_render_mip_batch: needs a resident integer volume in the paired layout, the TMU sampler "
                                 "(or a float32 volume) with linear filtering, and no skipping / statistics / slab / software-sampled path");
  int rc = ensure_batch(ctx, n);
  if (rc) return rc;
  const float miss_alpha = ctx->dtype == SPV_F32 ? -1.f : 0.f;
  const int set = ctx->batch_set ^ 1;
  MipAxisArgs ax;
  memset(&ax, 0, sizeof ax);
  memcpy(ax.invP, ctx->cam.invP, sizeof ax.invP);
  memcpy(ax.invM, invM, (size_t)n * 16 * sizeof(float));
  for (int f = 0; f < n; ++f) {
    int lax = 2, quad = 0;
    for (int tries = 0; tries < 3; ++tries) {  // a copy that cannot be allocated drops out of the choice
      choose_axis(ctx, ctx->cam.invP, invM + 16 * f, p->box, p->max_steps, ctx->axis_mode != 2, lax, quad);
      rc = ensure_axis(ctx, lax);
      if (rc && !ctx->axis_failed[lax]) return rc;
      if (!ctx->axis_failed[lax]) break;
    }
    if (ctx->axis_failed[lax]) return fail(ctx, SPV_EINVAL, "spv_render_mip_batch: no room for a layered copy of this float32 volume");
    ax.lax[f] = (unsigned char)lax;
    ax.quad[f] = (unsigned char)quad;
    ax.out[f] = ctx->d_batch[set] + (size_t)f * 2 * ctx->n();
    ax.alpha[f] = ax.out[f] + ctx->n();
  }
  ctx->last_axis = ax.lax[0];
  ctx->last_quad = ax.quad[0];
  ax.tex[0] = ctx->axis_tex[0]; ax.tex[1] = ctx->axis_tex[1];
  ax.tex[2] = ctx->dtype == SPV_F32 ? ctx->axis_tex[2] : ctx->tex_lin;
  ax.nx = ctx->nx; ax.ny = ctx->ny; ax.nz = ctx->gnz;
  ax.scale = ctx->dtype == SPV_F32 ? 1.f : (ctx->dtype == SPV_U16 ? 65535.f : 255.f);
  memcpy(ax.box, p->box, sizeof ax.box);
  ax.min_val = p->min_val; ax.max_val = p->max_val; ax.gamma = p->gamma; ax.max_steps = p->max_steps;
  ax.alpha_pow = p->alpha_pow;
  ax.width = ctx->width; ax.height = ctx->height; ax.n_frames = n;
  ax.y_begin = 0; ax.y_end = ctx->height;
  rc = join_post(ctx, -1);  // single frames beside the render stream may still read a copy that is about to be rebuilt
  if (rc) return rc;
  if (ctx->batch_copy_pending[set]) {  // the copies out of this set's planes, two calls ago
    CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_batch_copied[set], 0));
    if (to_host) CU(cudaEventSynchronize(ctx->ev_batch_copied[set]));  // the pinned planes are about to be cleaned by the host
    ctx->batch_copy_pending[set] = false;
  }
  if (ctx->batch_miss_alpha != miss_alpha) {  // the other set's copies may still be running: they read planes we re-clean
    CU(cudaStreamSynchronize(ctx->copy_stream));
    CU(cudaStreamSynchronize(ctx->copy_stream2));
    batch_miss_value(ctx, miss_alpha);
  }
  rc = begin_render(ctx);
  if (rc) return rc;
  // Only the CTA tiles the projected box can touch are launched (a third of the image on configs[1]; the ray setup of a
  // tile that misses costs as much as that of one that hits).  Every pixel outside a frame's rectangle is a miss: the
  // device planes hold the miss values there -- what an earlier frame of this slot left outside is cleared first.
  const int W = ctx->width, H = ctx->height;
  const size_t np = ctx->n();
  spv_ctx::Rect rect[MAX_BATCH];
  for (int f = 0; f < n; ++f) {
    Camera cam;
    memcpy(cam.invP, ctx->cam.invP, sizeof cam.invP);
    memcpy(cam.invM, invM + 16 * f, sizeof cam.invM);
    spv_ctx::Rect &r = rect[f];
    if (ctx->clip_copies) miss_free_rect(cam, p->box, W, H, r.x0, r.x1, r.y0, r.y1);
    else r = spv_ctx::Rect{0, W, 0, H};
    // launched: the tiles of the new rectangle and of the old one (whatever lies outside the new one is cleared by the
    // kernel: a tile of zeros costs a few stores, a separate memset per frame costs a launch)
    spv_ctx::Rect &d = ctx->batch_d_dirty[set][f];
    spv_ctx::Rect u = r;
    if (d.x1 > d.x0 && d.y1 > d.y0) {
      if (u.x1 > u.x0 && u.y1 > u.y0) {
        u.x0 = d.x0 < u.x0 ? d.x0 : u.x0; u.x1 = d.x1 > u.x1 ? d.x1 : u.x1;
        u.y0 = d.y0 < u.y0 ? d.y0 : u.y0; u.y1 = d.y1 > u.y1 ? d.y1 : u.y1;
      } else {
        u = d;
      }
    }
    ax.tile_x0[f] = (unsigned short)(u.x0 / 16);
    ax.tile_y0[f] = (unsigned short)(u.y0 / 8);
    ax.tile_nx[f] = (unsigned short)((u.x1 + 15) / 16 - u.x0 / 16);
    ax.tile_ny[f] = (unsigned short)((u.y1 + 7) / 8 - u.y0 / 8);
    ax.rend_x0[f] = (unsigned short)(r.x0 / 16);
    ax.rend_x1[f] = (unsigned short)((r.x1 + 15) / 16);
    ax.rend_y0[f] = (unsigned short)(r.y0 / 8);
    ax.rend_y1[f] = (unsigned short)((r.y1 + 7) / 8);
    d = r;
  }
  {  // the same tile rows for every frame: the frames' CTAs that run together cross the same slab of the volume
    unsigned y0 = 0xffff, y1 = 0;
    for (int f = 0; f < n; ++f)
      if (ax.tile_ny[f] && ax.tile_nx[f]) {
        y0 = ax.tile_y0[f] < y0 ? ax.tile_y0[f] : y0;
        y1 = (unsigned)ax.tile_y0[f] + ax.tile_ny[f] > y1 ? (unsigned)ax.tile_y0[f] + ax.tile_ny[f] : y1;
      }
    for (int f = 0; f < n; ++f)
      if (ax.tile_ny[f] && ax.tile_nx[f]) {  // rows added to a frame lie outside its rectangle: filled with the miss values
        ax.tile_y0[f] = (unsigned short)y0;
        ax.tile_ny[f] = (unsigned short)(y1 - y0);
      }
  }
  // Read-back behind the launch, in bands of tile rows where the driver offers stream memory operations: the CTAs count
  // themselves into their band's counter (all frames' CTAs of a tile row run together, so a band of every frame completes
  // at about the same time) and the copy streams wait for a counter before they move that band of every frame -- the
  // launch's own copies overlap its rendering, not only the next launch's.
  int nb = 1, band_rows = H;
  unsigned launch_y0 = 0, launch_ny = 0;
  for (int f = 0; f < n; ++f)
    if (ax.tile_ny[f] && ax.tile_nx[f]) { launch_y0 = ax.tile_y0[f]; launch_ny = ax.tile_ny[f]; }
  const bool banded = to_host && launch_ny > 0 && wait_value_fn() != nullptr && ctx->bands != 1;
  if (banded) {
    nb = ctx->bands > 0 ? ctx->bands : 4;
    if (nb > 16) nb = 16;
    band_rows = (((int)launch_ny * 8 + nb - 1) / nb + 15) / 16 * 16;
    nb = ((int)launch_ny * 8 + band_rows - 1) / band_rows;
    ax.band_done = ctx->d_band_done;
    ax.band_rows = band_rows;
    ax.row_mode = 2;  // tile rows in their natural order
    for (int b = 0; b < nb; ++b) {
      const unsigned t0 = (unsigned)(b * band_rows / 8), t1 = (unsigned)((b + 1) * band_rows / 8);
      const unsigned rows = (t1 < launch_ny ? t1 : launch_ny) - t0;
      for (int f = 0; f < n; ++f)
        if (ax.tile_ny[f] && ax.tile_nx[f]) ctx->band_expect[b] += (unsigned)ax.tile_nx[f] * rows;
    }
  }
  CU(launch_mip_axis(ax, ctx->dtype, ctx->stream));
  ctx->launches += 1;
  ctx->last_method = 0;
  rc = end_render(ctx);
  if (rc) return rc;
  CU(cudaEventRecord(ctx->ev_batch_rendered[set], ctx->stream));
  ctx->batch_render_pending[set] = true;
  ctx->batch_n[set] = n;
  ctx->batch_set = set;
  *set_out = set;
  if (!to_host) return 0;
  if (!banded) {
    CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_batch_rendered[set], 0));
    if (ctx->copy_streams > 1) CU(cudaStreamWaitEvent(ctx->copy_stream2, ctx->ev_batch_rendered[set], 0));
  }
  for (int f = 0; f < n; ++f) {  // the pinned planes: what an earlier frame left outside this frame's rectangle is cleared
    const spv_ctx::Rect &r = rect[f];
    float *hf = ctx->h_batch[set] + (size_t)f * 2 * np;
    spv_ctx::Rect &d = ctx->batch_h_dirty[set][f];
    if (!(d.x0 >= r.x0 && d.x1 <= r.x1 && d.y0 >= r.y0 && d.y1 <= r.y1))
      for (int y = d.y0; y < d.y1; ++y) {
        const bool row_in = y >= r.y0 && y < r.y1;
        const int segs[2][2] = {{d.x0, row_in ? (d.x1 < r.x0 ? d.x1 : r.x0) : d.x1}, {row_in ? (d.x0 > r.x1 ? d.x0 : r.x1) : d.x1, d.x1}};
        for (int k = 0; k < 2; ++k)
          if (segs[k][0] < segs[k][1]) {
            memset(hf + (size_t)y * W + segs[k][0], 0, (size_t)(segs[k][1] - segs[k][0]) * sizeof(float));
            float *al = hf + np + (size_t)y * W;
            if (miss_alpha == 0.f) memset(al + segs[k][0], 0, (size_t)(segs[k][1] - segs[k][0]) * sizeof(float));
            else std::fill(al + segs[k][0], al + segs[k][1], miss_alpha);
          }
      }
    d = r;
  }
  for (int b = 0; b < nb; ++b) {
    const int by0 = banded ? (int)launch_y0 * 8 + b * band_rows : 0, by1 = banded ? by0 + band_rows : H;
    if (banded)
      for (int k = 0; k < (ctx->copy_streams > 1 ? 2 : 1); ++k) {
        CUresult wr = wait_value_fn()(k ? ctx->copy_stream2 : ctx->copy_stream, (CUdeviceptr)(uintptr_t)(ctx->d_band_done + b),
                                      ctx->band_expect[b], CU_STREAM_WAIT_VALUE_GEQ);
        if (wr != CUDA_SUCCESS) return fail(ctx, (int)wr, "cuStreamWaitValue32 failed");
      }
    for (int f = 0; f < n; ++f) {
      const spv_ctx::Rect &r = rect[f];
      const int y0 = r.y0 > by0 ? r.y0 : by0, y1 = r.y1 < by1 ? r.y1 : by1;
      if (r.x1 <= r.x0 || y1 <= y0) continue;
      float *hf = ctx->h_batch[set] + (size_t)f * 2 * np;
      // both planes' rectangles in one 3-D copy (depth 2 = the two planes)
      cudaMemcpy3DParms cp;
      memset(&cp, 0, sizeof cp);
      cp.srcPtr = make_cudaPitchedPtr(ax.out[f], (size_t)W * sizeof(float), (size_t)W, (size_t)H);
      cp.dstPtr = make_cudaPitchedPtr(hf, (size_t)W * sizeof(float), (size_t)W, (size_t)H);
      cp.srcPos = make_cudaPos((size_t)r.x0 * sizeof(float), (size_t)y0, 0);
      cp.dstPos = cp.srcPos;
      cp.extent = make_cudaExtent((size_t)(r.x1 - r.x0) * sizeof(float), (size_t)(y1 - y0), 2);
      cp.kind = cudaMemcpyDeviceToHost;
      cudaStream_t cs = (ctx->copy_streams > 1 && (f & 1)) ? ctx->copy_stream2 : ctx->copy_stream;
      CU(cudaMemcpy3DAsync(&cp, cs));
      ctx->d2h_bytes += 2 * (size_t)(r.x1 - r.x0) * (size_t)(y1 - y0) * sizeof(float);
    }
  }
  if (ctx->copy_streams > 1) {
    CU(cudaEventRecord(ctx->ev_copy2, ctx->copy_stream2));
    CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_copy2, 0));
  }
  CU(cudaEventRecord(ctx->ev_batch_copied[set], ctx->copy_stream));
  ctx->batch_copy_pending[set] = true;
  return 0;
}

SPV_API int spv_mip_batch_possible(spv_ctx *ctx, const spv_mip_params *p) {
  if (!ctx || !p || !ctx->arr) return 0;
  if (p->num_parts != 1 || p->current_part != 0 || p->max_steps < 16 || p->flags != 0) return 0;
  return axis_path_possible(ctx) ? 1 : 0;
}

// Waits for set `set` of spv_render_mip_batch (its copies when it was rendered to the host, else its launch) and
// returns its planes: frame f's value plane at + f * 2 * width * height floats, its alpha plane one plane further.
SPV_API int spv_batch_wait(spv_ctx *ctx, int set, float **host, float **dev, int *n_frames) {
  BIND();
  if (set < 0 || set > 1 || !ctx->d_batch[set]) return fail(ctx, SPV_EINVAL, "spv_batch_wait: no such set");
  if (ctx->batch_copy_pending[set]) CU(cudaEventSynchronize(ctx->ev_batch_copied[set]));
  else if (ctx->batch_render_pending[set]) CU(cudaEventSynchronize(ctx->ev_batch_rendered[set]));
  ctx->batch_render_pending[set] = false;
  if (host) *host = ctx->h_batch[set];
  if (dev) *dev = ctx->d_batch[set];
  if (n_frames) *n_frames = ctx->batch_n[set];
  return 0;
}

// layer axis (0 x, 1 y, 2 z) and lane map (0: 2x2 quads, 1: row quads, 2: column quads) of the last plain projection;
// -1, -1 when it went through mip_fast_kernel
SPV_API int spv_mip_axis_used(spv_ctx *ctx, int *axis, int *quad) {
  if (!ctx) return SPV_EINVAL;
  if (axis) *axis = ctx->last_axis;
  if (quad) *quad = ctx->last_quad;
  return 0;
}

// ---- sort-last composite over peer memory -------------------------------------------------------------------------
SPV_API int spv_comp_init(spv_ctx *ctx, int rank, int world) {
  BIND();
  if (world < 1 || world > MAX_WORLD || rank < 0 || rank >= world)
    return fail(ctx, SPV_EINVAL, "spv_comp_init: need 0 <= rank < world <= 16");
  if (ctx->width % 4) return fail(ctx, SPV_EINVAL, "spv_comp_init: the image width must be a multiple of 4");
  CU(cudaStreamSynchronize(ctx->stream));
  free_comp(ctx);
  if (ctx->slot != 0) return fail(ctx, SPV_EINVAL, "spv_comp_init: select output slot 0 first");
  // CUDA loads a kernel at its first launch, and that load can wait for the kernels already running.  The composite's
  // kernels wait for each other: with several ranks in ONE process a first-time load behind a spinning comp_sync would
  // block the host thread that still has to enqueue the kernels the spin waits for.  Load them all now.
  CU(preload_mip_kernels());
  CU(preload_comp_kernels());
  CU(preload_iso_kernels());
  const int rows = ((ctx->height + world - 1) / world + 3) / 4 * 4;
  const size_t band = (size_t)rows * ctx->width;
  CU(cudaMalloc(&ctx->comp_part, 6 * (size_t)world * band * sizeof(float)));
  CU(cudaMalloc(&ctx->comp_flags, COMP_PHASES * MAX_WORLD * sizeof(unsigned)));
  CU(cudaMalloc(&ctx->comp_err, sizeof(unsigned)));
  CU(cudaMemsetAsync(ctx->comp_part, 0, 6 * (size_t)world * band * sizeof(float), ctx->stream));
  CU(cudaMemsetAsync(ctx->comp_flags, 0, COMP_PHASES * MAX_WORLD * sizeof(unsigned), ctx->stream));
  CU(cudaMemsetAsync(ctx->comp_err, 0, sizeof(unsigned), ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->comp_rank = rank;
  ctx->comp_world = world;
  ctx->comp_band_rows = rows;
  ctx->peer_part[rank] = ctx->comp_part;
  ctx->peer_flags[rank] = ctx->comp_flags;
  ctx->peer_out[rank] = ctx->dbuf_s[0];
  return 0;
}

SPV_API int spv_comp_export(spv_ctx *ctx, void *handles, size_t nbytes) {
  BIND();
  if (ctx->comp_world < 1) return fail(ctx, SPV_ENODATA, "spv_comp_export: spv_comp_init first");
  if (!handles || nbytes < 3 * sizeof(cudaIpcMemHandle_t)) return fail(ctx, SPV_EINVAL, "spv_comp_export: need 192 bytes");
  cudaIpcMemHandle_t *h = (cudaIpcMemHandle_t *)handles;
  CU(cudaIpcGetMemHandle(&h[0], ctx->comp_part));
  CU(cudaIpcGetMemHandle(&h[1], ctx->comp_flags));
  CU(cudaIpcGetMemHandle(&h[2], ctx->dbuf_s[0]));
  return 0;
}

SPV_API int spv_comp_import(spv_ctx *ctx, int peer, const void *handles, size_t nbytes) {
  BIND();
  if (ctx->comp_world < 1) return fail(ctx, SPV_ENODATA, "spv_comp_import: spv_comp_init first");
  if (peer < 0 || peer >= ctx->comp_world || peer == ctx->comp_rank) return fail(ctx, SPV_EINVAL, "spv_comp_import: bad peer rank");
  if (!handles || nbytes < 3 * sizeof(cudaIpcMemHandle_t)) return fail(ctx, SPV_EINVAL, "spv_comp_import: need 192 bytes");
  const cudaIpcMemHandle_t *h = (const cudaIpcMemHandle_t *)handles;
  for (int k = 0; k < 3; ++k) {
    if (ctx->ipc_opened[peer][k]) cudaIpcCloseMemHandle(ctx->ipc_opened[peer][k]);
    ctx->ipc_opened[peer][k] = nullptr;
    CU(cudaIpcOpenMemHandle(&ctx->ipc_opened[peer][k], h[k], cudaIpcMemLazyEnablePeerAccess));
  }
  ctx->peer_part[peer] = (float *)ctx->ipc_opened[peer][0];
  ctx->peer_flags[peer] = (unsigned *)ctx->ipc_opened[peer][1];
  ctx->peer_out[peer] = (float *)ctx->ipc_opened[peer][2];
  return 0;
}

SPV_API int spv_comp_import_local(spv_ctx *ctx, int peer, spv_ctx *other) {
  BIND();
  if (ctx->comp_world < 1 || !other || other->comp_world != ctx->comp_world || other->comp_rank != peer ||
      peer == ctx->comp_rank || other->width != ctx->width || other->height != ctx->height)
    return fail(ctx, SPV_EINVAL, "spv_comp_import_local: the peer context does not match (rank, world, image size)");
  if (other->device != ctx->device) {
    cudaError_t e = cudaDeviceEnablePeerAccess(other->device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
    else if (e != cudaSuccess) return cufail(ctx, e, "cudaDeviceEnablePeerAccess");
  }
  ctx->peer_part[peer] = other->comp_part;
  ctx->peer_flags[peer] = other->comp_flags;
  ctx->peer_out[peer] = other->dbuf_s[0];
  return 0;
}

SPV_API int spv_render_mip_composite(spv_ctx *ctx, const spv_mip_params *p) {
  BIND();
  const int W = ctx->comp_world, R = ctx->comp_rank;
  if (W < 1) return fail(ctx, SPV_ENODATA, "spv_render_mip_composite: spv_comp_init first");
  for (int r = 0; r < W; ++r)
    if (!ctx->peer_part[r]) return fail(ctx, SPV_ENODATA, "spv_render_mip_composite: a peer has not been imported");
  if (ctx->slot != 0) return fail(ctx, SPV_EINVAL, "spv_render_mip_composite: select output slot 0 first");
  if (p && (p->alpha_pow != 0.f || p->num_parts != 1))
    return fail(ctx, SPV_EINVAL, "spv_render_mip_composite: needs alpha_pow == 0 and num_parts == 1 (attenuation is order dependent)");
  const unsigned f = ++ctx->comp_frame;
  const unsigned parity = f & 1u;
  const size_t band = (size_t)ctx->comp_band_rows * ctx->width;
  PushArgs push;
  memset(&push, 0, sizeof push);
  for (int r = 0; r < W; ++r) push.part[r] = ctx->peer_part[r];
  push.band_rows = ctx->comp_band_rows;
  push.src_off = (unsigned)((parity * (unsigned)W + (unsigned)R) * band);
  int rc = render_mip_impl(ctx, p, 1, false, &push);
  if (rc) return rc;
  CU(launch_comp_sync(ctx->peer_flags, ctx->comp_flags, W, R, 0, f, ctx->comp_err, ctx->stream));
  CompFinishArgs fa;
  memset(&fa, 0, sizeof fa);
  fa.part = ctx->comp_part + (size_t)parity * W * band;
  for (int r = 0; r < W; ++r) fa.out[r] = ctx->peer_out[r];
  fa.world = W;
  fa.band_pixels = (unsigned)band;
  fa.first_pixel = (unsigned)((size_t)R * band);
  const size_t total = ctx->n();
  fa.n_pixels = fa.first_pixel >= total ? 0u : (unsigned)((total - fa.first_pixel) < band ? (total - fa.first_pixel) : band);
  fa.min_val = p->min_val; fa.max_val = p->max_val; fa.gamma = p->gamma;
  CU(launch_comp_finish(fa, ctx->stream));
  CU(launch_comp_sync(ctx->peer_flags, ctx->comp_flags, W, R, 1, f, ctx->comp_err, ctx->stream));
  CU(cudaEventRecord(ctx->ev1, ctx->stream));  // spv_last_timing_ms covers render + composite
  ctx->launches += 3;
  slot_clip_set(ctx, ctx->slot, ctx->cam, p->box, (ctx->dtype == SPV_F32 ? -1.f : 0.f));  // the composited frame: misses outside the projected box
  return 0;
}

SPV_API int spv_set_extra_slabs(spv_ctx *ctx, spv_ctx **others, int n) {
  if (!ctx) return SPV_EINVAL;
  if (n < 0 || n > MAX_EXTRA_SLABS || (n > 0 && !others)) return fail(ctx, SPV_EINVAL, "spv_set_extra_slabs: at most 3 extra slabs");
  for (int i = 0; i < n; ++i)
    if (!others[i] || others[i] == ctx || others[i]->device != ctx->device)
      return fail(ctx, SPV_EINVAL, "spv_set_extra_slabs: need other contexts on the same device");
  for (int i = 0; i < MAX_EXTRA_SLABS; ++i) ctx->extra[i] = i < n ? others[i] : nullptr;
  ctx->n_extra = n;
  return 0;
}

SPV_API int spv_comp_check(spv_ctx *ctx) {
  BIND();
  if (ctx->comp_world < 1) return fail(ctx, SPV_ENODATA, "spv_comp_check: spv_comp_init first");
  unsigned e = 0;
  CU(cudaMemcpyAsync(&e, ctx->comp_err, sizeof e, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  if (e) {
    char b[128];
    snprintf(b, sizeof b, "sort-last composite: timed out waiting for rank %u", e - 1);
    return fail(ctx, -110, b);
  }
  return 0;
}

SPV_API int spv_mip_finish(spv_ctx *ctx, const spv_mip_params *p) {
  BIND();
  if (!p) return fail(ctx, SPV_EINVAL, "spv_mip_finish: null params");
  CU(launch_mip_finish(ctx->raw(), ctx->out(), (int)ctx->n(), p->min_val, p->max_val, p->gamma, ctx->stream));
  ctx->launches += 1;
  ctx->last_method = 0;
  slot_clip_set(ctx, ctx->slot, ctx->cam, p->box, (ctx->dtype == SPV_F32 ? -1.f : 0.f));  // the composited frame: misses outside the projected box
  return 0;
}

// the occlusion tap table covers taps [0, taps_n); it is extended when a render asks for more
static int ensure_taps(spv_ctx *ctx, int n) {
  if (n <= ctx->taps_n) return 0;
  int cap = 64;
  while (cap < n) cap *= 2;
  CU(cudaStreamSynchronize(ctx->stream));
  CU(cudaStreamSynchronize(ctx->post_stream));
  if (ctx->d_taps) cudaFree(ctx->d_taps);
  ctx->d_taps = nullptr;
  ctx->taps_n = 0;
  CU(cudaMalloc(&ctx->d_taps, (size_t)cap * sizeof(float4)));
  CU(launch_occ_taps(ctx->d_taps, cap, ctx->stream));
  ctx->launches += 1;
  ctx->taps_n = cap;
  return 0;
}

// The table of tap offsets (launch_occ_table) for the current image size and these occlusion parameters: built when
// they change (one launch of the per-tap hashing over the whole image), reused by every frame after that.
// *table = nullptr where the table form does not apply (knob 17 off, radius > 44, table beyond 1 GiB, no memory).
static int ensure_occ_table(spv_ctx *ctx, int radius, int n_points, const void **table) {
  *table = nullptr;
  if (!ctx->occ_table_on || !ctx->d_occ_queue) return 0;
  const int key[4] = {ctx->width, ctx->height, radius, n_points};
  if (ctx->d_occ_table && memcmp(key, ctx->occ_key, sizeof(key)) == 0) {
    *table = ctx->d_occ_table;
    return 0;
  }
  const size_t bytes = occ_table_bytes(ctx->width, ctx->height, radius, n_points);
  if (!bytes) return 0;
  CU(cudaStreamSynchronize(ctx->stream));  // frames in flight may still read the old table
  if (ctx->post_stream) CU(cudaStreamSynchronize(ctx->post_stream));
  if (ctx->d_occ_table) cudaFree(ctx->d_occ_table);
  ctx->d_occ_table = nullptr;
  ctx->occ_key[2] = -1;
  if (cudaMalloc(&ctx->d_occ_table, bytes) != cudaSuccess) {  // not enough memory: the queue form needs none
    cudaGetLastError();
    ctx->d_occ_table = nullptr;
    return 0;
  }
  unsigned *bad = nullptr, h_bad = 0;
  CU(cudaMalloc(&bad, sizeof(unsigned)));
  CU(cudaMemsetAsync(bad, 0, sizeof(unsigned), ctx->stream));
  CU(launch_occ_table(ctx->d_occ_table, ctx->width, ctx->height, radius, n_points, ctx->d_taps, bad, ctx->stream));
  ctx->launches += 1;
  CU(cudaMemcpyAsync(&h_bad, bad, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));  // the passes of later frames read it from other streams as well
  cudaFree(bad);
  if (h_bad) {  // an offset beyond the halo: keep hashing per frame
    cudaFree(ctx->d_occ_table);
    ctx->d_occ_table = nullptr;
    return 0;
  }
  memcpy(ctx->occ_key, key, sizeof(key));
  *table = ctx->d_occ_table;
  return 0;
}

static ConvWeights conv_weights(int Nh, float coef) {
  ConvWeights w;
  memset(&w, 0, sizeof w);
  w.nh = Nh;
  for (int ht = 0; ht < Nh; ++ht)  // convolve_2d.cl:23 / :86, evaluated left to right in fp32
    w.w[ht] = expf((float)(coef * ((float)ht - (float)Nh / 2.f) * ((float)ht - (float)Nh / 2.f) / (float)Nh / (float)Nh));
  return w;
}

// to_host: the alpha plane (final once the march has run) travels to the selected slot's pinned staging while the
// blur / occlusion / shading passes run, the value plane follows after the shading pass
static int render_iso_impl(spv_ctx *ctx, const spv_iso_params *p, bool to_host) {
  if (!p) return fail(ctx, SPV_EINVAL, "spv_render_iso: null params");
  if (!ctx->arr) return fail(ctx, SPV_ENODATA, "spv_render_iso: no volume set");
  if (ctx->slab) return fail(ctx, SPV_EINVAL, "spv_render_iso: not available on a slab context");
  if (p->max_steps < 2) return fail(ctx, SPV_EINVAL, "spv_render_iso: max_steps must be >= 2");
  if (p->occ_n_points < 0 || p->occ_radius < 0) return fail(ctx, SPV_EINVAL, "spv_render_iso: negative occlusion parameter");
  IsoArgs a;
  a.cam = ctx->cam;
  a.vol = volume_of(ctx);
  a.coarse = ctx->coarse;
  a.cgx = ctx->cgx; a.cgy = ctx->cgy; a.cgz = ctx->cgz;
  a.top = ctx->top;
  a.tgx = (ctx->cgx + 3) / 4; a.tgy = (ctx->cgy + 3) / 4; a.tgz = (ctx->cgz + 3) / 4;
  memcpy(a.box, p->box, sizeof a.box);
  a.iso_val = p->iso_val; a.gamma = p->gamma; a.max_steps = p->max_steps;
  a.segments = ctx->iso_segments;
  a.centre_out = ctx->iso_centre_out;
  const bool exact_iso = ctx->sampler == SPV_SAMPLER_EXACT;
  a.tile_hit = exact_iso ? nullptr : ctx->d_tile_hit_s[ctx->slot];
  a.skip = ctx->skipping != 0;  // auto (-1) = on: for iso surfaces the brick test is nearly free and exact
  if (a.skip && ctx->sampler != SPV_SAMPLER_EXACT) {
    int rcb = ensure_bricks(ctx);
    if (rcb) return rcb;
  }
  a.width = ctx->width; a.height = ctx->height;
  const bool post = !(p->flags & SPV_ISO_RAW_ONLY);
  // with post passes the march writes the raw normals into tmp_vec and the fused blur moves them to their plane
  a.out = ctx->out(); a.alpha = ctx->alpha(); a.depth = ctx->depth(); a.normals = post ? ctx->tmp_vec() : ctx->normals();
  a.stats = ctx->stats_on ? ctx->d_stats : nullptr;
  const bool linear = ctx->linear && (ctx->dtype == SPV_F32 || ctx->int_filter);
  int rc = post ? ensure_taps(ctx, p->occ_n_points) : 0;
  const void *occ_table = nullptr;
  if (!rc && post) rc = ensure_occ_table(ctx, p->occ_radius, p->occ_n_points, &occ_table);
  if (rc) return rc;
  rc = begin_render(ctx);
  if (rc) return rc;
  const int s = ctx->slot;
  const size_t n = ctx->n();
  // overlap (tuning knob 14): this frame's screen-space passes go to post_stream, so that the NEXT frame's search (other
  // slot, render stream) runs beside them -- the search leaves a third of the SM time idle in its tail, the passes are
  // five small launches.  Whoever reads the planes joins (spv_read*, spv_stream_join, spv_sync); the read-back of
  // spv_render_iso_to_host waits for the passes on the copy stream only.  (Putting every other frame's search on a stream of its own as well was measured: no further gain.)
  const bool overlap = ctx->iso_overlap && post && !ctx->stats_on;
  rc = join_post(ctx, s);  // the passes of the frame that used this slot last may still be running
  if (rc) return rc;
  if (to_host && ctx->copy_pending[s]) CU(cudaEventSynchronize(ctx->ev_copied[s]));  // staging about to be rewritten
  // a device-only render into a slot whose planes an asynchronous read-back may still be copying: the search waits for it
  // on the device (render_sequence issues frames three ahead of the one it hands out)
  if (!to_host && ctx->copy_pending[s]) CU(cudaStreamWaitEvent(ctx->stream, slot_free_event(ctx, s), 0));
  // read-back: only the rectangle the projected box can touch (outside it there is no surface: out 0, alpha 0, which the
  // pinned planes hold already)
  int cxa = 0, cxb = ctx->width, cya = 0, cyb = ctx->height;
  if (to_host) {
    if (ctx->clip_copies) miss_free_rect(ctx->cam, p->box, ctx->width, ctx->height, cxa, cxb, cya, cyb);
    else staging_dirty(ctx, s);
  }
  // planes [first, first + count) of the slot, clipped to that rectangle, device -> pinned on the copy stream
  auto copy_planes = [&](int first, int count) -> int {
    if (cxa >= cxb || cya >= cyb) return 0;
    const size_t W = (size_t)ctx->width, H = (size_t)ctx->height;
    cudaMemcpy3DParms cp;
    memset(&cp, 0, sizeof cp);
    cp.srcPtr = make_cudaPitchedPtr(ctx->dbuf_s[s], W * sizeof(float), W, H);
    cp.dstPtr = make_cudaPitchedPtr(ctx->hpin_s[s], W * sizeof(float), W, H);
    cp.srcPos = make_cudaPos((size_t)cxa * sizeof(float), (size_t)cya, (size_t)first);
    cp.dstPos = cp.srcPos;
    cp.extent = make_cudaExtent((size_t)(cxb - cxa) * sizeof(float), (size_t)(cyb - cya), (size_t)count);
    cp.kind = cudaMemcpyDeviceToHost;
    CU(cudaMemcpy3DAsync(&cp, ctx->copy_stream));
    ctx->d2h_bytes += (size_t)count * (size_t)(cxb - cxa) * (size_t)(cyb - cya) * sizeof(float);
    return 0;
  };
  CU(launch_iso(a, fmt_of(ctx), linear, ctx->sampler == SPV_SAMPLER_EXACT, ctx->stats_on != 0, ctx->stream));
  ctx->launches += 1;
  // (the pinned planes outside the rectangle go back to the miss values on the host while the search runs)
  if (to_host && ctx->clip_copies) staging_clean_outside_rect(ctx, s, cxa, cxb, cya, cyb, 0.f);
  cudaStream_t pst = ctx->stream;
  if (overlap) {
    // one set of scratch (occlusion queue; tmp planes are per slot, taps constant): the passes of successive frames
    // run in order on post_stream; they start when this frame's search is done
    CU(cudaEventRecord(ctx->ev_searched[s], ctx->stream));
    CU(cudaStreamWaitEvent(ctx->post_stream, ctx->ev_searched[s], 0));
    pst = ctx->post_stream;
  }
  if (to_host && post) {
    CU(cudaEventRecord(ctx->ev_rendered[s], ctx->stream));
    CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_rendered[s], 0));
    rc = copy_planes(1, 1);  // the alpha plane is final once the march has run
    if (rc) return rc;
  }
  if (post) {
    // volumerender.py:470-497
    CU(launch_conv_xy(ctx->tmp_vec(), ctx->normals(), ctx->width, ctx->height, 3, conv_weights(7, -5.f), a.tile_hit, 0, pst));
    CU(launch_occlusion(ctx->tmp(), ctx->width, ctx->height, p->occ_radius, p->occ_n_points, ctx->depth(), a.tile_hit,
                        ctx->d_taps, ctx->d_occ_queue, occ_table ? ctx->occ_frame : ctx->occ_frame++, ctx->sms, pst, 0, -1,
                        occ_table));  // (the table form leaves the queue's counters alone: the frame parity stays)
    if (ctx->fuse_shading) {  // knob 20: the shading rides on the epilogue of the occlusion blur
      ShadeArgs sh;
      sh.out = ctx->out(); sh.cam = ctx->cam; sh.occ_strength = p->occ_strength;
      sh.normals = ctx->normals(); sh.depth = ctx->depth();
      CU(launch_conv_xy(ctx->tmp(), ctx->occ(), ctx->width, ctx->height, 1, conv_weights(5, -10.f), a.tile_hit,
                        p->occ_radius, pst, 0, -1, &sh));
    } else {
      CU(launch_conv_xy(ctx->tmp(), ctx->occ(), ctx->width, ctx->height, 1, conv_weights(5, -10.f), a.tile_hit,
                        p->occ_radius, pst));
      CU(launch_shading(ctx->out(), ctx->width, ctx->height, ctx->cam, p->occ_strength, ctx->normals(), ctx->depth(),
                        ctx->occ(), pst));
    }
    ctx->launches += (a.tile_hit && !occ_table ? 5 : 4) - (ctx->fuse_shading ? 1 : 0);  // (the queue form of the occlusion is two launches)
    if (overlap) {
      CU(cudaEventRecord(ctx->ev_posted[s], ctx->post_stream));
      ctx->post_pending[s] = true;
    }
  }
  ctx->last_method = 1;
  slot_clip_set(ctx, s, ctx->cam, p->box, 0.f, post);  // no surface: out 0 (shading_kernel), alpha 0 (every element type)
  rc = end_render(ctx);
  if (rc) return rc;
  if (to_host) {
    if (overlap) {  // the passes ran beside the render stream: the value plane is final when they are
      CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_posted[s], 0));
    } else {
      CU(cudaEventRecord(ctx->ev_rendered[s], ctx->stream));
      CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_rendered[s], 0));
    }
    rc = copy_planes(0, post ? 1 : 2);
    if (rc) return rc;
    CU(cudaEventRecord(ctx->ev_copied[s], ctx->copy_stream));
    ctx->copy_pending[s] = true;
  }
  return 0;
}

SPV_API int spv_render_iso(spv_ctx *ctx, const spv_iso_params *p) {
  BIND();
  return render_iso_impl(ctx, p, false);
}

SPV_API int spv_render_iso_to_host(spv_ctx *ctx, const spv_iso_params *p, int wait, float **host) {
  BIND();
  int rc = render_iso_impl(ctx, p, true);
  if (rc) return rc;
  if (wait) {
    CU(cudaEventSynchronize(ctx->ev_copied[ctx->slot]));
    CU(cudaStreamSynchronize(ctx->stream));  // statistics / timing events of this frame
  }
  if (host) *host = ctx->hpin_s[ctx->slot];
  return 0;
}

// ---- sort-last iso surface ----------------------------------------------------------------------------------------
static int iso_args(spv_ctx *ctx, const spv_iso_params *p, IsoArgs &a, const char *who) {
  if (!p) return fail(ctx, SPV_EINVAL, "sort-last iso surface: null params");
  if (!ctx->arr || !ctx->slab) return fail(ctx, SPV_ENODATA, "sort-last iso surface: set a slab with spv_set_volume_slab_halo first");
  if (ctx->sampler != SPV_SAMPLER_TMU) return fail(ctx, SPV_EINVAL, "sort-last iso surface: needs the TMU sampler");
  if (p->max_steps < 2) return fail(ctx, SPV_EINVAL, "sort-last iso surface: max_steps must be >= 2");
  (void)who;
  a.cam = ctx->cam;
  a.vol = volume_of(ctx);
  a.coarse = ctx->coarse;
  a.cgx = ctx->cgx; a.cgy = ctx->cgy; a.cgz = ctx->cgz;
  a.top = ctx->top;
  a.tgx = (ctx->cgx + 3) / 4; a.tgy = (ctx->cgy + 3) / 4; a.tgz = (ctx->cgz + 3) / 4;
  memcpy(a.box, p->box, sizeof a.box);
  a.iso_val = p->iso_val; a.gamma = p->gamma; a.max_steps = p->max_steps;
  a.segments = ctx->iso_segments;
  a.centre_out = ctx->iso_centre_out;
  a.tile_hit = ctx->d_tile_hit;
  a.skip = ctx->skipping != 0;
  a.width = ctx->width; a.height = ctx->height;
  a.out = ctx->out(); a.alpha = ctx->alpha(); a.depth = ctx->depth(); a.normals = ctx->normals();
  a.stats = ctx->stats_on ? ctx->d_stats : nullptr;
  return 0;
}

SPV_API int spv_iso_slab_search(spv_ctx *ctx, const spv_iso_params *p) {
  BIND();
  IsoArgs a;
  int rc = iso_args(ctx, p, a, "search");
  if (rc) return rc;
  if (a.skip) {
    rc = ensure_bricks(ctx);
    if (rc) return rc;
  }
  rc = begin_render(ctx);
  if (rc) return rc;
  int *k = (int *)ctx->raw();  // the two candidate planes live in [raw | tmp]
  const bool linear = ctx->linear && (ctx->dtype == SPV_F32 || ctx->int_filter);
  CU(launch_iso_slab(a, fmt_of(ctx), linear, 0, k, k + ctx->n(), ctx->occ(), ctx->d_iso_err, ctx->stream));
  ctx->launches += 1;
  ctx->last_method = 1;
  ctx->clip_s[ctx->slot].valid = false;  // (sort-last iso frames: read back whole)
  return end_render(ctx);
}

SPV_API int spv_iso_slab_resolve(spv_ctx *ctx, const spv_iso_params *p) {
  BIND();
  IsoArgs a;
  int rc = iso_args(ctx, p, a, "resolve");
  if (rc) return rc;
  a.stats = nullptr;
  int *k = (int *)ctx->raw();  // the two candidate planes live in [raw | tmp]
  const bool linear = ctx->linear && (ctx->dtype == SPV_F32 || ctx->int_filter);
  CU(launch_iso_slab(a, fmt_of(ctx), linear, 1, k, k + ctx->n(), ctx->occ(), ctx->d_iso_err, ctx->stream));
  ctx->launches += 1;
  return 0;
}

SPV_API int spv_iso_slab_post(spv_ctx *ctx, const spv_iso_params *p) {
  BIND();
  IsoArgs a;
  int rc = iso_args(ctx, p, a, "post");
  if (rc) return rc;
  if (p->occ_n_points < 0 || p->occ_radius < 0) return fail(ctx, SPV_EINVAL, "sort-last iso surface: negative occlusion parameter");
  int *k = (int *)ctx->raw();  // the two candidate planes live in [raw | tmp]
  CU(launch_iso_slab(a, fmt_of(ctx), true, 2, k, k + ctx->n(), ctx->occ(), ctx->d_iso_err, ctx->stream));
  ctx->launches += 1;
  if (!(p->flags & SPV_ISO_RAW_ONLY)) {
    rc = ensure_taps(ctx, p->occ_n_points);
    if (rc) return rc;
    CU(launch_conv(ctx->normals(), ctx->tmp_vec(), ctx->width, ctx->height, 3, conv_weights(7, -5.f), ctx->stream));
    CU(launch_occlusion(ctx->occ(), ctx->width, ctx->height, p->occ_radius, p->occ_n_points, ctx->depth(), a.tile_hit,
                        ctx->d_taps, ctx->d_occ_queue, ctx->occ_frame++, ctx->sms, ctx->stream));
    CU(launch_conv(ctx->occ(), ctx->tmp(), ctx->width, ctx->height, 1, conv_weights(5, -10.f), ctx->stream));
    CU(launch_shading(ctx->out(), ctx->width, ctx->height, ctx->cam, p->occ_strength, ctx->normals(), ctx->depth(),
                      ctx->occ(), ctx->stream));
    ctx->launches += 7;
  }
  CU(cudaEventRecord(ctx->ev1, ctx->stream));
  return 0;
}

SPV_API int spv_iso_slab_check(spv_ctx *ctx) {
  BIND();
  unsigned e = 0;
  CU(cudaMemcpyAsync(&e, ctx->d_iso_err, sizeof e, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  if (e) {
    CU(cudaMemsetAsync(ctx->d_iso_err, 0, sizeof e, ctx->stream));
    return fail(ctx, SPV_EINVAL, "sort-last iso surface: the slab's halo does not cover the gradient taps of a surface "
                                 "pixel (2 * dt * gamma^2 * Nz + one ray step slices are needed): upload with a larger halo");
  }
  return 0;
}

// Sort-last iso surface with every exchange over peer memory (no NCCL on the data path), enqueue-only:
//   search: k1 / k0 candidates of band o's pixels are stored into owner o's staging   | barrier (arrival counters)
//   owner : MIN over the ranks, the band's final k planes stored into EVERY rank      | barrier
//   resolve: the rank owning a pixel's crossing stores the finished pixel into EVERY rank's planes | barrier
//   post passes on every rank
SPV_API int spv_render_iso_composite(spv_ctx *ctx, const spv_iso_params *p) {
  BIND();
  const int W = ctx->comp_world, R = ctx->comp_rank;
  if (W < 1) return fail(ctx, SPV_ENODATA, "spv_render_iso_composite: spv_comp_init first");
  for (int r = 0; r < W; ++r)
    if (!ctx->peer_part[r]) return fail(ctx, SPV_ENODATA, "spv_render_iso_composite: a peer has not been imported");
  if (ctx->slot != 0) return fail(ctx, SPV_EINVAL, "spv_render_iso_composite: select output slot 0 first");
  IsoArgs a;
  int rc = iso_args(ctx, p, a, "composite");
  if (rc) return rc;
  if (p->occ_n_points < 0 || p->occ_radius < 0) return fail(ctx, SPV_EINVAL, "sort-last iso surface: negative occlusion parameter");
  if (a.skip) {
    rc = ensure_bricks(ctx);
    if (rc) return rc;
  }
  const bool post = !(p->flags & SPV_ISO_RAW_ONLY);
  // (no tap-offset table here: building it allocates and synchronises, which must not happen while a peer's frame --
  // whose kernels spin on this rank's flags -- is in flight; sort-last frames hash their taps)
  if (post) {
    rc = ensure_taps(ctx, p->occ_n_points);
    if (rc) return rc;
  }
  const unsigned f = ++ctx->comp_frame;
  const unsigned parity = f & 1u;
  const size_t band = (size_t)ctx->comp_band_rows * ctx->width, n = ctx->n();
  IsoPeer peer;
  memset(&peer, 0, sizeof peer);
  peer.world = W;
  peer.band_rows = ctx->comp_band_rows;
  peer.src = (unsigned)R;
  peer.band = (unsigned)band;
  for (int r = 0; r < W; ++r) {
    peer.kpart[r] = (int *)(ctx->peer_part[r] + 2 * (size_t)W * band) + (size_t)parity * W * 2 * band;
    peer.planes[r] = ctx->peer_out[r];
  }
  // with post passes the resolved normals go to tmp_vec and the fused blur moves them to their plane
  peer.normals_plane = post ? 9 : 4;
  rc = begin_render(ctx);
  if (rc) return rc;
  ctx->n_ph = 0;
#define SPV_PHASE()                                                                  \
  do {                                                                               \
    if (ctx->time_phases && ctx->n_ph < 9) {                                            \
      if (!ctx->ev_ph[ctx->n_ph]) CU(cudaEventCreate(&ctx->ev_ph[ctx->n_ph]));       \
      CU(cudaEventRecord(ctx->ev_ph[ctx->n_ph++], ctx->stream));                     \
    }                                                                                \
  } while (0)
  SPV_PHASE();
  int *k = (int *)ctx->raw();  // the two candidate planes live in [raw | tmp]
  const bool linear = ctx->linear && (ctx->dtype == SPV_F32 || ctx->int_filter);
  CU(launch_iso_slab(a, fmt_of(ctx), linear, 0, k, k + n, ctx->occ(), ctx->d_iso_err, ctx->stream, &peer));
  SPV_PHASE();
  CU(launch_comp_sync(ctx->peer_flags, ctx->comp_flags, W, R, 2, f, ctx->comp_err, ctx->stream));
  SPV_PHASE();
  KReduceArgs ka;
  memset(&ka, 0, sizeof ka);
  ka.part = peer.kpart[R];
  for (int r = 0; r < W; ++r) ka.kplanes[r] = (int *)(ctx->peer_out[r] + 7 * n);  // [raw | tmp]: the k planes
  ka.world = W;
  ka.band = (unsigned)band;
  ka.first_pixel = (unsigned)((size_t)R * band);
  ka.n_pixels = ka.first_pixel >= n ? 0u : (unsigned)((n - ka.first_pixel) < band ? (n - ka.first_pixel) : band);
  ka.n_image = n;
  CU(launch_k_reduce(ka, ctx->stream));
  SPV_PHASE();
  CU(launch_comp_sync(ctx->peer_flags, ctx->comp_flags, W, R, 3, f, ctx->comp_err, ctx->stream));
  SPV_PHASE();
  a.stats = nullptr;
  CU(launch_iso_slab(a, fmt_of(ctx), linear, 1, k, k + n, ctx->occ(), ctx->d_iso_err, ctx->stream, &peer));
  SPV_PHASE();
  CU(launch_comp_sync(ctx->peer_flags, ctx->comp_flags, W, R, 4, f, ctx->comp_err, ctx->stream));
  SPV_PHASE();
  ctx->launches += 6;
  if (post) {
    // The screen-space passes (volumerender.py:470-497) on this rank's band of rows only -- every rank holds the
    // complete raw planes, and a pass reads whatever rows around the band its taps reach -- then the band's finished
    // planes go to every peer.  Per rank the passes cost 1 / world of what they cost on one GPU.
    const int ya = R * ctx->comp_band_rows < ctx->height ? R * ctx->comp_band_rows : ctx->height;
    const int yb = ya + ctx->comp_band_rows < ctx->height ? ya + ctx->comp_band_rows : ctx->height;
    const bool shard = W > 1 && ctx->iso_post_sharded;
    const int y0 = shard ? ya : 0, y1 = shard ? yb : ctx->height;
    // the occlusion blur of row y reads the raw occlusion of rows y - 2 .. y + 2
    const int oy0 = shard ? (y0 - 4 > 0 ? y0 - 4 : 0) : 0, oy1 = shard ? (y1 + 4 < ctx->height ? y1 + 4 : ctx->height) : ctx->height;
    CU(launch_conv_xy(ctx->tmp_vec(), ctx->normals(), ctx->width, ctx->height, 3, conv_weights(7, -5.f), a.tile_hit, 0,
                      ctx->stream, y0, y1));
    CU(launch_occlusion(ctx->tmp(), ctx->width, ctx->height, p->occ_radius, p->occ_n_points, ctx->depth(), a.tile_hit,
                        ctx->d_taps, ctx->d_occ_queue, ctx->occ_frame++, ctx->sms, ctx->stream, oy0, oy1));
    CU(launch_conv_xy(ctx->tmp(), ctx->occ(), ctx->width, ctx->height, 1, conv_weights(5, -10.f), a.tile_hit,
                      p->occ_radius, ctx->stream, y0, y1));
    CU(launch_shading(ctx->out(), ctx->width, ctx->height, ctx->cam, p->occ_strength, ctx->normals(), ctx->depth(),
                      ctx->occ(), ctx->stream, y0, y1));
    ctx->launches += 5;
    SPV_PHASE();
    if (shard) {
      BandGatherArgs g;
      memset(&g, 0, sizeof g);
      for (int r = 0; r < W; ++r) g.planes[r] = ctx->peer_out[r];
      g.world = W; g.rank = R; g.width = ctx->width; g.height = ctx->height; g.y_first = y0; g.y_end = y1;
      CU(launch_band_gather(g, ctx->stream));
      CU(launch_comp_sync(ctx->peer_flags, ctx->comp_flags, W, R, 5, f, ctx->comp_err, ctx->stream));
      ctx->launches += 2;
      SPV_PHASE();
    }
  } else {
    CU(cudaMemsetAsync(ctx->occ(), 0, n * sizeof(float), ctx->stream));  // what the NCCL path's resolve leaves there
  }
  ctx->last_method = 1;
  ctx->clip_s[ctx->slot].valid = false;  // (sort-last iso frames: read back whole)
  return end_render(ctx);
#undef SPV_PHASE
}

// durations between the phase boundaries of the last spv_render_iso_composite that ran with tuning knob 13 on:
// search | wait for the candidates | MIN + redistribution | wait | resolve | wait | screen-space passes | band gather + wait
SPV_API int spv_last_phases_ms(spv_ctx *ctx, float *ms, int n, int *count) {
  BIND();
  if (!ms || !count) return fail(ctx, SPV_EINVAL, "spv_last_phases_ms: null argument");
  CU(cudaStreamSynchronize(ctx->stream));
  *count = ctx->n_ph > 1 ? ctx->n_ph - 1 : 0;
  for (int i = 0; i + 1 < ctx->n_ph && i < n; ++i) CU(cudaEventElapsedTime(&ms[i], ctx->ev_ph[i], ctx->ev_ph[i + 1]));
  return 0;
}

static float *buf_of(spv_ctx *c, int which, size_t *count) {
  const size_t n = c->n();
  *count = n;
  switch (which) {
    case SPV_BUF_OUT: return c->out();
    case SPV_BUF_ALPHA: return c->alpha();
    case SPV_BUF_DEPTH: return c->depth();
    case SPV_BUF_OCC: return c->occ();
    case SPV_BUF_RAW: return c->raw();
    case SPV_BUF_KPLANES: *count = 2 * n; return c->raw();  // int32 [2][h][w] in [raw | tmp], sort-last iso surface
    case SPV_BUF_NORMALS: *count = 3 * n; return c->normals();
    default: return nullptr;
  }
}

SPV_API int spv_read(spv_ctx *ctx, int which, float *host_dst, size_t n) {
  BIND();
  size_t count = 0;
  float *src = buf_of(ctx, which, &count);
  if (!src || !host_dst) return fail(ctx, SPV_EINVAL, "spv_read: bad buffer id or null destination");
  if (n != count) return fail(ctx, SPV_EINVAL, "spv_read: element count does not match the buffer");
  int rcj = join_post(ctx, ctx->slot);
  if (rcj) return rcj;
  CU(cudaMemcpyAsync(host_dst, src, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  ctx->d2h_bytes += n * sizeof(float);
  CU(cudaStreamSynchronize(ctx->stream));
  return 0;
}

SPV_API int spv_read_many(spv_ctx *ctx, float *out, float *alpha, float *depth, float *normals, float *occ) {
  BIND();
  const size_t n = ctx->n();
  // [out | alpha | depth | occ | normals]: MIP wrote the first two planes, iso all seven
  const size_t planes = (depth || normals || occ) ? 7 : 2;
  staging_dirty(ctx, ctx->slot);
  int rcj = join_post(ctx, ctx->slot);
  if (rcj) return rcj;
  CU(cudaMemcpyAsync(ctx->hpin, ctx->dbuf, planes * n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  ctx->d2h_bytes += planes * n * sizeof(float);
  CU(cudaStreamSynchronize(ctx->stream));
  if (out) memcpy(out, ctx->hpin, n * sizeof(float));
  if (alpha) memcpy(alpha, ctx->hpin + n, n * sizeof(float));
  if (depth) memcpy(depth, ctx->hpin + 2 * n, n * sizeof(float));
  if (occ) memcpy(occ, ctx->hpin + 3 * n, n * sizeof(float));
  if (normals) memcpy(normals, ctx->hpin + 4 * n, 3 * n * sizeof(float));
  return 0;
}

// Enqueue the copy of the leading `planes` result planes of slot s into its pinned staging on stream st (the staging is
// quiescent).  output + alpha of a finished frame (what a display needs): pixels the projected box cannot touch are misses
// -- out 0; alpha 0 for integer max projections and for iso surfaces (no crossing), -1 for float32 max projections -- and
// are not copied; the pinned planes hold those values there already (tuning knob 9, as spv_render_mip_to_host).
static int copy_slot_planes(spv_ctx *ctx, int s, int planes, cudaStream_t st, bool stage = false) {
  const spv_ctx::SlotClip &c = ctx->clip_s[s];
  ctx->freed_by_stage[s] = false;
  if (c.valid && planes <= 2 && ctx->clip_copies) {
    const int W = ctx->width, H = ctx->height;
    int xa, xb, ya, yb;
    miss_free_rect(c.cam, c.box, W, H, xa, xb, ya, yb);
    const int fxa = xa, fxb = xb, fya = ya, fyb = yb;  // this frame's own rectangle
    bool widened = false;
    if (stage && ctx->stage_reads && c.full && planes == 2 && xa < xb && ya < yb && ctx->clean_alpha[s] == c.miss_alpha &&
        ctx->dirty_x0[s] < ctx->dirty_x1[s] && ctx->dirty_lo[s] < ctx->dirty_hi[s]) {
      // the pinned planes hold an earlier frame inside their dirty rectangle.  Where that sticks out of this frame's
      // rectangle by little (a camera path: a few pixels per frame), copying the union of the two from the device --
      // whose planes hold the miss values out there -- is cheaper than having the host clear thousands of row ends
      const int ux0 = xa < ctx->dirty_x0[s] ? xa : ctx->dirty_x0[s], ux1 = xb > ctx->dirty_x1[s] ? xb : ctx->dirty_x1[s];
      const int uy0 = ya < ctx->dirty_lo[s] ? ya : ctx->dirty_lo[s], uy1 = yb > ctx->dirty_hi[s] ? yb : ctx->dirty_hi[s];
      if ((double)(ux1 - ux0) * (uy1 - uy0) <= 1.125 * (double)(xb - xa) * (yb - ya)) {
        xa = ux0 < 0 ? 0 : ux0; xb = ux1 > W ? W : ux1;
        ya = uy0 < 0 ? 0 : uy0; yb = uy1 > H ? H : uy1;
        widened = true;
      }
    }
    staging_clean_outside_rect(ctx, s, xa, xb, ya, yb, c.miss_alpha);
    if (widened) {  // what is copied outside the frame's own rectangle are miss values: the planes are dirty inside it only
      ctx->dirty_x0[s] = fxa; ctx->dirty_x1[s] = fxb;
      ctx->dirty_lo[s] = fya; ctx->dirty_hi[s] = fyb;
    }
    if (stage && ctx->stage_reads && xa < xb && ya < yb && !ctx->dstage_s[s] &&
        cudaMalloc(&ctx->dstage_s[s], 2 * ctx->n() * sizeof(float)) != cudaSuccess) {
      cudaGetLastError();  // no memory for the staging planes: copy straight out of the slot
      ctx->dstage_s[s] = nullptr;
      stage = false;
    }
    if (stage && ctx->stage_reads && xa < xb && ya < yb) {
      // rectangle -> device staging (same layout), slot free; then staging -> pinned host memory, one 2-D copy per plane
      CU(launch_rect_copy(ctx->dbuf_s[s], ctx->dstage_s[s], W, H, xa, xb, ya, yb, planes, st));
      ctx->launches += 1;
      CU(cudaEventRecord(ctx->ev_freed[s], st));
      ctx->freed_by_stage[s] = true;
      const size_t off = (size_t)ya * W + xa;
      for (int pl = 0; pl < planes; ++pl)
        CU(cudaMemcpy2DAsync(ctx->hpin_s[s] + pl * ctx->n() + off, (size_t)W * sizeof(float), ctx->dstage_s[s] + pl * ctx->n() + off,
                             (size_t)W * sizeof(float), (size_t)(xb - xa) * sizeof(float), (size_t)(yb - ya),
                             cudaMemcpyDeviceToHost, st));
      ctx->d2h_bytes += (size_t)planes * (size_t)(xb - xa) * (size_t)(yb - ya) * sizeof(float);
      return 0;
    }
    if (xa < xb && ya < yb) {
      cudaMemcpy3DParms cp;  // the rectangle of the leading plane(s) in one 3-D copy
      memset(&cp, 0, sizeof cp);
      cp.srcPtr = make_cudaPitchedPtr(ctx->dbuf_s[s], (size_t)W * sizeof(float), (size_t)W, (size_t)H);
      cp.dstPtr = make_cudaPitchedPtr(ctx->hpin_s[s], (size_t)W * sizeof(float), (size_t)W, (size_t)H);
      cp.srcPos = make_cudaPos((size_t)xa * sizeof(float), (size_t)ya, 0);
      cp.dstPos = cp.srcPos;
      cp.extent = make_cudaExtent((size_t)(xb - xa) * sizeof(float), (size_t)(yb - ya), (size_t)planes);
      cp.kind = cudaMemcpyDeviceToHost;
      CU(cudaMemcpy3DAsync(&cp, st));
      ctx->d2h_bytes += (size_t)planes * (size_t)(xb - xa) * (size_t)(yb - ya) * sizeof(float);
    }
    return 0;
  }
  staging_dirty(ctx, s);
  CU(cudaMemcpyAsync(ctx->hpin_s[s], ctx->dbuf_s[s], (size_t)planes * ctx->n() * sizeof(float), cudaMemcpyDeviceToHost, st));
  ctx->d2h_bytes += (size_t)planes * ctx->n() * sizeof(float);
  return 0;
}

SPV_API int spv_read_pinned(spv_ctx *ctx, int planes, float **host) {
  BIND();
  if (planes < 1 || planes > 7 || !host) return fail(ctx, SPV_EINVAL, "spv_read_pinned: planes must be 1..7");
  if (ctx->copy_pending[ctx->slot]) CU(cudaEventSynchronize(ctx->ev_copied[ctx->slot]));  // same staging memory
  int rcj = join_post(ctx, ctx->slot);
  if (rcj) return rcj;
  rcj = copy_slot_planes(ctx, ctx->slot, planes, ctx->stream);
  if (rcj) return rcj;
  CU(cudaStreamSynchronize(ctx->stream));
  *host = ctx->hpin;
  return 0;
}

SPV_API int spv_select_slot(spv_ctx *ctx, int slot) {
  BIND();
  if (slot != 0 && slot != 1) return fail(ctx, SPV_EINVAL, "spv_select_slot: slot must be 0 or 1");
  if (!ctx->dbuf_s[slot]) {
    int rc = alloc_slot(ctx, slot);
    if (rc) return rc;
  }
  // renders into this slot must not overtake an asynchronous read of it that is still in flight
  if (ctx->copy_pending[slot]) CU(cudaStreamWaitEvent(ctx->stream, slot_free_event(ctx, slot), 0));
  ctx->slot = slot;
  ctx->dbuf = ctx->dbuf_s[slot];
  ctx->hpin = ctx->hpin_s[slot];
  return 0;
}

SPV_API int spv_read_pinned_async(spv_ctx *ctx, int planes) {
  BIND();
  if (planes < 1 || planes > 7) return fail(ctx, SPV_EINVAL, "spv_read_pinned_async: planes must be 1..7");
  const int s = ctx->slot;
  CU(cudaEventRecord(ctx->ev_rendered[s], ctx->stream));
  CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_rendered[s], 0));
  // iso overlap: the slot's screen-space passes run beside the render stream; the copy waits for them, the render
  // stream (and with it the next frame's search) does not
  if (ctx->post_pending[s]) CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_posted[s], 0));
  if (ctx->copy_pending[s]) CU(cudaEventSynchronize(ctx->ev_copied[s]));  // the staging may be cleaned by the host
  {
    int rcc = copy_slot_planes(ctx, s, planes, ctx->copy_stream, true);
    if (rcc) return rcc;
  }
  CU(cudaEventRecord(ctx->ev_copied[s], ctx->copy_stream));
  ctx->copy_pending[s] = true;
  return 0;
}

SPV_API int spv_wait_slot(spv_ctx *ctx, int slot, float **host) {
  BIND();
  if ((slot != 0 && slot != 1) || !host) return fail(ctx, SPV_EINVAL, "spv_wait_slot: bad argument");
  if (!ctx->hpin_s[slot]) return fail(ctx, SPV_ENODATA, "spv_wait_slot: slot was never used");
  if (ctx->copy_pending[slot]) CU(cudaEventSynchronize(ctx->ev_copied[slot]));
  *host = ctx->hpin_s[slot];
  return 0;
}

SPV_API int spv_set_lut(spv_ctx *ctx, const float *rgb, int n) {
  BIND();
  if (!rgb || n < 1 || n > 65536) return fail(ctx, SPV_EINVAL, "spv_set_lut: need 1..65536 RGB triples");
  CU(cudaStreamSynchronize(ctx->stream));
  if (n != ctx->n_lut) {
    if (ctx->d_lut) cudaFree(ctx->d_lut);
    ctx->d_lut = nullptr;
    ctx->n_lut = 0;
    CU(cudaMalloc(&ctx->d_lut, (size_t)n * 3 * sizeof(float)));
    ctx->n_lut = n;
  }
  CU(cudaMemcpyAsync(ctx->d_lut, rgb, (size_t)n * 3 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));  // the host pointer is only borrowed for this call
  return 0;
}

SPV_API int spv_read_rgba8(spv_ctx *ctx, int mode_black, unsigned char *host_dst, size_t nbytes) {
  BIND();
  if (!ctx->d_lut) return fail(ctx, SPV_ENODATA, "spv_read_rgba8: no colour map (spv_set_lut)");
  const size_t n = ctx->n();
  if (!host_dst || nbytes != 4 * n) return fail(ctx, SPV_EINVAL, "spv_read_rgba8: need a destination of width*height*4 bytes");
  if (!ctx->d_rgba) {
    CU(cudaMalloc(&ctx->d_rgba, 4 * n));
    CU(cudaMallocHost(&ctx->h_rgba, 4 * n));
  }
  int rcj = join_post(ctx, ctx->slot);
  if (rcj) return rcj;
  CU(launch_display(ctx->out(), ctx->alpha(), ctx->d_lut, ctx->n_lut, mode_black != 0, ctx->d_rgba, n, ctx->stream));
  ctx->launches += 1;
  CU(cudaMemcpyAsync(ctx->h_rgba, ctx->d_rgba, 4 * n, cudaMemcpyDeviceToHost, ctx->stream));
  ctx->d2h_bytes += 4 * n;
  CU(cudaStreamSynchronize(ctx->stream));
  memcpy(host_dst, ctx->h_rgba, 4 * n);
  return 0;
}

SPV_API int spv_device_ptr(spv_ctx *ctx, int which, void **dev_ptr) {
  if (!ctx || !dev_ptr) return fail(ctx, SPV_EINVAL, "spv_device_ptr: null argument");
  size_t count = 0;
  float *src = buf_of(ctx, which, &count);
  if (!src) return fail(ctx, SPV_EINVAL, "spv_device_ptr: bad buffer id");
  *dev_ptr = src;
  return 0;
}

SPV_API int spv_last_timing_ms(spv_ctx *ctx, float *ms) {
  BIND();
  if (!ms) return fail(ctx, SPV_EINVAL, "spv_last_timing_ms: null argument");
  if (!ctx->timed) return fail(ctx, SPV_ENODATA, "spv_last_timing_ms: nothing rendered yet");
  CU(cudaEventSynchronize(ctx->ev1));
  CU(cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
  return 0;
}

SPV_API int spv_last_stats(spv_ctx *ctx, unsigned long long *v, int n) {
  BIND();
  if (!v || n < 2) return fail(ctx, SPV_EINVAL, "spv_last_stats: need room for 2 counters");
  if (!ctx->stats_on) return fail(ctx, SPV_ENODATA, "spv_last_stats: statistics are not enabled");
  CU(cudaStreamSynchronize(ctx->stream));
  for (int k = 0; k < (n < 40 ? n : 40); ++k) v[k] = ctx->h_stats[k];
  return 0;
}

SPV_API int spv_sample_points(spv_ctx *ctx, const float *host_pos, int n, float *host_out) {
  BIND();
  if (!ctx->arr) return fail(ctx, SPV_ENODATA, "spv_sample_points: no volume set");
  if (!host_pos || !host_out || n < 1) return fail(ctx, SPV_EINVAL, "spv_sample_points: bad argument");
  if (ctx->slab) return fail(ctx, SPV_EINVAL, "spv_sample_points: not available on a slab context");
  float *d = nullptr;
  CU(cudaMalloc(&d, (size_t)n * 4 * sizeof(float)));
  cudaError_t e = cudaMemcpyAsync(d, host_pos, (size_t)n * 3 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream);
  const bool linear = ctx->linear && (ctx->dtype == SPV_F32 || ctx->int_filter);
  if (e == cudaSuccess)
    e = launch_sample_points(volume_of(ctx), fmt_of(ctx), linear, ctx->sampler == SPV_SAMPLER_EXACT, d, n, d + 3 * (size_t)n,
                             ctx->stream);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(host_out, d + 3 * (size_t)n, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d);
  ctx->launches += 1;
  if (e != cudaSuccess) return cufail(ctx, e, "spv_sample_points");
  return 0;
}

static int texrate_probe_impl(spv_ctx *ctx, int iters, const float *vec9, double *samples_per_s) {
  if (!ctx->arr) return fail(ctx, SPV_ENODATA, "spv_texrate_probe: no volume set");
  if (iters < 1 || !samples_per_s) return fail(ctx, SPV_EINVAL, "spv_texrate_probe: bad argument");
  if (vec9)
    for (int i = 0; i < 9; ++i)
      if (!(fabsf(vec9[i]) <= 8.f)) return fail(ctx, SPV_EINVAL, "spv_texrate_probe_footprint: |component| <= 8 texels");
  int sms = 0;
  CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
  const int blocks = sms * 8;
  const bool linear = ctx->linear && (ctx->dtype == SPV_F32 || ctx->int_filter);
  Volume V = volume_of(ctx);
  int fmt = fmt_of(ctx);
  // float32 volumes that render through the layered pair copies (spv_mip_axis.cu): the rate of THAT fetch -- bilinear on
  // two-channel float texels + the lerp -- not of the 3-D array's trilinear fetch
  if (ctx->dtype == SPV_F32 && linear && axis_path_possible(ctx)) {
    int rca = ensure_axis(ctx, 2);
    if (rca && !ctx->axis_failed[2]) return rca;
    if (!ctx->axis_failed[2]) {
      V.filt = ctx->axis_tex[2];
      fmt = SPV_F32 + 3 * LAYOUT_ZPAIR;
    }
  }
  CU(launch_texrate_probe(V, fmt, linear, blocks, 8, vec9, ctx->tmp(), ctx->stream));  // warm-up
  CU(cudaEventRecord(ctx->ev0, ctx->stream));
  CU(launch_texrate_probe(V, fmt, linear, blocks, iters, vec9, ctx->tmp(), ctx->stream));
  CU(cudaEventRecord(ctx->ev1, ctx->stream));
  CU(cudaEventSynchronize(ctx->ev1));
  float ms = 0.f;
  CU(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
  ctx->launches += 2;
  *samples_per_s = (double)blocks * 256.0 * 16.0 * (double)iters / ((double)ms * 1e-3);
  return 0;
}

SPV_API int spv_texrate_probe(spv_ctx *ctx, int iters, double *samples_per_s) {
  BIND();
  return texrate_probe_impl(ctx, iters, nullptr, samples_per_s);
}

SPV_API int spv_texrate_probe_footprint(spv_ctx *ctx, int iters, const float *vec9, double *samples_per_s) {
  BIND();
  if (!vec9) return fail(ctx, SPV_EINVAL, "spv_texrate_probe_footprint: null vectors");
  return texrate_probe_impl(ctx, iters, vec9, samples_per_s);
}

SPV_API int spv_launch_count(spv_ctx *ctx, unsigned long long *n) {
  if (!ctx || !n) return SPV_EINVAL;
  *n = ctx->launches;
  return 0;
}

SPV_API int spv_d2h_bytes(spv_ctx *ctx, unsigned long long *n) {
  if (!ctx || !n) return SPV_EINVAL;
  *n = ctx->d2h_bytes;
  return 0;
}

}  // extern "C"
