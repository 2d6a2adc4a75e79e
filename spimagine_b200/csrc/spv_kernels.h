// spv_kernels.h -- argument blocks and host-side launchers shared by the kernel files and the C ABI.
#pragma once
#include "../../include/spimcuda.h"
#include "spv_common.cuh"

namespace spv {

// Sort-last composite over peer memory: rank r stores the raw partial maxima of the pixels in image band o straight
// into the staging of band o's owner (its own memory when o == r, NVLink peer memory otherwise).
constexpr int MAX_WORLD = 16;
constexpr int MAX_EXTRA_SLABS = 3;
struct PushArgs {
  float *part[MAX_WORLD];  // owner o's staging [parity][src rank][band_rows * width]
  int band_rows;           // image rows per band (multiple of 4: a warp's 8x4 tile lies in one band)
  unsigned src_off;        // (parity * world + my rank) * band_rows * width
};

struct MipArgs {
  Camera cam;
  Volume vol;
  const float2 *coarse;  // coarse min/max grid: one cell = 4^3 bricks
  int cgx, cgy, cgz;
  float box[6];
  float min_val, max_val, gamma, alpha_pow;
  int num_parts, current_part, max_steps, flags;
  int tile_variant;  // CTA shape of the fast kernel (tuning)
  int width, height;
  int y_begin, y_end;  // pixel rows this launch renders (a band; y_begin is a multiple of 16), fast kernel only
  float *out, *alpha, *raw;
  unsigned long long *stats;  // [hit rays, texture samples issued] or nullptr
  unsigned *tile_counter;     // non-null: persistent CTAs pull tiles from this counter
  unsigned *band_done;        // non-null (static grid only): band_done[b] counts the CTAs of band b (band_rows image
  int band_rows;              // rows each) whose results are in memory -- the copy stream waits on these counters
  int row_mode;               // order the tile rows are dealt in when band_done is set: 0 = from the top and bottom edges
                              // inwards, 1 = the rows outside [hit_tile_a, hit_tile_b) first, then that range top to bottom
  unsigned hit_tile_a, hit_tile_b;  // tile rows (8 pixels) the projected box can touch (conservative; any value is correct)
  Volume extra[MAX_EXTRA_SLABS];  // further slabs of the same global volume resident on this GPU (slab renders):
  int n_extra;                    // one ray setup, every slab's owned interval marched in turn
  PushArgs push;              // used when flags has SPV_MIP_PUSH
};

struct IsoArgs {
  Camera cam;
  Volume vol;
  const float2 *coarse;
  int cgx, cgy, cgz;
  const float2 *top;  // min/max over 4^3 coarse cells (128^3 texels)
  int tgx, tgy, tgz;
  float box[6];
  float iso_val, gamma;
  int max_steps;
  int skip;  // empty-space skipping on the min/max grids (texture-unit path)
  int segments;   // warps that share a ray of the texture-unit search, each taking a segment of its samples: 1, 2 or 4
                  // (tuning knob 4)
  int centre_out; // CTAs are dealt from the image centre outwards (tuning knob 5, default on)
  int width, height;
  float *out, *alpha, *depth, *normals;
  unsigned char *tile_hit;  // one flag per 8x4 warp tile (texture-unit path) or nullptr
  unsigned long long *stats;
};

struct ConvWeights {
  float w[32];  // exp(coef*(ht - Nh/2)^2/Nh^2), ht = 0..Nh-1, evaluated on the host
  int nh;
};

cudaError_t launch_mip(const MipArgs &a, int dtype, bool linear, bool fast, bool exact, bool skip, bool slab,
                       bool stats, cudaStream_t st);
cudaError_t launch_mip_finish(const float *raw, float *out, int n, float minVal, float maxVal, float gamma,
                              cudaStream_t st);

// software-sampled max projection (spv_mip_smem.cu): TMA-staged shared-memory slabs of permuted linear uint16 copies
int mip_smem_configs();               // box / ring geometries compiled in (tuning knob 11)
void mip_smem_box(int cfg, int *abp);  // box extents (contiguous axis, second axis, planes) the tensor maps of cfg need
// tex_of8: of every 8 tiles (fixed pattern) this many are rendered through the texture unit instead (hybrid mode)
cudaError_t launch_mip_smem(const MipArgs &a, int fmt, int cfg, const void *maps /* CUtensorMap[3] of cfg */, int tex_of8,
                            cudaStream_t st);
// dst[(d * NB + b) * pitchA + a] = the volume with slowest axis D (x, y or z) and contiguous axis z, x, x
cudaError_t launch_permute(const Volume &V, int fmt, int D, int NA, int NB, int ND, size_t pitchA, void *dst,
                           cudaStream_t st);


// View-aligned layered copies and multi-frame launches (spv_mip_axis.cu).
// The integer volume is resident as 2-D layered arrays of pairs {v[l], v[l+1]} along a LAYER AXIS (0 x, 1 y, 2 z; the
// z copy is the primary array).  A frame picks the axis and the way a warp's lanes map to pixels so that the four lanes
// of a hardware quad (one texture request) and the consecutive samples of a ray stay inside one layer:
//   quad 0: 2x2-pixel quads, 8x4 warp tiles, 2x2 warps      (layer axis along the view direction)
//   quad 1: 4x1 row quads,  16x2 warp tiles, 1x4 warps      (layer axis along the camera's up direction)
//   quad 2: 1x4 column quads, 4x8 warp tiles, 4x1 warps     (layer axis along the camera's right direction)
// A CTA is 16x8 pixels in every mode.  One launch renders up to MAX_BATCH frames that share everything but the model
// view; CTAs are dealt (tile row, frame, tile x), so the frames' CTAs of one tile row run together and what one of them
// pulls into L2 the others find there.
constexpr int MAX_BATCH = 16;
struct MipAxisArgs {
  float invP[16];
  float invM[MAX_BATCH][16];
  float *out[MAX_BATCH], *alpha[MAX_BATCH];
  cudaTextureObject_t tex[3];  // filtered pair textures by layer axis (0: not built)
  unsigned char lax[MAX_BATCH], quad[MAX_BATCH];
  // CTA tiles (16x8 pixels) [tile_x0, tile_x0 + tile_nx) x [tile_y0, tile_y0 + tile_ny) of frame f are launched.  Of
  // those, the tiles [rend_x0, rend_x1) x [rend_y0, rend_y1) -- the ones the projected box can touch -- are rendered,
  // the others are filled with the miss values (what an earlier frame left in these planes outside the new rectangle).
  unsigned short tile_x0[MAX_BATCH], tile_y0[MAX_BATCH], tile_nx[MAX_BATCH], tile_ny[MAX_BATCH];
  unsigned short rend_x0[MAX_BATCH], rend_x1[MAX_BATCH], rend_y0[MAX_BATCH], rend_y1[MAX_BATCH];
  int nx, ny, nz;              // volume extent
  float scale;                 // 65535 or 255
  float box[6];
  float min_val, max_val, gamma;
  float alpha_pow;             // != 0: front-to-back attenuation (the ALPHA instantiation)
  int max_steps;
  int width, height;
  int n_frames;
  int y_begin, y_end;          // pixel rows this launch renders (single-frame band launches; multiples of 16)
  unsigned *band_done;         // as MipArgs (single frames only)
  int band_rows, row_mode;
  unsigned hit_tile_a, hit_tile_b;
};
cudaError_t launch_mip_axis(const MipAxisArgs &a, int dtype, cudaStream_t st);
cudaError_t preload_mip_axis();
// dst (surface of a layered pair array) = the volume re-layered along axis lax (0: width nz, height ny, nx layers;
// 1: width nx, height nz, ny layers), read from the primary z-pair array through its point texture
cudaError_t launch_axis_pair(const Volume &V, int dtype, int lax, cudaSurfaceObject_t dst, cudaStream_t st);

// peer composite (spv_comp.cu)
struct CompFinishArgs {
  const float *part;          // my band's staging of this parity: [world][band_pixels]
  float *out[MAX_WORLD];      // every rank's SPV_BUF_OUT plane
  int world;
  unsigned band_pixels;       // band_rows * width (multiple of 4)
  unsigned first_pixel;       // my band's first pixel in the image
  unsigned n_pixels;          // pixels of my band that exist (the last band may be cut by the image edge)
  float min_val, max_val, gamma;
};
cudaError_t launch_comp_sync(unsigned *const *peer_flags, const unsigned *flags, int world, int rank, int phase,
                             unsigned value, unsigned *err, cudaStream_t st);
cudaError_t launch_comp_finish(const CompFinishArgs &a, cudaStream_t st);
// rows [y_first, y_end) of the finished planes (out, occlusion, blurred normals) of a sort-last iso frame go from the rank
// that computed them into every other rank's planes (128-bit peer stores)
struct BandGatherArgs {
  float *planes[MAX_WORLD];  // rank r's [out | alpha | depth | occ | normals(3) | ...]
  int world, rank, width, height, y_first, y_end;
};
cudaError_t launch_band_gather(const BandGatherArgs &a, cudaStream_t st);
struct PeerFlagPtrs { unsigned *p[MAX_WORLD]; };
// load every kernel a sort-last frame launches before any of them can be spinning on a peer (lazy module loading)
cudaError_t preload_mip_kernels();
cudaError_t preload_comp_kernels();
cudaError_t preload_iso_kernels();

cudaError_t launch_sample_points(const Volume &V, int dtype, bool linear, bool exact, const float *pos, int n,
                                 float *out, cudaStream_t st);
cudaError_t launch_texrate_probe(const Volume &V, int dtype, bool linear, int blocks, int iters, const float *vec9,
                                 float *sink, cudaStream_t st);

cudaError_t launch_iso(const IsoArgs &a, int dtype, bool linear, bool exact, bool stats, cudaStream_t st);
// Sort-last iso surface over peer memory (spv_render_iso_composite): world > 0 switches the two slab kernels from
// their local planes to the peers' memory.
//   search   k1 / k0 of band o's pixels go to owner o's staging  kpart[o][src][plane][band]
//   resolve  the rank owning a pixel's crossing stores the finished pixel into EVERY rank's result planes; pixels
//            without a crossing are cleared locally (depth = INFINITY), pixels resolved elsewhere are left alone
struct IsoPeer {
  int world;                 // 0: off (local planes; the caller reduces with NCCL)
  int band_rows;             // image rows per band
  unsigned src;              // my rank
  unsigned band;             // band_rows * width
  int *kpart[MAX_WORLD];     // owner o's staging of this frame's parity
  float *planes[MAX_WORLD];  // rank r's [out | alpha | depth | occ | normals(3) | ...] planes
  int normals_plane;         // plane index the resolved normals go to: 4 (normals) or 9 (tmp_vec, blurred into 4 later)
};
// sort-last iso surface: phase 0 = search (writes k1 / k0), 1 = resolve (reads them), 2 = fix-up after the SUM
cudaError_t launch_iso_slab(const IsoArgs &a, int dtype, bool linear, int phase, int *k1, int *k0, float *occ,
                            unsigned *err, cudaStream_t st, const IsoPeer *peer = nullptr);
// owner of a band: element-wise MIN of the k1 / k0 partials of all ranks, stored into every rank's k planes
struct KReduceArgs {
  const int *part;           // my staging of this parity: [world][2][band]
  int *kplanes[MAX_WORLD];   // rank r's k planes: [2][n_image]
  int world;
  unsigned band, first_pixel, n_pixels;  // as CompFinishArgs
  size_t n_image;            // width * height
};
cudaError_t launch_k_reduce(const KReduceArgs &a, cudaStream_t st);
// buf -> tmp (x pass), tmp -> buf (y pass); ncomp = 1 (conv_x/conv_y) or 3 (conv_vec_x/conv_vec_y)
cudaError_t launch_conv(float *buf, float *tmp, int width, int height, int ncomp, const ConvWeights &w,
                        cudaStream_t st);
// both passes in one launch (in != out), bit-identical to launch_conv.  tile_hit (may be null): `in` is +0 on every
// pixel further than `reach` pixels from a flagged 8x4 tile; such regions are filled with zeros without the arithmetic
// y_first / y_end: only these rows are written (default: all) -- one rank's band of a sort-last frame
// shade (ncomp == 1 only, may be null): the shading pass in the blur's epilogue -- out[p] = launch_shading's value for
// the pixel, computed from the blurred occlusion just stored, normals and depth (same result as the separate launch)
struct ShadeArgs {
  float *out;
  Camera cam;
  float occ_strength;
  const float *normals, *depth;
};
cudaError_t launch_conv_xy(const float *in, float *out, int width, int height, int ncomp, const ConvWeights &w,
                           const unsigned char *tile_hit, int reach, cudaStream_t st, int y_first = 0, int y_end = -1,
                           const ShadeArgs *shade = nullptr);
// taps[i] = the four rand_int() values of occlusion tap i (they depend on i only)
cudaError_t launch_occ_taps(float4 *taps, int n, cudaStream_t st);
// queue (occ_queue_bytes, zero-initialised) + tile flags: only the pixel blocks in reach of a surface are computed, handed
// out dynamically; `frame` must change parity from call to call on the same queue.  Without: every block, in place.
size_t occ_queue_bytes(int width, int height);
extern int occ_ctas_per_sm;
// `table` (may be null; launch_occ_table for this image size, radius and tap count): the taps' pixel offsets are read
// from it instead of being hashed per frame, depth gathers go through a shared-memory tile (same result, bit for bit)
cudaError_t launch_occlusion(float *occ, int width, int height, int radius, int n_points, const float *depth,
                             const unsigned char *tile_hit, const float4 *taps, unsigned *queue, unsigned frame, int sms,
                             cudaStream_t st, int y_first = 0, int y_end = -1, const void *table = nullptr);
size_t occ_table_bytes(int width, int height, int radius, int n_points);  // 0: no table for these parameters
cudaError_t launch_occ_table(void *table, int width, int height, int radius, int n_points, const float4 *taps, unsigned *bad,
                             cudaStream_t st);
cudaError_t launch_shading(float *out, int width, int height, const Camera &cam, float occ_strength,
                           const float *normals, const float *depth, const float *occ, cudaStream_t st, int y_first = 0,
                           int y_end = -1);

// min/max brick grids of the resident volume (built from the point texture)
cudaError_t launch_build_bricks(const Volume &vol, int dtype, int local_nz, float2 *bricks, float2 *coarse, int cgx,
                                int cgy, int cgz, float2 *top, float *minmax /* [2] device */, cudaStream_t st);

// LAYOUT_ZPAIR ingest: dst (linear, z-major) = {src[z], src[min(z+1, nz-1)]} for z in [zbeg, zend)
cudaError_t launch_pair(const void *src, void *dst, int dtype, size_t slice, int nz, int zbeg, int zend,
                        cudaStream_t st);

// display hand-off (spv_display.cu): value plane + LUT -> packed RGBA8
cudaError_t launch_rect_copy(const float *src, float *dst, int W, int H, int xa, int xb, int ya, int yb, int planes, cudaStream_t st);
cudaError_t launch_display(const float *value, const float *alpha, const float *lut, int n_lut, int mode_black,
                           void *rgba, size_t n, cudaStream_t st);

// ingest of host arrays of another element type: dst[i] = (dst type) src[i], n elements (src_type: SPV_SRC_*)
size_t src_elem_size(int src_type);
cudaError_t launch_convert(const void *src, void *dst, int src_type, int dtype, size_t n, cudaStream_t st);

// separable 3-D convolution of a volume (spv_filter.cu; gputools.convolve_sep3 behind BlurProcessor.apply,
// models/imageprocessor.py:47-71).  h: host taps; d_taps: the same taps in device memory (used for nh > FILTER_MAX_TAPS)
constexpr int FILTER_MAX_TAPS = 63;
struct FilterTaps { float w[FILTER_MAX_TAPS + 1]; };
void parallel_memcpy(void *dst, const void *src, size_t n);  // several host threads (spv_filter.cu)
cudaError_t launch_filter_x(const void *in, int dtype, float *out, int nx, int ny, int nz, const float *h, int nh,
                            const float *d_taps, cudaStream_t st);
cudaError_t launch_filter_axis(const float *in, float *out, int nx, int ny, int nz, int axis, const float *h, int nh,
                               const float *d_taps, cudaStream_t st);

}  // namespace spv
