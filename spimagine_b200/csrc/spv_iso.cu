// spv_iso.cu -- iso-surface ray march and its screen-space post passes.  Replaces
//   iso_surface            spimagine/volumerender/kernels/iso_kernel.cl:17-225
//   conv_x/y, conv_vec_x/y spimagine/volumerender/kernels/convolve_2d.cl:6-137
//   occlusion              spimagine/volumerender/kernels/occlusion.cl:41-82  (+ utils.cl:10-38)
//   shading                spimagine/volumerender/kernels/iso_kernel.cl:505-588
// in the launch order of VolumeRenderer._render_isosurface (volumerender.py:446-506).
#include <string.h>

#include "spv_kernels.h"

namespace spv {

#ifndef SPV_ISO_MINB
#define SPV_ISO_MINB 6  // resident CTAs per SM the texture-unit search is compiled for (register budget)
#endif

// -------------------------------------------------------------------------------------------------------------
// iso_surface.  One warp = one 8x4 pixel tile (same lane order as the MIP kernel).  The coarse search keeps the
// reference's sample positions (pos += delta accumulation) and its first-crossing rule; the bracket refinement
// and the 12-tap gradient follow iso_kernel.cl:140-200 operation by operation.
template <int FMT, bool LINEAR, bool EXACT, bool STATS>
__global__ void __launch_bounds__(128) iso_kernel(const IsoArgs a) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int lx = (lane & 1) | ((lane >> 1) & 2) | ((lane >> 2) & 4);
  const int ly = ((lane >> 1) & 1) | ((lane >> 2) & 2);
  const unsigned x = blockIdx.x * 16 + (warp & 1) * 8 + lx, y = blockIdx.y * 8 + (warp >> 1) * 4 + ly;
  const unsigned Nx = a.width, Ny = a.height;
  if (x >= Nx || y >= Ny) return;
  const size_t p = x + (size_t)Nx * y;
  const Volume &V = a.vol;
  const float INF = __int_as_float(0x7f800000);
  unsigned long long nfetch = 0;

  Ray r = make_ray(x, y, Nx, Ny, a.cam, a.box);
  bool hitIso = false;
  float tnear = r.tnear, t_hit = INF;
  v4 normal = mk4(0.f, 0.f, 0.f, 0.f);
  float colVal = 0.f;
  if (r.hit) {
    const v4 orig = r.orig, direc = r.direc;
    if (tnear < 0.0f) tnear = 0.0f;
    const float isoVal = a.iso_val;
    const int maxSteps = a.max_steps;
    const float dt = 1.f * (r.tfar - tnear) / ((float)maxSteps - 1.f);
    const v4 delta_pos = scl4(.5f * dt, direc);
    const v4 pos0 = scl4(0.5f, add4(sadd4(1.f, orig), scl4(tnear, direc)));
    v4 pos = pos0;
    // :106 reads the first sample with read_imagef whatever the image type; taken as a proper read
    float newVal = sample<FMT, LINEAR, EXACT>(V, pos.x, pos.y, pos.z);
    const bool isGreater = newVal > isoVal;
    int i = 1;
    for (i = 1; i < maxSteps; i++) {
      pos = add4(pos, delta_pos);
      t_hit = tnear + (float)i * dt;
      newVal = sample<FMT, LINEAR, EXACT>(V, pos.x, pos.y, pos.z);
      if ((newVal > isoVal) != isGreater) {
        hitIso = true;
        break;
      }
    }
    if (STATS) nfetch += (unsigned long long)min(i, maxSteps - 1) + 1ull;
    if (hitIso) {
      const int maxBisect = 10;
      const v4 delta_pos2 = mk4(delta_pos.x / (float)maxBisect, delta_pos.y / (float)maxBisect,
                                delta_pos.z / (float)maxBisect, delta_pos.w / (float)maxBisect);
      const float dt2 = dt / (float)maxBisect;
      pos = add4(pos0, scl4((float)(i - 1), delta_pos));
      for (int j = 1; j <= maxBisect; j++) {
        newVal = sample<FMT, LINEAR, EXACT>(V, pos.x, pos.y, pos.z);
        pos = add4(pos, delta_pos2);
        t_hit += dt2;
        if (STATS) ++nfetch;
        if ((newVal > isoVal) != isGreater) break;
      }
      v4 light = mk4(2.f, -1.f, -2.f, 0.f);
      const float c_ambient = .3f, c_diffuse = .4f, c_specular = .3f;
      light = mult(a.cam.invM, light);
      light = normalize4(light);
      float h = dt;
      h *= (a.gamma * a.gamma);  // pow(gamma, 2.f)
      const float h2 = 2.f * h;
#define SPV_S(dx, dy, dz) sample<FMT, LINEAR, EXACT>(V, pos.x + (dx), pos.y + (dy), pos.z + (dz))
      normal.x = 2.f * SPV_S(h, 0.f, 0.f) - 2.f * SPV_S(-h, 0.f, 0.f) + SPV_S(h2, 0.f, 0.f) - SPV_S(-h2, 0.f, 0.f);
      normal.y = 2.f * SPV_S(0.f, h, 0.f) - 2.f * SPV_S(0.f, -h, 0.f) + SPV_S(0.f, h2, 0.f) - SPV_S(0.f, -h2, 0.f);
      normal.z = SPV_S(0.f, 0.f, h) - SPV_S(0.f, 0.f, -h) + SPV_S(0.f, 0.f, h2) - SPV_S(0.f, 0.f, -h2);
#undef SPV_S
      if (STATS) nfetch += 12;
      normal.w = 0.f;
      normal = scl4(1.f - (float)(2 * (int)isGreater), normalize4(normal));
      const v4 reflect = sub4(scl4(2.f * dot4(light, normal), normal), light);
      const float diffuse = fmaxf(0.f, dot4(light, normal));
      const float specular = powf(fmaxf(0.f, dot4(normalize4(reflect), normalize4(direc))), 10.f);
      colVal = c_ambient + c_diffuse * diffuse + (diffuse > 0.f ? 1.f : 0.f) * c_specular * specular;
    }
  }
  if (hitIso) {
    a.out[p] = colVal;
    a.alpha[p] = tnear;
    a.depth[p] = t_hit;
    a.normals[3 * p + 0] = normal.x;
    a.normals[3 * p + 1] = normal.y;
    a.normals[3 * p + 2] = normal.z;
  } else {
    a.out[p] = 0.f;
    a.alpha[p] = 0.f;
    a.depth[p] = INF;
    a.normals[3 * p + 0] = 0.f;
    a.normals[3 * p + 1] = 0.f;
    a.normals[3 * p + 2] = 0.f;
  }
  if (STATS && a.stats) {
    atomicAdd(a.stats + 0, r.hit ? 1ull : 0ull);
    atomicAdd(a.stats + 1, nfetch);
  }
}

// -------------------------------------------------------------------------------------------------------------
// The texture-unit path.  Same first-crossing rule, bracket refinement and 12-tap gradient as iso_kernel, but
//   * sample k sits at pos0 + k*delta (one fma per axis, unnormalised texel coordinates) instead of being
//     accumulated, so that
//   * the coarse search fetches BATCH samples at a time (independent fetches in flight instead of one
//     fetch -> compare -> branch round trip per sample) and then looks for the first crossing in the batch;
//     at most BATCH-1 samples behind the crossing are fetched in vain,
//   * the ten refinement samples are fetched together as well.
// The per-ray state is shared by the single-GPU kernel and the two sort-last kernels (search / resolve), which
// therefore evaluate bit-identical expressions.
struct IsoRay {
  bool hit;                // the ray meets the box
  float tnear, dt;
  v4 direc, pos0, delta_pos;
  float u0, v0, w0, du, dv, dw;  // unnormalised texel coordinates of sample t: u0 + t*du
};

__device__ __forceinline__ IsoRay iso_ray(const IsoArgs &a, unsigned x, unsigned y, bool inb) {
  IsoRay q;
  Ray r = make_ray(x, y, a.width, a.height, a.cam, a.box);
  q.hit = r.hit && inb;
  q.tnear = r.tnear;
  if (q.tnear < 0.0f) q.tnear = 0.0f;
  q.direc = r.direc;
  q.dt = 1.f * (r.tfar - q.tnear) / ((float)a.max_steps - 1.f);
  q.delta_pos = scl4(.5f * q.dt, r.direc);
  q.pos0 = scl4(0.5f, add4(sadd4(1.f, r.orig), scl4(q.tnear, r.direc)));
  const Volume &V = a.vol;
  q.u0 = q.pos0.x * V.fnx; q.v0 = q.pos0.y * V.fny; q.w0 = q.pos0.z * V.fnz;
  q.du = q.delta_pos.x * V.fnx; q.dv = q.delta_pos.y * V.fny; q.dw = q.delta_pos.z * V.fnz;
  return q;
}

template <int FMT, bool LINEAR>
__device__ __forceinline__ float iso_at(const Volume &V, const IsoRay &q, float t) {
  return sample_tmu_uvw<FMT, LINEAR>(V, fmaf(t, q.du, q.u0), fmaf(t, q.dv, q.v0), fmaf(t, q.dw, q.w0));
}

// exact empty-space test for sample t: a sample whose footprint lies in a cell with max <= iso cannot be "> iso",
// one in a cell with min > iso cannot be "<= iso".  want_greater: are we looking for a sample > iso?
// COARSE_FIRST: look at the 32^3 cell first (cheaper when most of them are decided); false where the caller has
// just established that the ray is inside an undecided coarse cell.
template <bool COARSE_FIRST = true>
__device__ __forceinline__ bool iso_cell_may_hold(const IsoArgs &a, const IsoRay &q, float t, bool want_greater) {
  const Volume &V = a.vol;
  const float cx0 = q.u0 - 0.5f, cy0 = q.v0 - 0.5f, cz0 = q.w0 - 0.5f - (float)V.z_lo;
  const int ix = __float2int_rd(fmaf(t, q.du, cx0)), iy = __float2int_rd(fmaf(t, q.dv, cy0)),
            iz = __float2int_rd(fmaf(t, q.dw, cz0));
  bool need = true;
  if (COARSE_FIRST) {
    const int cx = min(max(ix >> (BRICK_SHIFT + 2), 0), a.cgx - 1), cy = min(max(iy >> (BRICK_SHIFT + 2), 0), a.cgy - 1),
              cz = min(max(iz >> (BRICK_SHIFT + 2), 0), a.cgz - 1);
    const float2 c = __ldg(a.coarse + ((size_t)cz * a.cgy + cy) * a.cgx + cx);
    need = want_greater ? (c.y > a.iso_val) : !(c.x > a.iso_val);
  }
  if (need) {
    const int bx = min(max(ix >> BRICK_SHIFT, 0), V.gx - 1), by = min(max(iy >> BRICK_SHIFT, 0), V.gy - 1),
              bz = min(max(iz >> BRICK_SHIFT, 0), V.gz - 1);
    const float2 b = brick_at(V, bx, by, bz);
    need = want_greater ? (b.y > a.iso_val) : !(b.x > a.iso_val);
  }
  return need;
}

// class of sample t by the 8^3 brick its footprint starts in: 2 = certainly <= iso, 1 = certainly > iso, 0 = fetch it
__device__ __forceinline__ int iso_brick_class(const IsoArgs &a, const IsoRay &q, float t) {
  const Volume &V = a.vol;
  const float cx0 = q.u0 - 0.5f, cy0 = q.v0 - 0.5f, cz0 = q.w0 - 0.5f - (float)V.z_lo;
  const int ix = __float2int_rd(fmaf(t, q.du, cx0)), iy = __float2int_rd(fmaf(t, q.dv, cy0)),
            iz = __float2int_rd(fmaf(t, q.dw, cz0));
  const int bx = min(max(ix >> BRICK_SHIFT, 0), V.gx - 1), by = min(max(iy >> BRICK_SHIFT, 0), V.gy - 1),
            bz = min(max(iz >> BRICK_SHIFT, 0), V.gz - 1);
  const float2 b = brick_at(V, bx, by, bz);
  return !(b.y > a.iso_val) ? 2 : ((b.x > a.iso_val) ? 1 : 0);
}

// ---- hierarchical empty-space traversal --------------------------------------------------------------------------
// The per-sample test above costs ~25 instructions for every one of the max_steps samples of a ray, although most
// rays spend most of their length in cells that cannot hold a crossing.  The traversal below jumps over such a cell
// in one step.  It stays exact: the texel index floor(fma(t, d, c)) of every axis is monotone in t, so if samples k and
// e lie in the same cell, every sample in between does; the parametric exit is only a guess for e, which is then
// checked with the very expression that classifies a single sample.
struct IsoDda {
  float cx0, cy0, cz0;  // footprint origin of sample 0 in texel units (z: local slices)
  float rdu, rdv, rdw;  // 1 / step per axis (unused where the step is 0)
};
__device__ __forceinline__ IsoDda iso_dda(const IsoArgs &a, const IsoRay &q) {
  IsoDda d;
  d.cx0 = q.u0 - 0.5f; d.cy0 = q.v0 - 0.5f; d.cz0 = q.w0 - 0.5f - (float)a.vol.z_lo;
  d.rdu = 1.f / q.du; d.rdv = 1.f / q.dv; d.rdw = 1.f / q.dw;
  return d;
}
struct IsoLevel {
  const float2 *grid;
  int gx, gy, gz;
};
template <int SHIFT>
__device__ __forceinline__ void iso_cell_of(const IsoRay &q, const IsoDda &d, const IsoLevel &L, float t, int &cx, int &cy,
                                            int &cz) {
  cx = min(max(__float2int_rd(fmaf(t, q.du, d.cx0)) >> SHIFT, 0), L.gx - 1);
  cy = min(max(__float2int_rd(fmaf(t, q.dv, d.cy0)) >> SHIFT, 0), L.gy - 1);
  cz = min(max(__float2int_rd(fmaf(t, q.dw, d.cz0)) >> SHIFT, 0), L.gz - 1);
}
// first sample after k (at most kend) that lies in another cell than sample k, whose cell is (cx, cy, cz)
template <int SHIFT>
__device__ __forceinline__ int iso_cell_exit(const IsoRay &q, const IsoDda &d, const IsoLevel &L, int k, int kend, int cx,
                                             int cy, int cz) {
  float te = (float)(kend - 1);
  if (q.du > 0.f) { if (cx < L.gx - 1) te = fminf(te, ((float)((cx + 1) << SHIFT) - d.cx0) * d.rdu); }
  else if (q.du < 0.f) { if (cx > 0) te = fminf(te, ((float)(cx << SHIFT) - d.cx0) * d.rdu); }
  if (q.dv > 0.f) { if (cy < L.gy - 1) te = fminf(te, ((float)((cy + 1) << SHIFT) - d.cy0) * d.rdv); }
  else if (q.dv < 0.f) { if (cy > 0) te = fminf(te, ((float)(cy << SHIFT) - d.cy0) * d.rdv); }
  if (q.dw > 0.f) { if (cz < L.gz - 1) te = fminf(te, ((float)((cz + 1) << SHIFT) - d.cz0) * d.rdw); }
  else if (q.dw < 0.f) { if (cz > 0) te = fminf(te, ((float)(cz << SHIFT) - d.cz0) * d.rdw); }
  int e = max(__float2int_rd(te), k);  // guess: the last sample inside
  int ex, ey, ez;
  if (e > k) {
    iso_cell_of<SHIFT>(q, d, L, (float)e, ex, ey, ez);
    if (ex != cx || ey != cy || ez != cz) {
      --e;
      if (e > k) {
        iso_cell_of<SHIFT>(q, d, L, (float)e, ex, ey, ez);
        if (ex != cx || ey != cy || ez != cz) e = k;
      }
    }
  }
  return e + 1;
}
// class of the cell sample k lies in: 2 = every sample in it is <= iso, 1 = every sample is > iso, 0 = undecided;
// kexit = first sample outside the cell when the class is not 0
template <int SHIFT>
__device__ __forceinline__ int iso_cell_class(const IsoArgs &a, const IsoRay &q, const IsoDda &d, const IsoLevel &L, int k,
                                              int kend, int &kexit) {
  int cx, cy, cz;
  iso_cell_of<SHIFT>(q, d, L, (float)k, cx, cy, cz);
  const float2 c = __ldg(L.grid + ((size_t)cz * L.gy + cy) * L.gx + cx);
  const int cls = !(c.y > a.iso_val) ? 2 : ((c.x > a.iso_val) ? 1 : 0);
  if (cls) kexit = iso_cell_exit<SHIFT>(q, d, L, k, kend, cx, cy, cz);
  return cls;
}
constexpr int COARSE_SHIFT = BRICK_SHIFT + 2, TOP_SHIFT = BRICK_SHIFT + 4;
// One traversal step over both levels: the two cells of sample k are looked up together (independent loads: a warp
// deep in the traversal is bound by their latency).  Returns k if neither cell is of class `skip_cls`, else the first
// sample outside the larger skippable cell.
__device__ __forceinline__ int iso_skip_step(const IsoArgs &a, const IsoRay &q, const IsoDda &d, const IsoLevel &Lt,
                                             const IsoLevel &Lc, int k, int kend, int skip_cls) {
  const float t = (float)k;
  const int ix = __float2int_rd(fmaf(t, q.du, d.cx0)), iy = __float2int_rd(fmaf(t, q.dv, d.cy0)),
            iz = __float2int_rd(fmaf(t, q.dw, d.cz0));
  const int tx = min(max(ix >> TOP_SHIFT, 0), Lt.gx - 1), ty = min(max(iy >> TOP_SHIFT, 0), Lt.gy - 1),
            tz = min(max(iz >> TOP_SHIFT, 0), Lt.gz - 1);
  const int cx = min(max(ix >> COARSE_SHIFT, 0), Lc.gx - 1), cy = min(max(iy >> COARSE_SHIFT, 0), Lc.gy - 1),
            cz = min(max(iz >> COARSE_SHIFT, 0), Lc.gz - 1);
  const float2 vt = __ldg(Lt.grid + ((size_t)tz * Lt.gy + ty) * Lt.gx + tx);
  const float2 vc = __ldg(Lc.grid + ((size_t)cz * Lc.gy + cy) * Lc.gx + cx);
  const int cls_t = !(vt.y > a.iso_val) ? 2 : ((vt.x > a.iso_val) ? 1 : 0);
  const int cls_c = !(vc.y > a.iso_val) ? 2 : ((vc.x > a.iso_val) ? 1 : 0);
  if (cls_t == skip_cls) return iso_cell_exit<TOP_SHIFT>(q, d, Lt, k, kend, tx, ty, tz);
  if (cls_c == skip_cls) return iso_cell_exit<COARSE_SHIFT>(q, d, Lc, k, kend, cx, cy, cz);
  return k;
}

// bracket refinement, 12-tap gradient and Phong shading at crossing sample i (iso_kernel.cl:140-215)
template <int FMT, bool LINEAR>
__device__ __forceinline__ void iso_resolve(const IsoArgs &a, const IsoRay &q, int i, bool isGreater, float &t_hit,
                                            v4 &normal, float &colVal) {
  const Volume &V = a.vol;
  const float isoVal = a.iso_val;
  t_hit = q.tnear + (float)i * q.dt;
  const int maxBisect = 10;
  const float dt2 = q.dt / (float)maxBisect;
  float v[maxBisect];
#pragma unroll
  for (int j = 0; j < maxBisect; ++j) v[j] = iso_at<FMT, LINEAR>(V, q, (float)(i - 1) + (float)j / (float)maxBisect);
  int J = maxBisect;
#pragma unroll
  for (int j = maxBisect - 1; j >= 0; --j)
    if ((v[j] > isoVal) != isGreater) J = j + 1;
  for (int j = 0; j < J; ++j) t_hit += dt2;  // accumulated like the reference does
  // where the reference's `pos` stands after the refinement loop, in normalised coordinates
  const float ts = (float)(i - 1) + (float)J / (float)maxBisect;
  const float px = fmaf(ts, q.delta_pos.x, q.pos0.x), py = fmaf(ts, q.delta_pos.y, q.pos0.y),
              pz = fmaf(ts, q.delta_pos.z, q.pos0.z);
  v4 light = mk4(2.f, -1.f, -2.f, 0.f);
  const float c_ambient = .3f, c_diffuse = .4f, c_specular = .3f;
  light = mult(a.cam.invM, light);
  light = normalize4(light);
  float h = q.dt;
  h *= (a.gamma * a.gamma);
  const float h2 = 2.f * h;
#define SPV_S(dx, dy, dz) sample_tmu<FMT, LINEAR>(V, px + (dx), py + (dy), pz + (dz))
  const float xa = SPV_S(h, 0.f, 0.f), xb = SPV_S(-h, 0.f, 0.f), xc = SPV_S(h2, 0.f, 0.f), xd = SPV_S(-h2, 0.f, 0.f);
  const float ya = SPV_S(0.f, h, 0.f), yb = SPV_S(0.f, -h, 0.f), yc = SPV_S(0.f, h2, 0.f), yd = SPV_S(0.f, -h2, 0.f);
  const float za = SPV_S(0.f, 0.f, h), zb = SPV_S(0.f, 0.f, -h), zc = SPV_S(0.f, 0.f, h2), zd = SPV_S(0.f, 0.f, -h2);
#undef SPV_S
  normal.x = 2.f * xa - 2.f * xb + xc - xd;
  normal.y = 2.f * ya - 2.f * yb + yc - yd;
  normal.z = za - zb + zc - zd;
  normal.w = 0.f;
  normal = scl4(1.f - (float)(2 * (int)isGreater), normalize4(normal));
  const v4 reflect = sub4(scl4(2.f * dot4(light, normal), normal), light);
  const float diffuse = fmaxf(0.f, dot4(light, normal));
  const float specular = powf(fmaxf(0.f, dot4(normalize4(reflect), normalize4(q.direc))), 10.f);
  colVal = c_ambient + c_diffuse * diffuse + (diffuse > 0.f ? 1.f : 0.f) * c_specular * specular;
}

// CTA of 4 warps = 16x8 pixels (2x2 warp tiles), 2 warps = 16x4, 1 warp = 8x4.
// CTA (bx, by) of the grid renders tile (tile_x, tile_y).  centre_out: the grid's rows and columns are dealt from the
// middle of the image outwards (0 -> mid, 1 -> mid-1, 2 -> mid+1, ...), so the CTAs launched first are the ones whose
// rays cross the most of the volume and the last ones launched are cheap: the launch ends without a tail of long CTAs.
__device__ __forceinline__ unsigned centre_out(unsigned r, unsigned n) {
  const unsigned mid = n >> 1;
  return (r & 1u) ? mid - ((r + 1u) >> 1) : mid + (r >> 1);
}
__device__ __forceinline__ void tile_of_cta(bool centre, unsigned &tile_x, unsigned &tile_y) {
  tile_x = centre ? centre_out(blockIdx.x, gridDim.x) : blockIdx.x;
  tile_y = centre ? centre_out(blockIdx.y, gridDim.y) : blockIdx.y;
}
__device__ __forceinline__ void tile_pixel(unsigned &x, unsigned &y, bool centre = false) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int lx = (lane & 1) | ((lane >> 1) & 2) | ((lane >> 2) & 4);
  const int ly = ((lane >> 1) & 1) | ((lane >> 2) & 2);
  const int cw = blockDim.x >= 64 ? 16 : 8, ch = blockDim.x >= 128 ? 8 : 4;
  unsigned tx, ty;
  tile_of_cta(centre, tx, ty);
  x = tx * cw + (warp & 1) * 8 + lx;
  y = ty * ch + (warp >> 1) * 4 + ly;
}
static dim3 iso_grid(int width, int height, int cta_warps) {
  const int cw = cta_warps >= 2 ? 16 : 8, ch = cta_warps >= 4 ? 8 : 4;
  return dim3((width + cw - 1) / cw, (height + ch - 1) / ch);
}

// One flag per 8x4 warp tile: does it contain a surface pixel?  (lets the occlusion pass skip empty regions.)  Every
// warp owns its flag, so no CTA barrier and no atomics are needed.  All 32 lanes must call this.
__device__ __forceinline__ void set_tile_flag(unsigned char *tile_hit, int width, int height, bool hit, bool centre = false) {
  const unsigned any = __any_sync(0xffffffffu, hit);
  if (!tile_hit || (threadIdx.x & 31) != 0) return;
  const int warp = threadIdx.x >> 5;
  unsigned cx, cy;
  tile_of_cta(centre, cx, cy);
  const int tx = cx * (blockDim.x >= 64 ? 2 : 1) + (warp & 1), ty = cy * (blockDim.x >= 128 ? 2 : 1) + (warp >> 1);
  const int tiles_x = (width + 7) / 8, tiles_y = (height + 3) / 4;
  if (tx < tiles_x && ty < tiles_y) tile_hit[ty * tiles_x + tx] = (unsigned char)(any != 0);
}

constexpr int K_NONE = 0x7fffffff;  // no such sample

// One step of the search for the first crossing among samples [k0, kend) of a ray: cross the cells that cannot hold
// one (hierarchical skipping, exact), then classify the next BATCH samples by their bricks and fetch the survivors
// together (independent fetches in flight instead of one fetch -> compare -> branch round trip per sample; at most
// BATCH-1 samples behind the crossing are fetched in vain).  Returns true while the range is not exhausted and nothing
// was found; `found` receives the crossing sample.
template <int FMT, bool LINEAR, bool SKIP>
__device__ __forceinline__ bool iso_search_step(const IsoArgs &a, const IsoRay &q, const IsoDda &d, const IsoLevel &Ltop,
                                                const IsoLevel &Lco, bool isGreater, int &k0, int kend, int &found,
                                                unsigned &nfetch, unsigned &nskip) {
  constexpr int BATCH = 8;
  const Volume &V = a.vol;
  const float INF = __int_as_float(0x7f800000);
  if (SKIP) {
    const int skip_cls = isGreater ? 1 : 2;  // cells of this class hold no crossing
    while (k0 < kend) {
      const int kn = iso_skip_step(a, q, d, Ltop, Lco, k0, kend, skip_cls);
      if (kn == k0) break;
      k0 = kn;
      ++nskip;
    }
    if (k0 >= kend) return false;
  }
  float v[BATCH];
  bool need[BATCH];
#pragma unroll
  for (int j = 0; j < BATCH; ++j) {
    const float t = (float)min(k0 + j, kend - 1);
    need[j] = SKIP ? iso_cell_may_hold<false>(a, q, t, !isGreater) : true;
  }
#pragma unroll
  for (int j = 0; j < BATCH; ++j) {
    v[j] = need[j] ? iso_at<FMT, LINEAR>(V, q, (float)min(k0 + j, kend - 1)) : (isGreater ? INF : -INF);
    nfetch += need[j];
  }
  bool hit = false;
#pragma unroll
  for (int j = BATCH - 1; j >= 0; --j)
    if (k0 + j < kend && ((v[j] > a.iso_val) != isGreater)) {
      found = k0 + j;
      hit = true;
    }
  if (hit) return false;
  k0 += BATCH;
  return k0 < kend;
}

// One lane per ray (a.segments == 1, the default) or per ray segment: a CTA of 4 warps covers 4 / SEG tiles of 8x4
// pixels, the samples [1, max_steps) of every ray are cut into SEG equal segments searched by SEG different warps, and
// the minimum over the segments' first crossings is the ray's first crossing.  Segments were meant to shorten the
// launch's longest warps (rays that graze blobs for most of their length: 230 k cycles against a mean of 13 k, see
// spv_last_stats); measured, the SEG ray setups per ray and the barrier cost more than the shorter searches save
// (97 / 110 / 178 us for 1 / 2 / 4 segments on configs[2]), so it stays a tuning knob.
template <int FMT, bool LINEAR, bool SKIP>
__global__ void __launch_bounds__(128, SPV_ISO_MINB) iso_fast_kernel(const IsoArgs a) {
  __shared__ int s_found[4][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int SEG = a.segments;                    // 1, 2 or 4
  const int seg = warp % SEG, tile = warp / SEG, tiles = 4 / SEG;
  // tile geometry of the CTA: 4 tiles = 2x2 (16x8 pixels), 2 = 2x1, 1 = 1x1
  const int lx = (lane & 1) | ((lane >> 1) & 2) | ((lane >> 2) & 4);
  const int ly = ((lane >> 1) & 1) | ((lane >> 2) & 2);
  unsigned cx, cy;
  tile_of_cta(a.centre_out != 0, cx, cy);
  const unsigned tx = cx * (tiles >= 2 ? 2 : 1) + (tile & 1), ty = cy * (tiles >= 4 ? 2 : 1) + (tile >> 1);
  const unsigned x = tx * 8 + lx, y = ty * 4 + ly;
  const unsigned Nx = a.width, Ny = a.height;
  const bool inb = x < Nx && y < Ny;
  const size_t p = x + (size_t)Nx * y;
  const Volume &V = a.vol;
  const float INF = __int_as_float(0x7f800000);
  const int maxSteps = a.max_steps;
  unsigned nfetch = 0;
  const long long t_begin = a.stats ? clock64() : 0;
  unsigned dbg_iters = 0, dbg_walk = 0;  // statistics: search iterations of this warp, skip steps it waited for

  const IsoRay q = iso_ray(a, x, y, inb);
  const IsoDda d = iso_dda(a, q);
  const IsoLevel Ltop = {a.top, a.tgx, a.tgy, a.tgz}, Lco = {a.coarse, a.cgx, a.cgy, a.cgz};
  const bool isGreater = q.hit && iso_at<FMT, LINEAR>(V, q, 0.f) > a.iso_val;
  const int seg_len = (maxSteps - 1 + SEG - 1) / SEG;
  int i = K_NONE, k0 = 1 + seg * seg_len;
  const int kend = min(k0 + seg_len, maxSteps);
  bool active = q.hit && k0 < kend;
  for (;;) {
    unsigned nskip = 0;
    if (active) active = iso_search_step<FMT, LINEAR, SKIP>(a, q, d, Ltop, Lco, isGreater, k0, kend, i, nfetch, nskip);
    const unsigned act = __ballot_sync(0xffffffffu, active);
    if (a.stats) {
      ++dbg_iters;
      dbg_walk += __reduce_max_sync(0xffffffffu, nskip);
    }
    if (act == 0u) break;
  }
  if (SEG > 1) {
    s_found[warp][lane] = i;
    __syncthreads();
    if (seg == 0)
      for (int s2 = 1; s2 < SEG; ++s2) i = min(i, s_found[warp + s2][lane]);
  }
  if (seg == 0) {
    const bool hitIso = i != K_NONE;
    float t_hit = INF;
    v4 normal = mk4(0.f, 0.f, 0.f, 0.f);
    float colVal = 0.f;
    if (hitIso) {
      iso_resolve<FMT, LINEAR>(a, q, i, isGreater, t_hit, normal, colVal);
      nfetch += 22;
    }
    if (inb) {
      a.out[p] = hitIso ? colVal : 0.f;
      a.alpha[p] = hitIso ? q.tnear : 0.f;
      a.depth[p] = hitIso ? t_hit : INF;
      a.normals[3 * p + 0] = normal.x;
      a.normals[3 * p + 1] = normal.y;
      a.normals[3 * p + 2] = normal.z;
    }
    const unsigned any = __any_sync(0xffffffffu, hitIso);
    if (lane == 0 && a.tile_hit) {
      const int tiles_x = (a.width + 7) / 8, tiles_y = (a.height + 3) / 4;
      if ((int)tx < tiles_x && (int)ty < tiles_y) a.tile_hit[ty * tiles_x + tx] = (unsigned char)(any != 0);
    }
  }
  if (a.stats) {
    const unsigned long long dur = (unsigned long long)(clock64() - t_begin);
    unsigned nh = (q.hit && seg == 0) ? 1u : 0u;
    for (int o = 16; o > 0; o >>= 1) {
      nh += __shfl_down_sync(0xffffffffu, nh, o);
      nfetch += __shfl_down_sync(0xffffffffu, nfetch, o);
    }
    if (lane == 0) {
      atomicAdd(a.stats + 0, (unsigned long long)nh);
      atomicAdd(a.stats + 1, (unsigned long long)nfetch);
      // how long this warp lived (the longest one bounds the launch from below), with what it did:
      // cycles << 32 | search iterations << 16 | skip steps the warp executed in lockstep
      atomicMax(a.stats + 2, (dur << 32) | ((unsigned long long)min(dbg_iters, 65535u) << 16) | min(dbg_walk, 65535u));
      atomicAdd(a.stats + 3, dur);
      atomicAdd(a.stats + 4 + min(63 - __clzll(dur | 1ull), 35), 1ull);  // histogram of log2(cycles)
    }
  }
}

// -------------------------------------------------------------------------------------------------------------
// Sort-last iso surface: every rank holds one z-slab (+ halo) of the volume and marches the SAME rays.
//   iso_slab_search   per pixel, over the samples this slab owns: k1 = first k with s_k > iso, k0 = first k with
//                     s_k <= iso (INT_MAX if none).  After an element-wise MIN over the ranks, s_0 > iso <=> k1 == 0
//                     and the first crossing is i = (k1 == 0) ? k0 : k1 -- exactly the sample the single-GPU search
//                     stops at.  Samples whose brick bounds decide the comparison are classified without a fetch.
//   iso_slab_resolve  the rank owning sample i refines the bracket, takes the gradient and shades (it needs
//                     2 h N_z + one ray step of halo slices; checked per pixel, *err is raised if the resident
//                     slices do not cover the taps); every other rank writes zeros, so that an element-wise SUM over
//                     the ranks assembles the planes bit for bit.
//   iso_slab_fix      after the SUM: depth = INFINITY on pixels without a crossing, tile flags for the occlusion pass

template <int FMT, bool LINEAR, bool SKIP>
__global__ void __launch_bounds__(128) iso_slab_search_kernel(const IsoArgs a, int *__restrict__ k1_plane,
                                                              int *__restrict__ k0_plane, const IsoPeer peer) {
  constexpr int BATCH = 8;
  unsigned x, y;
  tile_pixel(x, y);
  const bool inb = x < (unsigned)a.width && y < (unsigned)a.height;
  const Volume &V = a.vol;
  const IsoRay q = iso_ray(a, x, y, inb);
  int k1 = K_NONE, k0 = K_NONE;
  unsigned nfetch = 0;
  if (q.hit) {
    const float isoVal = a.iso_val;
    int ka, kb;
    owned_interval_w(V, q.w0, q.dw, a.max_steps, ka, kb);
    const IsoDda d = iso_dda(a, q);
    const IsoLevel Ltop = {a.top, a.tgx, a.tgy, a.tgz}, Lco = {a.coarse, a.cgx, a.cgy, a.cgz};
    int kk = ka;
    while (kk < kb && (k1 == K_NONE || k0 == K_NONE)) {
      if (SKIP) {
        // a decided cell classifies all of its samples at once: only the first one can lower k1 / k0
        int kexit, c = iso_cell_class<TOP_SHIFT>(a, q, d, Ltop, kk, kb, kexit);
        if (!c) c = iso_cell_class<COARSE_SHIFT>(a, q, d, Lco, kk, kb, kexit);
        if (c) {
          if (c == 1) k1 = min(k1, kk); else k0 = min(k0, kk);
          kk = kexit;
          continue;
        }
      }
      float v[BATCH];
      int cls[BATCH];  // 0: fetch, 1: certainly > iso, 2: certainly <= iso, 3: not a sample
#pragma unroll
      for (int j = 0; j < BATCH; ++j) {
        const int k = kk + j;
        cls[j] = k < kb ? 0 : 3;
        if (SKIP && k < kb) cls[j] = iso_brick_class(a, q, (float)k);
      }
#pragma unroll
      for (int j = 0; j < BATCH; ++j) {
        v[j] = cls[j] == 0 ? iso_at<FMT, LINEAR>(V, q, (float)(kk + j)) : 0.f;
        nfetch += cls[j] == 0;
      }
#pragma unroll
      for (int j = BATCH - 1; j >= 0; --j) {
        if (cls[j] == 3) continue;
        const bool greater = cls[j] == 0 ? (v[j] > isoVal) : (cls[j] == 1);
        if (greater) k1 = min(k1, kk + j); else k0 = min(k0, kk + j);
      }
      kk += BATCH;
    }
  }
  if (inb) {
    if (peer.world > 0) {  // straight into the staging of the band's owner (peer memory over NVLink)
      const unsigned o = y / (unsigned)peer.band_rows;
      const size_t idx = (size_t)(y - o * (unsigned)peer.band_rows) * a.width + x;
      int *dst = peer.kpart[o] + (size_t)peer.src * 2 * peer.band;
      dst[idx] = k1;
      dst[peer.band + idx] = k0;
    } else {
      const size_t p = x + (size_t)a.width * y;
      k1_plane[p] = k1;
      k0_plane[p] = k0;
    }
  }
  if (a.stats) {
    atomicAdd(a.stats + 0, q.hit ? 1ull : 0ull);
    atomicAdd(a.stats + 1, (unsigned long long)nfetch);
  }
}

template <int FMT, bool LINEAR>
__global__ void __launch_bounds__(128) iso_slab_resolve_kernel(const IsoArgs a, const int *__restrict__ k1_plane,
                                                               const int *__restrict__ k0_plane, float *__restrict__ occ,
                                                               unsigned *err, const IsoPeer peer) {
  unsigned x, y;
  tile_pixel(x, y);
  const bool inb = x < (unsigned)a.width && y < (unsigned)a.height;
  const size_t p = inb ? x + (size_t)a.width * y : 0;
  const Volume &V = a.vol;
  const int k1 = inb ? k1_plane[p] : K_NONE, k0 = inb ? k0_plane[p] : K_NONE;
  const bool isGreater = k1 == 0;
  const int i = isGreater ? k0 : k1;
  float colVal = 0.f, t_hit = 0.f, tn = 0.f;
  v4 normal = mk4(0.f, 0.f, 0.f, 0.f);
  bool mine = false;  // this slab owns the crossing sample
  if (i != K_NONE) {
    const IsoRay q = iso_ray(a, x, y, true);
    const float s = slice_of_k(V, q.w0, q.dw, i);
    if (s >= (float)V.z0 && s < (float)V.z1) {  // this slab owns the crossing sample
      // slices the refinement samples (t in [i-1, i)) and the gradient taps (+-2h along z around them) touch
      const float wa = fmaf((float)(i - 1), q.dw, q.w0), wb = fmaf((float)i, q.dw, q.w0);
      const float h2w = 2.f * fabsf(q.dt * (a.gamma * a.gamma)) * V.fnz;
      const float wmin = fminf(wa, wb) - h2w - 0.5f, wmax = fmaxf(wa, wb) + h2w - 0.5f;
      const float lo_need = fmaxf(floorf(wmin) - 1.f, 0.f), hi_need = fminf(floorf(wmax) + 2.f, (float)(V.nz - 1));
      if (lo_need < (float)V.z_lo || hi_need > (float)(V.z_lo + V.local_nz - 1)) atomicExch(err, 1u);
      iso_resolve<FMT, LINEAR>(a, q, i, isGreater, t_hit, normal, colVal);
      tn = q.tnear;
      mine = true;
    }
  }
  if (peer.world > 0) {
    // exactly one rank owns a crossing: it stores the finished pixel into every rank's planes (its own included);
    // a pixel without a crossing is cleared by every rank for itself; nobody else touches a pixel
    const size_t n = (size_t)a.width * a.height;
    if (inb && (mine || i == K_NONE)) {
      if (i == K_NONE) t_hit = __int_as_float(0x7f800000);
      for (int r = 0; r < (mine ? peer.world : 1); ++r) {
        float *P = mine ? peer.planes[r] : peer.planes[peer.src];
        P[p] = colVal;
        P[n + p] = tn;
        P[2 * n + p] = t_hit;
        float *N3 = P + (size_t)peer.normals_plane * n + 3 * p;
        N3[0] = normal.x;
        N3[1] = normal.y;
        N3[2] = normal.z;
      }
    }
    set_tile_flag(a.tile_hit, a.width, a.height, inb && i != K_NONE);
    return;
  }
  if (!inb) return;
  a.out[p] = colVal;
  a.alpha[p] = tn;
  a.depth[p] = t_hit;
  occ[p] = 0.f;
  a.normals[3 * p + 0] = normal.x;
  a.normals[3 * p + 1] = normal.y;
  a.normals[3 * p + 2] = normal.z;
}

__global__ void __launch_bounds__(128) iso_slab_fix_kernel(int width, int height, const int *__restrict__ k1_plane,
                                                           const int *__restrict__ k0_plane, float *__restrict__ depth,
                                                           unsigned char *tile_hit) {
  unsigned x, y;
  tile_pixel(x, y);
  const bool inb = x < (unsigned)width && y < (unsigned)height;
  bool hit = false;
  if (inb) {
    const size_t p = x + (size_t)width * y;
    const int k1 = k1_plane[p], k0 = k0_plane[p];
    hit = (k1 == 0 ? k0 : k1) != K_NONE;
    if (!hit) depth[p] = __int_as_float(0x7f800000);
  }
  set_tile_flag(tile_hit, width, height, hit);
}

template <int FMT>
static cudaError_t launch_iso_slab_dt(const IsoArgs &a, bool linear, int phase, int *k1, int *k0, float *occ,
                                      unsigned *err, cudaStream_t st, const IsoPeer &peer) {
  dim3 grid((a.width + 15) / 16, (a.height + 7) / 8), block(128);
  if (phase == 0) {
    if (a.skip) {
      if (linear) iso_slab_search_kernel<FMT, true, true><<<grid, block, 0, st>>>(a, k1, k0, peer);
      else iso_slab_search_kernel<FMT, false, true><<<grid, block, 0, st>>>(a, k1, k0, peer);
    } else {
      if (linear) iso_slab_search_kernel<FMT, true, false><<<grid, block, 0, st>>>(a, k1, k0, peer);
      else iso_slab_search_kernel<FMT, false, false><<<grid, block, 0, st>>>(a, k1, k0, peer);
    }
  } else if (phase == 1) {
    if (linear) iso_slab_resolve_kernel<FMT, true><<<grid, block, 0, st>>>(a, k1, k0, occ, err, peer);
    else iso_slab_resolve_kernel<FMT, false><<<grid, block, 0, st>>>(a, k1, k0, occ, err, peer);
  } else {
    iso_slab_fix_kernel<<<grid, block, 0, st>>>(a.width, a.height, k1, k0, a.depth, a.tile_hit);
  }
  return cudaGetLastError();
}

cudaError_t launch_iso_slab(const IsoArgs &a, int dtype, bool linear, int phase, int *k1, int *k0, float *occ,
                            unsigned *err, cudaStream_t st, const IsoPeer *peer_in) {
  IsoPeer peer;
  if (peer_in) peer = *peer_in; else memset(&peer, 0, sizeof peer);
  switch (dtype) {  // FMT = dtype + 3 * layout
    case 0: return launch_iso_slab_dt<0>(a, linear, phase, k1, k0, occ, err, st, peer);
    case 1: return launch_iso_slab_dt<1>(a, linear, phase, k1, k0, occ, err, st, peer);
    case 2: return launch_iso_slab_dt<2>(a, linear, phase, k1, k0, occ, err, st, peer);
    case 4: return launch_iso_slab_dt<4>(a, linear, phase, k1, k0, occ, err, st, peer);
    case 5: return launch_iso_slab_dt<5>(a, linear, phase, k1, k0, occ, err, st, peer);
    default: return cudaErrorInvalidValue;
  }
}

template <int FMT>
static cudaError_t launch_iso_dt(const IsoArgs &a, bool linear, bool exact, bool stats, cudaStream_t st) {
  dim3 grid((a.width + 15) / 16, (a.height + 7) / 8), block(128);
  if (!exact) {
    // a.segments in {1, 2, 4}: the CTA's 4 warps cover 4, 2 or 1 tiles
    const int tiles = a.segments == 4 ? 1 : (a.segments == 2 ? 2 : 4);
    const dim3 fgrid = iso_grid(a.width, a.height, tiles), fblock(128);
    if (fgrid.y > 65535u) return cudaErrorInvalidValue;
    if (a.skip) {
      if (linear) iso_fast_kernel<FMT, true, true><<<fgrid, fblock, 0, st>>>(a);
      else iso_fast_kernel<FMT, false, true><<<fgrid, fblock, 0, st>>>(a);
    } else {
      if (linear) iso_fast_kernel<FMT, true, false><<<fgrid, fblock, 0, st>>>(a);
      else iso_fast_kernel<FMT, false, false><<<fgrid, fblock, 0, st>>>(a);
    }
    return cudaGetLastError();
  }
#define SPV_ISO(L, E, S) iso_kernel<FMT, L, E, S><<<grid, block, 0, st>>>(a)
  if (stats) {
    if (linear) { if (exact) SPV_ISO(true, true, true); else SPV_ISO(true, false, true); }
    else { if (exact) SPV_ISO(false, true, true); else SPV_ISO(false, false, true); }
  } else {
    if (linear) { if (exact) SPV_ISO(true, true, false); else SPV_ISO(true, false, false); }
    else { if (exact) SPV_ISO(false, true, false); else SPV_ISO(false, false, false); }
  }
#undef SPV_ISO
  return cudaGetLastError();
}

cudaError_t launch_iso(const IsoArgs &a, int dtype, bool linear, bool exact, bool stats, cudaStream_t st) {
  switch (dtype) {  // FMT = dtype + 3 * layout
    case 0: return launch_iso_dt<0>(a, linear, exact, stats, st);
    case 1: return launch_iso_dt<1>(a, linear, exact, stats, st);
    case 2: return launch_iso_dt<2>(a, linear, exact, stats, st);
    case 4: return launch_iso_dt<4>(a, linear, exact, stats, st);
    case 5: return launch_iso_dt<5>(a, linear, exact, stats, st);
    default: return cudaErrorInvalidValue;
  }
}

// -------------------------------------------------------------------------------------------------------------
// separable blur, one output pixel per thread.  Taps ht in [h_start, h_end) at offset ht - Nh/2 (integer
// division): the weight table is centred at Nh/2.f but the taps at Nh/2, as in the reference.
template <int NCOMP, bool ALONG_Y>
__global__ void conv_kernel(const float *__restrict__ input, float *__restrict__ output, int Nx, int Ny,
                            const ConvWeights cw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= Nx || j >= Ny) return;
  const int Nh = cw.nh;
  const int c = ALONG_Y ? j : i, N = ALONG_Y ? Ny : Nx;
  const int start = c - Nh / 2;
  const int h_start = ((c - Nh / 2) < 0) ? Nh / 2 - c : 0;
  const int h_end = ((c + Nh / 2) >= N) ? Nh - (c + Nh / 2 - N + 1) : Nh;
  float res[NCOMP];
#pragma unroll
  for (int k = 0; k < NCOMP; ++k) res[k] = 0.f;
  float sum_val = 0.f;
  for (int ht = h_start; ht < h_end; ++ht) {
    const float val = cw.w[ht];
    sum_val += val;
    const size_t q = ALONG_Y ? ((size_t)i + (size_t)(start + ht) * Nx) : ((size_t)(start + ht) + (size_t)j * Nx);
#pragma unroll
    for (int k = 0; k < NCOMP; ++k) res[k] += val * input[NCOMP * q + k];
  }
  const size_t o = (size_t)i + (size_t)j * Nx;
#pragma unroll
  for (int k = 0; k < NCOMP; ++k) output[NCOMP * o + k] = res[k] / sum_val;
}

cudaError_t launch_conv(float *buf, float *tmp, int width, int height, int ncomp, const ConvWeights &w,
                        cudaStream_t st) {
  dim3 block(32, 8), grid((width + 31) / 32, (height + 7) / 8);
  if (ncomp == 1) {
    conv_kernel<1, false><<<grid, block, 0, st>>>(buf, tmp, width, height, w);
    conv_kernel<1, true><<<grid, block, 0, st>>>(tmp, buf, width, height, w);
  } else {
    conv_kernel<3, false><<<grid, block, 0, st>>>(buf, tmp, width, height, w);
    conv_kernel<3, true><<<grid, block, 0, st>>>(tmp, buf, width, height, w);
  }
  return cudaGetLastError();
}

// The screen-space passes only have work near surface pixels, which the march flags per 8x4 tile:
//   occlusion  every tap lands within `radius` pixels of its pixel; if no tile in reach holds a surface pixel, every
//              depth read is INFINITY, `depth < depth0` is false for every tap and the result is exactly 0
//   blur       a window over zeros (the normals of non-surface pixels; the occlusion out of reach) is exactly +0
// Flags of the tiles in reach of pixels [px0, px1] x [py0, py1], OR-ed over the strided share `first, first + step, ...`
// of the calling thread.
__device__ __forceinline__ int tiles_in_reach(const unsigned char *__restrict__ tile_hit, int tiles_x, int Nx, int Ny,
                                              int px0, int px1, int py0, int py1, int radius, int first, int step) {
  const int tx0 = max(px0 - radius, 0) / 8, tx1 = min(px1 + radius, Nx - 1) / 8;
  const int ty0 = max(py0 - radius, 0) / 4, ty1 = min(py1 + radius, Ny - 1) / 4;
  const int tw = tx1 - tx0 + 1, n = tw * (ty1 - ty0 + 1);
  int mine = 0;
  for (int t = first; t < n; t += step) mine |= tile_hit[(ty0 + t / tw) * tiles_x + tx0 + t % tw];
  return mine;
}

// iso_kernel.cl:505-588 for one pixel: the shaded value from the blurred normal, the depth and the blurred occlusion
__device__ __forceinline__ float shade_pixel(unsigned x, unsigned y, int Nx, int Ny, const Camera &cam, float occ_strength,
                                             const float *__restrict__ input_normals, const float *__restrict__ input_depth,
                                             float occ) {
  const size_t p = x + (size_t)Nx * y;
  const float depth = input_depth[p];
  if (!(depth < __int_as_float(0x7f800000))) return 0.f;  // colVal * 0 below: nothing to compute
  v4 orig, direc;
  eye_ray(x, y, Nx, Ny, cam, orig, direc);
  v4 light = mk4(2.f, -1.f, -2.f, 0.f);
  const float c_ambient = .5f, c_diffuse = .3f, c_specular = .2f;
  light = mult(cam.invM, light);
  light = normalize4(light);
  v4 normal = mk4(input_normals[3 * p], input_normals[1 + 3 * p], input_normals[2 + 3 * p], 0.f);
  normal = normalize4(normal);
  const v4 reflect = sub4(scl4(2.f * dot4(light, normal), normal), light);
  const float diffuse = fmaxf(0.f, dot4(light, normal));
  const float specular = powf(fmaxf(0.f, dot4(normalize4(reflect), normalize4(direc))), 10.f);
  float colVal = c_ambient + c_diffuse * diffuse + (diffuse > 0.f ? 1.f : 0.f) * c_specular * specular;
  colVal = (1.f - occ_strength) * colVal + occ_strength * colVal * (1.f - occ);
  // iso_kernel.cl:580 has a double literal: the sum is formed in double and rounded once
  colVal = (float)((double)((1.f - occ_strength) * colVal) + 1.0 * (double)occ_strength * (double)colVal);
  colVal *= ((depth < __int_as_float(0x7f800000)) ? 1.f : 0.f);
  return colVal;
}

// Both passes in one launch: the x pass of the rows a 32x16 output tile needs (Nh-1 halo rows) goes to shared memory,
// the y pass reads it from there.  Every output is the same sequence of fp32 operations as conv_kernel<.,false>
// followed by conv_kernel<.,true>, so the result is bit-identical to the two-launch form; it needs `in` != `out`.
// A thread owns one FLOAT of the interleaved row (component k of pixel i sits at 3i+k; its x taps are 3 floats
// apart), so every global and shared access of a warp is unit-stride.
// NH > 0: the tap count is a compile-time constant and pixels whose taps all lie inside the image (nearly all) run an
// unrolled loop with the weights in registers; NH == 0: any tap count up to 32.
// tile_hit != nullptr: the input is known to be +0 on every pixel further than `reach` pixels from a flagged 8x4 tile;
// a CTA whose taps see only such pixels stores zeros (what the arithmetic would give) and is done.
// SHADE (NCOMP == 1: the blur of the occlusion plane): the shading pass rides on the epilogue -- the pixel's blurred
// occlusion goes to its plane and, with the blurred normal and the depth, straight into shade_pixel (one launch and one
// read of the occlusion plane less; the same expressions on the same values as shading_kernel).
constexpr int CONV_TH = 16;
template <int NCOMP, int NH, bool SHADE>
__global__ void __launch_bounds__(256) conv_xy_kernel(const float *__restrict__ input, float *__restrict__ output, int Nx,
                                                      int Ny, const ConvWeights cw,
                                                      const unsigned char *__restrict__ tile_hit, int tiles_x, int reach,
                                                      int y_first, int y_end, const ShadeArgs sh) {
  // rows [y_first, y_end) are written (the whole image, or one rank's band of a sort-last frame); taps read any row
  extern __shared__ float s_row[];  // [CONV_TH + Nh - 1][32]
  const int Nh = NH > 0 ? NH : cw.nh, R = Nh / 2;
  const int c = blockIdx.x * 32 + threadIdx.x;  // column of the interleaved image (NCOMP * Nx floats per row)
  const int i = c / NCOMP, y0 = y_first + blockIdx.y * CONV_TH;
  const int rows = CONV_TH + Nh - 1;
  const int pitch = NCOMP * Nx;
  if (tile_hit) {
    const int px0 = (int)(blockIdx.x * 32) / NCOMP, px1 = min((int)(blockIdx.x * 32 + 31) / NCOMP, Nx - 1);
    const int live = __syncthreads_or(tiles_in_reach(tile_hit, tiles_x, Nx, Ny, px0, px1, y0, min(y0 + CONV_TH, y_end) - 1,
                                                     reach + Nh, threadIdx.y * 32 + threadIdx.x, 256));
    if (!live) {
      if (c < pitch)
        for (int r = threadIdx.y; r < CONV_TH && y0 + r < y_end; r += blockDim.y) {
          output[(size_t)(y0 + r) * pitch + c] = 0.f;
          if (SHADE) sh.out[(size_t)(y0 + r) * pitch + c] = 0.f;  // no surface tile in reach: no surface pixel here
        }
      return;
    }
  }
  float w[NH > 0 ? NH : 1];
  float sum_full = 0.f;  // the weight sum of a pixel with all its taps, added up in tap order
  if (NH > 0) {
#pragma unroll
    for (int ht = 0; ht < NH; ++ht) {
      w[ht] = cw.w[ht];
      sum_full += w[ht];
    }
  }
  if (c < pitch) {
    const int h_start = ((i - R) < 0) ? R - i : 0;
    const int h_end = ((i + R) >= Nx) ? Nh - (i + R - Nx + 1) : Nh;
    const bool inner = NH > 0 && h_start == 0 && h_end == Nh;
    for (int r = threadIdx.y; r < rows; r += blockDim.y) {
      const int j = y0 - R + r;
      if (j < 0 || j >= Ny) continue;
      const float *row = input + (size_t)j * pitch + (c - NCOMP * R);  // tap ht at row[NCOMP * ht]
      float res = 0.f, sum_val = 0.f;
      if (inner) {
#pragma unroll
        for (int ht = 0; ht < (NH > 0 ? NH : 1); ++ht) res += w[ht] * row[NCOMP * ht];
        sum_val = sum_full;
      } else {
        for (int ht = h_start; ht < h_end; ++ht) {
          const float val = cw.w[ht];
          sum_val += val;
          res += val * row[NCOMP * ht];
        }
      }
      s_row[r * 32 + threadIdx.x] = res / sum_val;
    }
  }
  __syncthreads();
  if (c >= pitch) return;
  for (int r = threadIdx.y; r < CONV_TH; r += blockDim.y) {
    const int j = y0 + r;
    if (j >= y_end) break;
    const int h_start = ((j - R) < 0) ? R - j : 0;
    const int h_end = ((j + R) >= Ny) ? Nh - (j + R - Ny + 1) : Nh;
    const float *col = s_row + r * 32 + threadIdx.x;  // tap ht at col[32 * ht]
    float res = 0.f, sum_val = 0.f;
    if (NH > 0 && h_start == 0 && h_end == Nh) {
#pragma unroll
      for (int ht = 0; ht < (NH > 0 ? NH : 1); ++ht) res += w[ht] * col[32 * ht];
      sum_val = sum_full;
    } else {
      for (int ht = h_start; ht < h_end; ++ht) {
        const float val = cw.w[ht];
        sum_val += val;
        res += val * col[32 * ht];
      }
    }
    const float v = res / sum_val;
    output[(size_t)j * pitch + c] = v;
    if (SHADE) sh.out[(size_t)j * pitch + c] = shade_pixel((unsigned)c, (unsigned)j, Nx, Ny, sh.cam, sh.occ_strength, sh.normals, sh.depth, v);
  }
}

cudaError_t launch_conv_xy(const float *in, float *out, int width, int height, int ncomp, const ConvWeights &w,
                           const unsigned char *tile_hit, int reach, cudaStream_t st, int y_first, int y_end,
                           const ShadeArgs *shade) {
  if (in == out || w.nh < 1 || w.nh > 32 || (ncomp != 1 && ncomp != 3)) return cudaErrorInvalidValue;
  if (y_end < 0 || y_end > height) y_end = height;
  if (y_first < 0) y_first = 0;
  if (y_first >= y_end) return cudaSuccess;
  dim3 block(32, 8), grid((ncomp * width + 31) / 32, (y_end - y_first + CONV_TH - 1) / CONV_TH);
  if (grid.y > 65535u) return cudaErrorInvalidValue;
  const size_t smem = (size_t)(CONV_TH + w.nh - 1) * 32 * sizeof(float);
  const int tx = (width + 7) / 8;
  ShadeArgs sh;
  memset(&sh, 0, sizeof sh);
  if (shade) {
    if (ncomp != 1 || !shade->out || shade->out == out || shade->out == in) return cudaErrorInvalidValue;
    sh = *shade;
  }
#define SPV_CONV(C, N, S) conv_xy_kernel<C, N, S><<<grid, block, smem, st>>>(in, out, width, height, w, tile_hit, tx, reach, y_first, y_end, sh)
  if (ncomp == 3 && w.nh == 7) SPV_CONV(3, 7, false);
  else if (ncomp == 1 && w.nh == 5 && shade) SPV_CONV(1, 5, true);
  else if (ncomp == 1 && w.nh == 5) SPV_CONV(1, 5, false);
  else if (ncomp == 1 && shade) SPV_CONV(1, 0, true);
  else if (ncomp == 1) SPV_CONV(1, 0, false);
  else SPV_CONV(3, 0, false);
#undef SPV_CONV
  return cudaGetLastError();
}

// -------------------------------------------------------------------------------------------------------------
// utils.cl:10-38, uint32 wrap-around arithmetic
__device__ __forceinline__ uint32_t lcg_hash(uint32_t x, uint32_t y) {
  uint32_t a = 4421u + (1u + x) * (1u + y) + x + y;
#pragma unroll
  for (int i = 0; i < 10; i++) a = (1664525u * a + 1013904223u) % 79197919u;
  return a;
}
__device__ __forceinline__ float random_cl(uint32_t x, uint32_t y) {
  return ((float)lcg_hash(x, y) * 1.0f) / (float)(79197919);
}
__device__ __forceinline__ float rand_int_cl(uint32_t x, uint32_t y, int start, int end) {
  return (float)(int)((float)start + random_cl(x, y) * (float)(end - start));
}

// The four rand_int() arguments of a tap depend on the tap index only (occlusion.cl:60,62): they are evaluated once
// into a table (occ_taps_kernel, cached by the context), which halves the hash evaluations per pixel (2 instead of 4
// per tap, 10 LCG rounds each).  Values are identical to evaluating them per pixel.
__global__ void occ_taps_kernel(float4 *taps, int n) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (unsigned)n) return;
  taps[i] = make_float4(rand_int_cl(i, i * i, 0, 1000), rand_int_cl(i * i, i, 294, 97701), rand_int_cl(i * i, i, 0, 1997),
                        rand_int_cl(i, i * i, 569, 17633));
}
cudaError_t launch_occ_taps(float4 *taps, int n, cudaStream_t st) {
  occ_taps_kernel<<<(n + 127) / 128, 128, 0, st>>>(taps, n);
  return cudaGetLastError();
}

// occlusion.cl:54-77 for one pixel
__device__ __forceinline__ float occlusion_pixel(int x, int y, int Nx, int Ny, int radius, int number_points,
                                                 const float *__restrict__ input_depth, const float4 *__restrict__ taps) {
  const float MPI_2 = 6.2831853071795f;
  const float depth0 = input_depth[x + (size_t)y * Nx];
  float occ = 0.f;
  for (unsigned i = 0; i < (unsigned)number_points; ++i) {
    const float4 tp = __ldg(taps + i);
    const float r = (float)(unsigned)radius * random_cl((uint32_t)((float)x + tp.x), (uint32_t)((float)y + tp.y));
    const float phi = MPI_2 * random_cl((uint32_t)((float)x + tp.z), (uint32_t)((float)y + tp.w));
    const int x2 = clampi((int)((float)x + r * cosf(phi)), 0, Nx - 1);
    const int y2 = clampi((int)((float)y + r * sinf(phi)), 0, Ny - 1);
    const float depth = input_depth[x2 + (size_t)y2 * Nx];
    occ += (depth < depth0 ? 1.f : 0.f);
  }
  return occ / (float)(unsigned)number_points;
}

// Plain form, one CTA per block of 32x2 pixels (used when there are no tile flags: the exact-sampler path).
__global__ void __launch_bounds__(64) occlusion_kernel(float *__restrict__ d_output, int Nx, int Ny, int radius,
                                                       int number_points, const float *__restrict__ input_depth,
                                                       const unsigned char *__restrict__ tile_hit, int tiles_x,
                                                       const float4 *__restrict__ taps) {
  int reach = 1;
  if (tile_hit) {
    const int bx = blockIdx.x * blockDim.x, by = blockIdx.y * blockDim.y;
    reach = __syncthreads_or(tiles_in_reach(tile_hit, tiles_x, Nx, Ny, bx, bx + blockDim.x - 1, by, by + blockDim.y - 1,
                                            radius, threadIdx.y * blockDim.x + threadIdx.x, blockDim.x * blockDim.y));
  }
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= Nx || y >= Ny) return;
  d_output[x + (size_t)Nx * y] = reach ? occlusion_pixel(x, y, Nx, Ny, radius, number_points, input_depth, taps) : 0.f;
}

// Queue form.  Only the blocks near a surface do any work (60 hash chains per pixel) and they are few and clustered:
// handed out as CTAs they land unevenly on the SMs (measured: SMs busy 62 % of the launch).  Instead
//   occ_list_kernel   one warp per 32x8 pixels: out of reach -> writes the zeros, in reach -> appends its 32x2 blocks to a list
//   occ_queue_kernel  a fixed grid (a few CTAs per SM) pulls work items off the list until it is empty
// The counters {count, head} are double-buffered by frame parity; the list kernel clears the other pair for the next
// frame, so there is no memset on the stream.
constexpr int OCC_BW = 32, OCC_BH = 2, OCC_GROUP = 4;  // the list kernel classifies OCC_GROUP stacked blocks at once
__global__ void __launch_bounds__(256) occ_list_kernel(float *__restrict__ d_output, int Nx, int Ny, int radius,
                                                       const unsigned char *__restrict__ tile_hit, int tiles_x, int obx,
                                                       int oby, unsigned *__restrict__ cnt, unsigned *__restrict__ cnt_next,
                                                       unsigned *__restrict__ list, int y_first, int y_end) {
  // one warp per group of OCC_GROUP vertically adjacent blocks (32 x 8 pixels): their reach rectangles differ by a
  // few rows only, so one scan of the tile flags over the union decides all of them (conservatively: a block that
  // is computed although nothing is in reach still gets the right answer, 0)
  const int g = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (g == 0 && lane == 0) cnt_next[0] = cnt_next[1] = 0u;
  const int gy = (oby + OCC_GROUP - 1) / OCC_GROUP;
  if (g >= obx * gy) return;
  const int bx = g % obx, by0 = (g / obx) * OCC_GROUP, nb = min(OCC_GROUP, oby - by0);
  const int px = bx * OCC_BW, py = by0 * OCC_BH;
  if (py + nb * OCC_BH <= y_first || py >= y_end) return;  // not this rank's rows (sort-last frames): left untouched
  const int mine = tiles_in_reach(tile_hit, tiles_x, Nx, Ny, px, px + OCC_BW - 1, py, py + nb * OCC_BH - 1, radius, lane, 32);
  if (__any_sync(0xffffffffu, mine != 0)) {
    if (lane < nb) list[atomicAdd(cnt, 1u)] = (unsigned)((by0 + lane) * obx + bx);
  } else {
    const int x = px + lane;
    if (x < Nx)
      for (int r = 0; r < nb * OCC_BH; ++r)
        if (py + r < Ny) d_output[x + (size_t)Nx * (py + r)] = 0.f;
  }
}

// A work item is 8 pixels of a row; a warp gives 4 lanes to each pixel, lane g takes taps g, g+4, ... and the four
// partial counts are added with shuffles (sums of 0/1 floats: exact in any order, so the result is unchanged).  Small
// items (a quarter of a warp-row's 30 taps) keep the SMs evenly loaded up to the last ~2 us.
__global__ void __launch_bounds__(128) occ_queue_kernel(float *__restrict__ d_output, int Nx, int Ny, int radius,
                                                        int number_points, const float *__restrict__ input_depth,
                                                        const float4 *__restrict__ taps, int obx,
                                                        unsigned *__restrict__ cnt, const unsigned *__restrict__ list) {
  __shared__ unsigned s_base;
  constexpr unsigned ITEMS_PER_BLOCK = OCC_BW * OCC_BH / 8;
  const unsigned total = cnt[0] * ITEMS_PER_BLOCK;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane & 3;
  const float MPI_2 = 6.2831853071795f;
  for (;;) {
    if (threadIdx.x == 0) s_base = atomicAdd(cnt + 1, 4u);
    __syncthreads();
    const unsigned base = s_base;
    __syncthreads();
    if (base >= total) break;
    const unsigned item = base + warp;
    if (item >= total) continue;
    const int ob = (int)list[item / ITEMS_PER_BLOCK], sub = (int)(item % ITEMS_PER_BLOCK);
    const int x = (ob % obx) * OCC_BW + (sub % (OCC_BW / 8)) * 8 + (lane >> 2), y = (ob / obx) * OCC_BH + sub / (OCC_BW / 8);
    const bool inb = x < Nx && y < Ny;
    float occ = 0.f;
    if (inb) {  // occlusion.cl:54-77, taps g, g + 4, ...
      const float depth0 = input_depth[x + (size_t)y * Nx];
      for (unsigned i = g; i < (unsigned)number_points; i += 4) {
        const float4 tp = __ldg(taps + i);
        const float r = (float)(unsigned)radius * random_cl((uint32_t)((float)x + tp.x), (uint32_t)((float)y + tp.y));
        const float phi = MPI_2 * random_cl((uint32_t)((float)x + tp.z), (uint32_t)((float)y + tp.w));
        const int x2 = clampi((int)((float)x + r * cosf(phi)), 0, Nx - 1);
        const int y2 = clampi((int)((float)y + r * sinf(phi)), 0, Ny - 1);
        occ += (input_depth[x2 + (size_t)y2 * Nx] < depth0 ? 1.f : 0.f);
      }
    }
    occ += __shfl_xor_sync(0xffffffffu, occ, 1);
    occ += __shfl_xor_sync(0xffffffffu, occ, 2);
    if (inb && g == 0) d_output[x + (size_t)Nx * y] = occ / (float)(unsigned)number_points;
  }
}

// Table form.  Where a tap of pixel (x, y) lands -- clamp((int)(x + r cos(phi))), clamp((int)(y + r sin(phi))) with r and
// phi hashed from (x, y, tap) -- does not depend on the image: the offsets (x2 - x, y2 - y) are evaluated once per
// (image size, radius, tap count) by the very expressions of occlusion_pixel and kept as signed bytes.  A frame then
// costs one 16-byte table load per lane (its 8 taps) instead of 2 x 10 LCG rounds, cos and sin per tap (205
// instructions), and the depth gathers come out of a shared-memory tile of the 32 x 8 pixel group and its halo
// (random banks: ~3 cycles per request instead of one tag lookup per lane).  Same taps, same comparisons, sums of
// 0 / 1: the result is bit-identical to the queue form.
//   layout   table[((chunk * Npix + pixel) * 4 + g) * 8 + j] = offsets of tap 32 chunk + g + 4 j (lane g of a pixel's 4 lanes)
__global__ void __launch_bounds__(256) occ_table_kernel(char2 *__restrict__ table, int Nx, int Ny, int radius, int number_points,
                                                        const float4 *__restrict__ taps, int halo, unsigned *__restrict__ bad) {
  const size_t npix = (size_t)Nx * Ny;
  const int chunks = (number_points + 31) / 32;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= npix * 32 * chunks) return;
  const int j = (int)(idx & 7), g = (int)((idx >> 3) & 3);
  const size_t pc = idx >> 5;
  const int chunk = (int)(pc / npix);
  const size_t pix = pc % npix;
  const int x = (int)(pix % Nx), y = (int)(pix / Nx);
  const unsigned i = 32u * chunk + g + 4u * j;
  char2 o = make_char2(0, 0);
  if (i < (unsigned)number_points) {  // occlusion.cl:60-65
    const float MPI_2 = 6.2831853071795f;
    const float4 tp = __ldg(taps + i);
    const float r = (float)(unsigned)radius * random_cl((uint32_t)((float)x + tp.x), (uint32_t)((float)y + tp.y));
    const float phi = MPI_2 * random_cl((uint32_t)((float)x + tp.z), (uint32_t)((float)y + tp.w));
    const int dx = clampi((int)((float)x + r * cosf(phi)), 0, Nx - 1) - x;
    const int dy = clampi((int)((float)y + r * sinf(phi)), 0, Ny - 1) - y;
    if (dx < -halo || dx > halo || dy < -halo || dy > halo) atomicAdd(bad, 1u);  // (cannot happen: |r| <= radius)
    o = make_char2((signed char)dx, (signed char)dy);
  }
  table[idx] = o;
}

__global__ void __launch_bounds__(256) occ_tile_kernel(float *__restrict__ d_output, int Nx, int Ny, int radius, int number_points,
                                                       const float *__restrict__ input_depth, const char2 *__restrict__ table,
                                                       const unsigned char *__restrict__ tile_hit, int tiles_x, int obx) {
  // one CTA per group of 32 x 8 pixels; a group without a surface tile in reach stores its zeros and is done (as
  // occ_list_kernel decides for the queue form: no list, no counters here -- a CTA costs a few us at most)
  extern __shared__ float s_depth[];  // [8 + 2 halo][32 + 2 halo]
  const int halo = radius + 1;
  const int tw = OCC_BW + 2 * halo, th = OCC_GROUP * OCC_BH + 2 * halo;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane & 3;
  const int px = (blockIdx.x % obx) * OCC_BW, py = (blockIdx.x / obx) * OCC_GROUP * OCC_BH;
  const int reach = __syncthreads_or(tiles_in_reach(tile_hit, tiles_x, Nx, Ny, px, px + OCC_BW - 1, py,
                                                    py + OCC_GROUP * OCC_BH - 1, radius, threadIdx.x, 256));
  if (!reach) {
    const int x = px + lane, y = py + warp;
    if (x < Nx && y < Ny) d_output[x + (size_t)Nx * y] = 0.f;
    return;
  }
  const int ox = px - halo, oy = py - halo;  // image position of tile element (0, 0)
  for (int e = threadIdx.x; e < tw * th; e += 256) {
    const int tx = e % tw, ty = e / tw, ix = ox + tx, iy = oy + ty;
    if (ix >= 0 && ix < Nx && iy >= 0 && iy < Ny) s_depth[e] = input_depth[ix + (size_t)iy * Nx];  // taps are clamped into the image
  }
  __syncthreads();
  const size_t npix = (size_t)Nx * Ny;
  const int chunks = (number_points + 31) / 32;
  const int y = py + warp;
#pragma unroll
  for (int step = 0; step < OCC_BW / 8; ++step) {
    const int x = px + step * 8 + (lane >> 2);
    const bool inb = x < Nx && y < Ny;
    float occ = 0.f;
    if (inb) {
      const float *centre = s_depth + (warp + halo) * tw + (x - ox);
      const float depth0 = *centre;
      const size_t pix = x + (size_t)y * Nx;
      for (int chunk = 0; chunk < chunks; ++chunk) {
        const uint4 q = __ldg(reinterpret_cast<const uint4 *>(table + ((chunk * npix + pix) * 4 + g) * 8));
        const unsigned w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const unsigned i = 32u * chunk + g + 4u * j;
          const unsigned pair = (w[j >> 1] >> (16 * (j & 1))) & 0xffffu;
          const int dx = (int)(signed char)(pair & 0xffu), dy = (int)(signed char)(pair >> 8);
          if (i < (unsigned)number_points) occ += (centre[dy * tw + dx] < depth0 ? 1.f : 0.f);
        }
      }
    }
    occ += __shfl_xor_sync(0xffffffffu, occ, 1);
    occ += __shfl_xor_sync(0xffffffffu, occ, 2);
    if (inb && g == 0) d_output[x + (size_t)Nx * y] = occ / (float)(unsigned)number_points;
  }
}

// bytes of the tap-offset table of a width x height image (0: the table form does not apply -- radius beyond a signed
// byte or a shared-memory tile, or a table beyond 1 GiB)
size_t occ_table_bytes(int width, int height, int radius, int n_points) {
  if (radius < 0 || radius > 44 || n_points < 1) return 0;  // the group's tile and halo within 48 KB of shared memory
  const size_t bytes = (size_t)width * height * 64 * (size_t)((n_points + 31) / 32);
  return bytes <= ((size_t)1 << 30) ? bytes : 0;
}
static size_t occ_tile_smem(int radius) {
  const int halo = radius + 1;
  return (size_t)(OCC_BW + 2 * halo) * (OCC_GROUP * OCC_BH + 2 * halo) * sizeof(float);
}
// fills `table` (occ_table_bytes) for this image size, radius and tap count; *bad (device, zeroed by the caller) counts
// offsets beyond the halo (none: the caller falls back to the queue form if it is not 0)
cudaError_t launch_occ_table(void *table, int width, int height, int radius, int n_points, const float4 *taps, unsigned *bad,
                             cudaStream_t st) {
  const size_t n = (size_t)width * height * 32 * (size_t)((n_points + 31) / 32);
  if ((n + 255) / 256 > 0x7fffffffull) return cudaErrorInvalidValue;
  occ_table_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(reinterpret_cast<char2 *>(table), width, height, radius, n_points,
                                                             taps, radius + 1, bad);
  return cudaGetLastError();
}

size_t occ_queue_bytes(int width, int height) {
  const size_t n_ob = (size_t)((width + OCC_BW - 1) / OCC_BW) * ((height + OCC_BH - 1) / OCC_BH);
  return (4 + n_ob) * sizeof(unsigned);  // [2 parities][count, head] + the list
}

int occ_ctas_per_sm = 10;  // resident CTAs (4 warps each) per SM of the occlusion queue kernel (tuning knob 6)
cudaError_t launch_occlusion(float *occ, int width, int height, int radius, int n_points, const float *depth,
                             const unsigned char *tile_hit, const float4 *taps, unsigned *queue, unsigned frame, int sms,
                             cudaStream_t st, int y_first, int y_end, const void *table) {
  if (y_end < 0 || y_end > height) y_end = height;
  if (y_first < 0) y_first = 0;
  if (tile_hit && queue) {
    const int obx = (width + OCC_BW - 1) / OCC_BW, oby = (height + OCC_BH - 1) / OCC_BH;
    const int groups = obx * ((oby + OCC_GROUP - 1) / OCC_GROUP);
    unsigned *cnt = queue + 2 * (frame & 1u), *cnt_next = queue + 2 * ((frame + 1u) & 1u), *list = queue + 4;
    if (table && y_first == 0 && y_end == height) {  // one CTA per group of 32 x 8 pixels
      occ_tile_kernel<<<groups, 256, occ_tile_smem(radius), st>>>(occ, width, height, radius, n_points, depth,
                                                                  reinterpret_cast<const char2 *>(table), tile_hit, (width + 7) / 8, obx);
      return cudaGetLastError();
    }
    occ_list_kernel<<<(groups + 7) / 8, 256, 0, st>>>(occ, width, height, radius, tile_hit, (width + 7) / 8, obx, oby, cnt,
                                                     cnt_next, list, y_first, y_end);
    occ_queue_kernel<<<(sms > 0 ? sms : 148) * occ_ctas_per_sm, 128, 0, st>>>(occ, width, height, radius, n_points, depth, taps, obx, cnt,
                                                                 list);
    return cudaGetLastError();
  }
  dim3 block(32, 2), grid((width + 31) / 32, (height + 1) / 2);
  if (grid.y > 65535) return cudaErrorInvalidValue;
  occlusion_kernel<<<grid, block, 0, st>>>(occ, width, height, radius, n_points, depth, tile_hit, (width + 7) / 8, taps);
  return cudaGetLastError();
}

// -------------------------------------------------------------------------------------------------------------
__global__ void shading_kernel(float *__restrict__ d_output, int Nx, int Ny, const Camera cam, float occ_strength,
                               const float *__restrict__ input_normals, const float *__restrict__ input_depth,
                               const float *__restrict__ input_occlusion, int y_first, int y_end) {
  const unsigned x = blockIdx.x * blockDim.x + threadIdx.x, y = (unsigned)y_first + blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= (unsigned)Nx || y >= (unsigned)y_end) return;
  const size_t p = x + (size_t)Nx * y;
  if (!(input_depth[p] < __int_as_float(0x7f800000))) {  // (the occlusion plane is not read where there is no surface)
    d_output[p] = 0.f;
    return;
  }
  d_output[p] = shade_pixel(x, y, Nx, Ny, cam, occ_strength, input_normals, input_depth, input_occlusion[p]);
}

cudaError_t launch_shading(float *out, int width, int height, const Camera &cam, float occ_strength,
                           const float *normals, const float *depth, const float *occ, cudaStream_t st, int y_first,
                           int y_end) {
  if (y_end < 0 || y_end > height) y_end = height;
  if (y_first < 0) y_first = 0;
  if (y_first >= y_end) return cudaSuccess;
  dim3 block(32, 8), grid((width + 31) / 32, (y_end - y_first + 7) / 8);
  shading_kernel<<<grid, block, 0, st>>>(out, width, height, cam, occ_strength, normals, depth, occ, y_first, y_end);
  return cudaGetLastError();
}

// every kernel a sort-last iso surface launches (see preload_mip_kernels)
#define SPV_PRELOAD(k)                                         \
  do {                                                         \
    cudaFuncAttributes fa_;                                    \
    cudaError_t e_ = cudaFuncGetAttributes(&fa_, k);           \
    if (e_ != cudaSuccess) return e_;                          \
  } while (0)
template <int FMT>
static cudaError_t preload_iso_fmt() {
  SPV_PRELOAD((iso_slab_search_kernel<FMT, true, true>));
  SPV_PRELOAD((iso_slab_search_kernel<FMT, false, true>));
  SPV_PRELOAD((iso_slab_search_kernel<FMT, true, false>));
  SPV_PRELOAD((iso_slab_search_kernel<FMT, false, false>));
  SPV_PRELOAD((iso_slab_resolve_kernel<FMT, true>));
  SPV_PRELOAD((iso_slab_resolve_kernel<FMT, false>));
  return cudaSuccess;
}
cudaError_t preload_iso_kernels() {
  cudaError_t e;
  if ((e = preload_iso_fmt<0>()) != cudaSuccess) return e;
  if ((e = preload_iso_fmt<1>()) != cudaSuccess) return e;
  if ((e = preload_iso_fmt<2>()) != cudaSuccess) return e;
  if ((e = preload_iso_fmt<4>()) != cudaSuccess) return e;
  if ((e = preload_iso_fmt<5>()) != cudaSuccess) return e;
  SPV_PRELOAD(iso_slab_fix_kernel);
  SPV_PRELOAD((conv_xy_kernel<3, 7, false>));
  SPV_PRELOAD((conv_xy_kernel<1, 5, false>));
  SPV_PRELOAD((conv_xy_kernel<1, 0, false>));
  SPV_PRELOAD((conv_xy_kernel<3, 0, false>));
  SPV_PRELOAD((conv_xy_kernel<1, 5, true>));
  SPV_PRELOAD((conv_xy_kernel<1, 0, true>));
  SPV_PRELOAD(occ_taps_kernel);
  SPV_PRELOAD(occ_list_kernel);
  SPV_PRELOAD(occ_queue_kernel);
  SPV_PRELOAD(occ_table_kernel);
  SPV_PRELOAD(occ_tile_kernel);
  SPV_PRELOAD(occlusion_kernel);
  SPV_PRELOAD(shading_kernel);
  return cudaSuccess;
}

}  // namespace spv
