/*
 * spim_oracle.c -- CPU ORACLE for the spimagine volume-raycasting hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it.  The product path (spimagine_b200/, libspimcuda.so) never
 * links, imports or calls anything in oracle/.
 *
 * It is a plain-C (C99 + OpenMP) restatement of the reference's OpenCL kernels,
 * one work-item per loop iteration, in IEEE fp32 with contraction disabled
 * (build with -ffp-contract=off, no fast-math):
 *
 *   ray setup / box clip   spimagine/volumerender/kernels/volume_kernel.cl:45-93
 *                          spimagine/volumerender/kernels/utils.cl:41-70
 *   max_project_float      spimagine/volumerender/kernels/volume_kernel.cl:20-185
 *   max_project_short      spimagine/volumerender/kernels/volume_kernel.cl:190-345
 *   iso_surface            spimagine/volumerender/kernels/iso_kernel.cl:17-225
 *   shading                spimagine/volumerender/kernels/iso_kernel.cl:505-588
 *   conv_x/conv_y          spimagine/volumerender/kernels/convolve_2d.cl:6-61
 *   conv_vec_x/conv_vec_y  spimagine/volumerender/kernels/convolve_2d.cl:65-137
 *   occlusion, random      spimagine/volumerender/kernels/occlusion.cl:41-82,
 *                          spimagine/volumerender/kernels/utils.cl:10-38
 *   host launch order      spimagine/volumerender/volumerender.py:327-390, 446-506
 *
 * The image sampler (read_imagef / read_imageui) is NOT in the reference tree:
 * it belongs to the OpenCL implementation (pyopencl + an ICD such as POCL, both
 * unpinned in the reference's setup.py:27-35).  It is restated here from the
 * published OpenCL 1.2 specification, section 8.2 ("Image addressing and
 * filtering"): normalised coordinates, CLK_ADDRESS_CLAMP_TO_EDGE, and
 * CLK_FILTER_NEAREST / CLK_FILTER_LINEAR.
 *
 * PARITY PIN: the reference ships no golden vectors for this path and cannot
 * run here (no pyopencl / OpenCL device).  The restatement is pinned instead
 * against oracle/_ref/libspim_ref.so, which oracle/build.py compiles from
 * the reference's own kernel TEXT (the .cl files where they lie under
 * /root/reference, through a small OpenCL-C-on-g++ shim).  tests/ asserts that
 * both agree bit for bit, and tests/golden/ holds outputs of that build.
 *
 * Operation order (shared with oracle/ocl_shim.hpp and with the CUDA kernels'
 * "exact" mode so results can be compared bitwise):
 *   dot(a,b)      = ((a.x*b.x + a.y*b.y) + a.z*b.z) + a.w*b.w
 *   normalize(v)  = v / sqrt(dot(v,v))   (zero vector -> zero vector)
 *   min(a,b)      = b < a ? b : a        max(a,b) = a < b ? b : a
 *   clamp(x,l,h)  = fmin(fmax(x,l),h)    (NaN -> l)
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define SO_EXPORT __attribute__((visibility("default")))

typedef struct { float x, y, z, w; } v4;

/* ------------------------------------------------------------------ */
/* volume + sampler description (mirrors oracle/ocl_shim.hpp)          */
/* ------------------------------------------------------------------ */
typedef struct {
  const void *data;   /* C-order (z,y,x) */
  int dtype;          /* 0 = float32, 1 = uint16, 2 = uint8 */
  int nx, ny, nz;
  int filter;         /* 0 = CLK_FILTER_NEAREST, 1 = CLK_FILTER_LINEAR */
  int int_linear;     /* integer images under a LINEAR sampler: 1 = interpolate
                         (as a float image would), 0 = nearest (devices that
                         cannot filter integer formats); the OpenCL spec leaves
                         read_imageui + CLK_FILTER_LINEAR undefined */
  int weight_bits;    /* 0 = fp32 weights (OpenCL spec); 8 = weights rounded to
                         8 fractional bits (model of the CUDA texture unit) */
} so_volume;

static inline v4 mk4(float x, float y, float z, float w) { v4 r = {x, y, z, w}; return r; }
static inline v4 add4(v4 a, v4 b) { return mk4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
static inline v4 sub4(v4 a, v4 b) { return mk4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
static inline v4 mul4(v4 a, v4 b) { return mk4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
static inline v4 div4(v4 a, v4 b) { return mk4(a.x / b.x, a.y / b.y, a.z / b.z, a.w / b.w); }
static inline v4 scl4(float s, v4 a) { return mk4(s * a.x, s * a.y, s * a.z, s * a.w); }
static inline v4 sadd4(float s, v4 a) { return mk4(s + a.x, s + a.y, s + a.z, s + a.w); }
static inline float dot4(v4 a, v4 b) { return ((a.x * b.x + a.y * b.y) + a.z * b.z) + a.w * b.w; }
static inline float minf_cl(float a, float b) { return b < a ? b : a; }
static inline float maxf_cl(float a, float b) { return a < b ? b : a; }
static inline v4 min4(v4 a, v4 b) { return mk4(minf_cl(a.x, b.x), minf_cl(a.y, b.y), minf_cl(a.z, b.z), minf_cl(a.w, b.w)); }
static inline v4 max4(v4 a, v4 b) { return mk4(maxf_cl(a.x, b.x), maxf_cl(a.y, b.y), maxf_cl(a.z, b.z), maxf_cl(a.w, b.w)); }
static inline float clampf_cl(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
static inline int clampi_cl(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }
static inline v4 normalize4(v4 v) {
  float d = dot4(v, v);
  if (d == 0.f) return v;
  float s = sqrtf(d);
  return mk4(v.x / s, v.y / s, v.z / s, v.w / s);
}
/* utils.cl:63-70 */
static inline v4 mult(const float *M, v4 v) {
  v4 r;
  r.x = dot4(v, mk4(M[0], M[1], M[2], M[3]));
  r.y = dot4(v, mk4(M[4], M[5], M[6], M[7]));
  r.z = dot4(v, mk4(M[8], M[9], M[10], M[11]));
  r.w = dot4(v, mk4(M[12], M[13], M[14], M[15]));
  return r;
}

/* utils.cl:41-60 */
static inline int intersectBox(v4 r_o, v4 r_d, v4 boxmin, v4 boxmax, float *tnear, float *tfar) {
  v4 invR = div4(mk4(1.0f, 1.0f, 1.0f, 1.0f), r_d);
  v4 tbot = mul4(invR, sub4(boxmin, r_o));
  v4 ttop = mul4(invR, sub4(boxmax, r_o));
  v4 tmin = min4(ttop, tbot);
  v4 tmax = max4(ttop, tbot);
  float largest_tmin = maxf_cl(maxf_cl(tmin.x, tmin.y), maxf_cl(tmin.x, tmin.z));
  float smallest_tmax = minf_cl(minf_cl(tmax.x, tmax.y), minf_cl(tmax.x, tmax.z));
  *tnear = largest_tmin;
  *tfar = smallest_tmax;
  return smallest_tmax > largest_tmin;
}

/* ------------------------------------------------------------------ */
/* sampler: OpenCL 1.2 spec section 8.2                                */
/* ------------------------------------------------------------------ */
static inline float texel(const so_volume *V, int i, int j, int k) {
  size_t o = ((size_t)k * (size_t)V->ny + (size_t)j) * (size_t)V->nx + (size_t)i;
  switch (V->dtype) {
    case 0: return ((const float *)V->data)[o];
    case 1: return (float)((const uint16_t *)V->data)[o];
    default: return (float)((const uint8_t *)V->data)[o];
  }
}
/* floor() to int with the float clamped first so the cast is always defined */
static inline int floor_to_int(float f, int n) {
  float fl = floorf(f);
  if (!(fl >= -1.f)) fl = -1.f;      /* also catches NaN */
  if (fl > (float)n) fl = (float)n;
  return (int)fl;
}
static inline float quant_weight(float a, int bits) {
  if (bits <= 0) return a;
  float s = (float)(1 << bits);
  return floorf(a * s + 0.5f) / s;
}
static inline float sample_uvw(const so_volume *V, float u, float v, float w) {
  int linear = V->filter && (V->dtype == 0 || V->int_linear);
  if (!linear) {
    int i = clampi_cl(floor_to_int(u, V->nx), 0, V->nx - 1);
    int j = clampi_cl(floor_to_int(v, V->ny), 0, V->ny - 1);
    int k = clampi_cl(floor_to_int(w, V->nz), 0, V->nz - 1);
    return texel(V, i, j, k);
  }
  float ub = u - 0.5f, vb = v - 0.5f, wb = w - 0.5f;
  int i0 = floor_to_int(ub, V->nx), j0 = floor_to_int(vb, V->ny), k0 = floor_to_int(wb, V->nz);
  float a = quant_weight(ub - floorf(ub), V->weight_bits);
  float b = quant_weight(vb - floorf(vb), V->weight_bits);
  float c = quant_weight(wb - floorf(wb), V->weight_bits);
  if (!(a == a)) a = 0.f;
  if (!(b == b)) b = 0.f;
  if (!(c == c)) c = 0.f;
  int i1 = clampi_cl(i0 + 1, 0, V->nx - 1), j1 = clampi_cl(j0 + 1, 0, V->ny - 1), k1 = clampi_cl(k0 + 1, 0, V->nz - 1);
  i0 = clampi_cl(i0, 0, V->nx - 1); j0 = clampi_cl(j0, 0, V->ny - 1); k0 = clampi_cl(k0, 0, V->nz - 1);
  float a1 = 1.f - a, b1 = 1.f - b, c1 = 1.f - c;
  /* the eight-term sum exactly as the specification writes it */
  float T = a1 * b1 * c1 * texel(V, i0, j0, k0);
  T = T + a * b1 * c1 * texel(V, i1, j0, k0);
  T = T + a1 * b * c1 * texel(V, i0, j1, k0);
  T = T + a * b * c1 * texel(V, i1, j1, k0);
  T = T + a1 * b1 * c * texel(V, i0, j0, k1);
  T = T + a * b1 * c * texel(V, i1, j0, k1);
  T = T + a1 * b * c * texel(V, i0, j1, k1);
  T = T + a * b * c * texel(V, i1, j1, k1);
  return T;
}
static inline float sample(const so_volume *V, v4 pos) {
  return sample_uvw(V, pos.x * (float)V->nx, pos.y * (float)V->ny, pos.z * (float)V->nz);
}

/* ------------------------------------------------------------------ */
/* ray setup shared by the three kernels                               */
/*   volume_kernel.cl:45-93, iso_kernel.cl:39-91, iso_kernel.cl:517-542 */
/* ------------------------------------------------------------------ */
typedef struct { v4 orig, direc; float tnear, tfar; int hit; } ray_t;

static inline void eye_ray(unsigned x, unsigned y, unsigned Nx, unsigned Ny,
                           const float *invP, const float *invM, v4 *orig, v4 *direc) {
  float u = (x / (float)Nx) * 2.0f - 1.0f;
  float v = (y / (float)Ny) * 2.0f - 1.0f;
  v4 front = mk4(u, v, -1.f, 1.f);
  v4 back = mk4(u, v, 1.f, 1.f);
  v4 orig0 = mult(invP, front);
  orig0 = scl4(1.f / orig0.w, orig0);
  v4 o = mult(invM, orig0);
  o = scl4(1.f / o.w, o);
  v4 temp = mult(invP, back);
  temp = scl4(1.f / temp.w, temp);
  v4 d = mult(invM, normalize4(sub4(temp, orig0)));
  d.w = 0.0f;
  *orig = o;
  *direc = d;
}

static inline ray_t make_ray(unsigned x, unsigned y, unsigned Nx, unsigned Ny,
                             const float *invP, const float *invM, const float *box) {
  ray_t r;
  eye_ray(x, y, Nx, Ny, invP, invM, &r.orig, &r.direc);
  v4 boxMin = mk4(box[0], box[2], box[4], 1.f);
  v4 boxMax = mk4(box[1], box[3], box[5], 1.f);
  r.hit = intersectBox(r.orig, r.direc, boxMin, boxMax, &r.tnear, &r.tfar);
  return r;
}

/* Row sampling for timing runs (bench.py's bounded CPU samples): so_max_project visits only rows
 * y = so_row0, so_row0 + so_rowstep, ... ; the default (0, 1) is the whole image. */
static int so_row0 = 0, so_rowstep = 1;
SO_EXPORT void so_set_row_sampling(int row0, int rowstep) {
  so_row0 = row0 < 0 ? 0 : row0;
  so_rowstep = rowstep < 1 ? 1 : rowstep;
}

/* ------------------------------------------------------------------ */
/* max_project_float / max_project_short                               */
/* ------------------------------------------------------------------ */
/* pos_mode: 0 = accumulate pos += delta exactly like the reference loop;
 *           1 = pos_k = fma(k, delta, pos0)  (the form a kernel that skips or
 *               re-orders samples must use; differs by fp32 rounding drift)
 *           2 = as 1 but in unnormalised texel coordinates, u_k = fma(k, delta*N, pos0*N):
 *               the form libspimcuda's TMU kernel uses (alpha_pow == 0 branch only) */
SO_EXPORT int so_max_project(const so_volume *V, int width, int height,
                             const float *invP, const float *invM, const float *box,
                             float minVal, float maxVal, float gamma, float alpha_pow,
                             int numParts, int currentPart, int maxSteps, int pos_mode,
                             float *d_output, float *d_alpha_output) {
  const int isShort = V->dtype != 0;
  const unsigned Nx = (unsigned)width, Ny = (unsigned)height;
  if (numParts < 1) return -1;
#pragma omp parallel for schedule(dynamic, 4)
  for (int yy = so_row0; yy < height; yy += so_rowstep) {
    for (int xx = 0; xx < width; ++xx) {
      unsigned x = (unsigned)xx, y = (unsigned)yy;
      ray_t r = make_ray(x, y, Nx, Ny, invP, invM, box);
      if (!r.hit) {
        d_output[x + Nx * y] = 0.f;
        d_alpha_output[x + Nx * y] = isShort ? 0.f : -1.f;  /* volume_kernel.cl:88 vs :261 */
        continue;
      }
      float tnear = r.tnear, tfar = r.tfar;
      v4 orig = r.orig, direc = r.direc;
      if (tnear < 0.0f) tnear = 0.0f;
      float colVal = 0.f;
      const int reducedSteps = maxSteps / numParts;
      const int LOOPUNROLL = 16;
      const float dt = fabsf(tfar - tnear) / (float)((reducedSteps / LOOPUNROLL) * LOOPUNROLL);
      orig = add4(orig, scl4((float)currentPart * dt, direc));
      v4 delta_pos = scl4(.5f * dt, direc);
      v4 pos0 = scl4(0.5f, add4(sadd4(1.f, orig), scl4(tnear, direc)));
      v4 pos = pos0;
      float newVal;
      int k = 0;
      if (alpha_pow == 0) {
        for (int i = 0; i <= reducedSteps / LOOPUNROLL; ++i) {
          for (int j = 0; j < LOOPUNROLL; ++j) {
            if (pos_mode == 2)
              newVal = sample_uvw(V, fmaf((float)k, delta_pos.x * (float)V->nx, pos0.x * (float)V->nx),
                                  fmaf((float)k, delta_pos.y * (float)V->ny, pos0.y * (float)V->ny),
                                  fmaf((float)k, delta_pos.z * (float)V->nz, pos0.z * (float)V->nz));
            else
              newVal = sample(V, pos);
            colVal = fmaxf(colVal, newVal);
            ++k;
            if (pos_mode == 0) pos = add4(pos, delta_pos);
            else pos = mk4(fmaf((float)k, delta_pos.x, pos0.x), fmaf((float)k, delta_pos.y, pos0.y),
                           fmaf((float)k, delta_pos.z, pos0.z), 0.f);
          }
        }
        colVal = (maxVal == 0) ? colVal : (colVal - minVal) / (maxVal - minVal);
      } else {
        float cumsum = 1.f;
        for (int i = 0; i <= reducedSteps / LOOPUNROLL; ++i) {
          for (int j = 0; j < LOOPUNROLL; ++j) {
            newVal = sample(V, pos);
            newVal = (maxVal == 0) ? newVal : (newVal - minVal) / (maxVal - minVal);
            colVal = fmaxf(colVal, cumsum * newVal);
            if (isShort) cumsum *= (1.f - .1f * alpha_pow * alpha_pow * newVal);       /* :312 */
            else cumsum *= (1.f - alpha_pow * alpha_pow * clampf_cl(newVal, 0.f, 1.f)); /* :146 */
            ++k;
            if (pos_mode == 0) pos = add4(pos, delta_pos);
            else pos = mk4(fmaf((float)k, delta_pos.x, pos0.x), fmaf((float)k, delta_pos.y, pos0.y),
                           fmaf((float)k, delta_pos.z, pos0.z), 0.f);
            if (cumsum <= 0.01f) break;  /* leaves the inner loop only */
          }
        }
      }
      colVal = clampf_cl(powf(colVal, gamma), 0.f, 1.f);
      float alphaVal = isShort ? tnear : 1.f;  /* :329 vs :168 */
      if (currentPart == 0) {
        d_output[x + Nx * y] = colVal;
        d_alpha_output[x + Nx * y] = alphaVal;
      } else {
        d_output[x + Nx * y] = fmaxf(colVal, d_output[x + Nx * y]);
        d_alpha_output[x + Nx * y] = fmaxf(alphaVal, d_alpha_output[x + Nx * y]);
      }
    }
  }
  return 0;
}

/* Sort-last partial render (NOT in the reference; restates SURVEY.md 8e): the raw, un-windowed maximum over
 * the samples of each ray whose trilinear footprint STARTS in slices [z0, z1) of the volume (clamped slice
 * index), -1 for rays that miss the box, 0 for hit rays that own no sample.  Positions as pos_mode 2.  The
 * element-wise maximum over a partition of [0, nz) equals the z0=0, z1=nz render bit for bit. */
SO_EXPORT int so_max_project_raw(const so_volume *V, int width, int height, const float *invP, const float *invM,
                                 const float *box, int maxSteps, int z0, int z1, float *raw) {
  const unsigned Nx = (unsigned)width, Ny = (unsigned)height;
#pragma omp parallel for schedule(dynamic, 4)
  for (int yy = 0; yy < height; ++yy) {
    for (int xx = 0; xx < width; ++xx) {
      unsigned x = (unsigned)xx, y = (unsigned)yy;
      ray_t r = make_ray(x, y, Nx, Ny, invP, invM, box);
      if (!r.hit) { raw[x + Nx * y] = -1.f; continue; }
      float tnear = r.tnear < 0.0f ? 0.0f : r.tnear;
      const int S = (maxSteps / 16 + 1) * 16;
      const float dt = fabsf(r.tfar - tnear) / (float)((maxSteps / 16) * 16);
      v4 delta_pos = scl4(.5f * dt, r.direc);
      v4 pos0 = scl4(0.5f, add4(sadd4(1.f, r.orig), scl4(tnear, r.direc)));
      const float u0 = pos0.x * (float)V->nx, v0 = pos0.y * (float)V->ny, w0 = pos0.z * (float)V->nz;
      const float du = delta_pos.x * (float)V->nx, dv = delta_pos.y * (float)V->ny, dw = delta_pos.z * (float)V->nz;
      float colVal = 0.f;
      for (int k = 0; k < S; ++k) {
        float w = fmaf((float)k, dw, w0);
        float kf = fminf(fmaxf(floorf(w - 0.5f), 0.f), (float)(V->nz - 1));
        if (kf >= (float)z0 && kf < (float)z1)
          colVal = fmaxf(colVal, sample_uvw(V, fmaf((float)k, du, u0), fmaf((float)k, dv, v0), w));
      }
      raw[x + Nx * y] = colVal;
    }
  }
  return 0;
}

/* Number of rays that hit the box: bench.py counts algorithmic samples with it. */
SO_EXPORT long so_count_hit_rays(int width, int height, const float *invP, const float *invM,
                                 const float *box) {
  long n = 0;
#pragma omp parallel for reduction(+ : n)
  for (int yy = 0; yy < height; ++yy)
    for (int xx = 0; xx < width; ++xx) {
      ray_t r = make_ray((unsigned)xx, (unsigned)yy, (unsigned)width, (unsigned)height, invP, invM, box);
      n += r.hit ? 1 : 0;
    }
  return n;
}

/* ------------------------------------------------------------------ */
/* iso_surface   iso_kernel.cl:17-225                                   */
/* ------------------------------------------------------------------ */
SO_EXPORT int so_iso_surface(const so_volume *V, int width, int height,
                             const float *invP, const float *invM, const float *box,
                             float isoVal, float gamma, int maxSteps,
                             float *d_output, float *d_alpha_output, float *d_depth_output,
                             float *d_normals_output) {
  const unsigned Nx = (unsigned)width, Ny = (unsigned)height;
#pragma omp parallel for schedule(dynamic, 4)
  for (int yy = 0; yy < height; ++yy) {
    for (int xx = 0; xx < width; ++xx) {
      unsigned x = (unsigned)xx, y = (unsigned)yy;
      size_t p = x + (size_t)Nx * y;
      ray_t r = make_ray(x, y, Nx, Ny, invP, invM, box);
      if (!r.hit) {
        d_output[p] = 0.f; d_alpha_output[p] = 0.f; d_depth_output[p] = INFINITY;
        d_normals_output[3 * p + 0] = 0.f; d_normals_output[3 * p + 1] = 0.f; d_normals_output[3 * p + 2] = 0.f;
        continue;
      }
      float tnear = r.tnear, tfar = r.tfar;
      v4 orig = r.orig, direc = r.direc;
      if (tnear < 0.0f) tnear = 0.0f;
      float colVal = 0;
      float dt = 1.f * (tfar - tnear) / ((float)maxSteps - 1.f);
      v4 delta_pos = scl4(.5f * dt, direc);
      v4 pos = scl4(0.5f, add4(sadd4(1.f, orig), scl4(tnear, direc)));
      /* :106 read_imagef on what may be an integer image: restated as a proper read */
      float newVal = sample(V, pos);
      int isGreater = newVal > isoVal;
      int hitIso = 0;
      float t_hit = INFINITY;
      int i = 1;
      for (i = 1; i < maxSteps; i++) {
        pos = add4(pos, delta_pos);
        t_hit = tnear + (float)i * dt;
        newVal = sample(V, pos);
        if ((newVal > isoVal) != isGreater) { hitIso = 1; break; }
      }
      if (!hitIso) {
        d_output[p] = 0.f; d_alpha_output[p] = 0.f; d_depth_output[p] = INFINITY;
        d_normals_output[3 * p + 0] = 0.f; d_normals_output[3 * p + 1] = 0.f; d_normals_output[3 * p + 2] = 0.f;
        continue;
      }
      const int maxBisect = 10;
      v4 delta_pos2 = mk4(delta_pos.x / (float)maxBisect, delta_pos.y / (float)maxBisect,
                          delta_pos.z / (float)maxBisect, delta_pos.w / (float)maxBisect);
      float dt2 = dt / (float)maxBisect;
      pos = add4(scl4(0.5f, add4(sadd4(1.f, orig), scl4(tnear, direc))), scl4((float)(i - 1), delta_pos));
      for (int j = 1; j <= maxBisect; j++) {
        newVal = sample(V, pos);
        pos = add4(pos, delta_pos2);
        t_hit += dt2;
        if ((newVal > isoVal) != isGreater) break;
      }
      v4 light = mk4(2.f, -1.f, -2.f, 0.f);
      float c_ambient = .3f, c_diffuse = .4f, c_specular = .3f;
      light = mult(invM, light);
      light = normalize4(light);
      v4 normal;
      float h = dt;
      h *= powf(gamma, 2.f);
      float h2 = 2.f * h;
      normal.x = 2.f * sample(V, add4(pos, mk4(h, 0, 0, 0))) - 2.f * sample(V, add4(pos, mk4(-h, 0, 0, 0)))
               + sample(V, add4(pos, mk4(h2, 0, 0, 0))) - sample(V, add4(pos, mk4(-h2, 0, 0, 0)));
      normal.y = 2.f * sample(V, add4(pos, mk4(0, h, 0, 0))) - 2.f * sample(V, add4(pos, mk4(0, -h, 0, 0)))
               + sample(V, add4(pos, mk4(0, h2, 0, 0))) - sample(V, add4(pos, mk4(0, -h2, 0, 0)));
      normal.z = sample(V, add4(pos, mk4(0, 0, h, 0))) - sample(V, add4(pos, mk4(0, 0, -h, 0)))
               + sample(V, add4(pos, mk4(0, 0, h2, 0))) - sample(V, add4(pos, mk4(0, 0, -h2, 0)));
      normal.w = 0;
      normal = scl4(1.f - (float)(2 * isGreater), normalize4(normal));
      v4 reflect = sub4(scl4(2.f * dot4(light, normal), normal), light);
      float diffuse = fmaxf(0.f, dot4(light, normal));
      float specular = powf(fmaxf(0.f, dot4(normalize4(reflect), normalize4(direc))), 10.f);
      colVal = c_ambient + c_diffuse * diffuse + (float)(diffuse > 0) * c_specular * specular;
      d_output[p] = colVal;
      d_alpha_output[p] = tnear;
      d_depth_output[p] = t_hit;
      d_normals_output[3 * p + 0] = normal.x;
      d_normals_output[3 * p + 1] = normal.y;
      d_normals_output[3 * p + 2] = normal.z;
    }
  }
  return 0;
}

/* ------------------------------------------------------------------ */
/* separable blurs  convolve_2d.cl:6-137                                */
/*   ncomp = 1, coef = -10  : conv_x / conv_y                          */
/*   ncomp = 3, coef = -5   : conv_vec_x / conv_vec_y                  */
/* ------------------------------------------------------------------ */
static void conv_pass(const float *input, float *output, int Nx, int Ny, int Nh, int ncomp, float coef, int along_y) {
#pragma omp parallel for schedule(static)
  for (int j = 0; j < Ny; ++j) {
    for (int i = 0; i < Nx; ++i) {
      float res[3] = {0.f, 0.f, 0.f};
      float sum_val = 0.f;
      int c = along_y ? j : i, N = along_y ? Ny : Nx;
      int start = c - Nh / 2;
      const int h_start = ((c - Nh / 2) < 0) ? Nh / 2 - c : 0;
      const int h_end = ((c + Nh / 2) >= N) ? Nh - (c + Nh / 2 - N + 1) : Nh;
      for (int ht = h_start; ht < h_end; ++ht) {
        float val = expf((float)(coef * ((float)ht - (float)Nh / 2.f) * ((float)ht - (float)Nh / 2.f) / (float)Nh / (float)Nh));
        sum_val += val;
        size_t q = along_y ? ((size_t)i + (size_t)(start + ht) * Nx) : ((size_t)(start + ht) + (size_t)j * Nx);
        for (int k = 0; k < ncomp; ++k) res[k] += val * input[ncomp * q + k];
      }
      size_t o = (size_t)i + (size_t)j * Nx;
      for (int k = 0; k < ncomp; ++k) output[ncomp * o + k] = res[k] / sum_val;
    }
  }
}
/* volumerender.py:392-403: buf -> tmp (x), tmp -> buf (y) */
SO_EXPORT int so_convolve_scalar(float *buf, float *tmp, int width, int height, int Nh) {
  conv_pass(buf, tmp, width, height, Nh, 1, -10.f, 0);
  conv_pass(tmp, buf, width, height, Nh, 1, -10.f, 1);
  return 0;
}
/* volumerender.py:405-416 */
SO_EXPORT int so_convolve_vec(float *buf, float *tmp, int width, int height, int Nh) {
  conv_pass(buf, tmp, width, height, Nh, 3, -5.f, 0);
  conv_pass(tmp, buf, width, height, Nh, 3, -5.f, 1);
  return 0;
}

/* ------------------------------------------------------------------ */
/* hashed LCG  utils.cl:10-38                                           */
/* ------------------------------------------------------------------ */
static inline uint32_t lcg_hash(uint32_t x, uint32_t y) {
  uint32_t a = 4421u + (1u + x) * (1u + y) + x + y;
  for (int i = 0; i < 10; i++) a = (1664525u * a + 1013904223u) % 79197919u;
  return a;
}
static inline float random_cl(uint32_t x, uint32_t y) { return ((float)lcg_hash(x, y) * 1.0f) / (float)(79197919); }
static inline float rand_int_cl(uint32_t x, uint32_t y, int start, int end) {
  float rnd = random_cl(x, y);
  return (float)(int)((float)start + rnd * (float)(end - start));
}
SO_EXPORT uint32_t so_lcg_hash(uint32_t x, uint32_t y) { return lcg_hash(x, y); }
SO_EXPORT float so_random(uint32_t x, uint32_t y) { return random_cl(x, y); }
SO_EXPORT float so_rand_int(uint32_t x, uint32_t y, int start, int end) { return rand_int_cl(x, y, start, end); }

/* occlusion.cl:41-82 */
SO_EXPORT int so_occlusion(float *d_output, int width, int height, int radius, int number_points,
                           const float *input_depth) {
  const unsigned Nx = (unsigned)width, Ny = (unsigned)height;
  const float MPI_2 = 6.2831853071795f;
#pragma omp parallel for schedule(static)
  for (int y = 0; y < height; ++y) {
    for (int x = 0; x < width; ++x) {
      float depth0 = input_depth[x + y * Nx];
      float occ = 0.f;
      for (unsigned i = 0; i < (unsigned)number_points; ++i) {
        /* int x + float rand_int -> float; then float -> uint for random()'s parameters */
        float r = (float)(unsigned)radius *
                  random_cl((uint32_t)((float)x + rand_int_cl(i, i * i, 0, 1000)),
                            (uint32_t)((float)y + rand_int_cl(i * i, i, 294, 97701)));
        float phi = MPI_2 * random_cl((uint32_t)((float)x + rand_int_cl(i * i, i, 0, 1997)),
                                      (uint32_t)((float)y + rand_int_cl(i, i * i, 569, 17633)));
        int x2 = clampi_cl((int)((float)x + r * cosf(phi)), 0, (int)Nx - 1);
        int y2 = clampi_cl((int)((float)y + r * sinf(phi)), 0, (int)Ny - 1);
        float depth = input_depth[x2 + y2 * Nx];
        occ += (depth < depth0 ? 1.f : 0.f);
      }
      d_output[x + Nx * y] = occ / (float)(unsigned)number_points;
    }
  }
  return 0;
}

/* shading  iso_kernel.cl:505-588 */
SO_EXPORT int so_shading(float *d_output, int width, int height, const float *invP, const float *invM,
                         float occ_strength, const float *input_normals, const float *input_depth,
                         const float *input_occlusion) {
  const unsigned Nx = (unsigned)width, Ny = (unsigned)height;
#pragma omp parallel for schedule(static)
  for (int yy = 0; yy < height; ++yy) {
    for (int xx = 0; xx < width; ++xx) {
      unsigned x = (unsigned)xx, y = (unsigned)yy;
      v4 orig, direc;
      eye_ray(x, y, Nx, Ny, invP, invM, &orig, &direc);
      v4 light = mk4(2.f, -1.f, -2.f, 0.f);
      float c_ambient = .5f, c_diffuse = .3f, c_specular = .2f;
      light = mult(invM, light);
      light = normalize4(light);
      size_t p = x + (size_t)Nx * y;
      v4 normal = mk4(input_normals[3 * p], input_normals[1 + 3 * p], input_normals[2 + 3 * p], 0.f);
      float occ = input_occlusion[p];
      float depth = input_depth[p];
      normal = normalize4(normal);
      v4 reflect = sub4(scl4(2.f * dot4(light, normal), normal), light);
      float diffuse = fmaxf(0.f, dot4(light, normal));
      float specular = powf(fmaxf(0.f, dot4(normalize4(reflect), normalize4(direc))), 10.f);
      float colVal = c_ambient + c_diffuse * diffuse + (float)(diffuse > 0) * c_specular * specular;
      colVal = (1.f - occ_strength) * colVal + occ_strength * colVal * (1.f - occ);
      /* :580 mixes a double literal in: evaluated in double, rounded once on assignment */
      colVal = (float)((double)((1.f - occ_strength) * colVal) + 1.0 * (double)occ_strength * (double)colVal);
      colVal *= ((depth < INFINITY) ? 1.f : 0.f);
      d_output[p] = colVal;
    }
  }
  return 0;
}

/* The launch sequence of VolumeRenderer._render_isosurface (volumerender.py:446-506):
 * iso_surface -> conv_vec (Nh=7) -> occlusion -> conv scalar (Nh=5) -> shading. */
SO_EXPORT int so_render_isosurface(const so_volume *V, int width, int height,
                                   const float *invP, const float *invM, const float *box,
                                   float maxVal, float gamma, int maxSteps,
                                   float occ_strength, int occ_radius, int occ_n_points,
                                   float *out, float *alpha, float *depth, float *normals, float *occ,
                                   float *tmp, float *tmp_vec) {
  so_iso_surface(V, width, height, invP, invM, box, maxVal / 2, gamma, maxSteps, out, alpha, depth, normals);
  so_convolve_vec(normals, tmp_vec, width, height, 7);
  so_occlusion(occ, width, height, occ_radius, occ_n_points, depth);
  so_convolve_scalar(occ, tmp, width, height, 5);
  so_shading(out, width, height, invP, invM, occ_strength, normals, depth, occ);
  return 0;
}

SO_EXPORT float so_sample(const so_volume *V, float px, float py, float pz) { return sample(V, mk4(px, py, pz, 0.f)); }

SO_EXPORT int so_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
SO_EXPORT const char *so_kind(void) { return "port"; }
