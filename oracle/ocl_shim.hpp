// ocl_shim.hpp -- just enough OpenCL C 1.x on top of g++ to compile the
// REFERENCE's own kernel text (spimagine/volumerender/kernels/*.cl) for the host.
//
// TEST INFRASTRUCTURE ONLY (see oracle/spim_oracle.c header).  Nothing here is
// derived from the reference: it supplies what the OpenCL *implementation*
// would supply (vector types, built-ins, the image sampler of OpenCL 1.2 spec
// section 8.2, work-item ids).  oracle/build.py streams the reference .cl
// files from /root/reference through a one-line syntax rewrite
// ("(float4)(" -> "float4(") into g++ together with this header and
// ref_driver.inc; only the resulting oracle/_ref/libspim_ref.so is kept.
//
// Operation order of the built-ins is the one documented in spim_oracle.c so the
// restatement and this build can be compared bit for bit.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned int uint;

struct float4 {
  float x, y, z, w;
  float4() : x(0), y(0), z(0), w(0) {}
  float4(float x_, float y_, float z_, float w_) : x(x_), y(y_), z(z_), w(w_) {}
};
static inline float4 operator+(float4 a, float4 b) { return float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
static inline float4 operator-(float4 a, float4 b) { return float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
static inline float4 operator*(float4 a, float4 b) { return float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
static inline float4 operator/(float4 a, float4 b) { return float4(a.x / b.x, a.y / b.y, a.z / b.z, a.w / b.w); }
static inline float4 operator*(float s, float4 a) { return float4(s * a.x, s * a.y, s * a.z, s * a.w); }
static inline float4 operator*(float4 a, float s) { return float4(a.x * s, a.y * s, a.z * s, a.w * s); }
static inline float4 operator/(float4 a, float s) { return float4(a.x / s, a.y / s, a.z / s, a.w / s); }
static inline float4 operator+(float s, float4 a) { return float4(s + a.x, s + a.y, s + a.z, s + a.w); }
static inline float4 operator+(float4 a, float s) { return float4(a.x + s, a.y + s, a.z + s, a.w + s); }
static inline float4 &operator+=(float4 &a, float4 b) { a = a + b; return a; }
static inline float4 &operator*=(float4 &a, float s) { a = a * s; return a; }

static inline float ocl_min(float a, float b) { return b < a ? b : a; }
static inline float ocl_max(float a, float b) { return a < b ? b : a; }
static inline float4 ocl_min(float4 a, float4 b) { return float4(ocl_min(a.x, b.x), ocl_min(a.y, b.y), ocl_min(a.z, b.z), ocl_min(a.w, b.w)); }
static inline float4 ocl_max(float4 a, float4 b) { return float4(ocl_max(a.x, b.x), ocl_max(a.y, b.y), ocl_max(a.z, b.z), ocl_max(a.w, b.w)); }
static inline float ocl_fmax(float a, float b) { return fmaxf(a, b); }
static inline float ocl_clamp(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
static inline int ocl_clamp(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }
static inline float ocl_dot(float4 a, float4 b) { return ((a.x * b.x + a.y * b.y) + a.z * b.z) + a.w * b.w; }
static inline float4 ocl_normalize(float4 v) {
  float d = ocl_dot(v, v);
  if (d == 0.f) return v;
  float s = sqrtf(d);
  return float4(v.x / s, v.y / s, v.z / s, v.w / s);
}
static inline float ocl_length(float4 v) { return sqrtf(ocl_dot(v, v)); }
static inline float ocl_pow(float a, float b) { return powf(a, b); }
static inline float ocl_pow(float a, int b) { return powf(a, (float)b); }
static inline float ocl_exp(float a) { return expf(a); }
static inline float ocl_cos(float a) { return cosf(a); }
static inline float ocl_sin(float a) { return sinf(a); }
static inline float ocl_fabs(float a) { return fabsf(a); }

#define min ocl_min
#define max ocl_max
#define fmax ocl_fmax
#define clamp ocl_clamp
#define dot ocl_dot
#define normalize ocl_normalize
#define length ocl_length
#define pow ocl_pow
#define exp ocl_exp
#define native_exp ocl_exp
#define native_sqrt sqrtf
#define cos ocl_cos
#define sin ocl_sin
#define fabs ocl_fabs

// address-space / kernel qualifiers
#define __kernel extern "C"
#define __global
#define __constant const
#define __read_only
#define __QUALIFIER_CONSTANT __constant  // all_render_kernels.cl:13-17 (default branch)

// ---- images and samplers (OpenCL 1.2 spec section 8.2) ----
struct ocl_image3d {
  const void *data;  // C-order (z,y,x)
  int dtype;         // 0 float32, 1 uint16, 2 uint8
  int nx, ny, nz;
  int filter;        // run-time stand-in for -D SAMPLER_FILTER
  int int_linear;    // see so_volume.int_linear
  int weight_bits;   // see so_volume.weight_bits
};
typedef const ocl_image3d *image3d_t;
typedef int sampler_t;
enum { CLK_NORMALIZED_COORDS_TRUE = 1, CLK_ADDRESS_CLAMP_TO_EDGE = 2, CLK_FILTER_NEAREST = 0x10, CLK_FILTER_LINEAR = 0x20 };

static inline float ocl_texel(image3d_t V, int i, int j, int k) {
  size_t o = ((size_t)k * (size_t)V->ny + (size_t)j) * (size_t)V->nx + (size_t)i;
  if (V->dtype == 0) return ((const float *)V->data)[o];
  if (V->dtype == 1) return (float)((const uint16_t *)V->data)[o];
  return (float)((const uint8_t *)V->data)[o];
}
static inline int ocl_floor_to_int(float f, int n) {
  float fl = floorf(f);
  if (!(fl >= -1.f)) fl = -1.f;
  if (fl > (float)n) fl = (float)n;
  return (int)fl;
}
static inline float ocl_quant(float a, int bits) {
  if (bits <= 0) return a;
  float s = (float)(1 << bits);
  return floorf(a * s + 0.5f) / s;
}
static inline float ocl_sample(image3d_t V, sampler_t smp, float4 pos) {
  float u = pos.x * (float)V->nx, v = pos.y * (float)V->ny, w = pos.z * (float)V->nz;
  bool linear = (smp & CLK_FILTER_LINEAR) && (V->dtype == 0 || V->int_linear);
  if (!linear) {
    int i = ocl_clamp(ocl_floor_to_int(u, V->nx), 0, V->nx - 1);
    int j = ocl_clamp(ocl_floor_to_int(v, V->ny), 0, V->ny - 1);
    int k = ocl_clamp(ocl_floor_to_int(w, V->nz), 0, V->nz - 1);
    return ocl_texel(V, i, j, k);
  }
  float ub = u - 0.5f, vb = v - 0.5f, wb = w - 0.5f;
  int i0 = ocl_floor_to_int(ub, V->nx), j0 = ocl_floor_to_int(vb, V->ny), k0 = ocl_floor_to_int(wb, V->nz);
  float a = ocl_quant(ub - floorf(ub), V->weight_bits);
  float b = ocl_quant(vb - floorf(vb), V->weight_bits);
  float c = ocl_quant(wb - floorf(wb), V->weight_bits);
  if (!(a == a)) a = 0.f;
  if (!(b == b)) b = 0.f;
  if (!(c == c)) c = 0.f;
  int i1 = ocl_clamp(i0 + 1, 0, V->nx - 1), j1 = ocl_clamp(j0 + 1, 0, V->ny - 1), k1 = ocl_clamp(k0 + 1, 0, V->nz - 1);
  i0 = ocl_clamp(i0, 0, V->nx - 1); j0 = ocl_clamp(j0, 0, V->ny - 1); k0 = ocl_clamp(k0, 0, V->nz - 1);
  float a1 = 1.f - a, b1 = 1.f - b, c1 = 1.f - c;
  float T = a1 * b1 * c1 * ocl_texel(V, i0, j0, k0);
  T = T + a * b1 * c1 * ocl_texel(V, i1, j0, k0);
  T = T + a1 * b * c1 * ocl_texel(V, i0, j1, k0);
  T = T + a * b * c1 * ocl_texel(V, i1, j1, k0);
  T = T + a1 * b1 * c * ocl_texel(V, i0, j0, k1);
  T = T + a * b1 * c * ocl_texel(V, i1, j0, k1);
  T = T + a1 * b * c * ocl_texel(V, i0, j1, k1);
  T = T + a * b * c * ocl_texel(V, i1, j1, k1);
  return T;
}
// read_imageui's .x carries the value as float: exact for point sampling, and
// "as if filtered like a float image" when int_linear is set (the OpenCL spec
// leaves integer reads through a LINEAR sampler undefined).
struct ocl_pixel { float x, y, z, w; };
static inline ocl_pixel read_imagef(image3d_t V, sampler_t s, float4 pos) { float t = ocl_sample(V, s, pos); return ocl_pixel{t, 0.f, 0.f, 1.f}; }
static inline ocl_pixel read_imageui(image3d_t V, sampler_t s, float4 pos) { float t = ocl_sample(V, s, pos); return ocl_pixel{t, 0.f, 0.f, 1.f}; }

// ---- work-item functions ----
static thread_local int ocl_gid[2];
static int ocl_gsize[2];
static inline int get_global_id(int d) { return ocl_gid[d]; }
static inline int get_global_size(int d) { return ocl_gsize[d]; }

// ---- what volumerender.py passes with -D ----
static int ocl_maxSteps = 200;
static int ocl_filter = CLK_FILTER_LINEAR;
#define maxSteps ocl_maxSteps
#define SAMPLER_FILTER ocl_filter
