// spv_common.cuh -- shared device code of libspimcuda: the ray setup, the volume
// samplers and the min/max brick grid lookup.
//
// Arithmetic convention.  The whole library is compiled with -fmad=false and
// IEEE division / square root, and the operation ORDER below is the one the
// reference's OpenCL text implies when evaluated left to right in fp32
// (spimagine/volumerender/kernels/utils.cl:41-70, volume_kernel.cl:45-93):
//   dot(a,b)     = ((a.x*b.x + a.y*b.y) + a.z*b.z) + a.w*b.w
//   normalize(v) = v / sqrt(dot(v,v)),  zero vector -> zero vector
//   min(a,b)     = b < a ? b : a,   max(a,b) = a < b ? b : a
// so the hit mask, tnear/tfar and dt of every pixel can be compared bitwise
// with a host evaluation of the reference kernels.  Fused multiply-adds are
// written explicitly (fmaf) where the fast path wants them.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace spv {

struct v4 {
  float x, y, z, w;
};
__device__ __forceinline__ v4 mk4(float x, float y, float z, float w) {
  v4 r;
  r.x = x; r.y = y; r.z = z; r.w = w;
  return r;
}
__device__ __forceinline__ v4 add4(v4 a, v4 b) { return mk4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ v4 sub4(v4 a, v4 b) { return mk4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
__device__ __forceinline__ v4 mul4(v4 a, v4 b) { return mk4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ v4 scl4(float s, v4 a) { return mk4(s * a.x, s * a.y, s * a.z, s * a.w); }
__device__ __forceinline__ v4 sadd4(float s, v4 a) { return mk4(s + a.x, s + a.y, s + a.z, s + a.w); }
__device__ __forceinline__ float dot4(v4 a, v4 b) { return ((a.x * b.x + a.y * b.y) + a.z * b.z) + a.w * b.w; }
__device__ __forceinline__ float minf_cl(float a, float b) { return b < a ? b : a; }
__device__ __forceinline__ float maxf_cl(float a, float b) { return a < b ? b : a; }
__device__ __forceinline__ float clampf_cl(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
__device__ __forceinline__ int clampi(int x, int lo, int hi) { return min(max(x, lo), hi); }
__device__ __forceinline__ v4 normalize4(v4 v) {
  float d = dot4(v, v);
  if (d == 0.f) return v;
  float s = sqrtf(d);
  return mk4(v.x / s, v.y / s, v.z / s, v.w / s);
}

// camera: invP, invM row-major (utils.cl:63-70 `mult`)
struct Camera {
  float invP[16];
  float invM[16];
};
__device__ __forceinline__ v4 mult(const float *M, v4 v) {
  v4 r;
  r.x = dot4(v, mk4(M[0], M[1], M[2], M[3]));
  r.y = dot4(v, mk4(M[4], M[5], M[6], M[7]));
  r.z = dot4(v, mk4(M[8], M[9], M[10], M[11]));
  r.w = dot4(v, mk4(M[12], M[13], M[14], M[15]));
  return r;
}

struct Ray {
  v4 orig, direc;
  float tnear, tfar;
  bool hit;
};

// volume_kernel.cl:45-76 / iso_kernel.cl:39-71 / iso_kernel.cl:517-542: pixel-corner eye ray
__device__ __forceinline__ void eye_ray(unsigned x, unsigned y, unsigned Nx, unsigned Ny, const float *invP,
                                        const float *invM, v4 &orig, v4 &direc) {
  float u = ((float)x / (float)Nx) * 2.0f - 1.0f;
  float v = ((float)y / (float)Ny) * 2.0f - 1.0f;
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
  orig = o;
  direc = d;
}
__device__ __forceinline__ void eye_ray(unsigned x, unsigned y, unsigned Nx, unsigned Ny, const Camera &cam, v4 &orig,
                                        v4 &direc) {
  eye_ray(x, y, Nx, Ny, cam.invP, cam.invM, orig, direc);
}

// utils.cl:41-60 slab test (w lanes carried along like the float4 code does; they never matter)
__device__ __forceinline__ bool intersect_box(v4 r_o, v4 r_d, const float *box, float &tnear, float &tfar) {
  float ix = 1.0f / r_d.x, iy = 1.0f / r_d.y, iz = 1.0f / r_d.z;
  float bx = ix * (box[0] - r_o.x), by = iy * (box[2] - r_o.y), bz = iz * (box[4] - r_o.z);
  float tx = ix * (box[1] - r_o.x), ty = iy * (box[3] - r_o.y), tz = iz * (box[5] - r_o.z);
  float nx = minf_cl(tx, bx), ny = minf_cl(ty, by), nz = minf_cl(tz, bz);
  float mx = maxf_cl(tx, bx), my = maxf_cl(ty, by), mz = maxf_cl(tz, bz);
  float largest_tmin = maxf_cl(maxf_cl(nx, ny), maxf_cl(nx, nz));
  float smallest_tmax = minf_cl(minf_cl(mx, my), minf_cl(mx, mz));
  tnear = largest_tmin;
  tfar = smallest_tmax;
  return smallest_tmax > largest_tmin;
}

__device__ __forceinline__ Ray make_ray(unsigned x, unsigned y, unsigned Nx, unsigned Ny, const Camera &cam,
                                        const float *box) {
  Ray r;
  eye_ray(x, y, Nx, Ny, cam, r.orig, r.direc);
  r.hit = intersect_box(r.orig, r.direc, box, r.tnear, r.tfar);
  return r;
}
__device__ __forceinline__ Ray make_ray(unsigned x, unsigned y, unsigned Nx, unsigned Ny, const float *invP,
                                        const float *invM, const float *box) {
  Ray r;
  eye_ray(x, y, Nx, Ny, invP, invM, r.orig, r.direc);
  r.hit = intersect_box(r.orig, r.direc, box, r.tnear, r.tfar);
  return r;
}

// window and gamma of a projected maximum (volume_kernel.cl:160-170 / :320-330)
__device__ __forceinline__ float window_value(float col, float minVal, float maxVal, float gamma) {
  col = (maxVal == 0.f) ? col : (col - minVal) / (maxVal - minVal);
  if (gamma != 1.f) col = powf(col, gamma);
  return clampf_cl(col, 0.f, 1.f);
}

// ---------------------------------------------------------------------------------------------
// The resident volume, in one of two hardware layouts (both block-linear cudaArrays, i.e. bricked by the
// texture hardware), bound to texture objects with UNNORMALISED coordinates and clamp addressing:
//
//   LAYOUT_3D     3-D single-channel array; a sample is ONE trilinear fetch (two filter passes in the TMU).
//   LAYOUT_ZPAIR  (integer volumes) 2-D layered two-channel array whose layer k holds {v[k], v[k+1]}: a sample
//                 is ONE bilinear fetch (one filter pass: twice the texture-unit throughput) and an fp32 lerp
//                 along z in the ALU.  Costs twice the memory; the z weight is exact instead of 8-bit.
//
//   filt : the sampler of the current interpolation mode; integer volumes are read as normalised floats (the
//          only way the texture unit filters them) and scaled back by `scale`
//   pt   : point sampling, element type; exact texel fetches
// z_lo / z0 / z1 describe a slab of a larger volume (sort-last rendering): the array holds global slices
// [z_lo, z_lo + local_nz); this context owns samples whose footprint starts in [z0, z1).
// ---------------------------------------------------------------------------------------------
enum { LAYOUT_3D = 0, LAYOUT_ZPAIR = 1 };
struct Volume {
  cudaTextureObject_t filt;
  cudaTextureObject_t pt;
  int nx, ny, nz;          // GLOBAL extent (nz = gnz for a slab)
  int local_nz;            // slices resident here
  float fnx, fny, fnz;
  float scale;             // 1, 65535 or 255
  int z_lo, z0, z1;
  // brick grid: float2 {min,max} per brick of BRICK^3 texels, dilated by BRICK_DILATE texels
  const float2 *bricks;
  int gx, gy, gz;          // grid extent (gz counts LOCAL slices from z_lo)
};
constexpr int BRICK_SHIFT = 3;
constexpr int BRICK = 1 << BRICK_SHIFT;
constexpr int BRICK_DILATE = 2;  // footprint (+1) and coordinate rounding slack (+1)

// Kernels are specialised on FMT = dtype + 3 * layout  (dtype: 0 f32, 1 u16, 2 u8)
constexpr int NUM_FMT = 6;
template <int DT> struct TexelType;
template <> struct TexelType<0> { typedef float type; typedef float2 pair; };
template <> struct TexelType<1> { typedef unsigned short type; typedef ushort2 pair; };
template <> struct TexelType<2> { typedef unsigned char type; typedef uchar2 pair; };

// exact texel (i,j,k) with GLOBAL k
template <int FMT>
__device__ __forceinline__ float texel(const Volume &V, int i, int j, int k) {
  constexpr int DT = FMT % 3;
  if (FMT / 3 == LAYOUT_3D)
    return (float)tex3D<typename TexelType<DT>::type>(V.pt, (float)i + 0.5f, (float)j + 0.5f,
                                                       (float)(k - V.z_lo) + 0.5f);
  return (float)tex2DLayered<typename TexelType<DT>::pair>(V.pt, (float)i + 0.5f, (float)j + 0.5f, k - V.z_lo).x;
}

__device__ __forceinline__ int floor_to_int(float f, int n) {
  float fl = floorf(f);
  if (!(fl >= -1.f)) fl = -1.f;  // also NaN
  if (fl > (float)n) fl = (float)n;
  return (int)fl;
}

// SPV_SAMPLER_EXACT: the OpenCL 1.2 (section 8.2) sampler in fp32, eight-term sum in specification order.
// pos in normalised coordinates like read_imagef(volume, sampler, pos).
template <int FMT, bool LINEAR>
__device__ __forceinline__ float sample_exact(const Volume &V, float px, float py, float pz) {
  float u = px * V.fnx, v = py * V.fny, w = pz * V.fnz;
  if (!LINEAR) {
    int i = clampi(floor_to_int(u, V.nx), 0, V.nx - 1);
    int j = clampi(floor_to_int(v, V.ny), 0, V.ny - 1);
    int k = clampi(floor_to_int(w, V.nz), 0, V.nz - 1);
    return texel<FMT>(V, i, j, k);
  }
  float ub = u - 0.5f, vb = v - 0.5f, wb = w - 0.5f;
  int i0 = floor_to_int(ub, V.nx), j0 = floor_to_int(vb, V.ny), k0 = floor_to_int(wb, V.nz);
  float a = ub - floorf(ub), b = vb - floorf(vb), c = wb - floorf(wb);
  if (!(a == a)) a = 0.f;
  if (!(b == b)) b = 0.f;
  if (!(c == c)) c = 0.f;
  int i1 = clampi(i0 + 1, 0, V.nx - 1), j1 = clampi(j0 + 1, 0, V.ny - 1), k1 = clampi(k0 + 1, 0, V.nz - 1);
  i0 = clampi(i0, 0, V.nx - 1);
  j0 = clampi(j0, 0, V.ny - 1);
  k0 = clampi(k0, 0, V.nz - 1);
  float a1 = 1.f - a, b1 = 1.f - b, c1 = 1.f - c;
  float T = a1 * b1 * c1 * texel<FMT>(V, i0, j0, k0);
  T = T + a * b1 * c1 * texel<FMT>(V, i1, j0, k0);
  T = T + a1 * b * c1 * texel<FMT>(V, i0, j1, k0);
  T = T + a * b * c1 * texel<FMT>(V, i1, j1, k0);
  T = T + a1 * b1 * c * texel<FMT>(V, i0, j0, k1);
  T = T + a * b1 * c * texel<FMT>(V, i1, j0, k1);
  T = T + a1 * b * c * texel<FMT>(V, i0, j1, k1);
  T = T + a * b * c * texel<FMT>(V, i1, j1, k1);
  return T;
}

// SPV_SAMPLER_TMU: one hardware-filtered fetch.  (u,v,w) are UNNORMALISED texel coordinates of the
// global volume; subtracting the integer slab origin is exact in fp32, so a slab sees the same
// fractional weights as the whole volume would.
template <int FMT, bool LINEAR>
__device__ __forceinline__ float sample_tmu_uvw(const Volume &V, float u, float v, float w) {
  constexpr int DT = FMT % 3;
  const float wl = w - (float)V.z_lo;
  if (FMT / 3 == LAYOUT_3D) {
    if (DT != 0 && !LINEAR)  // integer + nearest: element-type point fetch, exact voxel values
      return (float)tex3D<typename TexelType<DT>::type>(V.pt, u, v, wl);
    float t = tex3D<float>(V.filt, u, v, wl);
    return DT == 0 ? t : t * V.scale;
  }
  const float top = (float)(V.local_nz - 1);
  if (!LINEAR) {
    const int layer = (int)fminf(fmaxf(floorf(wl), 0.f), top);
    return (float)tex2DLayered<typename TexelType<DT>::pair>(V.pt, u, v, layer).x;
  }
  // layer k = {v[k], v[k+1]} (the last layer repeats itself): bilinear in the texture unit, lerp along z here
  const float wb = wl - 0.5f;
  const float fl = floorf(wb);
  const float f = fl < 0.f ? 0.f : wb - fl;  // below the first slice centre: clamp-to-edge
  const int layer = (int)fminf(fmaxf(fl, 0.f), top);
  const float2 t = tex2DLayered<float2>(V.filt, u, v, layer);
  return fmaf(f, t.y - t.x, t.x) * V.scale;
}
template <int FMT, bool LINEAR>
__device__ __forceinline__ float sample_tmu(const Volume &V, float px, float py, float pz) {
  return sample_tmu_uvw<FMT, LINEAR>(V, px * V.fnx, py * V.fny, pz * V.fnz);
}

template <int FMT, bool LINEAR, bool EXACT>
__device__ __forceinline__ float sample(const Volume &V, float px, float py, float pz) {
  return EXACT ? sample_exact<FMT, LINEAR>(V, px, py, pz) : sample_tmu<FMT, LINEAR>(V, px, py, pz);
}

// ---- slab ownership (sort-last rendering) ----
// clamped slice index the footprint of the sample at unnormalised coordinate w(k) = w0 + k*dw starts in
__device__ __forceinline__ float slice_of_k(const Volume &V, float w0, float dw, int k) {
  return fminf(fmaxf(floorf(fmaf((float)k, dw, w0) - 0.5f), 0.f), (float)(V.nz - 1));
}
// [ka, kb) = the samples k in [0, S) whose slice lies in [z0, z1).  The slice index is monotone in k (fma,
// subtraction and floor all are), so the owned samples form one interval: two binary searches with the exact
// predicate.
__device__ __forceinline__ void owned_interval_w(const Volume &V, float w0, float dw, int S, int &ka, int &kb) {
  const float z0 = (float)V.z0, z1 = (float)V.z1;
  const bool up = dw >= 0.f;  // slice index non-decreasing in k
  int lo = 0, hi = S;
  while (lo < hi) {  // first k with (up ? slice >= z0 : slice < z1)
    const int mid = (lo + hi) >> 1;
    const float s = slice_of_k(V, w0, dw, mid);
    if (up ? (s >= z0) : (s < z1)) hi = mid; else lo = mid + 1;
  }
  ka = lo;
  hi = S;
  while (lo < hi) {  // first k >= ka with (up ? slice >= z1 : slice < z0)
    const int mid = (lo + hi) >> 1;
    const float s = slice_of_k(V, w0, dw, mid);
    if (up ? (s >= z1) : (s < z0)) hi = mid; else lo = mid + 1;
  }
  kb = lo;
}

// brick grid addressing: x fastest
__device__ __forceinline__ float2 brick_at(const Volume &V, int bx, int by, int bz) {
  return __ldg(V.bricks + ((size_t)bz * V.gy + by) * V.gx + bx);
}

}  // namespace spv
