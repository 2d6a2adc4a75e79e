// spv_bricks.cu -- min/max brick grids of the resident volume, built once per upload.
//
// A brick is BRICK^3 texels.  The value stored for brick b covers texels [BRICK*b - (D-1), BRICK*b + BRICK-1 + D]
// on every axis (clamped to the array): one texel for the upper neighbour of a trilinear footprint and one more
// either side as slack for coordinate rounding, so that "this sample's footprint starts in brick b, give or take
// one texel" is enough to bound the sample by the brick's {min,max}.  The coarse grid holds the {min,max} over
// 4^3 bricks, the top grid over 4^3 coarse cells.  The same pass yields the volume's global min/max (what GLWidget._get_min_max computes with a
// separate device reduction, spimagine/gui/glwidget.py:328-344).
#include "spv_kernels.h"

namespace spv {

template <int FMT>
__global__ void __launch_bounds__(128) brick_kernel(const Volume V, int local_nz, float2 *bricks) {
  const int bx = blockIdx.x, by = blockIdx.y, bz = blockIdx.z;
  const int D = BRICK_DILATE;
  const int x0 = max(BRICK * bx - (D - 1), 0), x1 = min(BRICK * bx + BRICK - 1 + D, V.nx - 1);
  const int y0 = max(BRICK * by - (D - 1), 0), y1 = min(BRICK * by + BRICK - 1 + D, V.ny - 1);
  const int z0 = max(BRICK * bz - (D - 1), 0), z1 = min(BRICK * bz + BRICK - 1 + D, local_nz - 1);
  const int ex = x1 - x0 + 1, ey = y1 - y0 + 1, ez = z1 - z0 + 1;
  const int n = ex * ey * ez;
  float lo = __int_as_float(0x7f800000), hi = -lo;
  for (int t = threadIdx.x; t < n; t += blockDim.x) {
    const int i = t % ex, j = (t / ex) % ey, k = t / (ex * ey);
    const float v = texel<FMT>(V, x0 + i, y0 + j, V.z_lo + z0 + k);  // texel() takes the global slice index
    lo = fminf(lo, v);
    hi = fmaxf(hi, v);
  }
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_down_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_down_sync(0xffffffffu, hi, o));
  }
  __shared__ float s_lo[4], s_hi[4];
  if ((threadIdx.x & 31) == 0) {
    s_lo[threadIdx.x >> 5] = lo;
    s_hi[threadIdx.x >> 5] = hi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 4; ++w) {
      lo = fminf(lo, s_lo[w]);
      hi = fmaxf(hi, s_hi[w]);
    }
    bricks[((size_t)bz * gridDim.y + by) * gridDim.x + bx] = make_float2(lo, hi);
  }
}

__global__ void coarse_kernel(const float2 *__restrict__ bricks, int gx, int gy, int gz, float2 *__restrict__ coarse,
                              int cgx, int cgy, int cgz) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cgx * cgy * cgz) return;
  const int cx = c % cgx, cy = (c / cgx) % cgy, cz = c / (cgx * cgy);
  float lo = __int_as_float(0x7f800000), hi = -lo;
  for (int k = 4 * cz; k < min(4 * cz + 4, gz); ++k)
    for (int j = 4 * cy; j < min(4 * cy + 4, gy); ++j)
      for (int i = 4 * cx; i < min(4 * cx + 4, gx); ++i) {
        const float2 b = bricks[((size_t)k * gy + j) * gx + i];
        lo = fminf(lo, b.x);
        hi = fmaxf(hi, b.y);
      }
  coarse[c] = make_float2(lo, hi);
}

__global__ void minmax_kernel(const float2 *__restrict__ coarse, int n, float *minmax) {
  float lo = __int_as_float(0x7f800000), hi = -lo;
  for (int t = threadIdx.x; t < n; t += blockDim.x) {
    lo = fminf(lo, coarse[t].x);
    hi = fmaxf(hi, coarse[t].y);
  }
  __shared__ float s_lo[1024], s_hi[1024];
  s_lo[threadIdx.x] = lo;
  s_hi[threadIdx.x] = hi;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      s_lo[threadIdx.x] = fminf(s_lo[threadIdx.x], s_lo[threadIdx.x + o]);
      s_hi[threadIdx.x] = fmaxf(s_hi[threadIdx.x], s_hi[threadIdx.x + o]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    minmax[0] = s_lo[0];
    minmax[1] = s_hi[0];
  }
}

cudaError_t launch_build_bricks(const Volume &vol, int dtype, int local_nz, float2 *bricks, float2 *coarse, int cgx,
                                int cgy, int cgz, float2 *top, float *minmax, cudaStream_t st) {
  dim3 grid(vol.gx, vol.gy, vol.gz);
  switch (dtype) {  // FMT = dtype + 3 * layout
    case 0: brick_kernel<0><<<grid, 128, 0, st>>>(vol, local_nz, bricks); break;
    case 1: brick_kernel<1><<<grid, 128, 0, st>>>(vol, local_nz, bricks); break;
    case 2: brick_kernel<2><<<grid, 128, 0, st>>>(vol, local_nz, bricks); break;
    case 4: brick_kernel<4><<<grid, 128, 0, st>>>(vol, local_nz, bricks); break;
    case 5: brick_kernel<5><<<grid, 128, 0, st>>>(vol, local_nz, bricks); break;
    default: return cudaErrorInvalidValue;
  }
  const int nc = cgx * cgy * cgz;
  coarse_kernel<<<(nc + 127) / 128, 128, 0, st>>>(bricks, vol.gx, vol.gy, vol.gz, coarse, cgx, cgy, cgz);
  const int tgx = (cgx + 3) / 4, tgy = (cgy + 3) / 4, tgz = (cgz + 3) / 4, nt = tgx * tgy * tgz;
  coarse_kernel<<<(nt + 127) / 128, 128, 0, st>>>(coarse, cgx, cgy, cgz, top, tgx, tgy, tgz);
  minmax_kernel<<<1, 1024, 0, st>>>(top, nt, minmax);
  return cudaGetLastError();
}

// LAYOUT_ZPAIR ingest: dst[z][y][x] = {src[z][y][x], src[min(z+1, nz-1)][y][x]} for z in [zbeg, zend), written to a
// C-order linear buffer that is then copied into the layered array.  One thread per 4 voxels of a row segment.
template <typename T, typename T2>
__global__ void pair_kernel(const T *__restrict__ src, T2 *__restrict__ dst, size_t slice, int nz, int zbeg, int zend) {
  const size_t n = slice * (size_t)(zend - zbeg);
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
    const size_t z = zbeg + t / slice, o = t % slice;
    const size_t z1 = z + 1 < (size_t)nz ? z + 1 : (size_t)nz - 1;
    T2 v;
    v.x = src[z * slice + o];
    v.y = src[z1 * slice + o];
    dst[t] = v;
  }
}

cudaError_t launch_pair(const void *src, void *dst, int dtype, size_t slice, int nz, int zbeg, int zend,
                        cudaStream_t st) {
  const size_t n = slice * (size_t)(zend - zbeg);
  const int blocks = (int)((n + 255) / 256 < 148 * 32 ? (n + 255) / 256 : 148 * 32);
  if (dtype == SPV_U16)
    pair_kernel<unsigned short, ushort2><<<blocks, 256, 0, st>>>((const unsigned short *)src, (ushort2 *)dst, slice, nz, zbeg, zend);
  else if (dtype == SPV_U8)
    pair_kernel<unsigned char, uchar2><<<blocks, 256, 0, st>>>((const unsigned char *)src, (uchar2 *)dst, slice, nz, zbeg, zend);
  else
    return cudaErrorInvalidValue;
  return cudaGetLastError();
}

// -------------------------------------------------------------------------------------------------------------
// Element-type conversion on the device (replaces the host `astype`, volumerender.py:245-246, 290-291).
template <typename D, typename S>
__device__ __forceinline__ D cast_like_numpy(S v) {
  return (D)v;  // integer -> integer wraps, anything -> float rounds to nearest
}
// floating point -> integer texels: truncate towards zero into int32, keep the low bits (what the x86 cvtt + narrowing
// store of numpy's cast loop yields for |v| < 2^31)
template <> __device__ __forceinline__ unsigned short cast_like_numpy<unsigned short, float>(float v) { return (unsigned short)(int)v; }
template <> __device__ __forceinline__ unsigned char cast_like_numpy<unsigned char, float>(float v) { return (unsigned char)(int)v; }
template <> __device__ __forceinline__ unsigned short cast_like_numpy<unsigned short, double>(double v) { return (unsigned short)(int)v; }
template <> __device__ __forceinline__ unsigned char cast_like_numpy<unsigned char, double>(double v) { return (unsigned char)(int)v; }

struct HalfBits { unsigned short b; };  // IEEE binary16, decoded without cuda_fp16.h
__device__ __forceinline__ float half_to_float(HalfBits h) {
  float f;
  asm("{ .reg .b16 t; mov.b16 t, %1; cvt.f32.f16 %0, t; }" : "=f"(f) : "h"(h.b));
  return f;
}
struct BoolByte { unsigned char b; };

template <typename D, typename S>
__device__ __forceinline__ D convert_one(S v) { return cast_like_numpy<D, S>(v); }
template <typename D>
__device__ __forceinline__ D convert_half(HalfBits v) { return cast_like_numpy<D, float>(half_to_float(v)); }
template <typename D>
__device__ __forceinline__ D convert_bool(BoolByte v) { return (D)(v.b != 0 ? 1 : 0); }

template <typename D, typename S, int KIND /* 0 plain, 1 half, 2 bool */>
__global__ void __launch_bounds__(256) convert_kernel(const S *__restrict__ src, D *__restrict__ dst, size_t n) {
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
    if (KIND == 1) dst[t] = convert_half<D>(*(const HalfBits *)(src + t));
    else if (KIND == 2) dst[t] = convert_bool<D>(*(const BoolByte *)(src + t));
    else dst[t] = convert_one<D, S>(src[t]);
  }
}

size_t src_elem_size(int src_type) {
  switch (src_type) {
    case SPV_SRC_I8: case SPV_SRC_U8: case SPV_SRC_BOOL: return 1;
    case SPV_SRC_I16: case SPV_SRC_U16: case SPV_SRC_F16: return 2;
    case SPV_SRC_I32: case SPV_SRC_U32: case SPV_SRC_F32: return 4;
    case SPV_SRC_I64: case SPV_SRC_U64: case SPV_SRC_F64: return 8;
    default: return 0;
  }
}

template <typename D>
static cudaError_t launch_convert_to(const void *src, D *dst, int src_type, size_t n, cudaStream_t st) {
  const int blocks = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
#define SPV_CV(S, KIND) convert_kernel<D, S, KIND><<<blocks, 256, 0, st>>>((const S *)src, dst, n)
  switch (src_type) {
    case SPV_SRC_I8: SPV_CV(signed char, 0); break;
    case SPV_SRC_U8: SPV_CV(unsigned char, 0); break;
    case SPV_SRC_I16: SPV_CV(short, 0); break;
    case SPV_SRC_U16: SPV_CV(unsigned short, 0); break;
    case SPV_SRC_I32: SPV_CV(int, 0); break;
    case SPV_SRC_U32: SPV_CV(unsigned int, 0); break;
    case SPV_SRC_I64: SPV_CV(long long, 0); break;
    case SPV_SRC_U64: SPV_CV(unsigned long long, 0); break;
    case SPV_SRC_F16: SPV_CV(unsigned short, 1); break;
    case SPV_SRC_F32: SPV_CV(float, 0); break;
    case SPV_SRC_F64: SPV_CV(double, 0); break;
    case SPV_SRC_BOOL: SPV_CV(unsigned char, 2); break;
    default: return cudaErrorInvalidValue;
  }
#undef SPV_CV
  return cudaGetLastError();
}

cudaError_t launch_convert(const void *src, void *dst, int src_type, int dtype, size_t n, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  if (dtype == SPV_F32) return launch_convert_to<float>(src, (float *)dst, src_type, n, st);
  if (dtype == SPV_U16) return launch_convert_to<unsigned short>(src, (unsigned short *)dst, src_type, n, st);
  if (dtype == SPV_U8) return launch_convert_to<unsigned char>(src, (unsigned char *)dst, src_type, n, st);
  return cudaErrorInvalidValue;
}

}  // namespace spv
