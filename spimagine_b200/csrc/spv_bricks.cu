// spv_bricks.cu -- min/max brick grids of the resident volume, built once per upload.
//
// A brick is BRICK^3 texels.  The value stored for brick b covers texels [BRICK*b - (D-1), BRICK*b + BRICK-1 + D]
// on every axis (clamped to the array): one texel for the upper neighbour of a trilinear footprint and one more
// either side as slack for coordinate rounding, so that "this sample's footprint starts in brick b, give or take
// one texel" is enough to bound the sample by the brick's {min,max}.  The coarse grid holds the {min,max} over
// 4^3 bricks.  The same pass yields the volume's global min/max (what GLWidget._get_min_max computes with a
// separate device reduction, spimagine/gui/glwidget.py:328-344).
#include "spv_kernels.h"

namespace spv {

template <int DT>
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
    const float v = (float)tex3D<typename TexelType<DT>::type>(V.pt, (float)(x0 + i) + 0.5f, (float)(y0 + j) + 0.5f,
                                                                 (float)(z0 + k) + 0.5f);
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
                                int cgy, int cgz, float *minmax, cudaStream_t st) {
  dim3 grid(vol.gx, vol.gy, vol.gz);
  switch (dtype) {
    case 0: brick_kernel<0><<<grid, 128, 0, st>>>(vol, local_nz, bricks); break;
    case 1: brick_kernel<1><<<grid, 128, 0, st>>>(vol, local_nz, bricks); break;
    default: brick_kernel<2><<<grid, 128, 0, st>>>(vol, local_nz, bricks); break;
  }
  const int nc = cgx * cgy * cgz;
  coarse_kernel<<<(nc + 127) / 128, 128, 0, st>>>(bricks, vol.gx, vol.gy, vol.gz, coarse, cgx, cgy, cgz);
  minmax_kernel<<<1, 1024, 0, st>>>(coarse, nc, minmax);
  return cudaGetLastError();
}

}  // namespace spv
