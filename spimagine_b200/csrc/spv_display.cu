// spv_display.cu -- display hand-off.  Replaces what the reference does on the way from the result buffers to the
// screen: buf.get() of the float planes (volumerender.py:388-390), the re-upload as GL textures
// (gui/gui_utils.py:121-162, gui/glwidget.py:412-444) and the colour look-up of gui/shaders/texture.frag:8-38.
// One pass turns the value plane (+ the alpha plane's sign) into packed RGBA8, a quarter of the bytes of the two
// float planes on the PCIe link.
//
// texture.frag per fragment:   col = texture(value plane).x    (a GL_RED upload: col = (v, 0, 0), clamped to [0,1])
//                              lut = texture_LUT(col.xy) in black mode, texture_LUT(1 - col.xy) otherwise
//                              frag = (lut.rgb, length(col.xyz)) ;  tnear < 0 -> (0, 0, 0, 0)
// with the LUT a 1 x N RGB texture under GL_LINEAR / CLAMP_TO_EDGE (gui_utils.py:136-145): texel coordinate
// u = s N - 1/2, i0 = floor(u), f = u - i0, both indices clamped, (1-f) lut[i0] + f lut[i1].  The 8-bit result is
// rint(255 x) per channel.  All in fp32, -fmad=false, so a plain fp32 host evaluation gives the same bytes.
#include "spv_kernels.h"

namespace spv {

__device__ __forceinline__ unsigned to_u8(float x) {
  return (unsigned)__float2int_rn(255.f * fminf(fmaxf(x, 0.f), 1.f));
}

__global__ void __launch_bounds__(256) display_kernel(const float *__restrict__ value, const float *__restrict__ alpha,
                                                      const float *__restrict__ lut, int n_lut, int mode_black,
                                                      uchar4 *__restrict__ rgba, size_t n) {
  const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  float v = value[p];
  v = fminf(fmaxf(v, 0.f), 1.f);  // unorm texture storage clamps; NaN -> 0
  if (!(v == v)) v = 0.f;
  const float s = mode_black ? v : 1.f - v;
  const float u = s * (float)n_lut - 0.5f;
  const float fl = floorf(u);
  const float f = u - fl;
  const int i0 = min(max((int)fl, 0), n_lut - 1), i1 = min(max((int)fl + 1, 0), n_lut - 1);
  const float w0 = 1.f - f;
  uchar4 o;
  o.x = (unsigned char)to_u8(w0 * lut[3 * i0 + 0] + f * lut[3 * i1 + 0]);
  o.y = (unsigned char)to_u8(w0 * lut[3 * i0 + 1] + f * lut[3 * i1 + 1]);
  o.z = (unsigned char)to_u8(w0 * lut[3 * i0 + 2] + f * lut[3 * i1 + 2]);
  o.w = (unsigned char)to_u8(v);
  if (alpha[p] < 0.f) o = make_uchar4(0, 0, 0, 0);  // texture.frag:37-38 (the float32 kernels mark misses with -1)
  rgba[p] = o;
}

cudaError_t launch_display(const float *value, const float *alpha, const float *lut, int n_lut, int mode_black,
                           void *rgba, size_t n, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  display_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(value, alpha, lut, n_lut, mode_black, (uchar4 *)rgba, n);
  return cudaGetLastError();
}

// columns [xa, xb) of rows [ya, yb) of the first `planes` planes (W x H floats each), same position in dst: the
// rectangle an asynchronous read-back moves into its device staging, which frees the slot's planes for the next render
__global__ void __launch_bounds__(256) rect_copy_kernel(const float *__restrict__ src, float *__restrict__ dst, int W, size_t n, int xa,
                                                        int xb, int ya) {
  const int x = xa + blockIdx.x * 256 + threadIdx.x;
  if (x >= xb) return;
  const size_t p = (size_t)blockIdx.z * n + (size_t)(ya + blockIdx.y) * W + x;
  dst[p] = src[p];
}
cudaError_t launch_rect_copy(const float *src, float *dst, int W, int H, int xa, int xb, int ya, int yb, int planes, cudaStream_t st) {
  if (xa >= xb || ya >= yb || planes < 1) return cudaSuccess;
  if (yb - ya > 65535) return cudaErrorInvalidValue;
  dim3 grid((unsigned)((xb - xa + 255) / 256), (unsigned)(yb - ya), (unsigned)planes);
  rect_copy_kernel<<<grid, 256, 0, st>>>(src, dst, W, (size_t)W * H, xa, xb, ya);
  return cudaGetLastError();
}

}  // namespace spv
