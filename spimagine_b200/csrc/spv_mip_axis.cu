// spv_mip_axis.cu -- max projection through view-aligned layered copies, several frames per launch (round 2).
// Replaces, for plain (alpha_pow == 0) projections of resident integer volumes, the sample loop of
//   max_project_short   spimagine/volumerender/kernels/volume_kernel.cl:185-335   (loop :293-298)
//
// What bounds mip_fast_kernel (profiles/r02_mip_tmu_ncu_summary.json): the L1TEX data stage is 78-82 % busy while an SM
// is active at 0.43-0.48 requests per clock, i.e. a quad request costs 1.6-1.9 data wavefronts, and the frame reads the
// whole z-paired volume from DRAM (four times the L2).  Measured on the benchmark camera (scripts/exp_multiframe.cu,
// profiles/r02_exp_multiframe_v*.txt):
//   * a request is cheapest when its four lanes AND the ray's consecutive samples stay inside one layer of the 2-D
//     layered array.  With layers along z that holds only for views along z.  With layers along the camera's up axis and
//     a quad = four pixels of one image row it holds for every angle of a sweep about that axis: 147 us per frame at
//     every angle instead of 136-172 us;
//   * frames of a sequence that share a launch share the volume in L2: CTAs are dealt (tile row, frame, tile x), the
//     F frames' CTAs of one tile row run together, and the wedge of layers that row's rays cross (~12 MB) is read from
//     DRAM once instead of F times.  20 frames 18 degrees apart, 10 per launch: 95.8 us per frame, 748 Gsamples/s =
//     0.645 of the texture unit's cache-resident rate (mip_fast_kernel, one frame per launch: 159 us, 0.39).
// So: up to three layered copies of the volume (pairs along x, y, z; built on the device from the primary z copy when a
// view first wants them), a per-frame choice of layer axis and lane-to-pixel map (spv_api.cu: choose_axis), and launches
// of up to MAX_BATCH frames.  Ray setup, sample positions and the z-lerp arithmetic are those of mip_fast_kernel; with the
// z copy and 2x2 quads the two kernels produce the same bits.
#include "spv_kernels.h"

namespace spv {

// ---- building a copy ----------------------------------------------------------------------------------------------
// One thread per destination texel, destination x fastest; a block covers 32 texels of a row in 8 consecutive layers, so
// the primary array is read in whole sectors whichever way the copy is turned.
template <int DT>
__global__ void __launch_bounds__(256) axis_pair_kernel(const Volume V, int lax, cudaSurfaceObject_t dst) {
  typedef typename TexelType<DT>::pair pair_t;
  const int dx = blockIdx.x * 32 + (threadIdx.x & 31);      // destination column
  const int dy = blockIdx.y;                                // destination row
  const int dl = blockIdx.z * 8 + (threadIdx.x >> 5);       // destination layer
  // lax 0: (column, row, layer) = (z, y, x), partner x + 1;   lax 1: (x, z, y), partner y + 1;   lax 2: (x, y, z), z + 1
  const int W = lax == 0 ? V.nz : V.nx, L = lax == 0 ? V.nx : (lax == 1 ? V.ny : V.nz);
  if (dx >= W || dl >= L) return;
  const int x = lax == 0 ? dl : dx, y = lax == 0 ? dy : (lax == 1 ? dl : dy), z = lax == 0 ? dx : (lax == 1 ? dy : dl);
  const int x1 = lax == 0 ? min(x + 1, V.nx - 1) : x, y1 = lax == 1 ? min(y + 1, V.ny - 1) : y,
            z1 = lax == 2 ? min(z + 1, V.nz - 1) : z;
  pair_t p;
  if (DT == 0) {  // float32 volumes live in a 3-D single-channel array
    p.x = tex3D<typename TexelType<DT>::type>(V.pt, (float)x + 0.5f, (float)y + 0.5f, (float)z + 0.5f);
    p.y = tex3D<typename TexelType<DT>::type>(V.pt, (float)x1 + 0.5f, (float)y1 + 0.5f, (float)z1 + 0.5f);
  } else {        // integer volumes in the z-paired layered array (lax 2 is that array itself: never built)
    p.x = tex2DLayered<pair_t>(V.pt, (float)x + 0.5f, (float)y + 0.5f, z).x;
    p.y = tex2DLayered<pair_t>(V.pt, (float)x1 + 0.5f, (float)y1 + 0.5f, z1).x;
  }
  surf2DLayeredwrite(p, dst, dx * (int)sizeof(pair_t), dy, dl);
}

cudaError_t launch_axis_pair(const Volume &V, int dtype, int lax, cudaSurfaceObject_t dst, cudaStream_t st) {
  const int W = lax == 0 ? V.nz : V.nx, H = lax == 0 ? V.ny : (lax == 1 ? V.nz : V.ny), L = lax == 0 ? V.nx : (lax == 1 ? V.ny : V.nz);
  const dim3 grid((W + 31) / 32, H, (L + 7) / 8);
  if (dtype == 0) axis_pair_kernel<0><<<grid, 256, 0, st>>>(V, lax, dst);
  else if (dtype == 1) axis_pair_kernel<1><<<grid, 256, 0, st>>>(V, lax, dst);
  else if (dtype == 2) axis_pair_kernel<2><<<grid, 256, 0, st>>>(V, lax, dst);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

// ---- the render kernel --------------------------------------------------------------------------------------------
// one CTA tile (bx, by) of frame f.  ALPHA: front-to-back attenuation (alpha_pow != 0, volume_kernel.cl:300-318) with the
// block structure of mip_alpha_kernel (spv_mip.cu): a block's 16 fetches in flight, the serial recurrence over the batch.
template <int DT, bool ALPHA>
__device__ __forceinline__ void mip_axis_tile(const MipAxisArgs &a, int f, unsigned bx, unsigned by, float (*s_out)[32],
                                              float (*s_alpha)[32]) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned Nx = a.width, Ny = a.height;
  if (bx < a.rend_x0[f] || bx >= a.rend_x1[f] || by < a.rend_y0[f] || by >= a.rend_y1[f]) {
    // outside the rectangle the box can project to: every ray misses (out 0, alpha 0 / -1 for float32), no ray is set up
    const unsigned px0 = bx * 16, py0 = by * 8;
    const float miss_alpha = DT == 0 ? -1.f : 0.f;
    if (Nx % 4 == 0 && px0 + 16 <= Nx && py0 + 8 <= Ny) {
      if (threadIdx.x < 64) {  // 32 float4 per plane
        float *plane = threadIdx.x < 32 ? a.out[f] : a.alpha[f];
        const float m = threadIdx.x < 32 ? 0.f : miss_alpha;
        const unsigned v = threadIdx.x & 31;
        *reinterpret_cast<float4 *>(plane + (size_t)(py0 + (v >> 2)) * Nx + px0 + (v & 3) * 4) = make_float4(m, m, m, m);
      }
    } else {
      const unsigned x = px0 + (threadIdx.x & 15), y = py0 + (threadIdx.x >> 4);
      if (x < Nx && y < Ny) {
        a.out[f][(size_t)y * Nx + x] = 0.f;
        a.alpha[f][(size_t)y * Nx + x] = miss_alpha;
      }
    }
    return;
  }
  // lanes -> pixels.  tw x th: the warp's tile; (lx, ly): this lane's pixel in it; quads are lanes 4i..4i+3.
  const int quad = a.quad[f], q = lane >> 2, i = lane & 3;
  int tw, lx, ly;
  unsigned tx0 = bx * 16, ty0 = by * 8;
  if (quad == 1) {         // 4x1 quads, 16x2 tiles, warps stacked
    tw = 16; lx = (q & 3) * 4 + i; ly = q >> 2;
    ty0 += warp * 2;
  } else if (quad == 2) {  // 1x4 quads, 4x8 tiles, warps side by side
    tw = 4; lx = q & 3; ly = (q >> 2) * 4 + i;
    tx0 += warp * 4;
  } else {                 // 2x2 quads, 8x4 tiles, 2x2 warps
    tw = 8; lx = (lane & 1) | ((lane >> 1) & 2) | ((lane >> 2) & 4); ly = ((lane >> 1) & 1) | ((lane >> 2) & 2);
    tx0 += (warp & 1) * 8; ty0 += (warp >> 1) * 4;
  }
  const int th = 32 / tw;
  const unsigned x = tx0 + lx, y = ty0 + ly;
  const bool inb = x < Nx && y < Ny;

  Ray r = make_ray(x, y, Nx, Ny, a.invP, a.invM[f], a.box);
  const bool hit = inb && r.hit;
  float tnear = r.tnear;
  if (tnear < 0.0f) tnear = 0.0f;
  float cur = 0.f;
  if (hit) {
    const int S = (a.max_steps / 16 + 1) * 16;
    const float dt = fabsf(r.tfar - tnear) / (float)((a.max_steps / 16) * 16);
    const v4 delta_pos = scl4(.5f * dt, r.direc);
    const v4 pos0 = scl4(0.5f, add4(sadd4(1.f, r.orig), scl4(tnear, r.direc)));
    const float fnx = (float)a.nx, fny = (float)a.ny, fnz = (float)a.nz;
    const float u0 = pos0.x * fnx, v0 = pos0.y * fny, w0 = pos0.z * fnz;
    const float du = delta_pos.x * fnx, dv = delta_pos.y * fny, dw = delta_pos.z * fnz;
    // (a, b): coordinates inside a layer, c: along the layer axis
    const int lax = a.lax[f];
    const float a0 = lax == 0 ? w0 : u0, da = lax == 0 ? dw : du;
    const float b0 = lax == 1 ? w0 : v0, db = lax == 1 ? dw : dv;
    const float c0 = lax == 0 ? u0 : (lax == 1 ? v0 : w0), dc = lax == 0 ? du : (lax == 1 ? dv : dw);
    const float top = (float)((lax == 0 ? a.nx : (lax == 1 ? a.ny : a.nz)) - 1);
    const cudaTextureObject_t tex = a.tex[lax];
    // layer l = {v[l], v[l+1]} (the last layer repeats itself): bilinear in the texture unit, lerp along c here
    auto fetch = [&](float kk, float2 &t, float &fr) {
      const float cb = fmaf(kk, dc, c0) - 0.5f;
      const float fl = floorf(cb);
      fr = fl < 0.f ? 0.f : cb - fl;  // below the first slice centre: clamp-to-edge
      const int layer = (int)fminf(fmaxf(fl, 0.f), top);
      t = tex2DLayered<float2>(tex, fmaf(kk, da, a0), fmaf(kk, db, b0), layer);
    };
    if (!ALPHA) {
      for (int k = 0; k < S; k += 16) {
        float2 t[16];
        float fr[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) fetch((float)(k + j), t[j], fr[j]);
#pragma unroll
        for (int j = 0; j < 16; ++j) cur = fmaxf(cur, fmaf(fr[j], t[j].y - t[j].x, t[j].x));
      }
      cur *= a.scale;  // rounding is monotone: max(s * t_k) == s * max(t_k), the value mip_fast_kernel computes
    } else {
      // v = (s - min) / (max - min);  col = max(col, cum * v);  cum *= 1 - 0.1 a^2 v   (integer volumes)
      // `if (cum <= .01) break` leaves the inner loop only: a ray that has gone dark still takes one sample at the start
      // of every remaining block, and the n-th executed sample sits at pos0 + n * delta.
      // (float32 volumes: cum *= 1 - a^2 clamp(v, 0, 1), volume_kernel.cl:146)
      const float minVal = a.min_val, maxVal = a.max_val;
      const float att = DT == 0 ? a.alpha_pow * a.alpha_pow : .1f * a.alpha_pow * a.alpha_pow;
      const int nblocks = a.max_steps / 16 + 1;
      float cum = 1.f, col = 0.f;
      int n = 0;  // samples executed so far
      for (int b = 0; b < nblocks; ++b) {
        float2 t[16];
        float fr[16];
        int j0 = 0;
        bool done = false;
        if (cum <= 0.01f) {  // dark on entry: one sample decides whether the block goes on
          fetch((float)n, t[0], fr[0]);
          float v = fmaf(fr[0], t[0].y - t[0].x, t[0].x) * a.scale;
          v = (maxVal == 0.f) ? v : (v - minVal) / (maxVal - minVal);
          col = fmaxf(col, cum * v);
          cum *= DT == 0 ? 1.f - att * clampf_cl(v, 0.f, 1.f) : 1.f - att * v;
          ++n;
          j0 = 1;
          done = cum <= 0.01f;
        }
        if (!done) {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (j >= j0) fetch((float)(n + j - j0), t[j], fr[j]);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if (j >= j0 && !done) {
              float v = fmaf(fr[j], t[j].y - t[j].x, t[j].x) * a.scale;
              v = (maxVal == 0.f) ? v : (v - minVal) / (maxVal - minVal);
              col = fmaxf(col, cum * v);
              cum *= DT == 0 ? 1.f - att * clampf_cl(v, 0.f, 1.f) : 1.f - att * v;
              ++n;
              done = cum <= 0.01f;
            }
          }
        }
      }
      if (a.gamma != 1.f) col = powf(col, a.gamma);
      cur = clampf_cl(col, 0.f, 1.f);
    }
  }

  // ---- epilogue: window, gamma; the tile goes through shared memory and leaves as 128-bit stores ----
  const float alphaVal = DT == 0 ? (hit ? 1.f : -1.f) : (hit ? tnear : 0.f);  // volume_kernel.cl:176-177 / :333-334
  const float outVal = hit ? (ALPHA ? cur : window_value(cur, a.min_val, a.max_val, a.gamma)) : 0.f;
  float *out_rows = a.out[f] + (size_t)ty0 * Nx, *alpha_rows = a.alpha[f] + (size_t)ty0 * Nx;
  const bool vec_ok = (Nx % 4 == 0) && (tx0 + tw <= Nx) && (ty0 + th <= Ny);
  if (vec_ok) {
    s_out[warp][ly * tw + lx] = outVal;
    s_alpha[warp][ly * tw + lx] = alphaVal;
    __syncwarp();
    if (lane < 16) {  // 8 float4 per plane: lanes 0-7 the value plane, lanes 8-15 the alpha plane
      const int v = lane & 7, per_row = tw >> 2, row = v / per_row, c4 = v - row * per_row;
      const float *src = (lane < 8 ? s_out[warp] : s_alpha[warp]) + row * tw + c4 * 4;
      float *base = lane < 8 ? out_rows : alpha_rows;
      *reinterpret_cast<float4 *>(base + (size_t)row * Nx + tx0 + c4 * 4) = *reinterpret_cast<const float4 *>(src);
    }
  } else if (inb) {
    const size_t p = x + (size_t)Nx * ly;
    out_rows[p] = outVal;
    alpha_rows[p] = alphaVal;
  }
}

template <int DT, bool ALPHA>
__global__ void __launch_bounds__(128) mip_axis_kernel(const __grid_constant__ MipAxisArgs a) {
  __shared__ __align__(16) float s_out[4][32];
  __shared__ __align__(16) float s_alpha[4][32];
  const int f = blockIdx.y;
  if (blockIdx.x >= a.tile_nx[f] || blockIdx.z >= a.tile_ny[f]) return;  // the grid covers the largest frame's tiles
  const unsigned bx = blockIdx.x + a.tile_x0[f];
  unsigned by = blockIdx.z + a.tile_y0[f] + a.y_begin / 8;
  if (a.band_done && a.row_mode < 2) {  // read-back overlap of a single frame: mip_fast_kernel's order of tile rows
    if (a.row_mode == 1) {
      const unsigned ta = a.hit_tile_a, tb = a.hit_tile_b, nout = gridDim.z - (tb - ta);
      by = blockIdx.z < nout ? (blockIdx.z < ta ? blockIdx.z : tb + (blockIdx.z - ta)) : ta + (blockIdx.z - nout);
    } else {
      by = (blockIdx.z & 1u) ? gridDim.z - 1u - (blockIdx.z >> 1) : (blockIdx.z >> 1);
    }
  }
  mip_axis_tile<DT, ALPHA>(a, f, bx, by, s_out, s_alpha);
  if (a.band_done) {  // this CTA's rows are stored: tell the copy streams (bands count from the first launched tile row)
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(a.band_done + ((by - a.tile_y0[f]) * 8) / (unsigned)a.band_rows, 1u);
    }
  }
}

cudaError_t launch_mip_axis(const MipAxisArgs &a, int dtype, cudaStream_t st) {
  unsigned gx = 0, gz = 0;
  for (int f = 0; f < a.n_frames; ++f) {
    gx = a.tile_nx[f] > gx ? a.tile_nx[f] : gx;
    gz = a.tile_ny[f] > gz ? a.tile_ny[f] : gz;
  }
  if (gx == 0 || gz == 0) return cudaSuccess;  // no frame's box is on screen
  const dim3 grid(gx, a.n_frames, gz);
  const bool att = a.alpha_pow != 0.f;
#define SPV_AXIS(DT) \
  do { if (att) mip_axis_kernel<DT, true><<<grid, 128, 0, st>>>(a); else mip_axis_kernel<DT, false><<<grid, 128, 0, st>>>(a); } while (0)
  if (dtype == 0) SPV_AXIS(0);
  else if (dtype == 1) SPV_AXIS(1);
  else if (dtype == 2) SPV_AXIS(2);
  else return cudaErrorInvalidValue;
#undef SPV_AXIS
  return cudaGetLastError();
}

cudaError_t preload_mip_axis() {
  cudaFuncAttributes fa;
  cudaError_t e = cudaFuncGetAttributes(&fa, mip_axis_kernel<0, false>);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, mip_axis_kernel<0, true>);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, mip_axis_kernel<1, false>);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, mip_axis_kernel<1, true>);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, mip_axis_kernel<2, false>);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, mip_axis_kernel<2, true>);
  return e;
}

}  // namespace spv
