// spv_mip.cu -- maximum-intensity projection kernels (replace max_project_float / max_project_short,
// spimagine/volumerender/kernels/volume_kernel.cl:20-345).
//
//   mip_ref_kernel   the reference's loop structure verbatim (16-unrolled blocks, positions accumulated with
//                    pos += delta, both attenuation laws, the inner-loop-only break): used for the EXACT
//                    sampler and whenever alpha_pow != 0.
//   mip_fast_kernel  alpha_pow == 0 only.  One warp = one 8x4 pixel tile (2x2 quads on consecutive lanes so a
//                    texture quad is spatially compact), sample positions pos0 + k*delta by fma (so samples can
//                    be visited in any order), hardware-filtered tex3D, optional empty-space skipping on the
//                    two-level min/max brick grid, optional slab ownership test for sort-last rendering, results
//                    staged through shared memory and written with 128-bit stores.
#include "spv_kernels.h"

namespace spv {

// -------------------------------------------------------------------------------------------------------------
template <int FMT, bool LINEAR, bool EXACT>
__global__ void __launch_bounds__(128) mip_ref_kernel(const MipArgs a) {
  const unsigned x = blockIdx.x * 16 + (threadIdx.x & 15);
  const unsigned y = blockIdx.y * 8 + (threadIdx.x >> 4);
  const unsigned Nx = a.width, Ny = a.height;
  if (x >= Nx || y >= Ny) return;
  const bool isShort = (FMT % 3) != 0;
  const size_t p = x + (size_t)Nx * y;
  Ray r = make_ray(x, y, Nx, Ny, a.cam, a.box);
  if (!r.hit) {
    a.out[p] = 0.f;
    a.alpha[p] = isShort ? 0.f : -1.f;  // volume_kernel.cl:88 vs :261
    return;
  }
  float tnear = r.tnear, tfar = r.tfar;
  v4 orig = r.orig, direc = r.direc;
  if (tnear < 0.0f) tnear = 0.0f;
  float colVal = 0.f;
  const int reducedSteps = a.max_steps / a.num_parts;
  const int LOOPUNROLL = 16;
  const float dt = fabsf(tfar - tnear) / (float)((reducedSteps / LOOPUNROLL) * LOOPUNROLL);
  orig = add4(orig, scl4((float)a.current_part * dt, direc));
  const v4 delta_pos = scl4(.5f * dt, direc);
  v4 pos = scl4(0.5f, add4(sadd4(1.f, orig), scl4(tnear, direc)));
  float newVal;
  const float minVal = a.min_val, maxVal = a.max_val, alpha_pow = a.alpha_pow;
  if (alpha_pow == 0.f) {
    for (int i = 0; i <= reducedSteps / LOOPUNROLL; ++i) {
#pragma unroll 4
      for (int j = 0; j < LOOPUNROLL; ++j) {
        newVal = sample<FMT, LINEAR, EXACT>(a.vol, pos.x, pos.y, pos.z);
        colVal = fmaxf(colVal, newVal);
        pos = add4(pos, delta_pos);
      }
    }
    colVal = (maxVal == 0.f) ? colVal : (colVal - minVal) / (maxVal - minVal);
  } else {
    float cumsum = 1.f;
    for (int i = 0; i <= reducedSteps / LOOPUNROLL; ++i) {
      for (int j = 0; j < LOOPUNROLL; ++j) {
        newVal = sample<FMT, LINEAR, EXACT>(a.vol, pos.x, pos.y, pos.z);
        newVal = (maxVal == 0.f) ? newVal : (newVal - minVal) / (maxVal - minVal);
        colVal = fmaxf(colVal, cumsum * newVal);
        if (isShort) cumsum *= (1.f - .1f * alpha_pow * alpha_pow * newVal);         // :312
        else cumsum *= (1.f - alpha_pow * alpha_pow * clampf_cl(newVal, 0.f, 1.f));  // :146
        pos = add4(pos, delta_pos);
        if (cumsum <= 0.01f) break;  // leaves the inner loop only, as in the reference
      }
    }
  }
  if (a.gamma != 1.f) colVal = powf(colVal, a.gamma);
  colVal = clampf_cl(colVal, 0.f, 1.f);
  const float alphaVal = isShort ? tnear : 1.f;  // :329 vs :168
  if (a.current_part == 0) {
    a.out[p] = colVal;
    a.alpha[p] = alphaVal;
  } else {
    a.out[p] = fmaxf(colVal, a.out[p]);
    a.alpha[p] = fmaxf(alphaVal, a.alpha[p]);
  }
}

// -------------------------------------------------------------------------------------------------------------
// Fast path.
struct Marcher {
  float u0, v0, w0, du, dv, dw;  // unnormalised texel coordinates of sample k: u0 + k*du
};

template <int FMT, bool LINEAR, bool SLAB>
__device__ __forceinline__ float fetch_k(const Volume &V, const Marcher &m, float k) {
  return sample_tmu_uvw<FMT, LINEAR>(V, fmaf(k, m.du, m.u0), fmaf(k, m.dv, m.v0), fmaf(k, m.dw, m.w0));
}

// does sample k belong to this slab?  (its footprint starts in global slices [z0, z1))
__device__ __forceinline__ bool owns_k(const Volume &V, const Marcher &m, float k) {
  float wb = fmaf(k, m.dw, m.w0) - 0.5f;
  float kf = fminf(fmaxf(floorf(wb), 0.f), (float)(V.nz - 1));
  return kf >= (float)V.z0 && kf < (float)V.z1;
}

// [ka, kb) = the samples k in [0, S) this slab owns
__device__ __forceinline__ void owned_interval(const Volume &V, const Marcher &m, int S, int &ka, int &kb) {
  owned_interval_w(V, m.w0, m.dw, S, ka, kb);
}

// clamped brick coordinate of a texel-centre coordinate c = u - 0.5
__device__ __forceinline__ int brick_coord(float c, int g, int shift) {
  // floor, then arithmetic shift; clamp handles everything outside (incl. huge/NaN -> 0)
  float f = fminf(fmaxf(floorf(c), -1.f), 2147483000.f);
  int i = (int)f >> shift;
  return min(max(i, 0), g - 1);
}

// Results of one warp tile: staged through shared memory and written as 128-bit stores (8 per plane per warp); tiles cut
// by the image edge and later parts of a multi-pass render (fmax merge, volume_kernel.cl:172-182) store per pixel.
__device__ __forceinline__ void store_tile(const MipArgs &a, float *dst_rows, float outVal, float alphaVal, int warp, int lane,
                                           int lx, int ly, unsigned x, unsigned tx0, unsigned ty0, bool inb,
                                           float (*s_out)[32], float (*s_alpha)[32]) {
  const unsigned Nx = a.width, Ny = a.height;
  float *alpha_rows = a.alpha + (size_t)ty0 * Nx;
  const bool vec_ok = (Nx % 4 == 0) && (tx0 + 8 <= Nx) && (ty0 + 4 <= Ny) && a.current_part == 0;
  if (vec_ok) {
    s_out[warp][ly * 8 + lx] = outVal;
    s_alpha[warp][ly * 8 + lx] = alphaVal;
    __syncwarp();
    // 8 float4 per plane: lanes 0-7 store the value plane, lanes 8-15 the alpha plane
    if (lane < 16) {
      const int q = lane & 7, row = q >> 1, half = q & 1;
      const float *src = (lane < 8 ? s_out[warp] : s_alpha[warp]) + row * 8 + half * 4;
      float *base = lane < 8 ? dst_rows : alpha_rows;
      *reinterpret_cast<float4 *>(base + (size_t)row * Nx + tx0 + half * 4) = *reinterpret_cast<const float4 *>(src);
    }
  } else if (inb) {
    const size_t p = x + (size_t)Nx * ly;
    if (a.current_part == 0) {
      dst_rows[p] = outVal;
      alpha_rows[p] = alphaVal;
    } else {  // multi-pass rendering: merge into what earlier parts left (volume_kernel.cl:172-182)
      dst_rows[p] = fmaxf(outVal, dst_rows[p]);
      alpha_rows[p] = fmaxf(alphaVal, alpha_rows[p]);
    }
  }
}

// One CTA tile: TX x TY warps cover (8*TX) x (4*TY) pixels starting at CTA tile (bx, by).
template <int FMT, bool LINEAR, bool SKIP, bool SLAB, int TX, int TY>
__device__ __forceinline__ void mip_fast_tile(const MipArgs &a, unsigned bx, unsigned by, float (*s_out)[32],
                                              float (*s_alpha)[32]) {
  const bool STATS = a.stats != nullptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int lx = (lane & 1) | ((lane >> 1) & 2) | ((lane >> 2) & 4);
  const int ly = ((lane >> 1) & 1) | ((lane >> 2) & 2);
  const unsigned tx0 = bx * (8 * TX) + (warp % TX) * 8, ty0 = by * (4 * TY) + (warp / TX) * 4;
  const unsigned x = tx0 + lx, y = ty0 + ly;
  const unsigned Nx = a.width, Ny = a.height;
  const bool inb = x < Nx && y < Ny;
  const bool isShort = (FMT % 3) != 0;
  const Volume &V = a.vol;

  Ray r = make_ray(x, y, Nx, Ny, a.cam, a.box);
  const bool hit = inb && r.hit;
  float tnear = r.tnear;
  if (tnear < 0.0f) tnear = 0.0f;
  float cur = 0.f;
  unsigned nfetch = 0;

  if (hit) {
    const int reducedSteps = a.max_steps / a.num_parts;
    const int S = (reducedSteps / 16 + 1) * 16;
    const float dt = fabsf(r.tfar - tnear) / (float)((reducedSteps / 16) * 16);
    v4 orig = add4(r.orig, scl4((float)a.current_part * dt, r.direc));
    const v4 delta_pos = scl4(.5f * dt, r.direc);
    const v4 pos0 = scl4(0.5f, add4(sadd4(1.f, orig), scl4(tnear, r.direc)));
    Marcher m;
    m.u0 = pos0.x * V.fnx; m.v0 = pos0.y * V.fny; m.w0 = pos0.z * V.fnz;
    m.du = delta_pos.x * V.fnx; m.dv = delta_pos.y * V.fny; m.dw = delta_pos.z * V.fnz;

    if (!SKIP) {
      if (!SLAB) {
        for (int k = 0; k < S; k += 16) {
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = fetch_k<FMT, LINEAR, SLAB>(V, m, (float)(k + j));
#pragma unroll
          for (int j = 0; j < 16; ++j) cur = fmaxf(cur, v[j]);
        }
        if (STATS) nfetch += S;
      } else {
        // The slice index floor(w(k) - 1/2) is monotone in k (fma, subtraction and floor all are), so the samples
        // a slab owns form one interval [ka, kb): two binary searches with the exact predicate, then march it.
        // Several slabs of the same volume can be resident on this GPU (a.extra): same ray, one interval each.
        for (int sl = 0; sl <= a.n_extra; ++sl) {
          const Volume &VS = sl == 0 ? V : a.extra[sl - 1];
          int ka, kb;
          owned_interval(VS, m, S, ka, kb);
          for (int k = ka; k < kb; k += 8) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = (k + j < kb) ? fetch_k<FMT, LINEAR, SLAB>(VS, m, (float)(k + j)) : 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) cur = fmaxf(cur, v[j]);
          }
          if (STATS) nfetch += (unsigned)(kb - ka);
        }
      }
    } else {
      // ---- pre-pass: every PRE-th sample gives a true lower bound of the ray maximum cheaply ----
      constexpr int PRE = 8;
      for (int k = 0; k < S; k += PRE * 8) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float kk = (float)(k + j * PRE);
          const bool ok = (k + j * PRE < S) && (!SLAB || owns_k(V, m, kk));
          v[j] = ok ? fetch_k<FMT, LINEAR, SLAB>(V, m, kk) : 0.f;
          if (STATS) nfetch += ok;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) cur = fmaxf(cur, v[j]);
      }
      // ---- coarse DDA over cells of CBRICK texels; descend to per-sample brick tests where needed ----
      constexpr int CSHIFT = BRICK_SHIFT + 2;
      constexpr float CB = (float)(1 << CSHIFT);
      // texel-centre coordinates c(k) = (u0 - .5) + k*du ; local z for the brick grid
      const float cx0 = m.u0 - 0.5f, cy0 = m.v0 - 0.5f, cz0 = m.w0 - 0.5f - (float)V.z_lo;
      int ix = (int)floorf(cx0 / CB), iy = (int)floorf(cy0 / CB), iz = (int)floorf(cz0 / CB);
      const float inf = __int_as_float(0x7f800000);
      const int sx = m.du > 0.f ? 1 : -1, sy = m.dv > 0.f ? 1 : -1, sz = m.dw > 0.f ? 1 : -1;
      const float tdx = m.du != 0.f ? CB / fabsf(m.du) : inf;
      const float tdy = m.dv != 0.f ? CB / fabsf(m.dv) : inf;
      const float tdz = m.dw != 0.f ? CB / fabsf(m.dw) : inf;
      float tmx = m.du != 0.f ? ((float)(ix + (sx > 0)) * CB - cx0) / m.du : inf;
      float tmy = m.dv != 0.f ? ((float)(iy + (sy > 0)) * CB - cy0) / m.dv : inf;
      float tmz = m.dw != 0.f ? ((float)(iz + (sz > 0)) * CB - cz0) / m.dw : inf;
      int k0 = 0;
      const int max_iter = 3 * S + a.cgx + a.cgy + a.cgz + 16;  // NaN/degenerate rays cannot spin here
      for (int it = 0; it < max_iter && k0 < S; ++it) {
        const float tout = fminf(fminf(tmx, tmy), tmz);
        // samples k < tout lie in this coarse cell (to within the dilation slack)
        int k1 = tout >= (float)S ? S : (int)ceilf(tout);
        k1 = min(max(k1, k0), S);
        if (k1 > k0) {
          const int cxi = min(max(ix, 0), a.cgx - 1), cyi = min(max(iy, 0), a.cgy - 1), czi = min(max(iz, 0), a.cgz - 1);
          const float cmax = __ldg(a.coarse + ((size_t)czi * a.cgy + cyi) * a.cgx + cxi).y;
          if (cmax > cur) {
            for (int k = k0; k < k1; k += 4) {
              float v[4];
              bool need[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float kk = (float)(k + j);
                need[j] = false;
                if (k + j < k1 && (!SLAB || owns_k(V, m, kk))) {
                  const int bx = brick_coord(fmaf(kk, m.du, cx0), V.gx, BRICK_SHIFT);
                  const int by = brick_coord(fmaf(kk, m.dv, cy0), V.gy, BRICK_SHIFT);
                  const int bz = brick_coord(fmaf(kk, m.dw, cz0), V.gz, BRICK_SHIFT);
                  need[j] = brick_at(V, bx, by, bz).y > cur;
                }
              }
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                v[j] = need[j] ? fetch_k<FMT, LINEAR, SLAB>(V, m, (float)(k + j)) : 0.f;
                if (STATS) nfetch += need[j];
              }
#pragma unroll
              for (int j = 0; j < 4; ++j) cur = fmaxf(cur, v[j]);
            }
          }
        }
        k0 = k1;
        // advance to the next coarse cell along the axis whose boundary comes first
        if (tmx <= tmy && tmx <= tmz) { ix += sx; tmx += tdx; }
        else if (tmy <= tmz) { iy += sy; tmy += tdy; }
        else { iz += sz; tmz += tdz; }
      }
      for (int k = k0; k < S; ++k) {  // only reached when the traversal was cut short: sample the rest plainly
        if (!SLAB || owns_k(V, m, (float)k)) {
          cur = fmaxf(cur, fetch_k<FMT, LINEAR, SLAB>(V, m, (float)k));
          if (STATS) ++nfetch;
        }
      }
    }
  }

  if (STATS) {
    unsigned nf = nfetch, nh = hit ? 1u : 0u;
    for (int o = 16; o > 0; o >>= 1) {
      nf += __shfl_down_sync(0xffffffffu, nf, o);
      nh += __shfl_down_sync(0xffffffffu, nh, o);
    }
    if (lane == 0) {
      atomicAdd(a.stats + 0, (unsigned long long)nh);
      atomicAdd(a.stats + 1, (unsigned long long)nf);
    }
  }

  // ---- epilogue ----
  const float alphaVal = hit ? (isShort ? tnear : 1.f) : (isShort ? 0.f : -1.f);
  float outVal;
  float *dst_rows;  // start of image row ty0 in the destination plane
  if (a.flags & SPV_MIP_RAW_ONLY) {
    outVal = hit ? cur : -1.f;  // un-windowed partial maximum (>= 0), -1 marks a miss; composited across GPUs
                                // with max before the window is applied
    if (a.flags & SPV_MIP_PUSH) {  // straight into the staging of the band's owner (peer memory over NVLink)
      const unsigned o = ty0 / (unsigned)a.push.band_rows;
      dst_rows = a.push.part[o] + a.push.src_off + (size_t)(ty0 - o * (unsigned)a.push.band_rows) * Nx;
    } else {
      dst_rows = a.raw + (size_t)ty0 * Nx;
    }
  } else {
    outVal = hit ? window_value(cur, a.min_val, a.max_val, a.gamma) : 0.f;
    dst_rows = a.out + (size_t)ty0 * Nx;
  }
  store_tile(a, dst_rows, outVal, alphaVal, warp, lane, lx, ly, x, tx0, ty0, inb, s_out, s_alpha);
}

// Static grid: one CTA per tile.
template <int FMT, bool LINEAR, bool SKIP, bool SLAB, int TX, int TY, int MINB>
__global__ void __launch_bounds__(32 * TX * TY, MINB) mip_fast_kernel(const MipArgs a) {
  __shared__ __align__(16) float s_out[TX * TY][32];
  __shared__ __align__(16) float s_alpha[TX * TY][32];
  unsigned by = blockIdx.y + a.y_begin / (4 * TY);
  // Read-back overlap (spv_render_mip_to_host): tile rows are dealt from the top and bottom edges inwards.  The rows whose
  // rays miss the volume (typically the outer ones) are done at once and travel while the rest renders, and the band
  // that completes last is a single one in the middle instead of the whole lower part of the image.
  // row_mode 1: the rows the projected box cannot touch first (top part, then bottom part), then the rows of the box
  // top to bottom -- one front moving through the volume (L2 locality of the plain order) and one band completing last.
  if (a.band_done) {
    if (a.row_mode == 1) {
      const unsigned ta = a.hit_tile_a, tb = a.hit_tile_b, nout = gridDim.y - (tb - ta);
      by = blockIdx.y < nout ? (blockIdx.y < ta ? blockIdx.y : tb + (blockIdx.y - ta)) : ta + (blockIdx.y - nout);
    } else {
      by = (blockIdx.y & 1u) ? gridDim.y - 1u - (blockIdx.y >> 1) : (blockIdx.y >> 1);
    }
  }
  mip_fast_tile<FMT, LINEAR, SKIP, SLAB, TX, TY>(a, blockIdx.x, by, s_out, s_alpha);
  if (a.band_done) {  // this CTA's rows are stored: tell the copy stream
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(a.band_done + (by * (4 * TY)) / (unsigned)a.band_rows, 1u);
    }
  }
}

// Persistent grid: (SMs x resident CTAs) CTAs pull tiles from a counter, so rays that miss the box or leave it
// early do not leave SMs idle behind a static schedule's last wave.  Tiles are handed out in row-major order:
// the CTAs in flight then cover a band of image rows, i.e. a slab of the volume that stays in L2.
template <int FMT, bool LINEAR, bool SKIP, bool SLAB, int TX, int TY, int MINB>
__global__ void __launch_bounds__(32 * TX * TY, MINB) mip_fast_persistent_kernel(const MipArgs a) {
  __shared__ __align__(16) float s_out[TX * TY][32];
  __shared__ __align__(16) float s_alpha[TX * TY][32];
  __shared__ unsigned s_tile;
  const unsigned tiles_x = (a.width + 8 * TX - 1) / (8 * TX);
  const unsigned tiles_y = (a.y_end - a.y_begin + 4 * TY - 1) / (4 * TY), ty0 = a.y_begin / (4 * TY);
  const unsigned ntiles = tiles_x * tiles_y;
  for (;;) {
    if (threadIdx.x == 0) s_tile = atomicAdd(a.tile_counter, 1u);
    __syncthreads();
    const unsigned t = s_tile;
    __syncthreads();
    if (t >= ntiles) break;
    mip_fast_tile<FMT, LINEAR, SKIP, SLAB, TX, TY>(a, t % tiles_x, ty0 + t / tiles_x, s_out, s_alpha);
  }
}

// -------------------------------------------------------------------------------------------------------------
// alpha_pow != 0 (volume_kernel.cl:134-158 float, :300-318 short): front-to-back attenuation.
//   v = (s - min) / (max - min);  col = max(col, cum * v);  cum *= 1 - a^2 clamp(v, 0, 1)  |  1 - 0.1 a^2 v  (short)
// in blocks of 16 samples; `if (cum <= .01) break` leaves the INNER loop only, so a ray that has gone dark still takes
// one sample at the start of each remaining block -- and every executed sample advances the position by one step, so the
// n-th executed sample sits at pos0 + n * delta whatever the block structure did.
// The recurrence is serial, the fetches are not: a block's 16 samples are fetched together (positions by fma, like
// mip_fast_kernel) and the recurrence runs over the register batch.  A ray that enters a block dark fetches one sample
// first and the other 15 only if that sample brought it back above the threshold (cum can grow again for short volumes:
// v < 0 below minVal).
template <int FMT, bool LINEAR>
__global__ void __launch_bounds__(128) mip_alpha_kernel(const MipArgs a) {
  __shared__ __align__(16) float s_out[4][32];
  __shared__ __align__(16) float s_alpha[4][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int lx = (lane & 1) | ((lane >> 1) & 2) | ((lane >> 2) & 4);
  const int ly = ((lane >> 1) & 1) | ((lane >> 2) & 2);
  const unsigned tx0 = blockIdx.x * 16 + (warp & 1) * 8, ty0 = blockIdx.y * 8 + (warp >> 1) * 4;
  const unsigned x = tx0 + lx, y = ty0 + ly;
  const unsigned Nx = a.width, Ny = a.height;
  const bool inb = x < Nx && y < Ny;
  const bool isShort = (FMT % 3) != 0;
  const Volume &V = a.vol;
  Ray r = make_ray(x, y, Nx, Ny, a.cam, a.box);
  const bool hit = inb && r.hit;
  float tnear = r.tnear;
  if (tnear < 0.0f) tnear = 0.0f;
  float col = 0.f;
  if (hit) {
    const int reducedSteps = a.max_steps / a.num_parts;
    const int nblocks = reducedSteps / 16 + 1;
    const float dt = fabsf(r.tfar - tnear) / (float)((reducedSteps / 16) * 16);
    const v4 orig = add4(r.orig, scl4((float)a.current_part * dt, r.direc));
    const v4 delta_pos = scl4(.5f * dt, r.direc);
    const v4 pos0 = scl4(0.5f, add4(sadd4(1.f, orig), scl4(tnear, r.direc)));
    Marcher m;
    m.u0 = pos0.x * V.fnx; m.v0 = pos0.y * V.fny; m.w0 = pos0.z * V.fnz;
    m.du = delta_pos.x * V.fnx; m.dv = delta_pos.y * V.fny; m.dw = delta_pos.z * V.fnz;
    const float minVal = a.min_val, maxVal = a.max_val;
    const float att = isShort ? .1f * a.alpha_pow * a.alpha_pow : a.alpha_pow * a.alpha_pow;  // :312 / :146
    float cum = 1.f;
    int n = 0;  // samples executed so far
    for (int b = 0; b < nblocks; ++b) {
      float v[16];
      int j0 = 0;
      bool done = false;
      if (cum <= 0.01f) {  // dark on entry: one sample decides whether the block goes on
        float s = fetch_k<FMT, LINEAR, false>(V, m, (float)n);
        s = (maxVal == 0.f) ? s : (s - minVal) / (maxVal - minVal);
        col = fmaxf(col, cum * s);
        cum *= isShort ? (1.f - att * s) : (1.f - att * clampf_cl(s, 0.f, 1.f));
        ++n;
        j0 = 1;
        done = cum <= 0.01f;
      }
      if (!done) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = (j >= j0) ? fetch_k<FMT, LINEAR, false>(V, m, (float)(n + j - j0)) : 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (j >= j0 && !done) {
            const float s = (maxVal == 0.f) ? v[j] : (v[j] - minVal) / (maxVal - minVal);
            col = fmaxf(col, cum * s);
            cum *= isShort ? (1.f - att * s) : (1.f - att * clampf_cl(s, 0.f, 1.f));
            ++n;
            done = cum <= 0.01f;
          }
        }
      }
    }
    if (a.gamma != 1.f) col = powf(col, a.gamma);
    col = clampf_cl(col, 0.f, 1.f);
  }
  const float alphaVal = hit ? (isShort ? tnear : 1.f) : (isShort ? 0.f : -1.f);
  store_tile(a, a.out + (size_t)ty0 * Nx, hit ? col : 0.f, alphaVal, warp, lane, lx, ly, x, tx0, ty0, inb, s_out, s_alpha);
}

// window + gamma of the composited raw maximum (sort-last renders)
__global__ void mip_finish_kernel(const float *raw, float *out, int n, float minVal, float maxVal, float gamma) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = raw[i];  // -1 marks a miss (every GPU agrees: the box test does not depend on the slab)
  out[i] = v < 0.f ? 0.f : window_value(v, minVal, maxVal, gamma);
}

// -------------------------------------------------------------------------------------------------------------
template <int FMT, bool LINEAR, bool SKIP, bool SLAB, int TX, int TY, int MINB>
static void launch_fast_shape(const MipArgs &a, cudaStream_t st) {
  dim3 grid((a.width + 8 * TX - 1) / (8 * TX), (a.y_end - a.y_begin + 4 * TY - 1) / (4 * TY)), block(32 * TX * TY);
  if (a.tile_counter) {
    static int resident = 0;  // CTAs per SM x SMs for this instantiation (queried once)
    if (resident == 0) {
      int per_sm = 0, dev = 0, sms = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mip_fast_persistent_kernel<FMT, LINEAR, SKIP, SLAB, TX, TY, MINB>,
                                                    32 * TX * TY, 0);
      resident = (per_sm > 0 ? per_sm : 1) * (sms > 0 ? sms : 1);
    }
    const unsigned ntiles = grid.x * grid.y;
    cudaMemsetAsync(a.tile_counter, 0, sizeof(unsigned), st);
    mip_fast_persistent_kernel<FMT, LINEAR, SKIP, SLAB, TX, TY, MINB>
        <<<ntiles < (unsigned)resident ? ntiles : (unsigned)resident, block, 0, st>>>(a);
  } else {
    mip_fast_kernel<FMT, LINEAR, SKIP, SLAB, TX, TY, MINB><<<grid, block, 0, st>>>(a);
  }
}

template <int FMT, bool LINEAR>
static cudaError_t launch_fast(const MipArgs &a, bool skip, bool slab, cudaStream_t st) {
  if (skip) {
    if (slab) launch_fast_shape<FMT, LINEAR, true, true, 2, 2, 1>(a, st);
    else launch_fast_shape<FMT, LINEAR, true, false, 2, 2, 1>(a, st);
  } else if (slab) {
    launch_fast_shape<FMT, LINEAR, false, true, 2, 2, 1>(a, st);
  } else {
    switch (a.tile_variant) {  // CTA shape / occupancy target: tuning knob (spv_set_tuning)
      case 1: launch_fast_shape<FMT, LINEAR, false, false, 2, 4, 1>(a, st); break;   // 16x16 px, 256 threads
      case 2: launch_fast_shape<FMT, LINEAR, false, false, 2, 2, 12>(a, st); break;  // 16x8 px, <= 40 registers
      case 3: launch_fast_shape<FMT, LINEAR, false, false, 2, 2, 16>(a, st); break;  // 16x8 px, <= 32 registers
      case 4: launch_fast_shape<FMT, LINEAR, false, false, 2, 1, 16>(a, st); break;  // 16x4 px, 64 threads
      // finer launch granularity at the default register budget: the last wave of box-hitting CTAs is shorter
      case 5: launch_fast_shape<FMT, LINEAR, false, false, 2, 1, 1>(a, st); break;   // 16x4 px, 64 threads
      case 6: launch_fast_shape<FMT, LINEAR, false, false, 1, 1, 1>(a, st); break;   // 8x4 px, one warp
      case 7: launch_fast_shape<FMT, LINEAR, false, false, 1, 2, 1>(a, st); break;   // 8x8 px, 64 threads
      default: launch_fast_shape<FMT, LINEAR, false, false, 2, 2, 1>(a, st); break;  // 16x8 px, 128 threads
    }
  }
  return cudaGetLastError();
}

template <int FMT, bool LINEAR>
static cudaError_t launch_ref(const MipArgs &a, bool exact, cudaStream_t st) {
  dim3 grid((a.width + 15) / 16, (a.height + 7) / 8), block(128);
  if (exact) mip_ref_kernel<FMT, LINEAR, true><<<grid, block, 0, st>>>(a);
  else mip_ref_kernel<FMT, LINEAR, false><<<grid, block, 0, st>>>(a);
  return cudaGetLastError();
}

template <int FMT>
static cudaError_t launch_dt(const MipArgs &a, bool linear, bool fast, bool exact, bool skip, bool slab, bool stats,
                             cudaStream_t st) {
  if (fast && a.alpha_pow != 0.f) {  // attenuated: batched fetches, serial recurrence
    dim3 grid((a.width + 15) / 16, (a.height + 7) / 8), block(128);
    if (linear) mip_alpha_kernel<FMT, true><<<grid, block, 0, st>>>(a);
    else mip_alpha_kernel<FMT, false><<<grid, block, 0, st>>>(a);
    return cudaGetLastError();
  }
  if (fast) return linear ? launch_fast<FMT, true>(a, skip, slab, st) : launch_fast<FMT, false>(a, skip, slab, st);
  return linear ? launch_ref<FMT, true>(a, exact, st) : launch_ref<FMT, false>(a, exact, st);
}

cudaError_t launch_mip(const MipArgs &a, int dtype, bool linear, bool fast, bool exact, bool skip, bool slab,
                       bool stats, cudaStream_t st) {
  switch (dtype) {  // FMT = dtype + 3 * layout
    case 0: return launch_dt<0>(a, linear, fast, exact, skip, slab, stats, st);
    case 1: return launch_dt<1>(a, linear, fast, exact, skip, slab, stats, st);
    case 2: return launch_dt<2>(a, linear, fast, exact, skip, slab, stats, st);
    case 4: return launch_dt<4>(a, linear, fast, exact, skip, slab, stats, st);
    case 5: return launch_dt<5>(a, linear, fast, exact, skip, slab, stats, st);
    default: return cudaErrorInvalidValue;
  }
}

// Direct sampler access for tests: out[i] = sample(volume, pos[i]) with pos in normalised coordinates, exactly
// as the render kernels sample (TMU or exact sampler).
template <int FMT, bool LINEAR, bool EXACT>
__global__ void sample_points_kernel(const Volume V, const float *__restrict__ pos, int n, float *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = sample<FMT, LINEAR, EXACT>(V, pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
}

cudaError_t launch_sample_points(const Volume &V, int dtype, bool linear, bool exact, const float *pos, int n,
                                 float *out, cudaStream_t st) {
  const int blocks = (n + 255) / 256;
#define SPV_SP(FMT) \
  if (linear) { if (exact) sample_points_kernel<FMT, true, true><<<blocks, 256, 0, st>>>(V, pos, n, out); \
                else sample_points_kernel<FMT, true, false><<<blocks, 256, 0, st>>>(V, pos, n, out); } \
  else { if (exact) sample_points_kernel<FMT, false, true><<<blocks, 256, 0, st>>>(V, pos, n, out); \
         else sample_points_kernel<FMT, false, false><<<blocks, 256, 0, st>>>(V, pos, n, out); }
  switch (dtype) {
    case 0: SPV_SP(0) break;
    case 1: SPV_SP(1) break;
    case 2: SPV_SP(2) break;
    case 4: SPV_SP(4) break;
    case 5: SPV_SP(5) break;
    default: return cudaErrorInvalidValue;
  }
#undef SPV_SP
  return cudaGetLastError();
}

// Calibration probe for the texture-sample roofline: every lane issues independent filtered fetches inside a
// small (cache-resident) region of the resident volume, with no dependent arithmetic between them.
// Lane (lx, ly) of the 8x4 tile samples at base + lx*a + ly*b + j*m (texels; j = 0..15 independent fetches in flight).
// The default vectors (a = .6 x, b = .6 y, m mostly along z, 8 slices deep) keep a 2x2 quad's footprint inside one
// cache line most of the time: the unit's peak.  With the benchmark camera's vectors (rays 1.15 texels apart, samples
// 1.6 texels apart along the view direction) it measures what the unit can deliver for THAT footprint when every
// fetch hits L1: all warps of all CTAs walk the same small region.
struct ProbeVecs {
  float a[3], b[3], m[3];
  int wrap8;  // z advances with (j & 7): the default probe
};
template <int FMT, bool LINEAR>
__global__ void __launch_bounds__(256) texrate_probe_kernel(const Volume V, int iters, const ProbeVecs pv, float *sink) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int lx = (lane & 1) | ((lane >> 1) & 2) | ((lane >> 2) & 4);
  const int ly = ((lane >> 1) & 1) | ((lane >> 2) & 2);
  const float fx = (float)lx, fy = (float)ly;
  float bx = 0.37f * (float)V.nx + fx * pv.a[0] + fy * pv.b[0];
  float by = 0.41f * (float)V.ny + fx * pv.a[1] + fy * pv.b[1];
  float bz = 0.29f * (float)V.nz + fx * pv.a[2] + fy * pv.b[2];
  if (pv.wrap8) {  // warps and CTAs next to each other, as the first calibration of this round measured it
    bx += 3.1f * (float)(warp & 3);
    by += 2.7f * (float)(warp >> 2);
    bz += 1.9f * (float)(blockIdx.x & 7);
  }
  float acc = 0.f;
  for (int it = 0; it < iters; ++it) {
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float fj = (float)j, fz = pv.wrap8 ? (float)(j & 7) : fj;
      v[j] = sample_tmu_uvw<FMT, LINEAR>(V, bx + pv.m[0] * fj, by + pv.m[1] * fj, bz + pv.m[2] * fz + 0.11f * (float)(it & 3));
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) acc = fmaxf(acc, v[j]);
  }
  if (acc == -12345.f) sink[0] = acc;  // never true: keeps the fetches alive
}

cudaError_t launch_texrate_probe(const Volume &V, int dtype, bool linear, int blocks, int iters, const float *vec9,
                                 float *sink, cudaStream_t st) {
  ProbeVecs pv = {{.6f, 0.f, 0.f}, {0.f, .6f, 0.f}, {.05f, .03f, .9f}, 1};
  if (vec9) {
    for (int i = 0; i < 3; ++i) {
      pv.a[i] = vec9[i];
      pv.b[i] = vec9[3 + i];
      pv.m[i] = vec9[6 + i];
    }
    pv.wrap8 = 0;
  }
#define SPV_PROBE(FMT) \
  if (linear) texrate_probe_kernel<FMT, true><<<blocks, 256, 0, st>>>(V, iters, pv, sink); \
  else texrate_probe_kernel<FMT, false><<<blocks, 256, 0, st>>>(V, iters, pv, sink)
  switch (dtype) {
    case 0: SPV_PROBE(0); break;
    case 1: SPV_PROBE(1); break;
    case 2: SPV_PROBE(2); break;
    case 3: SPV_PROBE(3); break;  // float32 pairs (the layered copies of spv_mip_axis.cu; V.filt is that copy's texture)
    case 4: SPV_PROBE(4); break;
    case 5: SPV_PROBE(5); break;
    default: return cudaErrorInvalidValue;
  }
#undef SPV_PROBE
  return cudaGetLastError();
}

// Load (CUDA loads kernels lazily, at their first launch, and that load can wait for running kernels) every kernel a
// sort-last max projection launches.  A composite's kernels wait for each other across contexts: none of them may be
// loaded for the first time while another one is already spinning (spv_comp_init calls this).
#define SPV_PRELOAD(k)                                         \
  do {                                                         \
    cudaFuncAttributes fa_;                                    \
    cudaError_t e_ = cudaFuncGetAttributes(&fa_, k);           \
    if (e_ != cudaSuccess) return e_;                          \
  } while (0)
template <int FMT>
static cudaError_t preload_mip_fmt() {
  SPV_PRELOAD((mip_fast_kernel<FMT, true, false, true, 2, 2, 1>));
  SPV_PRELOAD((mip_fast_kernel<FMT, false, false, true, 2, 2, 1>));
  SPV_PRELOAD((mip_fast_kernel<FMT, true, true, true, 2, 2, 1>));
  SPV_PRELOAD((mip_fast_kernel<FMT, false, true, true, 2, 2, 1>));
  SPV_PRELOAD((mip_fast_kernel<FMT, true, false, false, 2, 2, 1>));
  SPV_PRELOAD((mip_fast_kernel<FMT, false, false, false, 2, 2, 1>));
  return cudaSuccess;
}
cudaError_t preload_mip_kernels() {
  cudaError_t e;
  if ((e = preload_mip_fmt<0>()) != cudaSuccess) return e;
  if ((e = preload_mip_fmt<1>()) != cudaSuccess) return e;
  if ((e = preload_mip_fmt<2>()) != cudaSuccess) return e;
  if ((e = preload_mip_fmt<4>()) != cudaSuccess) return e;
  if ((e = preload_mip_fmt<5>()) != cudaSuccess) return e;
  SPV_PRELOAD(mip_finish_kernel);
  return cudaSuccess;
}

cudaError_t launch_mip_finish(const float *raw, float *out, int n, float minVal, float maxVal, float gamma,
                              cudaStream_t st) {
  mip_finish_kernel<<<(n + 255) / 256, 256, 0, st>>>(raw, out, n, minVal, maxVal, gamma);
  return cudaGetLastError();
}

}  // namespace spv
