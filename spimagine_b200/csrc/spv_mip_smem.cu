// spv_mip_smem.cu -- max projection (max_project_short, spimagine/volumerender/kernels/volume_kernel.cl:270-345, the
// alpha_pow == 0 loop :293-298) with the samples taken in SOFTWARE from shared memory: per-CTA ray-segment slabs of a
// linear uint16 copy of the volume are staged by the tensor memory accelerator (cp.async.bulk.tensor.3d box loads,
// SASS UTMALDG) through a three-stage mbarrier ring, and every sample is an fp32 trilinear blend of eight 16-bit
// shared-memory loads with exact fp32 weights -- no texture instruction on this path.
//
// Why: mip_fast_kernel issues one TEX per sample and is bound by the texture unit's data stage (1.9 data wavefronts per
// quad for this camera's footprint, profiles/r01_mip_ncu_summary_s4.json).  Here the bytes come from L2 / HBM as bulk
// boxes, independent of the rays' access pattern, and the arithmetic runs on the FMA / ALU pipes.
//
// Geometry.  A CTA owns a 16x16 pixel tile (8 consumer warps, each an 8x4 sub-tile like mip_fast_kernel) plus one
// producer warp.  The tile picks its dominant axis D (largest |component| of its centre ray in texel space) and works
// on the copy of the volume that has D as the slowest axis (three permuted copies; A = contiguous axis, B, D).  The
// tile's rays cross D-planes in a fixed order, so the volume is consumed in slabs of SM_T planes: slab s needs planes
// [dlo, dlo + SM_T] (base index and its upper neighbour) and, laterally, the hull of the tile's four corner rays
// between the slab's two bounding planes.  The producer turns that hull into the origin of one SM_A x SM_B x SM_P box
// (clamped so that the box lies inside the volume) and issues the TMA load; consumers march their own samples while the
// base plane index of a sample lies in the slab.
//
// Correctness does not depend on the producer's geometry: a sample whose 2x2x2 neighbourhood is not inside the staged
// box (hull overflow on very oblique views, samples on the volume's boundary layer where clamp-to-edge applies, the 16
// samples beyond tfar once they leave the volume) is fetched through the texture unit exactly like mip_fast_kernel.
#include <cuda.h>

#include "spv_kernels.h"

namespace spv {

// Box and ring geometry (template parameter of the kernel; spv_set_tuning knob 11 picks one)
template <int A_, int B_, int T_, int STAGES_, int MINB_>
struct SmemCfg {
  static constexpr int A = A_, B = B_, T = T_, P = T_ + 1, STAGES = STAGES_, MINB = MINB_;
  static constexpr int STAGE_ELEMS = A * B * P, STAGE_BYTES = STAGE_ELEMS * 2;
};
typedef SmemCfg<40, 32, 8, 3, 3> Cfg0;   // 23040 B per stage
typedef SmemCfg<32, 24, 8, 4, 3> Cfg1;   // 13824 B per stage: tighter box, one more stage
typedef SmemCfg<32, 24, 16, 3, 2> Cfg2;  // 26112 B per stage: thicker slabs (fewer hand-overs per sample)
typedef SmemCfg<32, 24, 4, 6, 3> Cfg3;   //  7680 B per stage: thin slabs, deep ring
typedef SmemCfg<40, 32, 8, 5, 1> Cfg4;   // the first box with a deeper ring (one CTA per SM)
constexpr int SM_NCFG = 5;
constexpr int SM_MAX_STAGES = 6;
constexpr int SM_CWARPS = 8;
constexpr int SM_THREADS = 32 * (SM_CWARPS + 1);
constexpr float SM_MAGIC = 12582912.f;                      // 1.5 * 2^23: x + MAGIC (round down) = MAGIC + floor(x)
constexpr int SM_MAGIC_BITS = 0x4B400000;

struct SlabDesc {
  int baseA, baseB, baseD;  // MAGIC_BITS + box origin: local index = float bits of (coordinate + MAGIC) - base
  unsigned ldmax;           // largest local base plane index a sample of this slab may have; when nothing was staged
                            // baseD is moved out of reach so that no sample passes the range test
};

struct SmemShared {
  unsigned long long full[SM_MAX_STAGES], empty[SM_MAX_STAGES];
  SlabDesc desc[SM_MAX_STAGES];
  int axis, up;        // D and the direction the tile's rays cross its planes
  int dmin, dmax;      // base plane indices the tile's samples span (clamped to the planes that can be staged)
  int anyhit;
  unsigned tile;
  alignas(16) float s_out[SM_CWARPS][32];  // read back as float4
  alignas(16) float s_alpha[SM_CWARPS][32];
};

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded: a wait that cannot end (a faulted bulk copy) becomes a trap, i.e. an error on the host, not a hang
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  for (unsigned spins = 0; !mbar_try_wait(bar, parity); ++spins)
    if (spins > (1u << 24)) __trap();
}
__device__ __forceinline__ void tma_load_3d(unsigned dst, const CUtensorMap *map, int c0, int c1, int c2, unsigned bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
      "l"(reinterpret_cast<unsigned long long>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}
__device__ __forceinline__ unsigned lds_u16(unsigned addr) {
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
  return v;
}

// coordinates of the permuted copy with slowest axis D: (A, B, D) = (z, y, x) | (x, z, y) | (x, y, z)
__device__ __forceinline__ void permute3(int D, float x, float y, float z, float &a, float &b, float &d) {
  a = D == 0 ? z : x;
  b = D == 1 ? z : y;
  d = D == 0 ? x : (D == 1 ? y : z);
}

struct Line {  // sample k of a ray in texel-centre coordinates of the permuted copy: p0 + k * dp
  float a0, b0, d0, da, db, dd;
};

// 2^23 + v as a float: differences of two such numbers are exact, so only one of each pair has to be un-biased
__device__ __forceinline__ float biased(unsigned v) { return __uint_as_float(v + 0x4B000000u); }
__device__ __forceinline__ float lerp_pair(float w, float A, float B) {  // a + w * (b - a) for A = 2^23 + a, B = 2^23 + b
  return fmaf(w, B - A, A - 8388608.f);
}

// Which sampler renders tile t: the software path, or (hybrid mode) the texture unit for `tex_of8` tiles of every 8 in a
// fixed pattern -- a static property of the tile, so an image does not depend on how the CTAs were scheduled.
__device__ __forceinline__ bool tile_is_tex(unsigned bx, unsigned by, int tex_of8) {
  return (int)((bx * 5u + by * 3u) & 7u) < tex_of8;
}

template <int FMT, bool STATS, class CFG>
__global__ void __launch_bounds__(SM_THREADS, CFG::MINB)
mip_smem_kernel(const MipArgs a, const int tex_of8, const __grid_constant__ CUtensorMap map0,
                const __grid_constant__ CUtensorMap map1, const __grid_constant__ CUtensorMap map2) {
  constexpr int SM_A = CFG::A, SM_B = CFG::B, SM_T = CFG::T, SM_STAGES = CFG::STAGES, SM_STAGE_ELEMS = CFG::STAGE_ELEMS,
                SM_STAGE_BYTES = CFG::STAGE_BYTES;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned short *stages = reinterpret_cast<unsigned short *>(smem_raw);
  SmemShared &sh = *reinterpret_cast<SmemShared *>(smem_raw + SM_STAGES * SM_STAGE_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool producer = warp == SM_CWARPS;
  const unsigned Nx = a.width, Ny = a.height;
  const Volume &V = a.vol;
  const unsigned tiles_x = (Nx + 15) / 16, tiles_y = (Ny + 15) / 16, ntiles = tiles_x * tiles_y;

  if (threadIdx.x == 0) {
    for (int s = 0; s < SM_STAGES; ++s) {
      mbar_init(smem_u32(&sh.full[s]), 1);
      mbar_init(smem_u32(&sh.empty[s]), SM_CWARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

  const int reducedSteps = a.max_steps;
  const int S = (reducedSteps / 16 + 1) * 16;
  const float Sf = (float)S;
  const int lx = (lane & 1) | ((lane >> 1) & 2) | ((lane >> 2) & 4);
  const int ly = ((lane >> 1) & 1) | ((lane >> 2) & 2);
  unsigned it_base = 0;  // slabs this CTA has been through: stage and phase of the ring carry on across tiles

  for (;;) {
    // ---- next tile (persistent CTAs pull tiles from a counter: no wave quantisation, rows stay L2-coherent) ----
    if (threadIdx.x == 0) {
      sh.tile = atomicAdd(a.tile_counter, 1u);
      sh.dmin = 0x7fffffff;
      sh.dmax = -0x7fffffff;
      sh.anyhit = 0;
    }
    __syncthreads();
    const unsigned t = sh.tile;
    if (t >= ntiles) break;
    const unsigned tbx = t % tiles_x, tby = t / tiles_x;
    const unsigned tile_x0 = tbx * 16, tile_y0 = tby * 16;
    const bool tex_tile = tile_is_tex(tbx, tby, tex_of8);

    // ---- ray setup: consumers their pixel, producer lanes 0-3 the tile's corner rays, lane 4 its centre ray ----
    unsigned x, y;
    if (!producer) {
      x = tile_x0 + (warp & 1) * 8 + lx;
      y = tile_y0 + (warp >> 1) * 4 + ly;
    } else {
      // pixel-corner rays of the tile's corner pixels bound every ray of the tile (rays are lines through the eye)
      const unsigned xe = min(tile_x0 + 15u, Nx - 1u), ye = min(tile_y0 + 15u, Ny - 1u);
      x = lane == 4 ? (tile_x0 + xe) / 2 : ((lane & 1) ? xe : tile_x0);
      y = lane == 4 ? (tile_y0 + ye) / 2 : ((lane & 2) ? ye : tile_y0);
    }
    const bool inb = x < Nx && y < Ny;
    Ray r = make_ray(x, y, Nx, Ny, a.cam, a.box);
    const bool hit = !producer && inb && r.hit;
    float tnear = r.tnear;
    if (tnear < 0.0f) tnear = 0.0f;
    // texel coordinates of sample k: u0 + k * du (as mip_fast_kernel); for rays that miss the box the line is still
    // the pixel's line (the producer needs its corner LINES whether or not they hit)
    float dt = fabsf(r.tfar - tnear) / (float)((reducedSteps / 16) * 16);
    if (!(dt > 0.f) || !r.hit) dt = 1.f / 192.f;
    float u0, v0, w0, du, dv, dw;
    {
      const v4 delta_pos = scl4(.5f * dt, r.direc);
      const v4 pos0 = scl4(0.5f, add4(sadd4(1.f, r.orig), scl4(r.hit ? tnear : 0.f, r.direc)));
      u0 = pos0.x * V.fnx; v0 = pos0.y * V.fny; w0 = pos0.z * V.fnz;
      du = delta_pos.x * V.fnx; dv = delta_pos.y * V.fny; dw = delta_pos.z * V.fnz;
    }
    float cur = 0.f;
    unsigned n_sw = 0, n_tex = 0;

    if (tex_tile) {
      // ================= texture-unit tile (hybrid mode): mip_fast_kernel's loop =================
      if (hit) {
        for (int k = 0; k < S; k += 8) {
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float kk = (float)(k + j);
            v[j] = sample_tmu_uvw<FMT, true>(V, fmaf(kk, du, u0), fmaf(kk, dv, v0), fmaf(kk, dw, w0));
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) cur = fmaxf(cur, v[j]);
        }
        if (STATS) n_tex += S;
      }
    } else {
      if (producer && lane == 4) {
        const float ax = fabsf(du), ay = fabsf(dv), az = fabsf(dw);
        const int D = (az >= ax && az >= ay) ? 2 : (ax >= ay ? 0 : 1);
        sh.axis = D;
        sh.up = (D == 0 ? du : (D == 1 ? dv : dw)) >= 0.f;
      }
      __syncthreads();  // axis known
      const int D = sh.axis;
      const bool asc = sh.up != 0;
      const int NA = D == 0 ? V.nz : V.nx, NB = D == 1 ? V.nz : V.ny, ND = D == 0 ? V.nx : (D == 1 ? V.ny : V.nz);
      Line L;
      permute3(D, u0 - 0.5f, v0 - 0.5f, w0 - 0.5f, L.a0, L.b0, L.d0);
      permute3(D, du, dv, dw, L.da, L.db, L.dd);

      if (hit) {  // base plane indices of my first and last sample
        const float f0 = floorf(L.d0), f1 = floorf(fmaf((float)(S - 1), L.dd, L.d0));
        const int lo = (int)fminf(fmaxf(fminf(f0, f1), 0.f), (float)(ND - 2));
        const int hi = (int)fminf(fmaxf(fmaxf(f0, f1), 0.f), (float)(ND - 2));
        atomicMin(&sh.dmin, lo);
        atomicMax(&sh.dmax, hi);
        sh.anyhit = 1;
      }
      __syncthreads();
      const bool anyhit = sh.anyhit != 0;
      // slabs: planes [s * SM_T, s * SM_T + SM_T) as base indices, visited in the direction the rays cross them
      const int s_first = anyhit ? sh.dmin / SM_T : 0, s_last = anyhit ? sh.dmax / SM_T : -1;
      const int n_slabs = s_last - s_first + 1;

      if (producer) {
        // ================= producer warp: lanes 0-3 = corner lines =================
        const CUtensorMap *map = D == 0 ? &map0 : (D == 1 ? &map1 : &map2);
        for (int it = 0; it < n_slabs; ++it) {
          const int s = asc ? s_first + it : s_last - it;
          const int dlo = s * SM_T;
          const unsigned git = it_base + (unsigned)it;
          const int stage = git % SM_STAGES;
          const unsigned phase = (git / SM_STAGES) & 1;
          // lateral hull of the corner lines between the slab's bounding planes d = dlo and d = dlo + SM_T
          float amin = 3e38f, amax = -3e38f, bmin = 3e38f, bmax = -3e38f;
          if (lane < 4) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const float kk = ((float)(dlo + e * SM_T) - L.d0) / L.dd;
              const float pa = fmaf(kk, L.da, L.a0), pb = fmaf(kk, L.db, L.b0);
              amin = fminf(amin, pa); amax = fmaxf(amax, pa);
              bmin = fminf(bmin, pb); bmax = fmaxf(bmax, pb);
            }
          }
#pragma unroll
          for (int o = 2; o > 0; o >>= 1) {
            amin = fminf(amin, __shfl_xor_sync(0xffffffffu, amin, o));
            amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
            bmin = fminf(bmin, __shfl_xor_sync(0xffffffffu, bmin, o));
            bmax = fmaxf(bmax, __shfl_xor_sync(0xffffffffu, bmax, o));
          }
          if (lane == 0) {
            // centre the needed range [floor(min), floor(max) + 1] in the box, then keep the box inside the volume
            const float ca = 0.5f * (floorf(amin) + floorf(amax)) + 0.5f - 0.5f * SM_A;
            const float cb = 0.5f * (floorf(bmin) + floorf(bmax)) + 0.5f - 0.5f * SM_B;
            // the box starts on a 16-byte boundary of its row (the bulk tensor copy faults on other start addresses:
            // measured, compute-sanitizer "illegal instruction" at UTMALDG); NA - SM_A is a multiple of 8 or rounded down
            const int oa = (int)fminf(fmaxf(ceilf(ca), 0.f), (float)(NA - SM_A)) & ~7;
            const int ob = (int)fminf(fmaxf(ceilf(cb), 0.f), (float)(NB - SM_B));
            const bool useful = amax >= -1.f && amin <= (float)NA && bmax >= -1.f && bmin <= (float)NB;
            mbar_wait(smem_u32(&sh.empty[stage]), phase ^ 1);
            SlabDesc dsc;
            dsc.baseA = SM_MAGIC_BITS + oa;
            dsc.baseB = SM_MAGIC_BITS + ob;
            dsc.baseD = SM_MAGIC_BITS + dlo;
            dsc.ldmax = useful ? (unsigned)(min(SM_T, ND - 1 - dlo) - 1) : 0x80000000u;  // flag: nothing staged
            sh.desc[stage] = dsc;
            const unsigned bar = smem_u32(&sh.full[stage]);
            if (useful) {
              mbar_arrive_expect_tx(bar, SM_STAGE_BYTES);
              tma_load_3d(smem_u32(stages + (size_t)stage * SM_STAGE_ELEMS), map, oa, ob, dlo, bar);
            } else {
              mbar_arrive(bar);
            }
          }
          __syncwarp();
        }
      } else {
        // ================= consumer warps =================
        float kf = 0.f;
        for (int it = 0; it < n_slabs; ++it) {
          const unsigned git = it_base + (unsigned)it;
          const int stage = git % SM_STAGES;
          const unsigned phase = (git / SM_STAGES) & 1;
          mbar_wait(smem_u32(&sh.full[stage]), phase);
          const SlabDesc dsc = sh.desc[stage];
          const bool staged = dsc.ldmax != 0x80000000u;
          const unsigned ldmax = staged ? dsc.ldmax : 0u;
          const int baseDr = staged ? dsc.baseD : dsc.baseD + (1 << 28);  // nothing staged: every range test fails
          const unsigned st_addr = smem_u32(stages + (size_t)stage * SM_STAGE_ELEMS);
          // samples whose base plane lies in this slab (or before it: those were not staged -> texture unit)
          const int edge = asc ? dsc.baseD + SM_T : dsc.baseD;  // plane bits beyond / at the start of the slab
          if (hit) {
            while (kf < Sf) {
              const float pd = fmaf(kf, L.dd, L.d0);
              const float td = __fadd_rd(pd, SM_MAGIC);
              const int bd = __float_as_int(td);
              if (asc ? (bd >= edge) : (bd < edge)) break;
              const float pa = fmaf(kf, L.da, L.a0), pb = fmaf(kf, L.db, L.b0);
              const float ta = __fadd_rd(pa, SM_MAGIC), tb = __fadd_rd(pb, SM_MAGIC);
              const unsigned la = (unsigned)(__float_as_int(ta) - dsc.baseA), lb = (unsigned)(__float_as_int(tb) - dsc.baseB),
                             ld = (unsigned)(bd - baseDr);
              float val;
              if (la <= (unsigned)(SM_A - 2) && lb <= (unsigned)(SM_B - 2) && ld <= ldmax) {
                const float wa = pa - (ta - SM_MAGIC), wb = pb - (tb - SM_MAGIC), wd = pd - (td - SM_MAGIC);
                const unsigned p = st_addr + 2u * ((ld * SM_B + lb) * SM_A + la);
                const float x00 = lerp_pair(wa, biased(lds_u16(p)), biased(lds_u16(p + 2)));
                const float x10 = lerp_pair(wa, biased(lds_u16(p + 2 * SM_A)), biased(lds_u16(p + 2 * SM_A + 2)));
                const float x01 = lerp_pair(wa, biased(lds_u16(p + 2 * SM_A * SM_B)), biased(lds_u16(p + 2 * SM_A * SM_B + 2)));
                const float x11 = lerp_pair(wa, biased(lds_u16(p + 2 * SM_A * SM_B + 2 * SM_A)),
                                            biased(lds_u16(p + 2 * SM_A * SM_B + 2 * SM_A + 2)));
                const float y0 = fmaf(wb, x10 - x00, x00), y1 = fmaf(wb, x11 - x01, x01);
                val = fmaf(wd, y1 - y0, y0);
                if (STATS) ++n_sw;
              } else {
                val = sample_tmu_uvw<FMT, true>(V, fmaf(kf, du, u0), fmaf(kf, dv, v0), fmaf(kf, dw, w0));
                if (STATS) ++n_tex;
              }
              cur = fmaxf(cur, val);
              kf += 1.f;
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&sh.empty[stage]));
        }
        if (hit) {  // what no slab covered (the samples beyond the last staged plane): texture unit, 8 in flight
          while (kf < Sf) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float kk = kf + (float)j;
              v[j] = kk < Sf ? sample_tmu_uvw<FMT, true>(V, fmaf(kk, du, u0), fmaf(kk, dv, v0), fmaf(kk, dw, w0)) : 0.f;
              if (STATS) n_tex += kk < Sf;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) cur = fmaxf(cur, v[j]);
            kf += 8.f;
          }
        }
      }
      it_base += (unsigned)n_slabs;
    }

    if (!producer) {
      if (STATS) {
        unsigned ns = n_sw, nt = n_tex, nh = hit ? 1u : 0u;
        for (int o = 16; o > 0; o >>= 1) {
          ns += __shfl_down_sync(0xffffffffu, ns, o);
          nt += __shfl_down_sync(0xffffffffu, nt, o);
          nh += __shfl_down_sync(0xffffffffu, nh, o);
        }
        if (lane == 0) {
          atomicAdd(a.stats + 0, (unsigned long long)nh);
          atomicAdd(a.stats + 1, (unsigned long long)(ns + nt));
          atomicAdd(a.stats + 2, (unsigned long long)ns);
        }
      }
      // ---- epilogue (as mip_fast_kernel): window, 128-bit stores ----
      const unsigned tx0 = tile_x0 + (warp & 1) * 8, ty0 = tile_y0 + (warp >> 1) * 4;
      const float alphaVal = hit ? tnear : 0.f;  // integer volumes: volume_kernel.cl:329 / :261
      float outVal = 0.f;
      if (hit) {
        float col = (a.max_val == 0.f) ? cur : (cur - a.min_val) / (a.max_val - a.min_val);
        if (a.gamma != 1.f) col = powf(col, a.gamma);
        outVal = clampf_cl(col, 0.f, 1.f);
      }
      float *dst_rows = a.out + (size_t)ty0 * Nx, *alpha_rows = a.alpha + (size_t)ty0 * Nx;
      const bool vec_ok = (Nx % 4 == 0) && (tx0 + 8 <= Nx) && (ty0 + 4 <= Ny);
      if (vec_ok) {
        sh.s_out[warp][ly * 8 + lx] = outVal;
        sh.s_alpha[warp][ly * 8 + lx] = alphaVal;
        __syncwarp();
        if (lane < 16) {
          const int q = lane & 7, row = q >> 1, half = q & 1;
          const float *src = (lane < 8 ? sh.s_out[warp] : sh.s_alpha[warp]) + row * 8 + half * 4;
          float *base = lane < 8 ? dst_rows : alpha_rows;
          *reinterpret_cast<float4 *>(base + (size_t)row * Nx + tx0 + half * 4) = *reinterpret_cast<const float4 *>(src);
        }
        __syncwarp();
      } else if (inb) {
        const size_t p = x + (size_t)Nx * ly;
        dst_rows[p] = outVal;
        alpha_rows[p] = alphaVal;
      }
    }
    __syncthreads();  // everyone is done with this tile's shared state before thread 0 sets up the next one
  }
}

template <class CFG>
static size_t cfg_smem_bytes() { return (size_t)CFG::STAGES * CFG::STAGE_BYTES + sizeof(SmemShared) + 128; }

void mip_smem_box(int cfg, int *abp) {
  switch (cfg) {
    case 1: abp[0] = Cfg1::A; abp[1] = Cfg1::B; abp[2] = Cfg1::P; break;
    case 2: abp[0] = Cfg2::A; abp[1] = Cfg2::B; abp[2] = Cfg2::P; break;
    case 3: abp[0] = Cfg3::A; abp[1] = Cfg3::B; abp[2] = Cfg3::P; break;
    case 4: abp[0] = Cfg4::A; abp[1] = Cfg4::B; abp[2] = Cfg4::P; break;
    default: abp[0] = Cfg0::A; abp[1] = Cfg0::B; abp[2] = Cfg0::P; break;
  }
}
int mip_smem_configs() { return SM_NCFG; }

template <int FMT, bool STATS, class CFG>
static cudaError_t launch_cfg(const MipArgs &a, const CUtensorMap *m, int tex_of8, cudaStream_t st) {
  const size_t smem = cfg_smem_bytes<CFG>();
  static int resident = 0;  // CTAs per SM x SMs of this instantiation (queried once)
  if (resident == 0) {
    cudaError_t e = cudaFuncSetAttribute(mip_smem_kernel<FMT, STATS, CFG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0, dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mip_smem_kernel<FMT, STATS, CFG>, SM_THREADS, smem);
    resident = (per_sm > 0 ? per_sm : 1) * (sms > 0 ? sms : 1);
  }
  const unsigned ntiles = ((a.width + 15) / 16) * ((a.height + 15) / 16);
  const unsigned grid = ntiles < (unsigned)resident ? ntiles : (unsigned)resident;
  mip_smem_kernel<FMT, STATS, CFG><<<grid, SM_THREADS, smem, st>>>(a, tex_of8, m[0], m[1], m[2]);
  return cudaGetLastError();
}

template <int FMT, bool STATS>
static cudaError_t launch_fmt(const MipArgs &a, int cfg, const CUtensorMap *m, int tex_of8, cudaStream_t st) {
  switch (cfg) {
    case 1: return launch_cfg<FMT, STATS, Cfg1>(a, m, tex_of8, st);
    case 2: return launch_cfg<FMT, STATS, Cfg2>(a, m, tex_of8, st);
    case 3: return launch_cfg<FMT, STATS, Cfg3>(a, m, tex_of8, st);
    case 4: return launch_cfg<FMT, STATS, Cfg4>(a, m, tex_of8, st);
    default: return launch_cfg<FMT, STATS, Cfg0>(a, m, tex_of8, st);
  }
}

cudaError_t launch_mip_smem(const MipArgs &a, int fmt, int cfg, const void *maps /* CUtensorMap[3] of this cfg */, int tex_of8,
                            cudaStream_t st) {
  const CUtensorMap *m = static_cast<const CUtensorMap *>(maps);
  if (!a.tile_counter) return cudaErrorInvalidValue;
  cudaError_t e = cudaMemsetAsync(a.tile_counter, 0, sizeof(unsigned), st);
  if (e != cudaSuccess) return e;
  const bool stats = a.stats != nullptr;
  if (fmt == 4) return stats ? launch_fmt<4, true>(a, cfg, m, tex_of8, st) : launch_fmt<4, false>(a, cfg, m, tex_of8, st);
  if (fmt == 1) return stats ? launch_fmt<1, true>(a, cfg, m, tex_of8, st) : launch_fmt<1, false>(a, cfg, m, tex_of8, st);
  return cudaErrorInvalidValue;
}

// ---- the permuted linear copies the tensor maps describe: dst[(d * NB + b) * pitchA + a] ----
template <int FMT>
__global__ void permute_kernel(const Volume V, int D, int NA, int NB, int ND, size_t pitchA, unsigned short *dst) {
  const int ia = blockIdx.x * blockDim.x + threadIdx.x, ib = blockIdx.y, id = blockIdx.z;
  if (ia >= NA) return;
  int i, j, k;  // x, y, z
  if (D == 0) { k = ia; j = ib; i = id; }
  else if (D == 1) { i = ia; k = ib; j = id; }
  else { i = ia; j = ib; k = id; }
  dst[((size_t)id * NB + ib) * pitchA + ia] = (unsigned short)texel<FMT>(V, i, j, k);
}

cudaError_t launch_permute(const Volume &V, int fmt, int D, int NA, int NB, int ND, size_t pitchA, void *dst,
                           cudaStream_t st) {
  dim3 block(128), grid((NA + 127) / 128, NB, ND);
  if (fmt == 4) permute_kernel<4><<<grid, block, 0, st>>>(V, D, NA, NB, ND, pitchA, (unsigned short *)dst);
  else if (fmt == 1) permute_kernel<1><<<grid, block, 0, st>>>(V, D, NA, NB, ND, pitchA, (unsigned short *)dst);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

}  // namespace spv
