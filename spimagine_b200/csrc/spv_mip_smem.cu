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

constexpr int SM_A = 40, SM_B = 32, SM_T = 8, SM_P = SM_T + 1;
constexpr int SM_STAGES = 3;
constexpr int SM_STAGE_ELEMS = SM_A * SM_B * SM_P;          // 11520 texels
constexpr int SM_STAGE_BYTES = SM_STAGE_ELEMS * 2;          // 23040 B
constexpr int SM_CWARPS = 8;
constexpr int SM_THREADS = 32 * (SM_CWARPS + 1);
constexpr float SM_MAGIC = 12582912.f;                      // 1.5 * 2^23: x + MAGIC (round down) = MAGIC + floor(x)
constexpr int SM_MAGIC_BITS = 0x4B400000;

struct SlabDesc {
  int baseA, baseB, baseD;  // MAGIC_BITS + box origin: local index = float bits of (coordinate + MAGIC) - base
  int ldmax;                // largest local base plane index a sample of this slab may have (-1: nothing staged)
};

struct SmemShared {
  unsigned long long full[SM_STAGES], empty[SM_STAGES];
  SlabDesc desc[SM_STAGES];
  int axis;            // D
  int dmin, dmax;      // base plane indices the tile's samples span (clamped to the planes that can be staged)
  int anyhit;
  float s_out[SM_CWARPS][32], s_alpha[SM_CWARPS][32];
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

// coordinates of the permuted copy with slowest axis D: (A, B, D) = (z, y, x) | (x, z, y) | (x, y, z)
__device__ __forceinline__ void permute3(int D, float x, float y, float z, float &a, float &b, float &d) {
  a = D == 0 ? z : x;
  b = D == 1 ? z : y;
  d = D == 0 ? x : (D == 1 ? y : z);
}

struct Line {  // sample k of a ray in texel-centre coordinates of the permuted copy: p0 + k * dp
  float a0, b0, d0, da, db, dd;
};

__device__ __forceinline__ float u16_to_float(unsigned v) { return __uint_as_float(v | 0x4B000000u) - 8388608.f; }

template <int FMT>
__global__ void __launch_bounds__(SM_THREADS, 3)
mip_smem_kernel(const MipArgs a, const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1,
                const __grid_constant__ CUtensorMap map2) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned short *stages = reinterpret_cast<unsigned short *>(smem_raw);
  SmemShared &sh = *reinterpret_cast<SmemShared *>(smem_raw + SM_STAGES * SM_STAGE_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool producer = warp == SM_CWARPS;
  const unsigned Nx = a.width, Ny = a.height;
  const unsigned tile_x0 = blockIdx.x * 16, tile_y0 = blockIdx.y * 16;
  const Volume &V = a.vol;
  const bool STATS = a.stats != nullptr;

  if (threadIdx.x == 0) {
    for (int s = 0; s < SM_STAGES; ++s) {
      mbar_init(smem_u32(&sh.full[s]), 1);
      mbar_init(smem_u32(&sh.empty[s]), SM_CWARPS);
    }
    sh.dmin = 0x7fffffff;
    sh.dmax = -0x7fffffff;
    sh.anyhit = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

  const int reducedSteps = a.max_steps;
  const int S = (reducedSteps / 16 + 1) * 16;

  // ---- ray setup: consumers their pixel, producer lanes 0-3 the tile's corner rays, lane 4 its centre ray ----
  const int lx = (lane & 1) | ((lane >> 1) & 2) | ((lane >> 2) & 4);
  const int ly = ((lane >> 1) & 1) | ((lane >> 2) & 2);
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
  // texel coordinates of sample k: u0 + k * du (as mip_fast_kernel); for rays that miss the box the line is still the
  // pixel's line (the producer needs its corner LINES whether or not they hit)
  float dt = fabsf(r.tfar - tnear) / (float)((reducedSteps / 16) * 16);
  if (!(dt > 0.f) || !r.hit) dt = 1.f / 192.f;
  float u0, v0, w0, du, dv, dw;
  {
    const v4 delta_pos = scl4(.5f * dt, r.direc);
    const v4 pos0 = scl4(0.5f, add4(sadd4(1.f, r.orig), scl4(r.hit ? tnear : 0.f, r.direc)));
    u0 = pos0.x * V.fnx; v0 = pos0.y * V.fny; w0 = pos0.z * V.fnz;
    du = delta_pos.x * V.fnx; dv = delta_pos.y * V.fny; dw = delta_pos.z * V.fnz;
  }
  if (producer && lane == 4) {
    const float ax = fabsf(du), ay = fabsf(dv), az = fabsf(dw);
    sh.axis = (az >= ax && az >= ay) ? 2 : (ax >= ay ? 0 : 1);
  }
  __syncthreads();  // barriers initialised, axis known
  const int D = sh.axis;
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
  // slabs: planes [s * SM_T, s * SM_T + SM_T) as base indices, visited in the direction the tile's rays cross them
  const bool up = __shfl_sync(0xffffffffu, L.dd, 0) >= 0.f;  // lane 0's ray; the centre ray decides for the producer
  const bool tile_up = D >= 0 ? (producer ? __shfl_sync(0xffffffffu, L.dd, 4) >= 0.f : up) : up;
  (void)tile_up;
  __shared__ int s_up;
  if (producer && lane == 4) s_up = L.dd >= 0.f;
  __syncthreads();
  const bool asc = s_up != 0;
  const int s_first = anyhit ? sh.dmin / SM_T : 0, s_last = anyhit ? sh.dmax / SM_T : -1;
  const int n_slabs = s_last - s_first + 1;

  float cur = 0.f;
  unsigned n_sw = 0, n_tex = 0;

  if (producer) {
    // ================= producer warp: lanes 0-3 = corner lines =================
    const CUtensorMap *map = D == 0 ? &map0 : (D == 1 ? &map1 : &map2);
    for (int it = 0; it < n_slabs; ++it) {
      const int s = asc ? s_first + it : s_last - it;
      const int dlo = s * SM_T;
      const int stage = it % SM_STAGES;
      const unsigned phase = (it / SM_STAGES) & 1;
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
        // centre the needed range [floor(min) - 1, floor(max) + 2] in the box, then keep the box inside the volume
        const float ca = 0.5f * (floorf(amin) + floorf(amax)) + 0.5f - 0.5f * SM_A;
        const float cb = 0.5f * (floorf(bmin) + floorf(bmax)) + 0.5f - 0.5f * SM_B;
        int oa = (int)fminf(fmaxf(ceilf(ca), 0.f), (float)(NA - SM_A));
        int ob = (int)fminf(fmaxf(ceilf(cb), 0.f), (float)(NB - SM_B));
        oa &= ~7;  // 16-byte aligned rows in global memory (the tensor map only needs element granularity; this keeps
                   // every row of the box inside as few 32-byte sectors as possible)
        const bool useful = amax >= -1.f && amin <= (float)NA && bmax >= -1.f && bmin <= (float)NB;
        mbar_wait(smem_u32(&sh.empty[stage]), phase ^ 1);
        SlabDesc dsc;
        dsc.baseA = SM_MAGIC_BITS + oa;
        dsc.baseB = SM_MAGIC_BITS + ob;
        dsc.baseD = SM_MAGIC_BITS + dlo;
        dsc.ldmax = useful ? min(SM_T, ND - 1 - dlo) - 1 : -1;
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
    const float Sf = (float)S;
    for (int it = 0; it < n_slabs; ++it) {
      const int stage = it % SM_STAGES;
      const unsigned phase = (it / SM_STAGES) & 1;
      mbar_wait(smem_u32(&sh.full[stage]), phase);
      const SlabDesc dsc = sh.desc[stage];
      const unsigned short *st = stages + (size_t)stage * SM_STAGE_ELEMS;
      // samples whose base plane lies in this slab (or before it: those were not staged and go to the texture unit)
      const int edge = asc ? dsc.baseD + SM_T : dsc.baseD;  // first plane bits beyond / first plane bits of the slab
      if (hit) {
        while (kf < Sf) {
          const float pd = fmaf(kf, L.dd, L.d0);
          const float td = __fadd_rd(pd, SM_MAGIC);
          const int bd = __float_as_int(td);
          if (asc ? (bd >= edge) : (bd < edge)) break;
          const float pa = fmaf(kf, L.da, L.a0), pb = fmaf(kf, L.db, L.b0);
          const float ta = __fadd_rd(pa, SM_MAGIC), tb = __fadd_rd(pb, SM_MAGIC);
          const int la = __float_as_int(ta) - dsc.baseA, lb = __float_as_int(tb) - dsc.baseB, ld = bd - dsc.baseD;
          float val;
          if ((unsigned)la <= (unsigned)(SM_A - 2) && (unsigned)lb <= (unsigned)(SM_B - 2) &&
              (unsigned)ld <= (unsigned)dsc.ldmax && dsc.ldmax >= 0) {
            const float wa = pa - (ta - SM_MAGIC), wb = pb - (tb - SM_MAGIC), wd = pd - (td - SM_MAGIC);
            const unsigned short *p = st + (ld * SM_B + lb) * SM_A + la;
            const float c000 = u16_to_float(p[0]), c100 = u16_to_float(p[1]);
            const float c010 = u16_to_float(p[SM_A]), c110 = u16_to_float(p[SM_A + 1]);
            const float c001 = u16_to_float(p[SM_A * SM_B]), c101 = u16_to_float(p[SM_A * SM_B + 1]);
            const float c011 = u16_to_float(p[SM_A * SM_B + SM_A]), c111 = u16_to_float(p[SM_A * SM_B + SM_A + 1]);
            const float x00 = fmaf(wa, c100 - c000, c000), x10 = fmaf(wa, c110 - c010, c010);
            const float x01 = fmaf(wa, c101 - c001, c001), x11 = fmaf(wa, c111 - c011, c011);
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
    if (hit) {  // what no slab covered (beyond the last staged plane, clamped coordinates): the texture unit
      while (kf < Sf) {
        cur = fmaxf(cur, sample_tmu_uvw<FMT, true>(V, fmaf(kf, du, u0), fmaf(kf, dv, v0), fmaf(kf, dw, w0)));
        if (STATS) ++n_tex;
        kf += 1.f;
      }
    }
  }

  if (producer) return;

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
  } else if (inb) {
    const size_t p = x + (size_t)Nx * ly;
    dst_rows[p] = outVal;
    alpha_rows[p] = alphaVal;
  }
}

size_t mip_smem_bytes() { return (size_t)SM_STAGES * SM_STAGE_BYTES + sizeof(SmemShared) + 128; }
void mip_smem_box(int *abp) { abp[0] = SM_A; abp[1] = SM_B; abp[2] = SM_P; }

cudaError_t launch_mip_smem(const MipArgs &a, int fmt, const void *maps /* CUtensorMap[3] */, cudaStream_t st) {
  const CUtensorMap *m = static_cast<const CUtensorMap *>(maps);
  dim3 grid((a.width + 15) / 16, (a.height + 15) / 16), block(SM_THREADS);
  const size_t smem = mip_smem_bytes();
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(mip_smem_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(mip_smem_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  if (fmt == 4) mip_smem_kernel<4><<<grid, block, smem, st>>>(a, m[0], m[1], m[2]);
  else if (fmt == 1) mip_smem_kernel<1><<<grid, block, smem, st>>>(a, m[0], m[1], m[2]);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
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
