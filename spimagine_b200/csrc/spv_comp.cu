// spv_comp.cu -- sort-last composite over peer memory (NVLink / NVSwitch), new relative to the reference, which is
// single-device.  The image is cut into `world` bands of rows; rank o owns band o.
//
//   render (spv_mip.cu, SPV_MIP_PUSH)  every rank stores the raw partial maxima of band o's pixels straight into
//                                      owner o's staging plane [parity][src rank] -- plain 128-bit stores into peer
//                                      memory, issued tile by tile while the ray march is still running elsewhere
//   comp_sync                          arrival counters in the owners' memory (release / acquire at system scope)
//   comp_finish                        the owner takes the max over the `world` partials of its band, applies
//                                      window + gamma and stores the finished pixels into EVERY rank's output plane
//   comp_sync                          second phase: every band has landed in my output plane
//
// max is associative, commutative and idempotent, so the result equals the single-GPU render bit for bit.  The
// staging is double-buffered by frame parity: a rank can only start pushing frame f+2 after it has seen every
// owner's phase-1 counter of frame f+1, which the owner raises after it has consumed frame f.
#include "spv_kernels.h"

namespace spv {

__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// One round of the arrival protocol in one launch.  Lane r raises counter [phase][my rank] in rank r's memory to
// `value` -- everything this stream wrote before (the preceding kernels' peer stores) is ordered in front of it --
// and then spins until rank r's counter [phase][r] in MY memory has reached `value`.  Bounded: after ~4 s a lane
// gives up and raises *err instead of hanging the GPU.
__global__ void comp_sync_kernel(PeerFlagPtrs peers, const unsigned *flags, int world, int rank, int phase,
                                 unsigned value, unsigned *err) {
  const int r = threadIdx.x;
  if (r >= world) return;
  __threadfence_system();
  st_release_sys(peers.p[r] + phase * MAX_WORLD + rank, value);
  const unsigned *f = flags + phase * MAX_WORLD + r;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    if ((int)(ld_acquire_sys(f) - value) >= 0) break;
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (t - t0 > 4000000000ull) {
      atomicExch(err, 1u + (unsigned)r);
      break;
    }
    __nanosleep(100);
  }
  __threadfence_system();
}

__device__ __forceinline__ float window1(float v, float minVal, float maxVal, float gamma) {
  if (v < 0.f) return 0.f;  // -1 marks a miss (every GPU agrees: the box test does not depend on the slab)
  v = (maxVal == 0.f) ? v : (v - minVal) / (maxVal - minVal);
  if (gamma != 1.f) v = powf(v, gamma);
  return fminf(fmaxf(v, 0.f), 1.f);
}

// one thread = 4 consecutive pixels of my band
__global__ void __launch_bounds__(256) comp_finish_kernel(const CompFinishArgs a) {
  const unsigned i = (blockIdx.x * blockDim.x + threadIdx.x) * 4u;
  if (i >= a.n_pixels) return;
  float4 m = *reinterpret_cast<const float4 *>(a.part + i);
  for (int s = 1; s < a.world; ++s) {
    const float4 v = *reinterpret_cast<const float4 *>(a.part + (size_t)s * a.band_pixels + i);
    m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
  }
  m.x = window1(m.x, a.min_val, a.max_val, a.gamma);
  m.y = window1(m.y, a.min_val, a.max_val, a.gamma);
  m.z = window1(m.z, a.min_val, a.max_val, a.gamma);
  m.w = window1(m.w, a.min_val, a.max_val, a.gamma);
  for (int r = 0; r < a.world; ++r) *reinterpret_cast<float4 *>(a.out[r] + a.first_pixel + i) = m;
}

// Sort-last iso surface: one thread = 4 consecutive pixels of my band; MIN over the ranks' candidates (INT_MAX = none)
__global__ void __launch_bounds__(256) k_reduce_kernel(const KReduceArgs a) {
  const unsigned i = (blockIdx.x * blockDim.x + threadIdx.x) * 4u;
  if (i >= a.n_pixels) return;
#pragma unroll
  for (int plane = 0; plane < 2; ++plane) {
    int4 m = *reinterpret_cast<const int4 *>(a.part + (size_t)plane * a.band + i);
    for (int s = 1; s < a.world; ++s) {
      const int4 v = *reinterpret_cast<const int4 *>(a.part + ((size_t)s * 2 + plane) * a.band + i);
      m.x = min(m.x, v.x); m.y = min(m.y, v.y); m.z = min(m.z, v.z); m.w = min(m.w, v.w);
    }
    for (int r = 0; r < a.world; ++r)
      *reinterpret_cast<int4 *>(a.kplanes[r] + (size_t)plane * a.n_image + a.first_pixel + i) = m;
  }
}

cudaError_t launch_k_reduce(const KReduceArgs &a, cudaStream_t st) {
  if (a.n_pixels == 0) return cudaSuccess;
  const unsigned threads = (a.n_pixels + 3) / 4;
  k_reduce_kernel<<<(threads + 255) / 256, 256, 0, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_comp_sync(unsigned *const *peer_flags, const unsigned *flags, int world, int rank, int phase,
                             unsigned value, unsigned *err, cudaStream_t st) {
  PeerFlagPtrs p;
  for (int r = 0; r < MAX_WORLD; ++r) p.p[r] = r < world ? peer_flags[r] : nullptr;
  comp_sync_kernel<<<1, 32, 0, st>>>(p, flags, world, rank, phase, value, err);
  return cudaGetLastError();
}

cudaError_t launch_comp_finish(const CompFinishArgs &a, cudaStream_t st) {
  if (a.n_pixels == 0) return cudaSuccess;
  const unsigned threads = (a.n_pixels + 3) / 4;
  comp_finish_kernel<<<(threads + 255) / 256, 256, 0, st>>>(a);
  return cudaGetLastError();
}

// Sort-last iso surface, last exchange: every rank ran the screen-space passes on its own band of rows only; the band's
// finished planes go to every peer.  One float4 per thread and plane segment; W * H is a multiple of 4 (spv_comp_init).
__global__ void __launch_bounds__(256) band_gather_kernel(const BandGatherArgs a) {
  const size_t n = (size_t)a.width * a.height;
  const size_t first = (size_t)a.y_first * a.width, count = (size_t)(a.y_end - a.y_first) * a.width;
  // segments: out [0, n), occ [3n, 4n) -> band pixels each; normals [4n, 7n) -> 3 floats per pixel
  const size_t q1 = count / 4, q3 = 3 * count / 4, total = 2 * q1 + q3;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    size_t off;  // float offset into the planes
    if (t < q1) off = first + 4 * t;
    else if (t < 2 * q1) off = 3 * n + first + 4 * (t - q1);
    else off = 4 * n + 3 * first + 4 * (t - 2 * q1);
    const float4 v = *reinterpret_cast<const float4 *>(a.planes[a.rank] + off);
    for (int r = 0; r < a.world; ++r)
      if (r != a.rank) *reinterpret_cast<float4 *>(a.planes[r] + off) = v;
  }
}
cudaError_t launch_band_gather(const BandGatherArgs &a, cudaStream_t st) {
  if (a.world < 2 || a.y_first >= a.y_end) return cudaSuccess;
  band_gather_kernel<<<592, 256, 0, st>>>(a);
  return cudaGetLastError();
}

cudaError_t preload_comp_kernels() {
  cudaFuncAttributes fa;
  cudaError_t e;
  if ((e = cudaFuncGetAttributes(&fa, comp_sync_kernel)) != cudaSuccess) return e;
  if ((e = cudaFuncGetAttributes(&fa, comp_finish_kernel)) != cudaSuccess) return e;
  if ((e = cudaFuncGetAttributes(&fa, k_reduce_kernel)) != cudaSuccess) return e;
  return cudaFuncGetAttributes(&fa, band_gather_kernel);
}

}  // namespace spv
