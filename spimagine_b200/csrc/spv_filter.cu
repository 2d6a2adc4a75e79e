// spv_filter.cu -- separable 3-D convolution of a volume on the device: what
// BlurProcessor / BlurXYZProcessor.apply get from gputools.convolve_sep3 (spimagine/models/imageprocessor.py:47-71),
// whose result the GUI hands to renderer.update_data (spimagine/gui/mainwidget.py:455-465).
//
// gputools is a third-party dependency that is not vendored in the reference tree (setup.py:31, unpinned).  Its
// published algorithm (gputools/convolve/kernels/convolve_sep.cl: conv_sep3_x / _y / _z, run in that order through
// float32 buffers) is restated here:
//
//     out[i] = sum_{ht = h_start}^{h_end - 1} h[ht] * in[i + Nh/2 - ht]        (per axis; integer Nh/2)
//     h_start = i + Nh/2 >= N ? i + Nh/2 + 1 - N : 0,   h_end = i - Nh/2 < 0 ? i + Nh/2 + 1 : Nh
//
// i.e. a true convolution centred on tap Nh/2 whose taps outside the volume are dropped, accumulated in float32 in
// ascending tap order.  `res += h * in` is evaluated as ONE fused multiply-add per tap (what OpenCL's default
// contraction gives on a GPU); the CPU restatement the tests compare against follows the same convention, bit for bit.
//
// Kernels.  A thread keeps R consecutive outputs of one line in registers and streams the R + Nh - 1 inputs they need
// from the highest position down, so that every output meets its taps in ascending order; after full unrolling every
// tap index is a compile-time constant and the weight is a constant-bank operand of the FMA (the taps travel as a
// kernel argument).  Dropped taps become fma(h, 0, acc) = acc.  Kernels are instantiated for the tap counts
// NHMAX in FILT_SIZES; other counts run in the next larger instantiation, whose taps beyond the real count are
// predicated off (a handful of predicates `nh > ht`, computed once), so a voxel never meets a tap the reference
// would not give it (NaN / Inf voxels spread exactly as far as in the reference).
//   conv_axis_kernel  y and z passes: lanes along x (coalesced), the line runs along the strided axis
//   conv_x_kernel     x pass: 32 lines x (XW + Nh - 1) inputs staged in shared memory (odd pitch: a lane per line
//                     reads conflict-free), results staged back for coalesced stores; reads the source element type
#include <string.h>

#include <string>
#include <thread>

#include "spimcuda.h"
#include "spv_kernels.h"

namespace spv {

// base + step * k as ONE instruction (IMAD.WIDE.U32 with an immediate k after unrolling); kept opaque so that the
// compiler does not turn the unrolled address sequence back into chains of 64-bit additions
__device__ __forceinline__ const char *mad_wide(unsigned step, unsigned k, const char *base) {
  unsigned long long r;
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(step), "r"(k), "l"((unsigned long long)base));
  return reinterpret_cast<const char *>(r);
}

// One pass along a strided axis.  in/out: float volumes; the line of (x, o) starts at base + o * so + x, positions
// along the axis are `sa` elements apart.  grid = (ceil(nx / 128), ceil(na / R), no).  The instantiation serves tap
// counts NHMIN <= nh <= NHMAX: steps beyond the real window (jj >= R + nh - 1, possible only for jj >= R + NHMIN - 1)
// load nothing, so a voxel never meets a padded tap.
template <int NHMAX, int NHMIN, int R>
__global__ void __launch_bounds__(128) conv_axis_kernel(const float *__restrict__ in, float *__restrict__ out, int nx, int na,
                                                        size_t sa, size_t so, int nh, const FilterTaps t) {
  const int x = blockIdx.x * 128 + threadIdx.x;
  if (x >= nx) return;
  const int p0 = blockIdx.y * R;
  const float *src = in + (size_t)blockIdx.z * so + x;
  float *dst = out + (size_t)blockIdx.z * so + x;
  const int half = nh / 2;
  const int qtop = p0 + R - 1 + half;  // input position met first (tap 0 of the last output)
  const int nsteps = R + nh - 1;       // positions qtop, qtop - 1, ... the R outputs read
  float acc[R];
#pragma unroll
  for (int r = 0; r < R; ++r) acc[r] = 0.f;
  if (qtop < na && qtop - (nsteps - 1) >= 0) {  // the whole window lies inside the volume (uniform per CTA)
    // addresses as low + step * constant with a 32-bit byte step: one IMAD.WIDE.U32 per load (the launcher checks
    // that the step fits); `low` is the position of the last unrolled step and is only dereferenced where a real tap
    // reads it
    constexpr int K = R + NHMAX - 2;
    const unsigned step = (unsigned)(sa * sizeof(float));
    const char *low = reinterpret_cast<const char *>(src) + ((long long)qtop - K) * (long long)(sa * sizeof(float));
#pragma unroll
    for (int jj = 0; jj < R + NHMAX - 1; ++jj) {
      const float *ptr = reinterpret_cast<const float *>(mad_wide(step, (unsigned)(K - jj), low));
      float v;
      if (jj < R + NHMIN - 1) v = __ldg(ptr);
      else v = jj < nsteps ? __ldg(ptr) : 0.f;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int ht = jj - (R - 1 - r);
        if (ht >= 0 && ht < NHMAX && (ht < NHMIN || ht < nh)) acc[r] = fmaf(t.w[ht], v, acc[r]);
      }
    }
  } else {  // window cut by a volume face: positions outside contribute fma(h, 0, acc) = acc
#pragma unroll
    for (int jj = 0; jj < R + NHMAX - 1; ++jj) {
      const int q = qtop - jj;
      float v = 0.f;
      if (jj < nsteps && q >= 0 && q < na) v = __ldg(src + (size_t)q * sa);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int ht = jj - (R - 1 - r);
        if (ht >= 0 && ht < NHMAX && (ht < NHMIN || ht < nh)) acc[r] = fmaf(t.w[ht], v, acc[r]);
      }
    }
  }
  char *obase = reinterpret_cast<char *>(dst + (size_t)p0 * sa);
  const unsigned ostep = (unsigned)(sa * sizeof(float));
#pragma unroll
  for (int r = 0; r < R; ++r)
    if (p0 + r < na) *reinterpret_cast<float *>(const_cast<char *>(mad_wide(ostep, (unsigned)r, obase))) = acc[r];
}

// The same pass W = 2 or 4 columns wide: a thread owns x = W t .. W t + W - 1 (one 64- or 128-bit load / store per
// position: 1/W of the address arithmetic and memory instructions per output) and R outputs of each column.  The
// two columns of a 64-bit pair are updated by ONE packed instruction (FFMA2 of sm_100: fma.rn.f32x2, whose weight
// operand is a broadcast scalar out of a uniform register) -- each half is the same IEEE fused multiply-add as the
// scalar kernel's, so the results are bit-identical while the issue slots per output halve.  Needs nx % W == 0 and
// volumes aligned to 4 W bytes.  grid = (ceil(nx / (128 W)), ceil(na / R), no).
template <int W> struct VecOf;
template <> struct VecOf<2> { typedef float2 type; };
template <> struct VecOf<4> { typedef float4 type; };

template <int NHMAX, int NHMIN, int R, int W>
__global__ void __launch_bounds__(128) conv_axisw_kernel(const float *__restrict__ in, float *__restrict__ out, int nx, int na,
                                                         size_t sa, size_t so, int nh, const FilterTaps t) {
  typedef typename VecOf<W>::type vec;
  constexpr int P = W / 2;  // packed pairs per position
  const int x = (blockIdx.x * 128 + threadIdx.x) * W;
  if (x >= nx) return;
  const int p0 = blockIdx.y * R;
  const float *src = in + (size_t)blockIdx.z * so + x;
  float *dst = out + (size_t)blockIdx.z * so + x;
  const int half = nh / 2;
  const int qtop = p0 + R - 1 + half;
  const int nsteps = R + nh - 1;
  float2 acc[R][P];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int k = 0; k < P; ++k) acc[r][k] = make_float2(0.f, 0.f);
  const bool interior = qtop < na && qtop - (nsteps - 1) >= 0;  // uniform per CTA
  constexpr int K = R + NHMAX - 2;
  const unsigned step = (unsigned)(sa * sizeof(float));
  const char *low = reinterpret_cast<const char *>(src) + ((long long)qtop - K) * (long long)(sa * sizeof(float));
  // one step of the stream: the position's W columns meet the taps of the R outputs that see it
#define SPV_AXISW_STEP(LOADED)                                                                                   \
  {                                                                                                              \
    float2 v[P];                                                                                                 \
    _Pragma("unroll") for (int k = 0; k < P; ++k) v[k] = make_float2(0.f, 0.f);                                  \
    if (LOADED) {                                                                                                \
      const vec l = __ldg(reinterpret_cast<const vec *>(mad_wide(step, (unsigned)(K - jj), low)));              \
      v[0] = make_float2(l.x, l.y);                                                                              \
      if (W == 4) v[P - 1] = make_float2(reinterpret_cast<const float *>(&l)[2], reinterpret_cast<const float *>(&l)[3]); \
    }                                                                                                            \
    _Pragma("unroll") for (int r = 0; r < R; ++r) {                                                              \
      const int ht = jj - (R - 1 - r);                                                                           \
      if (ht >= 0 && ht < NHMAX && (ht < NHMIN || ht < nh)) {                                                    \
        const float w = t.w[ht];                                                                                 \
        _Pragma("unroll") for (int k = 0; k < P; ++k) acc[r][k] = __ffma2_rn(make_float2(w, w), v[k], acc[r][k]); \
      }                                                                                                          \
    }                                                                                                            \
  }
  if (interior) {
#pragma unroll
    for (int jj = 0; jj < R + NHMAX - 1; ++jj) SPV_AXISW_STEP(jj < R + NHMIN - 1 || jj < nsteps)
  } else {
#pragma unroll
    for (int jj = 0; jj < R + NHMAX - 1; ++jj) SPV_AXISW_STEP(jj < nsteps && qtop - jj >= 0 && qtop - jj < na)
  }
#undef SPV_AXISW_STEP
  char *obase = reinterpret_cast<char *>(dst + (size_t)p0 * sa);
#pragma unroll
  for (int r = 0; r < R; ++r) {
    if (p0 + r < na) {
      vec sv;
      sv.x = acc[r][0].x;
      sv.y = acc[r][0].y;
      if (W == 4) {
        reinterpret_cast<float *>(&sv)[2] = acc[r][P - 1].x;
        reinterpret_cast<float *>(&sv)[3] = acc[r][P - 1].y;
      }
      *reinterpret_cast<vec *>(const_cast<char *>(mad_wide(step, (unsigned)r, obase))) = sv;
    }
  }
}

// loaded values wait in 32-bit registers (no sub-word packing)
template <typename T> struct Held { typedef unsigned type; };
template <> struct Held<float> { typedef float type; };
// exact uint8 / uint16 -> float on the FMA / ALU pipes (I2F runs on the slow conversion pipe): 2^23 + v has v in its
// mantissa for v < 2^23
__device__ __forceinline__ float to_float(unsigned v) { return __int_as_float(0x4b000000u | v) - 8388608.f; }
__device__ __forceinline__ float to_float(float v) { return v; }

// The x pass.  Lines are the nrows = ny * nz rows of the volume, nx elements each; a CTA of NW warps takes 32
// consecutive rows x XW = NW * R outputs.  grid = ceil(nx / XW) * ceil(nrows / 32), x tiles fastest.
template <typename TIN, int NHMAX, int NHMIN, int R, int NW>
__global__ void __launch_bounds__(32 * NW) conv_x_kernel(const TIN *__restrict__ in, float *__restrict__ out, int nx,
                                                         long long nrows, int nh, const FilterTaps t, int words) {
  constexpr int E = 4 / (int)sizeof(TIN);  // voxels per 32-bit word
  constexpr int XW = NW * R, TW = XW + NHMAX - 1, PITCH = (TW + E - 1) | 1, OPITCH = XW | 1;
  __shared__ float s_in[32][PITCH];
  __shared__ float s_out[32][OPITCH];
  const unsigned ntx = (unsigned)((nx + XW - 1) / XW);  // x tiles are dealt fastest: concurrent CTAs read whole rows
  const int x0 = (int)(blockIdx.x % ntx) * XW;
  const long long row0 = (long long)(blockIdx.x / ntx) * 32;
  const int half = nh / 2;
  const int qbase = x0 + half - (NHMAX - 1);  // input position of the first tile column the taps read
  static_assert(TW + E - 1 <= 32 * NW, "one thread per tile column");
  const int nr = nrows - row0 < 32 ? (int)(nrows - row0) : 32;
  int shift = 0;  // tile column 0 holds position qbase - shift
  if (E > 1 && words) {
    // uint8 / uint16 rows whose pitch and base are multiples of 4 bytes: a thread loads whole 32-bit words (E voxels),
    // so that a warp reads full 128-byte lines like the float32 path; a word lies entirely inside or outside a row
    shift = ((qbase % E) + E) % E;
    const int qa = qbase - shift;  // multiple of E
    constexpr int NWORDS = (TW + 2 * (E - 1)) / E;
    if (threadIdx.x < NWORDS) {
      const int q = qa + (int)threadIdx.x * E;
      const bool okx = q >= 0 && q < nx;
      const unsigned *p = reinterpret_cast<const unsigned *>(in + (size_t)row0 * nx + (okx ? q : 0));
      const size_t pitch_words = (size_t)nx / E;
      unsigned held[32];
#pragma unroll
      for (int rr = 0; rr < 32; ++rr) {
        held[rr] = (okx && rr < nr) ? __ldg(p) : 0u;
        p += pitch_words;
      }
#pragma unroll
      for (int rr = 0; rr < 32; ++rr)
#pragma unroll
        for (int k = 0; k < E; ++k) {
          const int c = (int)threadIdx.x * E + k;
          constexpr unsigned BITS = (8u * sizeof(TIN)) & 31u;  // 8 or 16 here (0 for float32: branch not taken)
          if (c < PITCH) s_in[rr][c] = to_float((held[rr] >> (BITS * k)) & ((1u << BITS) - 1u));
        }
    }
  } else if (threadIdx.x < TW) {  // a thread per tile column, walking down the 32 rows (coalesced across the warp)
    const int q = qbase + (int)threadIdx.x;
    const bool okx = q >= 0 && q < nx;
    const TIN *p = in + (size_t)row0 * nx + (okx ? q : 0);
    typename Held<TIN>::type held[32];  // every load of the column is in flight before the first one is needed
#pragma unroll
    for (int rr = 0; rr < 32; ++rr) {
      held[rr] = (okx && rr < nr) ? __ldg(p) : (TIN)0;
      p += nx;
    }
#pragma unroll
    for (int rr = 0; rr < 32; ++rr) s_in[rr][threadIdx.x] = to_float(held[rr]);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nsteps = R + nh - 1;
  float acc[R];
#pragma unroll
  for (int r = 0; r < R; ++r) acc[r] = 0.f;
  const float *line = s_in[lane] + shift + warp * R + R - 1 + NHMAX - 1;  // column of the position met first
#pragma unroll
  for (int jj = 0; jj < R + NHMAX - 1; ++jj) {
    float v = line[-jj];
    if (jj >= R + NHMIN - 1 && jj >= nsteps) v = 0.f;  // beyond the real window: only padded taps would meet it
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int ht = jj - (R - 1 - r);
      if (ht >= 0 && ht < NHMAX && (ht < NHMIN || ht < nh)) acc[r] = fmaf(t.w[ht], v, acc[r]);
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) s_out[lane][warp * R + r] = acc[r];
  __syncthreads();
  {  // a thread per output column and every (256 / XW)-th row
    const int c = threadIdx.x % XW, g = threadIdx.x / XW;
    constexpr int G = 32 * NW / XW;
    static_assert(32 % G == 0, "rows split evenly");
    if (x0 + c < nx) {
      float *o = out + (size_t)(row0 + g) * nx + x0 + c;
#pragma unroll 8
      for (int rr = g; rr < 32; rr += G) {
        if (row0 + rr < nrows) *o = s_out[rr][c];
        o += (size_t)G * nx;
      }
    }
  }
}

// The x pass on row PAIRS, software-pipelined: a CTA of 8 warps takes tiles of 64 consecutive rows x XW = 8 R outputs,
// a fixed grid of resident CTAs walks the tiles (x tiles fastest: concurrent CTAs read whole rows).
//   loads    every warp loads 8 rows of the tile as whole 32-bit words (full 128-byte lines per request), all of a
//            thread's loads in flight at once; with PIPE the NEXT tile's loads are issued before the current tile's
//            arithmetic and wait in registers, so a CTA has a tile of reads in flight all the time (the separate
//            load -> barrier -> compute -> barrier -> store phases of conv_x_kernel left the memory system idle for
//            most of a CTA's life: 3.3 of 6.5 TB/s)
//   taps     lane p owns rows p and p + 32, whose inputs sit side by side in shared memory (one 64-bit load per
//            position); both rows are updated by ONE packed FFMA2 per tap (fma.rn.f32x2, each half the scalar
//            kernel's IEEE fused multiply-add: bit-identical results, half the issue slots and shared-memory loads)
//   stores   the output tile reuses the input tile's memory and leaves as coalesced rows
// Needs rows of whole words (nx * sizeof(TIN) % 4 == 0, base aligned to 4 bytes); other volumes run conv_x_kernel.
template <typename TIN> __device__ __forceinline__ float word_elem(unsigned w, int k) {
  constexpr unsigned BITS = (8u * sizeof(TIN)) & 31u;
  return to_float((w >> (BITS * k)) & ((1u << BITS) - 1u));
}
template <> __device__ __forceinline__ float word_elem<float>(unsigned w, int) { return __uint_as_float(w); }

template <typename TIN, int NHMAX, int NHMIN, int R, int PIPE>
__global__ void __launch_bounds__(256, PIPE ? (sizeof(TIN) == 4 ? 2 : 3) : 4) conv_x2_kernel(const TIN *__restrict__ in, float *__restrict__ out, int nx, long long nrows,
                                                      int nh, const FilterTaps t, unsigned ntiles) {
  constexpr int NW = 8, E = 4 / (int)sizeof(TIN);  // voxels per 32-bit word
  constexpr int XW = NW * R, TW = XW + NHMAX - 1, PITCH = (TW + E - 1) | 1, OPITCH = XW | 1;
  constexpr int NWORDS = (TW + 2 * (E - 1)) / E, NCH = (NWORDS + 31) / 32;  // words per tile row, 32-word chunks
  constexpr int SMEM_FLOATS = 64 * PITCH > 64 * OPITCH ? 64 * PITCH : 64 * OPITCH;
  static_assert(SMEM_FLOATS * sizeof(float) <= 48 * 1024, "static shared memory");
  static_assert(NWORDS * E <= PITCH + E - 1, "tile row");
  __shared__ __align__(8) float s_mem[SMEM_FLOATS];
  float2(*s_in)[PITCH] = reinterpret_cast<float2(*)[PITCH]>(s_mem);   // [32][PITCH]: (row p, row p + 32)
  float(*s_out)[OPITCH] = reinterpret_cast<float(*)[OPITCH]>(s_mem);  // [64][OPITCH], after the taps have been read
  const unsigned ntx = (unsigned)((nx + XW - 1) / XW);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int half = nh / 2, nsteps = R + nh - 1;
  const unsigned pitch8 = 8u * (unsigned)nx * (unsigned)sizeof(TIN);  // bytes between the rows a warp loads
  unsigned held[8 * NCH];

  // issue the loads of `tile`: warp w takes rows w, w + 8, ..., w + 56
  auto issue = [&](unsigned tile) {
    const int x0 = (int)(tile % ntx) * XW;
    const long long row0 = (long long)(tile / ntx) * 64;
    const int qbase = x0 + half - (NHMAX - 1);
    const int qa = qbase - (((qbase % E) + E) % E);  // first word of the tile row (position, multiple of E)
    const int nr = nrows - row0 < 64 ? (int)(nrows - row0) : 64;
    const char *base = reinterpret_cast<const char *>(in) + ((row0 + warp) * (long long)nx + qa) * (long long)sizeof(TIN) + lane * 4;
    if (nr == 64 && qa >= 0 && qa + NWORDS * E <= nx) {  // uniform: the tile lies inside the volume
#pragma unroll
      for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch)
          held[k * NCH + ch] = (ch * 32 + 31 < NWORDS || ch * 32 + lane < NWORDS)
                                   ? __ldg(reinterpret_cast<const unsigned *>(mad_wide(pitch8, (unsigned)k, base) + ch * 128))
                                   : 0u;
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          const int wi = ch * 32 + lane, q = qa + wi * E;
          held[k * NCH + ch] = (wi < NWORDS && q >= 0 && q < nx && warp + 8 * k < nr)
                                   ? __ldg(reinterpret_cast<const unsigned *>(mad_wide(pitch8, (unsigned)k, base) + ch * 128))
                                   : 0u;
        }
    }
  };

  unsigned tile = blockIdx.x;
  if (tile >= ntiles) return;
  issue(tile);
  for (;;) {
    const int x0 = (int)(tile % ntx) * XW;
    const long long row0 = (long long)(tile / ntx) * 64;
    const int qbase = x0 + half - (NHMAX - 1);
    const int shift = ((qbase % E) + E) % E;  // tile column 0 holds position qbase - shift
    const int nr = nrows - row0 < 64 ? (int)(nrows - row0) : 64;
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        const int wi = ch * 32 + lane;
        if (ch * 32 + 31 < NWORDS || wi < NWORDS) {
#pragma unroll
          for (int e = 0; e < E; ++e) {
            const int c = wi * E + e;
            if (c < PITCH)
              s_in[warp + 8 * k][c] = make_float2(word_elem<TIN>(held[k * NCH + ch], e), word_elem<TIN>(held[(k + 4) * NCH + ch], e));
          }
        }
      }
    __syncthreads();
    const unsigned next = tile + gridDim.x;
    if (PIPE && next < ntiles) issue(next);  // in flight during this tile's arithmetic and stores
    float2 acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = make_float2(0.f, 0.f);
    const float2 *line = s_in[lane] + shift + warp * R + R - 1 + NHMAX - 1;  // column of the position met first
#pragma unroll
    for (int jj = 0; jj < R + NHMAX - 1; ++jj) {
      float2 v = line[-jj];
      if (jj >= R + NHMIN - 1 && jj >= nsteps) v = make_float2(0.f, 0.f);  // beyond the real window: only padded taps would meet it
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int ht = jj - (R - 1 - r);
        if (ht >= 0 && ht < NHMAX && (ht < NHMIN || ht < nh)) {
          const float w = t.w[ht];
          acc[r] = __ffma2_rn(make_float2(w, w), v, acc[r]);
        }
      }
    }
    __syncthreads();  // every warp has read its taps: the tile's memory becomes the output tile
#pragma unroll
    for (int r = 0; r < R; ++r) {
      s_out[lane][warp * R + r] = acc[r].x;
      s_out[lane + 32][warp * R + r] = acc[r].y;
    }
    __syncthreads();
    {  // a thread per output column and every G-th row
      const int c = threadIdx.x % XW, g = threadIdx.x / XW;
      constexpr int G = 256 / XW;
      static_assert(64 % G == 0, "rows split evenly");
      if (x0 + c < nx) {
        float *o = out + (size_t)(row0 + g) * nx + x0 + c;
#pragma unroll 8
        for (int rr = g; rr < 64; rr += G) {
          if (rr < nr) *o = s_out[rr][c];
          o += (size_t)G * nx;
        }
      }
    }
    if (next >= ntiles) break;
    if (!PIPE) issue(next);
    __syncthreads();  // the output tile has left: its memory takes the next tile's inputs
    tile = next;
  }
}

// x and y pass in one kernel (both tap counts in the same size class NHMAX <= 31): a CTA of 8 warps produces
// 128 (x) x YT (y) outputs of one slice, YT = 65 - NHMAX, from the 64 rows x (128 + NHMAX - 1) inputs they need.
//   stage 1  the x pass of conv_x_kernel for 2 x 32 rows (a lane per row), results into s_mid (odd pitch)
//   stage 2  the y pass out of s_mid: a thread per column and half of the YT outputs, streamed as in
//            conv_axis_kernel, stored straight to global memory (coalesced along x)
// Every intermediate value is the one the separate passes would have stored (same fma sequence), rows outside the
// volume are zeros as the dropped taps of the y pass require.  Saves one float32 write + read of the volume.
template <typename TIN, int NHMAX, int NHMIN>
__global__ void __launch_bounds__(256) conv_xy_fused_kernel(const TIN *__restrict__ in, float *__restrict__ out, int nx, int ny,
                                                            int nhx, int nhy, const FilterTaps tx, const FilterTaps ty,
                                                            int words) {
  constexpr int E = 4 / (int)sizeof(TIN);  // voxels per 32-bit word
  constexpr int R = 16, NW = 8, XW = NW * R, TW = XW + NHMAX - 1, PITCH = (TW + E - 1) | 1, MP = XW + 1;
  constexpr int YT = 65 - NHMAX, RY = YT / 2;
  extern __shared__ float smem[];
  float(*s_in)[PITCH] = reinterpret_cast<float(*)[PITCH]>(smem);             // [32][PITCH]
  float(*s_mid)[MP] = reinterpret_cast<float(*)[MP]>(smem + 32 * PITCH);      // [64][MP]
  const int x0 = blockIdx.x * XW, y0 = blockIdx.y * YT;
  const size_t slice = (size_t)blockIdx.z * nx * ny;
  const int halfx = nhx / 2, halfy = nhy / 2;
  const int qxbase = x0 + halfx - (NHMAX - 1);  // x position of tile column 0
  const int qybase = y0 + halfy - (NHMAX - 1);  // y position of tile row 0
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nstepsx = R + nhx - 1;
  const int shift = (E > 1 && words) ? ((qxbase % E) + E) % E : 0;  // tile column 0 holds x position qxbase - shift
  for (int round = 0; round < 2; ++round) {
    if (round) __syncthreads();  // everyone is done reading s_in
    const int qy0 = qybase + round * 32;
    if (E > 1 && words) {  // whole 32-bit words of uint8 / uint16 rows, as in conv_x_kernel
      constexpr int NWORDS = (TW + 2 * (E - 1)) / E;
      if (threadIdx.x < NWORDS) {
        const int qx = qxbase - shift + (int)threadIdx.x * E;
        const bool okx = qx >= 0 && qx < nx;
        const unsigned *p = reinterpret_cast<const unsigned *>(in + slice + (long long)qy0 * nx + (okx ? qx : 0));
        const long long pitch_words = nx / E;
#pragma unroll 8
        for (int rr = 0; rr < 32; ++rr) {
          const int qy = qy0 + rr;
          const unsigned w = (okx && qy >= 0 && qy < ny) ? __ldg(p) : 0u;
          p += pitch_words;
#pragma unroll
          for (int k = 0; k < E; ++k) {
            constexpr unsigned BITS = (8u * sizeof(TIN)) & 31u;
            const int c = (int)threadIdx.x * E + k;
            if (c < PITCH) s_in[rr][c] = to_float((w >> (BITS * k)) & ((1u << BITS) - 1u));
          }
        }
      }
    } else if (threadIdx.x < TW) {  // a thread per tile column, walking down the 32 rows of this round
      const int qx = qxbase + (int)threadIdx.x;
      const bool okx = qx >= 0 && qx < nx;
      const TIN *p = in + slice + (long long)qy0 * nx + (okx ? qx : 0);  // only dereferenced for rows inside the slice
#pragma unroll 8
      for (int rr = 0; rr < 32; ++rr) {
        const int qy = qy0 + rr;
        s_in[rr][threadIdx.x] = (okx && qy >= 0 && qy < ny) ? to_float((typename Held<TIN>::type)__ldg(p)) : 0.f;
        p += nx;
      }
    }
    __syncthreads();
    float acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0.f;
    const float *line = s_in[lane] + shift + warp * R + R - 1 + NHMAX - 1;
#pragma unroll
    for (int jj = 0; jj < R + NHMAX - 1; ++jj) {
      float v = line[-jj];
      if (jj >= R + NHMIN - 1 && jj >= nstepsx) v = 0.f;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int ht = jj - (R - 1 - r);
        if (ht >= 0 && ht < NHMAX && (ht < NHMIN || ht < nhx)) acc[r] = fmaf(tx.w[ht], v, acc[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) s_mid[round * 32 + lane][warp * R + r] = acc[r];
  }
  __syncthreads();
  // stage 2: column c of the tile, outputs y0 + half * RY + r
  const int c = threadIdx.x & (XW - 1), hchunk = threadIdx.x >> 7;
  const int nstepsy = RY + nhy - 1;
  float acc[RY];
#pragma unroll
  for (int r = 0; r < RY; ++r) acc[r] = 0.f;
  const int top = hchunk * RY + RY - 1 + NHMAX - 1;  // tile row of the position met first
#pragma unroll
  for (int jj = 0; jj < RY + NHMAX - 1; ++jj) {
    float v = s_mid[top - jj][c];
    if (jj >= RY + NHMIN - 1 && jj >= nstepsy) v = 0.f;
#pragma unroll
    for (int r = 0; r < RY; ++r) {
      const int ht = jj - (RY - 1 - r);
      if (ht >= 0 && ht < NHMAX && (ht < NHMIN || ht < nhy)) acc[r] = fmaf(ty.w[ht], v, acc[r]);
    }
  }
  if (x0 + c < nx) {
    const int ya = y0 + hchunk * RY;
    float *o = out + slice + (size_t)ya * nx + x0 + c;
#pragma unroll
    for (int r = 0; r < RY; ++r) {
      if (ya + r < ny) *o = acc[r];
      o += nx;
    }
  }
}

// any tap count: one output per thread, taps read from device memory with a runtime index (slow path; lanes run
// along the index with stride sx, the line along the axis with stride sa)
template <typename TIN>
__global__ void __launch_bounds__(256) conv_generic_kernel(const TIN *__restrict__ in, float *__restrict__ out, int nl, int na,
                                                           size_t sx, size_t sa, size_t so, int nh,
                                                           const float *__restrict__ taps) {
  const int l = blockIdx.x * 256 + threadIdx.x;
  if (l >= nl) return;
  const int p = blockIdx.y;
  const size_t base = (size_t)blockIdx.z * so + (size_t)l * sx;
  const int half = nh / 2;
  const int h_start = (p + half >= na) ? p + half + 1 - na : 0;
  const int h_end = (p - half < 0) ? p + half + 1 : nh;
  float res = 0.f;
  for (int ht = h_start; ht < h_end; ++ht) res = fmaf(taps[ht], (float)in[base + (size_t)(p + half - ht) * sa], res);
  out[base + (size_t)p * sa] = res;
}

int filter_axis_wide = 1;  // variant of the y / z passes (tuning knob 1): 1 = automatic, 16 / 32 = one column per thread with that many
                            // outputs, 1602 / 1604 = two / four columns per thread (packed FFMA2)

static const int FILT_SIZES[] = {3, 7, 11, 15, 19, 23, 27, 31, 35, 39, 47, 63};

static int pick_size(int nh) {
  for (int s : FILT_SIZES)
    if (nh <= s) return s;
  return 0;
}

template <int NHMAX, int NHMIN, int R, int W>
static void launch_axisw(const float *in, float *out, int nx, int na, int no, size_t sa, size_t so, int nh,
                         const FilterTaps &t, cudaStream_t st) {
  dim3 grid((nx + 128 * W - 1) / (128 * W), (na + R - 1) / R, no);
  conv_axisw_kernel<NHMAX, NHMIN, R, W><<<grid, 128, 0, st>>>(in, out, nx, na, sa, so, nh, t);
}

template <int NHMAX, int NHMIN>
static void launch_axis_n(const float *in, float *out, int nx, int na, int no, size_t sa, size_t so, int nh,
                          const FilterTaps &t, cudaStream_t st) {
  // default (1): four columns per thread with packed FFMA2 where rows are multiples of 16 bytes, two where of 8 bytes
  // (measured on B200, 512^3, 19 taps: y / z pass 180 / 195 us with one column, 172 / 177 with two, 164 / 171 with four =
  // 6.3-6.5 TB/s, the copy rate; profiles/r02s3_exp_blur2.txt), else one column
  const int wide = filter_axis_wide == 1 ? 4 : (filter_axis_wide == 1604 ? 4 : (filter_axis_wide == 1602 ? 2 : 0));
  const bool only = filter_axis_wide != 1;  // a variant asked for by the tuning knob is not replaced by a narrower one
  if (wide == 4 && nx % 4 == 0 && (((uintptr_t)in | (uintptr_t)out) & 15) == 0)
    return launch_axisw<NHMAX, NHMIN, 16, 4>(in, out, nx, na, no, sa, so, nh, t, st);
  if ((wide == 2 || (wide == 4 && !only)) && nx % 2 == 0 && (((uintptr_t)in | (uintptr_t)out) & 7) == 0)
    return launch_axisw<NHMAX, NHMIN, 16, 2>(in, out, nx, na, no, sa, so, nh, t, st);
  // one column: 32 outputs per thread from 11 taps up (1.56 instead of 2.1 loads per output: 3-5 % faster), 16 below
  if (filter_axis_wide == 32 || (filter_axis_wide != 16 && NHMAX >= 11)) {
    constexpr int R = 32;
    dim3 grid((nx + 127) / 128, (na + R - 1) / R, no);
    conv_axis_kernel<NHMAX, NHMIN, R><<<grid, 128, 0, st>>>(in, out, nx, na, sa, so, nh, t);
    return;
  }
  constexpr int R = 16;
  dim3 grid((nx + 127) / 128, (na + R - 1) / R, no);
  conv_axis_kernel<NHMAX, NHMIN, R><<<grid, 128, 0, st>>>(in, out, nx, na, sa, so, nh, t);
}

int filter_x_pairs = 2;  // x pass on row pairs with packed FFMA2 (conv_x2_kernel): 2 = software-pipelined (default), 1 = loads at the
                         // top of every tile, 0 = conv_x_kernel (tuning knob 2)

// resident CTAs of a kernel on the current device (cached per instantiation)
template <typename K> static int resident_ctas(K kernel, int threads) {
  int dev = 0, sms = 148, per_sm = 1;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
  return sms * per_sm;
}

template <typename TIN, int NHMAX, int NHMIN>
static void launch_x_n(const TIN *in, float *out, int nx, long long nrows, int nh, const FilterTaps &t, cudaStream_t st) {
  constexpr int R = 16, NW = 8;
  const int words = ((size_t)nx * sizeof(TIN)) % 4 == 0 && ((uintptr_t)in & 3) == 0;
  if constexpr (NHMAX <= 47) {
    const unsigned long long ntiles = (unsigned long long)((nrows + 63) / 64) * ((nx + NW * R - 1) / (NW * R));
    if (filter_x_pairs && words && ntiles < 0xffffffffull) {
      if (filter_x_pairs == 2) {
        static const int resident = resident_ctas(conv_x2_kernel<TIN, NHMAX, NHMIN, R, 1>, 256);
        const unsigned grid = ntiles < (unsigned long long)resident ? (unsigned)ntiles : (unsigned)resident;
        conv_x2_kernel<TIN, NHMAX, NHMIN, R, 1><<<grid, 256, 0, st>>>(in, out, nx, nrows, nh, t, (unsigned)ntiles);
      } else {
        static const int resident = resident_ctas(conv_x2_kernel<TIN, NHMAX, NHMIN, R, 0>, 256);
        const unsigned grid = ntiles < (unsigned long long)resident ? (unsigned)ntiles : (unsigned)resident;
        conv_x2_kernel<TIN, NHMAX, NHMIN, R, 0><<<grid, 256, 0, st>>>(in, out, nx, nrows, nh, t, (unsigned)ntiles);
      }
      return;
    }
  }
  dim3 grid((unsigned)(((nrows + 31) / 32) * ((nx + NW * R - 1) / (NW * R))));
  conv_x_kernel<TIN, NHMAX, NHMIN, R, NW><<<grid, 32 * NW, 0, st>>>(in, out, nx, nrows, nh, t, words && sizeof(TIN) < 4);
}

// CALL(NHMAX, NHMIN): the instantiation for tap counts NHMIN..NHMAX
#define SPV_FILT_DISPATCH(CALL)                 \
  switch (size) {                               \
    case 3: CALL(3, 1); break;                  \
    case 7: CALL(7, 4); break;                  \
    case 11: CALL(11, 8); break;                \
    case 15: CALL(15, 12); break;               \
    case 19: CALL(19, 16); break;               \
    case 23: CALL(23, 20); break;               \
    case 27: CALL(27, 24); break;               \
    case 31: CALL(31, 28); break;               \
    case 35: CALL(35, 32); break;               \
    case 39: CALL(39, 36); break;               \
    case 47: CALL(47, 40); break;               \
    default: CALL(63, 48); break;               \
  }

static FilterTaps padded(const float *h, int nh) {
  FilterTaps t;
  for (int i = 0; i < FILTER_MAX_TAPS; ++i) t.w[i] = i < nh ? h[i] : 0.f;
  return t;
}

// y (axis 1) or z (axis 2) pass over a float volume
cudaError_t launch_filter_axis(const float *in, float *out, int nx, int ny, int nz, int axis, const float *h, int nh,
                               const float *d_taps, cudaStream_t st) {
  const int na = axis == 1 ? ny : nz, no = axis == 1 ? nz : ny;
  const size_t sa = axis == 1 ? (size_t)nx : (size_t)nx * ny, so = axis == 1 ? (size_t)nx * ny : (size_t)nx;
  int size = pick_size(nh);
  if (sa * sizeof(float) > 0xffffffffull) size = 0;  // the unrolled kernel steps with a 32-bit byte stride
  if (size == 0) {
    // grid y / z limits (65535) cannot be hit by volumes that fit a texture (<= 16384 per axis)
    dim3 grid((nx + 255) / 256, na, no);
    conv_generic_kernel<float><<<grid, 256, 0, st>>>(in, out, nx, na, 1, sa, so, nh, d_taps);
    return cudaGetLastError();
  }
  const FilterTaps t = padded(h, nh);
#define SPV_AXIS(N, M) launch_axis_n<N, M>(in, out, nx, na, no, sa, so, nh, t, st)
  SPV_FILT_DISPATCH(SPV_AXIS)
#undef SPV_AXIS
  return cudaGetLastError();
}

template <typename TIN>
static cudaError_t filter_x_typed(const TIN *in, float *out, int nx, int ny, int nz, const float *h, int nh,
                                  const float *d_taps, cudaStream_t st) {
  const int size = pick_size(nh);
  if (size == 0) {  // lanes over y (stride nx), the line along x (stride 1), one grid layer per slice
    dim3 grid((ny + 255) / 256, nx, nz);
    conv_generic_kernel<TIN><<<grid, 256, 0, st>>>(in, out, ny, nx, (size_t)nx, 1, (size_t)nx * ny, nh, d_taps);
    return cudaGetLastError();
  }
  const FilterTaps t = padded(h, nh);
  const long long nrows = (long long)ny * nz;
#define SPV_X(N, M) launch_x_n<TIN, N, M>(in, out, nx, nrows, nh, t, st)
  SPV_FILT_DISPATCH(SPV_X)
#undef SPV_X
  return cudaGetLastError();
}

template <typename TIN, int NHMAX, int NHMIN>
static cudaError_t launch_xy_n(const TIN *in, float *out, int nx, int ny, int nz, int nhx, int nhy, const FilterTaps &tx,
                               const FilterTaps &ty, cudaStream_t st) {
  constexpr int E = 4 / (int)sizeof(TIN), TW = 128 + NHMAX - 1, PITCH = (TW + E - 1) | 1, YT = 65 - NHMAX;
  const size_t smem = (size_t)(32 * PITCH + 64 * 129) * sizeof(float);
  const int words = sizeof(TIN) < 4 && ((size_t)nx * sizeof(TIN)) % 4 == 0 && ((uintptr_t)in & 3) == 0;
  static bool attr_set = false;  // per instantiation
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_xy_fused_kernel<TIN, NHMAX, NHMIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  dim3 grid((nx + 127) / 128, (ny + YT - 1) / YT, nz);
  conv_xy_fused_kernel<TIN, NHMAX, NHMIN><<<grid, 256, smem, st>>>(in, out, nx, ny, nhx, nhy, tx, ty, words);
  return cudaGetLastError();
}

template <typename TIN>
static cudaError_t filter_xy_typed(const TIN *in, float *out, int nx, int ny, int nz, const float *hx, int nhx,
                                   const float *hy, int nhy, cudaStream_t st) {
  const int size = pick_size(nhx);
  const FilterTaps tx = padded(hx, nhx), ty = padded(hy, nhy);
  switch (size) {
    case 3: return launch_xy_n<TIN, 3, 1>(in, out, nx, ny, nz, nhx, nhy, tx, ty, st);
    case 7: return launch_xy_n<TIN, 7, 4>(in, out, nx, ny, nz, nhx, nhy, tx, ty, st);
    case 11: return launch_xy_n<TIN, 11, 8>(in, out, nx, ny, nz, nhx, nhy, tx, ty, st);
    case 15: return launch_xy_n<TIN, 15, 12>(in, out, nx, ny, nz, nhx, nhy, tx, ty, st);
    case 19: return launch_xy_n<TIN, 19, 16>(in, out, nx, ny, nz, nhx, nhy, tx, ty, st);
    case 23: return launch_xy_n<TIN, 23, 20>(in, out, nx, ny, nz, nhx, nhy, tx, ty, st);
    case 27: return launch_xy_n<TIN, 27, 24>(in, out, nx, ny, nz, nhx, nhy, tx, ty, st);
    case 31: return launch_xy_n<TIN, 31, 28>(in, out, nx, ny, nz, nhx, nhy, tx, ty, st);
    default: return cudaErrorInvalidValue;
  }
}

// can the x and y passes run as one kernel?  (both tap counts in the same size class, at most 31 taps)
bool filter_xy_fusable(int nhx, int nhy) {
  const int sx = pick_size(nhx), sy = pick_size(nhy);
  return sx != 0 && sx == sy && sx <= 31;
}

// fused x + y pass (filter_xy_fusable) from a volume of element type dtype into a float volume
cudaError_t launch_filter_xy(const void *in, int dtype, float *out, int nx, int ny, int nz, const float *hx, int nhx,
                             const float *hy, int nhy, cudaStream_t st) {
  switch (dtype) {
    case 0: return filter_xy_typed((const float *)in, out, nx, ny, nz, hx, nhx, hy, nhy, st);
    case 1: return filter_xy_typed((const unsigned short *)in, out, nx, ny, nz, hx, nhx, hy, nhy, st);
    case 2: return filter_xy_typed((const unsigned char *)in, out, nx, ny, nz, hx, nhx, hy, nhy, st);
    default: return cudaErrorInvalidValue;
  }
}

// x pass from a volume of element type dtype (SPV_F32 / SPV_U16 / SPV_U8) into a float volume
cudaError_t launch_filter_x(const void *in, int dtype, float *out, int nx, int ny, int nz, const float *h, int nh,
                            const float *d_taps, cudaStream_t st) {
  switch (dtype) {
    case 0: return filter_x_typed((const float *)in, out, nx, ny, nz, h, nh, d_taps, st);
    case 1: return filter_x_typed((const unsigned short *)in, out, nx, ny, nz, h, nh, d_taps, st);
    case 2: return filter_x_typed((const unsigned char *)in, out, nx, ny, nz, h, nh, d_taps, st);
    default: return cudaErrorInvalidValue;
  }
}

// memcpy on several host threads: pageable memory moves into the page-locked staging rings of the ingest paths at
// several times the rate of one thread (shared with spv_api.cu)
void parallel_memcpy(void *dst, const void *src, size_t n) {
  const unsigned hw = std::thread::hardware_concurrency();
  int T = n >= ((size_t)8 << 20) ? (hw >= 16 ? 8 : (hw >= 8 ? 4 : 2)) : 1;
  const size_t part = ((n / T) + 4095) & ~(size_t)4095;
  std::thread th[8];
  int started = 0;
  for (int t = 1; t < T; ++t) {
    const size_t off = (size_t)t * part;
    if (off >= n) break;
    const size_t len = n - off < part ? n - off : part;
    th[started++] = std::thread([=] { memcpy((char *)dst + off, (const char *)src + off, len); });
  }
  memcpy(dst, src, n < part ? n : part);
  for (int t = 0; t < started; ++t) th[t].join();
}


}  // namespace spv

// ---- C ABI -------------------------------------------------------------------------------------------------------
using namespace spv;

struct spv_filter {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t ev_pass[2] = {nullptr, nullptr};  // after the first / second kernel of the last convolution
  int passes = 0;                               // kernels of the last convolution (3, or 2 with the fused x + y pass)
  void *d_src = nullptr;   // the loaded volume in its own element type (host sources and converted ones)
  size_t src_cap = 0;
  float *buf[2] = {nullptr, nullptr};
  size_t buf_cap = 0;      // floats per buffer
  float *d_taps = nullptr; // 3 x FILT_LONG_TAPS, for tap counts beyond the unrolled instantiations
  char *h_ring = nullptr;  // 2 x 32 MiB page-locked staging for pageable host sources
  cudaEvent_t ev_ring[2] = {nullptr, nullptr};
  int nx = 0, ny = 0, nz = 0;
  const void *cur = nullptr;  // what the next convolution reads: the loaded volume or the last result
  int cur_dtype = 0;          // SPV_F32 / SPV_U16 / SPV_U8
  bool have_result = false, timed = false;
  int fuse_xy = 0;  // x and y pass in one kernel wherever the tap counts allow it (spv_filter_set_tuning knob 0; off: see
                    // spv_filter_convolve_sep3)
  unsigned long long launches = 0;
  std::string err;
};

static const int FILT_LONG_TAPS = 1024;
static thread_local std::string g_filter_create_err;

static int ffail(spv_filter *f, int code, const char *what) {
  (f ? f->err : g_filter_create_err) = what;
  return code;
}
static int fcufail(spv_filter *f, cudaError_t e, const char *where) {
  (f ? f->err : g_filter_create_err) = std::string(where) + ": " + cudaGetErrorString(e);
  cudaGetLastError();
  return (int)e;
}
#define FCU(call)                                          \
  do {                                                     \
    cudaError_t e_ = (call);                               \
    if (e_ != cudaSuccess) return fcufail(f, e_, #call);   \
  } while (0)
#define FBIND()                                                       \
  if (!f) return SPV_EINVAL;                                          \
  do {                                                                \
    cudaError_t e_ = cudaSetDevice(f->device);                        \
    if (e_ != cudaSuccess) return fcufail(f, e_, "cudaSetDevice");    \
  } while (0)

SPV_API const char *spv_filter_last_error(spv_filter *f) { return f ? f->err.c_str() : g_filter_create_err.c_str(); }

SPV_API int spv_filter_create(int device, spv_filter **out) {
  if (!out) return SPV_EINVAL;
  *out = nullptr;
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return fcufail(nullptr, e, "spv_filter_create: cudaSetDevice (libspimcuda needs a CUDA device; there is no CPU path)");
  spv_filter *f = new spv_filter;
  f->device = device;
  if ((e = cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking)) != cudaSuccess ||
      (e = cudaEventCreate(&f->ev0)) != cudaSuccess || (e = cudaEventCreate(&f->ev1)) != cudaSuccess ||
      (e = cudaEventCreate(&f->ev_pass[0])) != cudaSuccess || (e = cudaEventCreate(&f->ev_pass[1])) != cudaSuccess ||
      (e = cudaMalloc(&f->d_taps, 3 * FILT_LONG_TAPS * sizeof(float))) != cudaSuccess) {
    int rc = fcufail(nullptr, e, "spv_filter_create");
    spv_filter_destroy(f);
    return rc;
  }
  *out = f;
  return 0;
}

SPV_API int spv_filter_destroy(spv_filter *f) {
  if (!f) return 0;
  cudaSetDevice(f->device);
  if (f->stream) cudaStreamSynchronize(f->stream);
  if (f->d_src) cudaFree(f->d_src);
  for (int i = 0; i < 2; ++i)
    if (f->buf[i]) cudaFree(f->buf[i]);
  if (f->d_taps) cudaFree(f->d_taps);
  if (f->h_ring) cudaFreeHost(f->h_ring);
  for (int i = 0; i < 2; ++i)
    if (f->ev_ring[i]) cudaEventDestroy(f->ev_ring[i]);
  if (f->ev0) cudaEventDestroy(f->ev0);
  if (f->ev1) cudaEventDestroy(f->ev1);
  for (cudaEvent_t e : f->ev_pass)
    if (e) cudaEventDestroy(e);
  if (f->stream) cudaStreamDestroy(f->stream);
  cudaGetLastError();
  delete f;
  return 0;
}

static int filter_reserve(spv_filter *f, size_t n, size_t src_bytes) {
  if (n > f->buf_cap) {
    FCU(cudaStreamSynchronize(f->stream));
    for (int i = 0; i < 2; ++i) {
      if (f->buf[i]) cudaFree(f->buf[i]);
      f->buf[i] = nullptr;
    }
    f->buf_cap = 0;
    FCU(cudaMalloc(&f->buf[0], n * sizeof(float)));
    FCU(cudaMalloc(&f->buf[1], n * sizeof(float)));
    f->buf_cap = n;
  }
  if (src_bytes > f->src_cap) {
    FCU(cudaStreamSynchronize(f->stream));
    if (f->d_src) cudaFree(f->d_src);
    f->d_src = nullptr;
    f->src_cap = 0;
    FCU(cudaMalloc(&f->d_src, src_bytes));
    f->src_cap = src_bytes;
  }
  return 0;
}

SPV_API int spv_filter_load(spv_filter *f, const void *src, int on_device, int src_type, int nx, int ny, int nz) {
  FBIND();
  if (!src) return ffail(f, SPV_EINVAL, "spv_filter_load: null data");
  if (nx <= 0 || ny <= 0 || nz <= 0 || ny > 65535 || nz > 65535)
    return ffail(f, SPV_EINVAL, "spv_filter_load: bad extent (1 <= nx, 1 <= ny, nz <= 65535)");
  const size_t es = src_elem_size(src_type);
  if (es == 0) return ffail(f, SPV_EINVAL, "spv_filter_load: unknown source element type");
  const size_t n = (size_t)nx * ny * nz;
  const int native = src_type == SPV_SRC_F32 ? SPV_F32 : (src_type == SPV_SRC_U16 ? SPV_U16 : (src_type == SPV_SRC_U8 ? SPV_U8 : -1));
  const bool borrow = on_device && native >= 0;  // read in place by the next convolution
  int rc = filter_reserve(f, n, borrow ? 0 : n * es);
  if (rc) return rc;
  f->nx = nx; f->ny = ny; f->nz = nz;
  f->have_result = false;
  if (borrow) {
    f->cur = src;
    f->cur_dtype = native;
    return 0;
  }
  if (on_device) {
    FCU(cudaMemcpyAsync(f->d_src, src, n * es, cudaMemcpyDeviceToDevice, f->stream));
  } else {
    cudaPointerAttributes at;
    cudaError_t pe = cudaPointerGetAttributes(&at, src);
    if (pe != cudaSuccess) cudaGetLastError();
    if (pe == cudaSuccess && at.type == cudaMemoryTypeHost) {  // page-locked: the DMA engine reads it at PCIe rate
      FCU(cudaMemcpyAsync(f->d_src, src, n * es, cudaMemcpyHostToDevice, f->stream));
    } else {
      // pageable memory: host threads copy chunk i+1 into a page-locked ring while chunk i is on the PCIe link (the
      // ingest pipeline of spv_set_volume; a plain cudaMemcpy from pageable memory runs at a third of this rate)
      const size_t chunk = (size_t)32 << 20;
      if (!f->h_ring) {
        FCU(cudaMallocHost(&f->h_ring, 2 * chunk));
        FCU(cudaEventCreateWithFlags(&f->ev_ring[0], cudaEventDisableTiming));
        FCU(cudaEventCreateWithFlags(&f->ev_ring[1], cudaEventDisableTiming));
      }
      const size_t total = n * es;
      int i = 0;
      for (size_t off = 0; off < total; off += chunk, ++i) {
        const size_t len = total - off < chunk ? total - off : chunk;
        const int h = i & 1;
        if (i >= 2) FCU(cudaEventSynchronize(f->ev_ring[h]));  // the DMA that last read this half
        parallel_memcpy(f->h_ring + (size_t)h * chunk, (const char *)src + off, len);
        FCU(cudaMemcpyAsync((char *)f->d_src + off, f->h_ring + (size_t)h * chunk, len, cudaMemcpyHostToDevice, f->stream));
        FCU(cudaEventRecord(f->ev_ring[h], f->stream));
      }
    }
  }
  if (native >= 0) {
    f->cur = f->d_src;
    f->cur_dtype = native;
  } else {  // other element types become float32 first, as gputools' data.astype(np.float32) does on the host
    FCU(launch_convert(f->d_src, f->buf[1], src_type, SPV_F32, n, f->stream));
    f->cur = f->buf[1];
    f->cur_dtype = SPV_F32;
  }
  if (!on_device) FCU(cudaStreamSynchronize(f->stream));  // the host pointer is only borrowed for this call
  return 0;
}

SPV_API int spv_filter_convolve_sep3(spv_filter *f, const float *hx, int nhx, const float *hy, int nhy, const float *hz,
                                     int nhz) {
  FBIND();
  if (!f->cur) return ffail(f, SPV_ENODATA, "spv_filter_convolve_sep3: no volume loaded");
  if (!hx || !hy || !hz || nhx < 1 || nhy < 1 || nhz < 1 || nhx > FILT_LONG_TAPS || nhy > FILT_LONG_TAPS || nhz > FILT_LONG_TAPS)
    return ffail(f, SPV_EINVAL, "spv_filter_convolve_sep3: need 1 <= taps <= 1024 per axis");
  if (nhx > FILTER_MAX_TAPS || nhy > FILTER_MAX_TAPS || nhz > FILTER_MAX_TAPS) {
    FCU(cudaMemcpyAsync(f->d_taps, hx, nhx * sizeof(float), cudaMemcpyHostToDevice, f->stream));
    FCU(cudaMemcpyAsync(f->d_taps + FILT_LONG_TAPS, hy, nhy * sizeof(float), cudaMemcpyHostToDevice, f->stream));
    FCU(cudaMemcpyAsync(f->d_taps + 2 * FILT_LONG_TAPS, hz, nhz * sizeof(float), cudaMemcpyHostToDevice, f->stream));
    FCU(cudaStreamSynchronize(f->stream));  // the tap arrays are only borrowed for this call
  }
  const int i = f->cur == f->buf[0] ? 1 : 0;  // x: cur -> buf[i], y: buf[i] -> buf[1-i], z: buf[1-i] -> buf[i]
  FCU(cudaEventRecord(f->ev0, f->stream));
  // measured on B200 (profiles/r01_exp_blur.txt): the three passes each run at 65-80 % of the HBM copy rate; the fused
  // kernel saves a float32 round trip of the volume but is instruction-bound (39 % more x-pass FMAs, two barriers per
  // tile) and does not beat them yet -- opt-in
  const bool fuse = f->fuse_xy != 0;
  if (fuse && filter_xy_fusable(nhx, nhy)) {  // x + y in one kernel: cur -> buf[i], then z: buf[i] -> buf[1-i]
    const int j = i;  // buf[i] is not the source
    FCU(launch_filter_xy(f->cur, f->cur_dtype, f->buf[j], f->nx, f->ny, f->nz, hx, nhx, hy, nhy, f->stream));
    FCU(cudaEventRecord(f->ev_pass[0], f->stream));
    f->passes = 2;
    FCU(launch_filter_axis(f->buf[j], f->buf[1 - j], f->nx, f->ny, f->nz, 2, hz, nhz, f->d_taps + 2 * FILT_LONG_TAPS, f->stream));
    f->launches += 2;
    FCU(cudaEventRecord(f->ev1, f->stream));
    f->cur = f->buf[1 - j];
    f->cur_dtype = SPV_F32;
    f->have_result = true;
    f->timed = true;
    return 0;
  }
  FCU(launch_filter_x(f->cur, f->cur_dtype, f->buf[i], f->nx, f->ny, f->nz, hx, nhx, f->d_taps, f->stream));
  FCU(cudaEventRecord(f->ev_pass[0], f->stream));
  FCU(launch_filter_axis(f->buf[i], f->buf[1 - i], f->nx, f->ny, f->nz, 1, hy, nhy, f->d_taps + FILT_LONG_TAPS, f->stream));
  FCU(cudaEventRecord(f->ev_pass[1], f->stream));
  f->passes = 3;
  FCU(launch_filter_axis(f->buf[1 - i], f->buf[i], f->nx, f->ny, f->nz, 2, hz, nhz, f->d_taps + 2 * FILT_LONG_TAPS, f->stream));
  f->launches += 3;
  FCU(cudaEventRecord(f->ev1, f->stream));
  f->cur = f->buf[i];
  f->cur_dtype = SPV_F32;
  f->have_result = true;
  f->timed = true;
  return 0;
}

SPV_API int spv_filter_sync(spv_filter *f) {
  FBIND();
  FCU(cudaStreamSynchronize(f->stream));
  return 0;
}

SPV_API int spv_filter_result_device(spv_filter *f, float **dev) {
  FBIND();
  if (!dev) return ffail(f, SPV_EINVAL, "spv_filter_result_device: null pointer");
  if (!f->have_result) return ffail(f, SPV_ENODATA, "spv_filter_result_device: nothing convolved yet");
  *dev = const_cast<float *>((const float *)f->cur);
  return 0;
}

SPV_API int spv_filter_read(spv_filter *f, float *host_dst, size_t n) {
  FBIND();
  if (!host_dst) return ffail(f, SPV_EINVAL, "spv_filter_read: null destination");
  if (!f->have_result) return ffail(f, SPV_ENODATA, "spv_filter_read: nothing convolved yet");
  if (n != (size_t)f->nx * f->ny * f->nz) return ffail(f, SPV_EINVAL, "spv_filter_read: n must be nx * ny * nz");
  FCU(cudaMemcpyAsync(host_dst, f->cur, n * sizeof(float), cudaMemcpyDeviceToHost, f->stream));
  FCU(cudaStreamSynchronize(f->stream));
  return 0;
}

SPV_API int spv_filter_last_ms(spv_filter *f, float *ms) {
  FBIND();
  if (!ms) return ffail(f, SPV_EINVAL, "spv_filter_last_ms: null pointer");
  if (!f->timed) return ffail(f, SPV_ENODATA, "spv_filter_last_ms: nothing convolved yet");
  FCU(cudaEventSynchronize(f->ev1));
  FCU(cudaEventElapsedTime(ms, f->ev0, f->ev1));
  return 0;
}

SPV_API int spv_filter_last_pass_ms(spv_filter *f, float ms[3], int *passes) {
  FBIND();
  if (!ms) return ffail(f, SPV_EINVAL, "spv_filter_last_pass_ms: null pointer");
  if (!f->timed) return ffail(f, SPV_ENODATA, "spv_filter_last_pass_ms: nothing convolved yet");
  FCU(cudaEventSynchronize(f->ev1));
  ms[0] = ms[1] = ms[2] = 0.f;
  FCU(cudaEventElapsedTime(&ms[0], f->ev0, f->ev_pass[0]));
  if (f->passes == 3) {
    FCU(cudaEventElapsedTime(&ms[1], f->ev_pass[0], f->ev_pass[1]));
    FCU(cudaEventElapsedTime(&ms[2], f->ev_pass[1], f->ev1));
  } else {
    FCU(cudaEventElapsedTime(&ms[1], f->ev_pass[0], f->ev1));
  }
  if (passes) *passes = f->passes;
  return 0;
}

/* knob 0: x and y pass in one kernel wherever the tap counts allow it (default 0: three passes);
 * knob 1: the y / z pass variant: 1 = automatic (4 / 2 / 1 columns per thread as the row length allows), 16 / 32 = one column
 * with that many outputs per thread, 1602 / 1604 = two / four columns; knob 2: the x pass on row pairs, pipelined (2,
 * default) or not (1), or on single rows (0) */
SPV_API int spv_filter_set_tuning(spv_filter *f, int knob, int value) {
  FBIND();
  if (knob == 0) f->fuse_xy = value != 0;
  else if (knob == 1)  // process-wide: 1 = automatic, 16 / 32 = one column, R outputs per thread, 1602 / 1604 = 2 / 4 columns
    filter_axis_wide = (value == 32 || value == 16 || value == 1602 || value == 1604) ? value : 1;
  else if (knob == 2) filter_x_pairs = value == 1 ? 1 : (value ? 2 : 0);  // process-wide
  else return ffail(f, SPV_EINVAL, "spv_filter_set_tuning: unknown knob");
  return 0;
}

SPV_API int spv_filter_launch_count(spv_filter *f, unsigned long long *n) {
  if (!f || !n) return SPV_EINVAL;
  *n = f->launches;
  return 0;
}
