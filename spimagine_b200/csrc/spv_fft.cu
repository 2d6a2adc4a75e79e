// spv_fft.cu -- libspimfft.so: FFTProcessor.apply of spimagine (models/imageprocessor.py:82-98) on the device.
// See include/spimfft.h for the reference expression this follows.  Two hand-written passes around a library FFT:
//   pad_kernel       padded[z][y][x] = (float) src[(z - bz) mod nz][(y - by) mod ny][(x - bx) mod nx], b = ceil(d / 2)
//                    (np.pad(mode="wrap") of gputools.pad_to_shape), element type converted on the fly
//   cufftExecR2C     (Pz, Py, Px) real -> (Pz, Py, Px/2 + 1) complex: the volume is real, so F[-k] = conj F[k]
//   spectrum_kernel  out[z][y][x] = s |F[k]|, k = (i + o + P/2) mod P per axis, o = floor(d / 2), s = 1/sqrt(Pz Py Px);
//                    outputs with kx > Px/2 take the magnitude of the mirrored (stored) coefficient; optional
//                    log2(0.001 + .)
// Both passes are HBM-bound streaming kernels (4 + es bytes per padded voxel, 8 + 4 bytes per output voxel).
#include <cuda_runtime.h>
#include <cufft.h>
#include <stdint.h>
#include <string>
#include "spimfft.h"

namespace {

enum { SRC_U8 = 1, SRC_I16 = 2, SRC_U16 = 3, SRC_F32 = 9 };  // SPV_SRC_* of spimcuda.h
enum { E_INVAL = -22, E_NODATA = -61 };

__host__ __device__ inline int next_pow2(int n) {
  int p = 1;
  while (p < n) p <<= 1;
  return p;
}

struct Dims {
  int nx, ny, nz;  // volume
  int px, py, pz;  // padded (powers of two)
};

template <typename T>
__global__ void __launch_bounds__(256) pad_kernel(const T *__restrict__ src, float *__restrict__ dst, Dims d) {
  const int bx = (d.px - d.nx + 1) >> 1, by = (d.py - d.ny + 1) >> 1, bz = (d.pz - d.nz + 1) >> 1;
  const size_t rows = (size_t)d.py * d.pz;
  for (size_t row = blockIdx.y; row < rows; row += gridDim.y) {
    const int y = (int)(row % (size_t)d.py), z = (int)(row / (size_t)d.py);
    int sy = y - by, sz = z - bz;  // in [-n, 2n): one correction wraps it (P < 2n)
    sy += sy < 0 ? d.ny : 0;
    sy -= sy >= d.ny ? d.ny : 0;
    sz += sz < 0 ? d.nz : 0;
    sz -= sz >= d.nz ? d.nz : 0;
    const T *s = src + ((size_t)sz * d.ny + sy) * d.nx;
    float *o = dst + row * d.px;
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < d.px; x += gridDim.x * blockDim.x) {
      int sx = x - bx;
      sx += sx < 0 ? d.nx : 0;
      sx -= sx >= d.nx ? d.nx : 0;
      o[x] = (float)s[sx];
    }
  }
}

// pad_kernel where x needs no padding (nx a power of two, a multiple of 4) and the rows are 4-element aligned: four
// elements per thread through one vector load (a 2-byte load per thread leaves most of every sector request unused:
// 462 us against 250 us for the 512^3 uint16 volume)
template <typename T> struct alignas(4 * sizeof(T)) Vec4 { T v[4]; };
template <typename T>
__global__ void __launch_bounds__(256) pad_rows_kernel(const T *__restrict__ src, float *__restrict__ dst, Dims d) {
  const int by = (d.py - d.ny + 1) >> 1, bz = (d.pz - d.nz + 1) >> 1;
  const size_t rows = (size_t)d.py * d.pz;
  const int quads = d.nx >> 2;                     // a power of two (nx is one)
  const int qt = quads < 256 ? quads : 256;        // threads along a row; the CTA covers 256 / qt rows at a time
  const int rpb = 256 / qt, tq = threadIdx.x % qt, tr = threadIdx.x / qt;
  for (size_t row = (size_t)blockIdx.x * rpb + tr; row < rows; row += (size_t)gridDim.x * rpb) {
    const int y = (int)(row % (size_t)d.py), z = (int)(row / (size_t)d.py);
    int sy = y - by, sz = z - bz;
    sy += sy < 0 ? d.ny : 0;
    sy -= sy >= d.ny ? d.ny : 0;
    sz += sz < 0 ? d.nz : 0;
    sz -= sz >= d.nz ? d.nz : 0;
    const Vec4<T> *s = reinterpret_cast<const Vec4<T> *>(src + ((size_t)sz * d.ny + sy) * d.nx);
    float4 *o = reinterpret_cast<float4 *>(dst + row * d.px);
    for (int q = tq; q < quads; q += qt) {
      const Vec4<T> t = s[q];
      o[q] = make_float4((float)t.v[0], (float)t.v[1], (float)t.v[2], (float)t.v[3]);
    }
  }
}

// One thread per stored coefficient F[kz][ky][kx], kx <= Px/2: its magnitude goes to the output voxel it maps to
// (i = (k - o - P/2) mod P per axis, if inside the crop) and, for 0 < kx < Px/2, to the voxel of the mirrored index -k
// as well (F[-k] = conj F[k] is not stored).  Every coefficient is read once and every output written once: the
// first version computed every output from its own read, i.e. read the half spectrum twice (1.03 GB instead of 0.54).
__global__ void __launch_bounds__(256) spectrum_kernel(const float2 *__restrict__ F, float *__restrict__ out, Dims d,
                                                       float scale, int take_log) {
  constexpr int R = 4;  // rows per pass: R independent loads in flight per thread (one row at a time the pass was
                        // bound by the latency of its single load: 2.4 TB/s)
  const int ox = (d.px - d.nx) >> 1, oy = (d.py - d.ny) >> 1, oz = (d.pz - d.nz) >> 1;
  const int hx = d.px >> 1, cx = hx + 1;  // complex coefficients per row
  const int mx = d.px - 1, my_ = d.py - 1, mz_ = d.pz - 1;
  const size_t rows = (size_t)d.py * d.pz;
  for (size_t row0 = (size_t)blockIdx.x * R; row0 < rows; row0 += (size_t)gridDim.x * R) {
    float *od[R], *om[R];
    bool drow[R], mrow[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const size_t row = row0 + r;
      const int ky = (int)(row % (size_t)d.py), kz = (int)(row / (size_t)d.py);
      const int y = (ky - oy - (d.py >> 1)) & my_, z = (kz - oz - (d.pz >> 1)) & mz_;                  // direct row
      const int y2 = (((d.py - ky) & my_) - oy - (d.py >> 1)) & my_, z2 = (((d.pz - kz) & mz_) - oz - (d.pz >> 1)) & mz_;
      drow[r] = row < rows && y < d.ny && z < d.nz;
      mrow[r] = row < rows && y2 < d.ny && z2 < d.nz;
      od[r] = out + ((size_t)z * d.ny + y) * d.nx;
      om[r] = out + ((size_t)z2 * d.ny + y2) * d.nx;
    }
    for (int kx = threadIdx.x; kx <= hx; kx += blockDim.x) {
      float2 c[R];
#pragma unroll
      for (int r = 0; r < R; ++r)
        c[r] = (drow[r] || mrow[r]) ? F[(row0 + r) * cx + kx] : make_float2(0.f, 0.f);
      const int x = (kx - ox - hx) & mx, x2 = (d.px - kx - ox - hx) & mx;
      const bool mirrored = kx > 0 && kx < hx;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        float v = scale * sqrtf(c[r].x * c[r].x + c[r].y * c[r].y);
        if (take_log) v = log2f(0.001f + v);
        if (drow[r] && x < d.nx) od[r][x] = v;
        if (mirrored && mrow[r] && x2 < d.nx) om[r][x2] = v;
      }
    }
  }
}

}  // namespace

struct spf_plan {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  void *d_src = nullptr;
  size_t src_cap = 0;
  float *d_real = nullptr;
  size_t real_cap = 0;
  float2 *d_freq = nullptr;
  size_t freq_cap = 0;
  float *d_out = nullptr;
  size_t out_cap = 0;
  cufftHandle fft = 0;
  bool have_fft = false;
  Dims dims = {0, 0, 0, 0, 0, 0};
  Dims fft_dims = {0, 0, 0, 0, 0, 0};
  bool have_result = false;
  unsigned long long launches = 0;
  std::string err;
};

static thread_local std::string g_create_err;

static int pfail(spf_plan *p, int code, const std::string &what) {
  (p ? p->err : g_create_err) = what;
  return code;
}
static int pcufail(spf_plan *p, cudaError_t e, const char *where) {
  (p ? p->err : g_create_err) = std::string(where) + ": " + cudaGetErrorString(e);
  cudaGetLastError();
  return (int)e;
}
#define PCU(call)                                        \
  do {                                                   \
    cudaError_t e_ = (call);                             \
    if (e_ != cudaSuccess) return pcufail(p, e_, #call); \
  } while (0)
#define PFFT(call)                                                                              \
  do {                                                                                          \
    cufftResult r_ = (call);                                                                    \
    if (r_ != CUFFT_SUCCESS) return pfail(p, 10000 + (int)r_, std::string(#call) + ": cufftResult " + std::to_string((int)r_)); \
  } while (0)
#define PBIND()                                                    \
  if (!p) return E_INVAL;                                          \
  do {                                                             \
    cudaError_t e_ = cudaSetDevice(p->device);                     \
    if (e_ != cudaSuccess) return pcufail(p, e_, "cudaSetDevice"); \
  } while (0)

template <typename T>
static int grow(spf_plan *p, T **buf, size_t *cap, size_t need) {
  if (need <= *cap) return 0;
  PCU(cudaStreamSynchronize(p->stream));
  if (*buf) cudaFree(*buf);
  *buf = nullptr;
  *cap = 0;
  PCU(cudaMalloc((void **)buf, need));
  *cap = need;
  return 0;
}

extern "C" {

SPF_API const char *spf_last_error(spf_plan *p) { return p ? p->err.c_str() : g_create_err.c_str(); }

SPF_API int spf_destroy(spf_plan *p) {
  if (!p) return 0;
  cudaSetDevice(p->device);
  if (p->stream) cudaStreamSynchronize(p->stream);
  if (p->have_fft) cufftDestroy(p->fft);
  if (p->d_src) cudaFree(p->d_src);
  if (p->d_real) cudaFree(p->d_real);
  if (p->d_freq) cudaFree(p->d_freq);
  if (p->d_out) cudaFree(p->d_out);
  if (p->ev0) cudaEventDestroy(p->ev0);
  if (p->ev1) cudaEventDestroy(p->ev1);
  if (p->stream) cudaStreamDestroy(p->stream);
  cudaGetLastError();
  delete p;
  return 0;
}

SPF_API int spf_create(int device, spf_plan **out) {
  if (!out) return E_INVAL;
  *out = nullptr;
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return pcufail(nullptr, e, "spf_create: cudaSetDevice (libspimfft needs a CUDA device; there is no CPU path)");
  spf_plan *p = new spf_plan;
  p->device = device;
  if ((e = cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking)) != cudaSuccess ||
      (e = cudaEventCreate(&p->ev0)) != cudaSuccess || (e = cudaEventCreate(&p->ev1)) != cudaSuccess) {
    int rc = pcufail(nullptr, e, "spf_create");
    spf_destroy(p);
    return rc;
  }
  *out = p;
  return 0;
}

SPF_API int spf_spectrum(spf_plan *p, const void *src, int on_device, int src_type, int nx, int ny, int nz, int take_log,
                         float *host_dst) {
  PBIND();
  if (!src) return pfail(p, E_INVAL, "spf_spectrum: null data");
  if (nx <= 0 || ny <= 0 || nz <= 0 || nx > (1 << 14) || ny > (1 << 14) || nz > (1 << 14))
    return pfail(p, E_INVAL, "spf_spectrum: bad extent (1 .. 16384 per axis)");
  size_t es;
  switch (src_type) {
    case SRC_U8: es = 1; break;
    case SRC_I16: case SRC_U16: es = 2; break;
    case SRC_F32: es = 4; break;
    default: return pfail(p, E_INVAL, "spf_spectrum: element type must be uint8, int16, uint16 or float32");
  }
  Dims d = {nx, ny, nz, next_pow2(nx), next_pow2(ny), next_pow2(nz)};
  const size_t n = (size_t)nx * ny * nz, np = (size_t)d.px * d.py * d.pz;
  const size_t nc = (size_t)(d.px / 2 + 1) * d.py * d.pz;
  p->have_result = false;
  int rc;
  if ((rc = grow(p, &p->d_real, &p->real_cap, np * sizeof(float)))) return rc;
  if ((rc = grow(p, &p->d_freq, &p->freq_cap, nc * sizeof(float2)))) return rc;
  if ((rc = grow(p, &p->d_out, &p->out_cap, n * sizeof(float)))) return rc;
  const void *dsrc = src;
  if (!on_device) {
    if ((rc = grow(p, &p->d_src, &p->src_cap, n * es))) return rc;
    PCU(cudaMemcpyAsync(p->d_src, src, n * es, cudaMemcpyHostToDevice, p->stream));
    dsrc = p->d_src;
  }
  if (!p->have_fft || p->fft_dims.px != d.px || p->fft_dims.py != d.py || p->fft_dims.pz != d.pz) {
    if (p->have_fft) cufftDestroy(p->fft);
    p->have_fft = false;
    PFFT(cufftPlan3d(&p->fft, d.pz, d.py, d.px, CUFFT_R2C));
    p->have_fft = true;
    p->fft_dims = d;
    PFFT(cufftSetStream(p->fft, p->stream));
  }
  PCU(cudaEventRecord(p->ev0, p->stream));
  {
    const dim3 block(256), grid((unsigned)((d.px + 255) / 256), (unsigned)((size_t)d.py * d.pz < 65535 ? (size_t)d.py * d.pz : 65535));
    const bool rows_fast = d.px == d.nx && d.nx % 4 == 0 && ((uintptr_t)dsrc % (4 * es)) == 0;
    if (rows_fast) {
      const int quads = d.nx / 4, rpb = 256 / (quads < 256 ? quads : 256);
      const size_t ctas = ((size_t)d.py * d.pz + rpb - 1) / rpb;
      const dim3 g4((unsigned)(ctas < (size_t)148 * 64 ? ctas : (size_t)148 * 64));
      switch (src_type) {
        case SRC_U8: pad_rows_kernel<uint8_t><<<g4, block, 0, p->stream>>>((const uint8_t *)dsrc, p->d_real, d); break;
        case SRC_I16: pad_rows_kernel<int16_t><<<g4, block, 0, p->stream>>>((const int16_t *)dsrc, p->d_real, d); break;
        case SRC_U16: pad_rows_kernel<uint16_t><<<g4, block, 0, p->stream>>>((const uint16_t *)dsrc, p->d_real, d); break;
        default: pad_rows_kernel<float><<<g4, block, 0, p->stream>>>((const float *)dsrc, p->d_real, d); break;
      }
    } else
    switch (src_type) {
      case SRC_U8: pad_kernel<uint8_t><<<grid, block, 0, p->stream>>>((const uint8_t *)dsrc, p->d_real, d); break;
      case SRC_I16: pad_kernel<int16_t><<<grid, block, 0, p->stream>>>((const int16_t *)dsrc, p->d_real, d); break;
      case SRC_U16: pad_kernel<uint16_t><<<grid, block, 0, p->stream>>>((const uint16_t *)dsrc, p->d_real, d); break;
      default: pad_kernel<float><<<grid, block, 0, p->stream>>>((const float *)dsrc, p->d_real, d); break;
    }
    PCU(cudaGetLastError());
    ++p->launches;
  }
  PFFT(cufftExecR2C(p->fft, p->d_real, (cufftComplex *)p->d_freq));
  {
    const size_t prow = (size_t)d.py * d.pz;
    const size_t groups = (prow + 3) / 4;  // spectrum_kernel takes 4 rows per pass
    const dim3 block(256), grid((unsigned)(groups < (size_t)148 * 32 ? groups : (size_t)148 * 32));
    const float scale = (float)(1. / sqrt((double)np));
    spectrum_kernel<<<grid, block, 0, p->stream>>>(p->d_freq, p->d_out, d, scale, take_log != 0);
    PCU(cudaGetLastError());
    ++p->launches;
  }
  PCU(cudaEventRecord(p->ev1, p->stream));
  p->dims = d;
  p->have_result = true;
  if (host_dst) PCU(cudaMemcpyAsync(host_dst, p->d_out, n * sizeof(float), cudaMemcpyDeviceToHost, p->stream));
  PCU(cudaStreamSynchronize(p->stream));
  return 0;
}

SPF_API int spf_result_device(spf_plan *p, float **dev) {
  if (!p || !dev) return E_INVAL;
  if (!p->have_result) return pfail(p, E_NODATA, "spf_result_device: no spectrum computed yet");
  *dev = p->d_out;
  return 0;
}

SPF_API int spf_read(spf_plan *p, float *host_dst, size_t n) {
  PBIND();
  if (!host_dst) return E_INVAL;
  if (!p->have_result) return pfail(p, E_NODATA, "spf_read: no spectrum computed yet");
  if (n != (size_t)p->dims.nx * p->dims.ny * p->dims.nz) return pfail(p, E_INVAL, "spf_read: n is not the volume's size");
  PCU(cudaMemcpyAsync(host_dst, p->d_out, n * sizeof(float), cudaMemcpyDeviceToHost, p->stream));
  PCU(cudaStreamSynchronize(p->stream));
  return 0;
}

SPF_API int spf_padded_shape(spf_plan *p, int *px, int *py, int *pz) {
  if (!p || !px || !py || !pz) return E_INVAL;
  if (!p->have_result) return pfail(p, E_NODATA, "spf_padded_shape: no spectrum computed yet");
  *px = p->dims.px;
  *py = p->dims.py;
  *pz = p->dims.pz;
  return 0;
}

SPF_API int spf_last_ms(spf_plan *p, float *ms) {
  PBIND();
  if (!ms) return E_INVAL;
  if (!p->have_result) return pfail(p, E_NODATA, "spf_last_ms: no spectrum computed yet");
  PCU(cudaEventSynchronize(p->ev1));
  PCU(cudaEventElapsedTime(ms, p->ev0, p->ev1));
  return 0;
}

SPF_API int spf_launch_count(spf_plan *p, unsigned long long *n) {
  if (!p || !n) return E_INVAL;
  *n = p->launches;
  return 0;
}

}  // extern "C"
