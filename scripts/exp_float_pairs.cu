// exp_float_pairs.cu -- design experiment (round 2): would float32 volumes gain from the layered-pair copies too?
// A: 3-D R32F array, hardware trilinear (what mip_fast_kernel<f32> samples), 2x2-pixel quads, one frame per launch
// B: 2-D layered RG32F array of pairs along y, bilinear + fp32 lerp, 4x1 row quads, 1 and 10 frames per launch
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o exp_float_pairs.bin exp_float_pairs.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

struct Cam { float ox, oy, oz, ux, uy, uz, vx, vy, vz, wx, wy, wz; };
constexpr int MAXF = 16;
struct Cams { Cam c[MAXF]; };

__device__ __forceinline__ bool setup(int x, int y, int W, int H, const Cam &c, float N, float &u0, float &v0, float &w0,
                                      float &du, float &dv, float &dw, int S) {
  float sx = ((float)x / W * 2.f - 1.f) * 0.57735f, sy = ((float)y / H * 2.f - 1.f) * 0.57735f;
  float dx = c.wx + sx * c.ux + sy * c.vx, dy = c.wy + sx * c.uy + sy * c.vy, dz = c.wz + sx * c.uz + sy * c.vz;
  float inv = rsqrtf(dx * dx + dy * dy + dz * dz);
  dx *= inv; dy *= inv; dz *= inv;
  float tn = -1e30f, tf = 1e30f;
  float o[3] = {c.ox, c.oy, c.oz}, d[3] = {dx, dy, dz};
  for (int a = 0; a < 3; ++a) {
    float i = 1.f / d[a];
    float t0 = (-1.f - o[a]) * i, t1 = (1.f - o[a]) * i;
    tn = fmaxf(tn, fminf(t0, t1));
    tf = fminf(tf, fmaxf(t0, t1));
  }
  if (!(tf > tn)) return false;
  float dt = (tf - tn) / (S - 16);
  u0 = (0.5f * (1.f + c.ox + tn * dx)) * N; v0 = (0.5f * (1.f + c.oy + tn * dy)) * N; w0 = (0.5f * (1.f + c.oz + tn * dz)) * N;
  du = 0.5f * dt * dx * N; dv = 0.5f * dt * dy * N; dw = 0.5f * dt * dz * N;
  return true;
}


struct Shape { int qw, qh, wx, wy, cx, cy; };
template <int MODE>
__global__ void __launch_bounds__(128) marchf(cudaTextureObject_t tex, const __grid_constant__ Cams cams, int W, int H, int N,
                                              int S, Shape sh, float *out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = lane >> 2, i = lane & 3, nqx = sh.wx / sh.qw;
  const int lx = (q % nqx) * sh.qw + (i % sh.qw), ly = (q / nqx) * sh.qh + (i / sh.qw);
  const int f = blockIdx.y;
  const int x = (blockIdx.x * sh.cx + (warp % sh.cx)) * sh.wx + lx;
  const int y = (blockIdx.z * sh.cy + (warp / sh.cx)) * sh.wy + ly;
  float u0, v0, w0, du, dv, dw, cur = 0.f;
  if (setup(x, y, W, H, cams.c[f], (float)N, u0, v0, w0, du, dv, dw, S)) {
    for (int k = 0; k < S; k += 16) {
      if (MODE == 0) {
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) { float kk = (float)(k + j); v[j] = tex3D<float>(tex, fmaf(kk, du, u0), fmaf(kk, dv, v0), fmaf(kk, dw, w0)); }
#pragma unroll
        for (int j = 0; j < 16; ++j) cur = fmaxf(cur, v[j]);
      } else {
        float2 v[16]; float fr[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float kk = (float)(k + j);
          float wb = fmaf(kk, dv, v0) - 0.5f;   // pairs along y
          float fl = floorf(wb);
          fr[j] = wb - fl;
          int layer = min(max((int)fl, 0), N - 1);
          v[j] = tex2DLayered<float2>(tex, fmaf(kk, du, u0), fmaf(kk, dw, w0), layer);
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) cur = fmaxf(cur, fmaf(fr[j], v[j].y - v[j].x, v[j].x));
      }
    }
  }
  out[((size_t)f * H + y) * W + x] = cur;
}

__global__ void count_hits(Cam c, int W, int H, int N, int S, unsigned long long *n) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  float a, b, d, e, f, g;
  if (x < W && setup(x, y, W, H, c, (float)N, a, b, d, e, f, g, S)) atomicAdd(n, 1ull);
}

static Cam cam_at(float deg) {
  float t = deg * 3.14159265358979f / 180.f, s = sinf(t), co = cosf(t);
  Cam c;
  c.ox = 4 * s; c.oy = 0; c.oz = 4 * co;
  c.wx = -s; c.wy = 0; c.wz = -co;
  c.ux = co; c.uy = 0; c.uz = -s;
  c.vx = 0; c.vy = 1; c.vz = 0;
  return c;
}

int main(int argc, char **argv) {
  setvbuf(stdout, nullptr, _IONBF, 0);
  const int N = argc > 1 ? atoi(argv[1]) : 512;
  const int W = argc > 2 ? atoi(argv[2]) : 1024, H = W, S = 208;
  CK(cudaSetDevice(0));
  const size_t ntex = (size_t)N * N * N;
  std::vector<float> vol(ntex);
  unsigned s = 12345u;
  for (size_t i = 0; i < ntex; ++i) { s = s * 1664525u + 1013904223u; vol[i] = (float)(s >> 8) / 16777216.f; }
  cudaArray_t a3, ap;
  cudaChannelFormatDesc c1 = cudaCreateChannelDesc(32, 0, 0, 0, cudaChannelFormatKindFloat), c2 = cudaCreateChannelDesc(32, 32, 0, 0, cudaChannelFormatKindFloat);
  CK(cudaMalloc3DArray(&a3, &c1, make_cudaExtent(N, N, N), cudaArrayDefault));
  CK(cudaMalloc3DArray(&ap, &c2, make_cudaExtent(N, N, N), cudaArrayLayered));
  {
    cudaMemcpy3DParms p; memset(&p, 0, sizeof p);
    p.srcPtr = make_cudaPitchedPtr(vol.data(), (size_t)N * 4, N, N); p.dstArray = a3; p.extent = make_cudaExtent(N, N, N); p.kind = cudaMemcpyHostToDevice;
    CK(cudaMemcpy3D(&p));
    std::vector<float2> pair(ntex);
    for (size_t i = 0; i < ntex; ++i) pair[i] = make_float2(vol[i], vol[(i + (size_t)N * N) % ntex]);
    p.srcPtr = make_cudaPitchedPtr(pair.data(), (size_t)N * 8, N, N); p.dstArray = ap;
    CK(cudaMemcpy3D(&p));
  }
  cudaResourceDesc rd; memset(&rd, 0, sizeof rd); rd.resType = cudaResourceTypeArray;
  cudaTextureDesc td; memset(&td, 0, sizeof td);
  td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
  td.filterMode = cudaFilterModeLinear; td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
  cudaTextureObject_t t3, tp;
  rd.res.array.array = a3; CK(cudaCreateTextureObject(&t3, &rd, &td, nullptr));
  rd.res.array.array = ap; CK(cudaCreateTextureObject(&tp, &rd, &td, nullptr));
  float *out; CK(cudaMalloc(&out, (size_t)MAXF * W * H * 4));
  unsigned long long *d_n; CK(cudaMalloc(&d_n, 8));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  auto hits = [&](Cam c) {
    CK(cudaMemset(d_n, 0, 8));
    count_hits<<<dim3((W + 255) / 256, H), 256>>>(c, W, H, N, S, d_n);
    unsigned long long nh; CK(cudaMemcpy(&nh, d_n, 8, cudaMemcpyDeviceToHost));
    return (double)nh;
  };
  const Shape s22 = {2, 2, 8, 4, 2, 2}, srow = {4, 1, 16, 2, 1, 4};
  for (int mode = 0; mode < 2; ++mode)
    for (int F : {1, 10}) {
      double samples = 0, total = 0;
      for (int g = 0; g * F < 20; ++g) {
        Cams cs;
        for (int f = 0; f < F; ++f) { cs.c[f] = cam_at(18.f * (g * F + f)); samples += hits(cs.c[f]) * S; }
        const Shape sh = mode == 0 ? s22 : srow;
        const dim3 grid(W / (sh.cx * sh.wx), F, H / (sh.cy * sh.wy));
        float ms = 0;
        for (int it = 0; it < 2; ++it) {
          CK(cudaEventRecord(e0));
          for (int r = 0; r < 3; ++r) {
            if (mode == 0) marchf<0><<<grid, 128>>>(t3, cs, W, H, N, S, sh, out);
            else marchf<1><<<grid, 128>>>(tp, cs, W, H, N, S, sh, out);
          }
          CK(cudaEventRecord(e1));
          CK(cudaEventSynchronize(e1));
          CK(cudaGetLastError());
          CK(cudaEventElapsedTime(&ms, e0, e1));
        }
        total += ms * 1000.f / 3;
      }
      printf("%d^3 float32 -> %d^2, %s, %2d frame(s) per launch: %.1f us per frame, %.0f Gsamples/s\n", N, W,
             mode == 0 ? "3-D R32F, trilinear, 2x2 quads      " : "layered RG32F pairs along y, row quads", F, total / 20, samples / total * 1e-3);
    }
  return 0;
}
