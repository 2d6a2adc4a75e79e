// exp_texrate.cu -- design experiment: is a z-paired 2-D layered texture (RG16: texel = {v[z], v[z+1]}, one
// BILINEAR fetch + one fp32 lerp per sample) faster than a 3-D R16 texture (one TRILINEAR fetch per sample) for
// a ray-march access pattern?   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o exp_texrate exp_texrate.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

struct Cam { float ox, oy, oz, ux, uy, uz, vx, vy, vz, wx, wy, wz; };

__device__ __forceinline__ bool setup(int x, int y, int W, int H, const Cam c, float N, float &u0, float &v0, float &w0,
                                      float &du, float &dv, float &dw, int S) {
  // pinhole camera at c.o looking along c.w, fov 60
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

__device__ __forceinline__ void pixel_of(int &x, int &y) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int lx = (lane & 1) | ((lane >> 1) & 2) | ((lane >> 2) & 4);
  const int ly = ((lane >> 1) & 1) | ((lane >> 2) & 2);
  x = blockIdx.x * 16 + (warp & 1) * 8 + lx;
  y = blockIdx.y * 8 + (warp >> 1) * 4 + ly;
}

__global__ void __launch_bounds__(128) march3d(cudaTextureObject_t tex, Cam c, int W, int H, float N, int S, float *out) {
  int x, y; pixel_of(x, y);
  float u0, v0, w0, du, dv, dw, cur = 0.f;
  if (setup(x, y, W, H, c, N, u0, v0, w0, du, dv, dw, S)) {
    for (int k = 0; k < S; k += 16) {
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) { float kk = (float)(k + j); v[j] = tex3D<float>(tex, fmaf(kk, du, u0), fmaf(kk, dv, v0), fmaf(kk, dw, w0)); }
#pragma unroll
      for (int j = 0; j < 16; ++j) cur = fmaxf(cur, v[j]);
    }
  }
  out[y * W + x] = cur * 65535.f;
}

__global__ void __launch_bounds__(128) march2dl(cudaTextureObject_t tex, Cam c, int W, int H, float N, int S, float *out) {
  int x, y; pixel_of(x, y);
  float u0, v0, w0, du, dv, dw, cur = 0.f;
  if (setup(x, y, W, H, c, N, u0, v0, w0, du, dv, dw, S)) {
    for (int k = 0; k < S; k += 16) {
      float2 v[16]; float f[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float kk = (float)(k + j);
        float wb = fmaf(kk, dw, w0) - 0.5f;
        float fl = floorf(wb);
        f[j] = wb - fl;
        int layer = min(max((int)fl, 0), (int)N - 1);
        v[j] = tex2DLayered<float2>(tex, fmaf(kk, du, u0), fmaf(kk, dv, v0), layer);
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) cur = fmaxf(cur, fmaf(f[j], v[j].y - v[j].x, v[j].x));
    }
  }
  out[y * W + x] = cur * 65535.f;
}

int main(int argc, char **argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 512, W = 1024, H = 1024, S = 208;
  size_t nvox = (size_t)N * N * N;
  std::vector<unsigned short> h(nvox);
  for (size_t i = 0; i < nvox; ++i) {
    int x = i % N, y = (i / N) % N, z = i / ((size_t)N * N);
    float fx = 2.f * x / N - 1, fy = 2.f * y / N - 1, fz = 2.f * z / N - 1;
    h[i] = (unsigned short)(60000.f * expf(-3.f * (fx * fx + fy * fy + fz * fz)) + (rand() & 255));
  }
  // 3-D R16
  cudaArray_t a3; cudaChannelFormatDesc c1 = cudaCreateChannelDesc(16, 0, 0, 0, cudaChannelFormatKindUnsigned);
  CK(cudaMalloc3DArray(&a3, &c1, make_cudaExtent(N, N, N), 0));
  cudaMemcpy3DParms p = {0};
  p.srcPtr = make_cudaPitchedPtr(h.data(), N * 2, N, N); p.dstArray = a3; p.extent = make_cudaExtent(N, N, N); p.kind = cudaMemcpyHostToDevice;
  CK(cudaMemcpy3D(&p));
  // layered RG16: layer z holds {v[z], v[min(z+1,N-1)]}
  std::vector<ushort2> h2(nvox);
  for (int z = 0; z < N; ++z) {
    int z1 = z + 1 < N ? z + 1 : N - 1;
    for (size_t i = 0; i < (size_t)N * N; ++i) h2[(size_t)z * N * N + i] = make_ushort2(h[(size_t)z * N * N + i], h[(size_t)z1 * N * N + i]);
  }
  cudaArray_t a2; cudaChannelFormatDesc c2 = cudaCreateChannelDesc(16, 16, 0, 0, cudaChannelFormatKindUnsigned);
  CK(cudaMalloc3DArray(&a2, &c2, make_cudaExtent(N, N, N), cudaArrayLayered));
  cudaMemcpy3DParms q = {0};
  q.srcPtr = make_cudaPitchedPtr(h2.data(), N * 4, N, N); q.dstArray = a2; q.extent = make_cudaExtent(N, N, N); q.kind = cudaMemcpyHostToDevice;
  CK(cudaMemcpy3D(&q));
  cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeArray;
  cudaTextureDesc td = {}; td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
  td.filterMode = cudaFilterModeLinear; td.readMode = cudaReadModeNormalizedFloat; td.normalizedCoords = 0;
  cudaTextureObject_t t3, t2;
  rd.res.array.array = a3; CK(cudaCreateTextureObject(&t3, &rd, &td, nullptr));
  rd.res.array.array = a2; CK(cudaCreateTextureObject(&t2, &rd, &td, nullptr));
  float *out3, *out2; CK(cudaMalloc(&out3, W * H * 4)); CK(cudaMalloc(&out2, W * H * 4));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  dim3 grid(W / 16, H / 8), block(128);
  for (int which = 0; which < 2; ++which) {
    float total = 0; int frames = 0;
    for (int rep = 0; rep < 2; ++rep) {
      if (rep == 1) cudaEventRecord(e0);
      for (int f = 0; f < 90; ++f) {
        float th = 2.f * 3.14159265f * f / 90 + 1e-3f;
        Cam c; c.ox = 4.f * sinf(th); c.oy = 0; c.oz = 4.f * cosf(th);
        c.wx = -sinf(th); c.wy = 0; c.wz = -cosf(th); c.ux = cosf(th); c.uy = 0; c.uz = -sinf(th); c.vx = 0; c.vy = 1; c.vz = 0;
        if (which == 0) march3d<<<grid, block>>>(t3, c, W, H, (float)N, S, out3);
        else march2dl<<<grid, block>>>(t2, c, W, H, (float)N, S, out2);
      }
      if (rep == 1) { cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&total, e0, e1); frames = 90; }
    }
    printf("%s: %.1f us/frame  %.0f frames/s\n", which == 0 ? "3-D R16 trilinear      " : "layered RG16 bilinear+lerp", 1e3f * total / frames, frames / (total * 1e-3f));
  }
  std::vector<float> o3(W * H), o2(W * H);
  CK(cudaMemcpy(o3.data(), out3, W * H * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(o2.data(), out2, W * H * 4, cudaMemcpyDeviceToHost));
  double md = 0; for (int i = 0; i < W * H; ++i) md = fmax(md, fabs(o3[i] - o2[i]));
  printf("max |3d - layered| = %.3f (of 60000)\n", md);
  return 0;
}
