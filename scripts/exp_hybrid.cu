// exp_hybrid.cu -- design experiment: can some ray tiles sample the z-paired volume through the LSU pipe (global
// loads of a linear copy + integer/fp32 filtering in the ALU) while others use the texture pipe, so that the two
// pipes add up?   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o exp_hybrid exp_hybrid.cu
//
// Variants timed on the C2 geometry (512^3 uint16, 1024^2 image, 208 samples per hit ray, 90-frame sweep):
//   tex      : every tile through the texture unit (trimmed ALU version of the product kernel's loop)
//   lsu      : every tile through global loads
//   hyb R    : tile t uses the LSU path iff t % R == 0
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

struct Cam { float ox, oy, oz, ux, uy, uz, vx, vy, vz, wx, wy, wz; };

struct Lin {
  const uint32_t *p;  // [layer][y][pitch] texel = v[z] | v[z+1] << 16
  int pitch, ny, nx, nz;
};

__device__ __forceinline__ bool setup(int x, int y, int W, int H, const Cam c, float N, float &u0, float &v0, float &w0,
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

__device__ __forceinline__ void pixel_of(int &x, int &y, int &tile) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int lx = (lane & 1) | ((lane >> 1) & 2) | ((lane >> 2) & 4);
  const int ly = ((lane >> 1) & 1) | ((lane >> 2) & 2);
  x = blockIdx.x * 16 + (warp & 1) * 8 + lx;
  y = blockIdx.y * 8 + (warp >> 1) * 4 + ly;
  tile = (blockIdx.y * 2 + (warp >> 1)) * (gridDim.x * 2) + blockIdx.x * 2 + (warp & 1);
}

// texture path, trimmed: ~14 ALU ops per sample
__device__ __forceinline__ float march_tex(cudaTextureObject_t tex, float u0, float v0, float w0, float du, float dv,
                                           float dw, int S, unsigned top) {
  float cur = 0.f;
  const float w0h = w0 - 0.5f;
  for (int k = 0; k < S; k += 16) {
    float2 v[16]; float f[16];
    const float kf = (float)k;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float kk = kf + (float)j;
      const float wb = fmaxf(fmaf(kk, dw, w0h), 0.f);
      const unsigned li = __float2uint_rd(wb);
      f[j] = wb - (float)li;
      v[j] = tex2DLayered<float2>(tex, fmaf(kk, du, u0), fmaf(kk, dv, v0), (int)min(li, top));
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) cur = fmaxf(cur, fmaf(f[j], v[j].y - v[j].x, v[j].x));
  }
  return cur * 65535.f;
}

// LSU path: positions in 1/256 texel fixed point (the texture unit's own weight resolution), z-lerp by one dp2a,
// bilinear in fp32
template <int UNROLL>
__device__ __forceinline__ float march_lsu(const Lin L, float u0, float v0, float w0, float du, float dv, float dw,
                                           int S) {
  float cur = 0.f;
  const float X0 = (u0 - 0.5f) * 256.f, Y0 = (v0 - 0.5f) * 256.f, Z0 = (w0 - 0.5f) * 256.f;
  const float DX = du * 256.f, DY = dv * 256.f, DZ = dw * 256.f;
  const float XM = (float)((L.nx - 1) * 256), YM = (float)((L.ny - 1) * 256), ZM = (float)((L.nz - 1) * 256);
  for (int k = 0; k < S; k += UNROLL) {
    uint32_t t00[UNROLL], t10[UNROLL], t01[UNROLL], t11[UNROLL];
    int fx[UNROLL], fy[UNROLL], fz[UNROLL];
    const float kf = (float)k;
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) {
      const float kk = kf + (float)j;
      const int xi = __float2int_rd(fminf(fmaxf(fmaf(kk, DX, X0), 0.f), XM));
      const int yi = __float2int_rd(fminf(fmaxf(fmaf(kk, DY, Y0), 0.f), YM));
      const int zi = __float2int_rd(fminf(fmaxf(fmaf(kk, DZ, Z0), 0.f), ZM));
      fx[j] = xi & 255; fy[j] = yi & 255; fz[j] = zi & 255;
      const uint32_t *p = L.p + ((size_t)((zi >> 8) * L.ny + (yi >> 8)) * L.pitch + (xi >> 8));
      t00[j] = __ldg(p); t10[j] = __ldg(p + 1);
      t01[j] = __ldg(p + L.pitch); t11[j] = __ldg(p + L.pitch + 1);
    }
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) {
      const unsigned wz = (255u - fz[j]) | ((unsigned)fz[j] << 8);  // x*(255-f) + y*f + x = x*(256-f) + y*f
      const float r00 = (float)__dp2a_lo(t00[j], wz, t00[j] & 0xffffu);
      const float r10 = (float)__dp2a_lo(t10[j], wz, t10[j] & 0xffffu);
      const float r01 = (float)__dp2a_lo(t01[j], wz, t01[j] & 0xffffu);
      const float r11 = (float)__dp2a_lo(t11[j], wz, t11[j] & 0xffffu);
      const float ax = (float)fx[j] * (1.f / 256.f), ay = (float)fy[j] * (1.f / 256.f);
      const float top = fmaf(ax, r10 - r00, r00), bot = fmaf(ax, r11 - r01, r01);
      cur = fmaxf(cur, fmaf(ay, bot - top, top));
    }
  }
  return cur * (1.f / 256.f);
}

template <int MODE, int R>  // MODE 0 tex, 1 lsu, 2 hybrid
__global__ void __launch_bounds__(128) march(cudaTextureObject_t tex, Lin L, Cam c, int W, int H, float N, int S, float *out) {
  int x, y, tile; pixel_of(x, y, tile);
  float u0, v0, w0, du, dv, dw, cur = 0.f;
  if (setup(x, y, W, H, c, N, u0, v0, w0, du, dv, dw, S)) {
    const bool lsu = MODE == 1 || (MODE == 2 && (tile % R) == 0);
    if (lsu) cur = march_lsu<8>(L, u0, v0, w0, du, dv, dw, S);
    else cur = march_tex(tex, u0, v0, w0, du, dv, dw, S, (unsigned)L.nz - 1);
  }
  out[y * W + x] = cur;
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
  cudaTextureObject_t t2;
  rd.res.array.array = a2; CK(cudaCreateTextureObject(&t2, &rd, &td, nullptr));
  // linear copy, rows padded by one texel (the x+1 tap of the last column has weight 0), one extra row at the end
  Lin L; L.nx = N; L.ny = N; L.nz = N; L.pitch = N + 8;
  size_t lin_elems = (size_t)N * N * L.pitch + L.pitch + 8;
  std::vector<uint32_t> hl(lin_elems, 0);
  for (int z = 0; z < N; ++z)
    for (int y = 0; y < N; ++y)
      for (int x = 0; x < N; ++x) {
        ushort2 t = h2[((size_t)z * N + y) * N + x];
        hl[((size_t)z * N + y) * L.pitch + x] = (uint32_t)t.x | ((uint32_t)t.y << 16);
      }
  uint32_t *dl; CK(cudaMalloc(&dl, lin_elems * 4)); CK(cudaMemcpy(dl, hl.data(), lin_elems * 4, cudaMemcpyHostToDevice));
  L.p = dl;
  float *out[8]; for (int i = 0; i < 8; ++i) CK(cudaMalloc(&out[i], W * H * 4));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  dim3 grid(W / 16, H / 8), block(128);
  const char *names[] = {"tex", "lsu", "hyb 1/2", "hyb 1/3", "hyb 1/4", "hyb 1/6"};
  for (int which = 0; which < 6; ++which) {
    float total = 0; int frames = 0;
    for (int rep = 0; rep < 2; ++rep) {
      if (rep == 1) cudaEventRecord(e0);
      for (int f = 0; f < 90; ++f) {
        float th = 2.f * 3.14159265f * f / 90 + 1e-3f;
        Cam c; c.ox = 4.f * sinf(th); c.oy = 0; c.oz = 4.f * cosf(th);
        c.wx = -sinf(th); c.wy = 0; c.wz = -cosf(th); c.ux = cosf(th); c.uy = 0; c.uz = -sinf(th); c.vx = 0; c.vy = 1; c.vz = 0;
        switch (which) {
          case 0: march<0, 1><<<grid, block>>>(t2, L, c, W, H, (float)N, S, out[which]); break;
          case 1: march<1, 1><<<grid, block>>>(t2, L, c, W, H, (float)N, S, out[which]); break;
          case 2: march<2, 2><<<grid, block>>>(t2, L, c, W, H, (float)N, S, out[which]); break;
          case 3: march<2, 3><<<grid, block>>>(t2, L, c, W, H, (float)N, S, out[which]); break;
          case 4: march<2, 4><<<grid, block>>>(t2, L, c, W, H, (float)N, S, out[which]); break;
          case 5: march<2, 6><<<grid, block>>>(t2, L, c, W, H, (float)N, S, out[which]); break;
        }
      }
      if (rep == 1) { cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&total, e0, e1); frames = 90; }
    }
    CK(cudaGetLastError());
    printf("%-8s: %.1f us/frame  %.0f frames/s\n", names[which], 1e3f * total / frames, frames / (total * 1e-3f));
  }
  // per-angle for tex / lsu / hyb 1/3
  for (int deg = 0; deg <= 90; deg += 15) {
    float th = deg * 3.14159265f / 180 + 1e-3f;
    Cam c; c.ox = 4.f * sinf(th); c.oy = 0; c.oz = 4.f * cosf(th);
    c.wx = -sinf(th); c.wy = 0; c.wz = -cosf(th); c.ux = cosf(th); c.uy = 0; c.uz = -sinf(th); c.vx = 0; c.vy = 1; c.vz = 0;
    float t[3];
    for (int which = 0; which < 3; ++which) {
      for (int rep = 0; rep < 6; ++rep) {
        if (rep == 1) cudaEventRecord(e0);
        if (which == 0) march<0, 1><<<grid, block>>>(t2, L, c, W, H, (float)N, S, out[0]);
        if (which == 1) march<1, 1><<<grid, block>>>(t2, L, c, W, H, (float)N, S, out[1]);
        if (which == 2) march<2, 3><<<grid, block>>>(t2, L, c, W, H, (float)N, S, out[3]);
      }
      cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&t[which], e0, e1); t[which] /= 5;
    }
    printf("deg %2d: tex %.1f us  lsu %.1f us  hyb1/3 %.1f us\n", deg, 1e3f * t[0], 1e3f * t[1], 1e3f * t[2]);
  }
  std::vector<float> o0(W * H), o1(W * H);
  CK(cudaMemcpy(o0.data(), out[0], W * H * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(o1.data(), out[1], W * H * 4, cudaMemcpyDeviceToHost));
  double md = 0; for (int i = 0; i < W * H; ++i) md = fmax(md, fabs(o0[i] - o1[i]));
  printf("max |tex - lsu| = %.3f (of 60000)\n", md);
  return 0;
}
