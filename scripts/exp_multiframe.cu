// exp_multiframe.cu -- design experiment (round 2): do the texture path's L1 misses get cheaper when they hit L2?
//
// mip_fast_kernel runs at 0.38-0.42 of the cache-resident texture rate although neither HBM (3 TB/s of 6.5) nor the
// L1TEX data stage is saturated: nearly every quad request touches a sector that comes from DRAM (the z-paired 512^3
// volume is four times the L2) and the texture unit keeps a bounded number of requests in flight.  A cudaArray cannot be
// prefetched into L2 from the side (no public route gives it a linear address: scripts/exp_alias_probe.cu), but several
// FRAMES can share what one of them pulled in: for a sweep about the y axis the rays of image rows [j, j+8) cross the
// same wedge of y-rows of the volume at every angle.  One launch renders F frames; CTAs are ordered (tile row, frame,
// tile x) or (tile row, tile x, frame), so the F frames' CTAs of one tile row run together and the wedge (~12 MB) is read
// from DRAM once instead of F times.
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o exp_multiframe.bin exp_multiframe.cu
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

// One launch renders F frames; blockIdx = (tile x, frame, tile row).
//   LAX   volume axis that becomes the array's layer index (0 x, 1 y, 2 z); the texel holds {v[l], v[l+1]} along it
//   TR    swap the two in-layer axes (array x <-> array y)
//   qw x qh  pixels of a hardware quad (lanes 4i..4i+3): 2x2, 4x1 or 1x4;  wx x wy  pixels of a warp tile; CTA = 2x2 warps
// The array itself is the same noise cube for every variant -- only the coordinates are permuted, which is what a
// permuted copy of the volume would look like to the texture unit.
struct Shape { int qw, qh, wx, wy, cx, cy; };  // cx x cy warps per CTA
template <int LAX, int TR, int UN, int NT, int MINB = 1>
__global__ void __launch_bounds__(NT, MINB) march(cudaTextureObject_t tex, const __grid_constant__ Cams cams, int W, int H, int N,
                                             int S, int row0, Shape sh, float *out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = lane >> 2, i = lane & 3, nqx = sh.wx / sh.qw;
  const int lx = (q % nqx) * sh.qw + (i % sh.qw), ly = (q / nqx) * sh.qh + (i / sh.qw);
  const int f = blockIdx.y;
  const int x = (blockIdx.x * sh.cx + (warp % sh.cx)) * sh.wx + lx;
  const int y = ((blockIdx.z + row0) * sh.cy + (warp / sh.cx)) * sh.wy + ly;
  float u0, v0, w0, du, dv, dw, cur = 0.f;
  if (setup(x, y, W, H, cams.c[f], (float)N, u0, v0, w0, du, dv, dw, S)) {
    // (a, b) in-layer coordinates, c the layer coordinate
    float a0, b0, c0, da, db, dc;
    if (LAX == 2) { a0 = u0; da = du; b0 = v0; db = dv; c0 = w0; dc = dw; }
    else if (LAX == 1) { a0 = u0; da = du; b0 = w0; db = dw; c0 = v0; dc = dv; }
    else { a0 = w0; da = dw; b0 = v0; db = dv; c0 = u0; dc = du; }
    if (TR) { float t = a0; a0 = b0; b0 = t; t = da; da = db; db = t; }
    for (int k = 0; k < S; k += UN) {
      float2 v[UN]; float fr[UN];
#pragma unroll
      for (int j = 0; j < UN; ++j) {
        float kk = (float)(k + j);
        float wb = fmaf(kk, dc, c0) - 0.5f;
        float fl = floorf(wb);
        fr[j] = wb - fl;
        int layer = min(max((int)fl, 0), N - 1);
        v[j] = tex2DLayered<float2>(tex, fmaf(kk, da, a0), fmaf(kk, db, b0), layer);
      }
#pragma unroll
      for (int j = 0; j < UN; ++j) cur = fmaxf(cur, fmaf(fr[j], v[j].y - v[j].x, v[j].x));
    }
  }
  if (MINB == 11) {  // the library kernel's 1 KB of static shared memory (tile staging): does the L1 carve-out matter?
    __shared__ float stage[4][64];
    stage[warp][lane] = cur;
    stage[warp][32 + lane] = cur + 1.f;
    __syncwarp();
    cur = stage[warp][lane ^ 1] + stage[warp][32 + (lane ^ 1)] * 0.f;
  }
  out[((size_t)f * H + y) * W + x] = cur * 65535.f;
}


// Quads of 2 pixels x 2 sample PHASES: lanes (4i, 4i+1) are two neighbouring pixels of a row taking the even samples,
// lanes (4i+2, 4i+3) the same two pixels taking the odd samples; the two halves of a ray's maximum meet in a shuffle.
// A warp covers 16 x 1 pixels, a CTA (4 warps) 16 x 4.  The request's footprint is 2.2 x 2.6 texels instead of 4.5 x 2.
template <int LAX>
__global__ void __launch_bounds__(128) march_phase(cudaTextureObject_t tex, const __grid_constant__ Cams cams, int W, int H, int N,
                                                   int S, float *out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = lane >> 2, ph = (lane >> 1) & 1;
  const int f = blockIdx.y;
  const int x = blockIdx.x * 16 + q * 2 + (lane & 1);
  const int y = blockIdx.z * 4 + warp;
  float u0, v0, w0, du, dv, dw, cur = 0.f;
  if (setup(x, y, W, H, cams.c[f], (float)N, u0, v0, w0, du, dv, dw, S)) {
    float a0, b0, c0, da, db, dc;
    if (LAX == 2) { a0 = u0; da = du; b0 = v0; db = dv; c0 = w0; dc = dw; }
    else { a0 = u0; da = du; b0 = w0; db = dw; c0 = v0; dc = dv; }
    for (int k = 0; k < S; k += 32) {
      float2 v[16]; float fr[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        int kj = k + 2 * j + ph;
        float kk = (float)(kj < S ? kj : S - 1);
        float wb = fmaf(kk, dc, c0) - 0.5f;
        float fl = floorf(wb);
        fr[j] = wb - fl;
        int layer = min(max((int)fl, 0), N - 1);
        v[j] = tex2DLayered<float2>(tex, fmaf(kk, da, a0), fmaf(kk, db, b0), layer);
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) cur = fmaxf(cur, fmaf(fr[j], v[j].y - v[j].x, v[j].x));
    }
  }
  cur = fmaxf(cur, __shfl_xor_sync(0xffffffffu, cur, 2));
  if (!ph) out[((size_t)f * H + y) * W + x] = cur * 65535.f;
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
// a sweep about the x axis instead: image rows no longer map to fixed y-rows of the volume
static Cam cam_at_x(float deg) {
  float t = deg * 3.14159265358979f / 180.f, s = sinf(t), co = cosf(t);
  Cam c;
  c.ox = 0; c.oy = 4 * s; c.oz = 4 * co;
  c.wx = 0; c.wy = -s; c.wz = -co;
  c.ux = 1; c.uy = 0; c.uz = 0;
  c.vx = 0; c.vy = co; c.vz = -s;
  return c;
}

int main(int argc, char **argv) {
  setvbuf(stdout, nullptr, _IONBF, 0);
  const int N = argc > 1 ? atoi(argv[1]) : 512;
  const int W = 1024, H = 1024, S = 208;
  CK(cudaSetDevice(0));
  const size_t ntex = (size_t)N * N * N;
  cudaArray_t arr; cudaChannelFormatDesc cd = cudaCreateChannelDesc(16, 16, 0, 0, cudaChannelFormatKindUnsigned);
  CK(cudaMalloc3DArray(&arr, &cd, make_cudaExtent(N, N, N), cudaArrayLayered));
  {
    std::vector<unsigned short> vol(ntex);
    unsigned s = 12345u;
    for (size_t i = 0; i < ntex; ++i) { s = s * 1664525u + 1013904223u; vol[i] = (unsigned short)(s >> 16); }
    std::vector<unsigned> pair(ntex);
    for (size_t l = 0; l < (size_t)N; ++l) {
      size_t l1 = l + 1 < (size_t)N ? l + 1 : l;
      for (size_t i = 0; i < (size_t)N * N; ++i) pair[l * N * N + i] = vol[l * N * N + i] | ((unsigned)vol[l1 * N * N + i] << 16);
    }
    cudaMemcpy3DParms p; memset(&p, 0, sizeof p);
    p.srcPtr = make_cudaPitchedPtr(pair.data(), (size_t)N * 4, N, N); p.dstArray = arr; p.extent = make_cudaExtent(N, N, N);
    p.kind = cudaMemcpyHostToDevice;
    CK(cudaMemcpy3D(&p));
  }
  cudaResourceDesc rd; memset(&rd, 0, sizeof rd);
  rd.resType = cudaResourceTypeArray; rd.res.array.array = arr;
  cudaTextureDesc td; memset(&td, 0, sizeof td);
  td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
  td.filterMode = cudaFilterModeLinear; td.readMode = cudaReadModeNormalizedFloat; td.normalizedCoords = 0;
  cudaTextureObject_t tex; CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));

  float *out; CK(cudaMalloc(&out, (size_t)MAXF * W * H * 4));
  unsigned long long *d_n; CK(cudaMalloc(&d_n, 8));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));

  auto hits = [&](Cam c) {
    CK(cudaMemset(d_n, 0, 8));
    count_hits<<<dim3(W / 256, H), 256>>>(c, W, H, N, S, d_n);
    unsigned long long nh; CK(cudaMemcpy(&nh, d_n, 8, cudaMemcpyDeviceToHost));
    return (double)nh;
  };
  struct Variant { const char *name; int lax, tr, un, nt; Shape sh; };
  const Variant vars[] = {
      {"z-layered 2x2 quads  8x4 warp 2x2 warps 16 in flight", 2, 0, 16, 128, {2, 2, 8, 4, 2, 2}},
      {"y-layered 4x1 quads 16x2 warp 2x2 warps 16 in flight", 1, 0, 16, 128, {4, 1, 16, 2, 2, 2}},
      {"y-layered 4x1 quads 16x2 warp 1x4 warps 16 in flight", 1, 0, 16, 128, {4, 1, 16, 2, 1, 4}},
      {"y-layered 4x1 quads 16x2 warp 4x1 warps 16 in flight", 1, 0, 16, 128, {4, 1, 16, 2, 4, 1}},
      {"y-layered 4x1 quads 16x2 warp 1x2 warps 16 in flight", 1, 0, 16, 64, {4, 1, 16, 2, 1, 2}},
      {"y-layered 4x1 quads 16x2 warp 2x4 warps 16 in flight", 1, 0, 16, 256, {4, 1, 16, 2, 2, 4}},
      {"y-layered 4x1 quads 16x2 warp 1x8 warps 16 in flight", 1, 0, 16, 256, {4, 1, 16, 2, 1, 8}},
      {"y-layered 4x1 quads 16x2 warp 2x2 warps  8 in flight", 1, 0, 8, 128, {4, 1, 16, 2, 2, 2}},
      {"y-layered 4x1 quads 16x2 warp 2x2 warps 26 in flight", 1, 0, 26, 128, {4, 1, 16, 2, 2, 2}},
      {"y-layered 4x1 quads 32x1 warp 1x4 warps 16 in flight", 1, 0, 16, 128, {4, 1, 32, 1, 1, 4}},
      {"y-layered 4x1 quads 32x1 warp 1x8 warps 16 in flight", 1, 0, 16, 256, {4, 1, 32, 1, 1, 8}},
      {"y-layered 4x1 quads  8x4 warp 2x2 warps 16 in flight", 1, 0, 16, 128, {4, 1, 8, 4, 2, 2}},
      {"y-layered 4x1 quads  8x4 warp 4x1 warps 16 in flight", 1, 0, 16, 128, {4, 1, 8, 4, 4, 1}},
  };
  const int NV = sizeof vars / sizeof vars[0];
  // us per launch of F frames
  auto run = [&](const Variant &v, const Cams &cs, int F, int reps) {
    float ms = 0;
    const dim3 grid(W / (v.sh.cx * v.sh.wx), F, H / (v.sh.cy * v.sh.wy));
    for (int it = 0; it < 2; ++it) {
      CK(cudaEventRecord(e0));
      for (int r = 0; r < (it ? reps : 2); ++r) {
        if (v.lax == 2) march<2, 0, 16, 128><<<grid, 128>>>(tex, cs, W, H, N, S, 0, v.sh, out);
        else if (v.un == 8) march<1, 0, 8, 128><<<grid, 128>>>(tex, cs, W, H, N, S, 0, v.sh, out);
        else if (v.un == 26) march<1, 0, 26, 128><<<grid, 128>>>(tex, cs, W, H, N, S, 0, v.sh, out);
        else if (v.nt == 64) march<1, 0, 16, 64><<<grid, 64>>>(tex, cs, W, H, N, S, 0, v.sh, out);
        else if (v.nt == 256) march<1, 0, 16, 256><<<grid, 256>>>(tex, cs, W, H, N, S, 0, v.sh, out);
        else march<1, 0, 16, 128><<<grid, 128>>>(tex, cs, W, H, N, S, 0, v.sh, out);
      }
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      CK(cudaGetLastError());
      CK(cudaEventElapsedTime(&ms, e0, e1));
    }
    return ms * 1000.f / reps;
  };
  if (argc > 2 && !strcmp(argv[2], "occ")) {  // resident CTAs per SM (register budget) of the y-layered row-quad kernel
    const Shape sh = {4, 1, 16, 2, 1, 4};
    for (int minb : {10, 11, 10, 11, 12}) {
      double samples = 0, total = 0;
      for (int g = 0; g < 2; ++g) {
        Cams cs;
        for (int f = 0; f < 10; ++f) { cs.c[f] = cam_at(18.f * (g * 10 + f)); samples += hits(cs.c[f]) * S; }
        const dim3 grid(W / 16, 10, H / 8);
        float ms = 0;
        for (int it = 0; it < 2; ++it) {
          CK(cudaEventRecord(e0));
          for (int r = 0; r < 3; ++r) {
            switch (minb) {
              case 2: march<1, 0, 16, 128, 2><<<grid, 128, 40 * 1024>>>(tex, cs, W, H, N, S, 0, sh, out); break;   // + shared memory: 2 / 4 CTAs really
              case 4: march<1, 0, 16, 128, 4><<<grid, 128, 40 * 1024>>>(tex, cs, W, H, N, S, 0, sh, out); break;
              case 6: march<1, 0, 16, 128, 6><<<grid, 128, 32 * 1024>>>(tex, cs, W, H, N, S, 0, sh, out); break;
              case 8: march<1, 0, 16, 128, 8><<<grid, 128, 24 * 1024>>>(tex, cs, W, H, N, S, 0, sh, out); break;
              case 10: march<1, 0, 16, 128, 10><<<grid, 128>>>(tex, cs, W, H, N, S, 0, sh, out); break;
              case 11: march<1, 0, 16, 128, 11><<<grid, 128>>>(tex, cs, W, H, N, S, 0, sh, out); break;
              case 12: march<1, 0, 16, 128, 12><<<grid, 128>>>(tex, cs, W, H, N, S, 0, sh, out); break;
              default: march<1, 0, 16, 128, 16><<<grid, 128>>>(tex, cs, W, H, N, S, 0, sh, out); break;
            }
          }
          CK(cudaEventRecord(e1));
          CK(cudaEventSynchronize(e1));
          CK(cudaGetLastError());
          CK(cudaEventElapsedTime(&ms, e0, e1));
        }
        total += ms * 1000.f / 3;
      }
      printf("launch bounds (128, %2d): %.1f us per frame, %.0f Gs/s\n", minb, total / 20, samples / total * 1e-3);
    }
    return 0;
  }
  if (argc > 3) {  // one variant, one launch size, for a profiler: exp_multiframe.bin 512 <variant> <F>
    const int vi = atoi(argv[2]), F = atoi(argv[3]);
    Cams cs;
    for (int f = 0; f < F; ++f) cs.c[f] = cam_at(18.f * f);
    printf("%s F=%d: %.1f us per launch\n", vars[vi].name, F, run(vars[vi], cs, F, 3));
    return 0;
  }
  // the bench's own frame set: 20 frames 18 degrees apart, in launches of F
  printf("20 frames 18 degrees apart (the 20-step bench sweep), us per frame | Gsamples/s, by frames per launch\n%-54s", "");
  const int Fs[] = {1, 2, 4, 5, 10, 16};
  for (int F : Fs) printf("    F = %2d    ", F);
  printf("\n");
  for (int vi = 0; vi < NV; ++vi) {
    printf("%s", vars[vi].name);
    for (int F : Fs) {
      double samples = 0, total = 0;
      for (int g = 0; g * F < 20; ++g) {
        Cams cs; int n = 0;
        for (int f = 0; f < F && g * F + f < 20; ++f, ++n) { cs.c[f] = cam_at(18.f * (g * F + f)); samples += hits(cs.c[f]) * S; }
        total += run(vars[vi], cs, n, 3);
      }
      printf("  %6.1f | %4.0f", total / 20, samples / total * 1e-3);
    }
    printf("\n");
  }
  // ---- quads of 2 pixels x 2 sample phases ----
  for (int lax : {1, 2}) {
    printf("%s-layered, quads of 2 pixels x 2 sample phases, 16x1 warps:", lax == 1 ? "y" : "z");
    for (int F : {1, 2, 4, 5, 10, 16}) {
      double samples = 0, total = 0;
      for (int g = 0; g * F < 20; ++g) {
        Cams cs; int n = 0;
        for (int f = 0; f < F && g * F + f < 20; ++f, ++n) { cs.c[f] = cam_at(18.f * (g * F + f)); samples += hits(cs.c[f]) * S; }
        const dim3 grid(W / 16, n, H / 4);
        float ms = 0;
        for (int it = 0; it < 2; ++it) {
          CK(cudaEventRecord(e0));
          for (int r = 0; r < 3; ++r) {
            if (lax == 1) march_phase<1><<<grid, 128>>>(tex, cs, W, H, N, S, out);
            else march_phase<2><<<grid, 128>>>(tex, cs, W, H, N, S, out);
          }
          CK(cudaEventRecord(e1));
          CK(cudaEventSynchronize(e1));
          CK(cudaGetLastError());
          CK(cudaEventElapsedTime(&ms, e0, e1));
        }
        total += ms * 1000.f / 3;
      }
      printf("  F=%d %6.1f | %4.0f", F, total / 20, samples / total * 1e-3);
    }
    printf("\n");
  }
  return 0;
}
