#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#define TRY(x) ({ CUresult e_ = (x); const char *s_ = nullptr; cuGetErrorString(e_, &s_); printf("  %-80.80s -> %d %s\n", #x, (int)e_, e_ ? (s_ ? s_ : "?") : "ok"); e_; })
int main() {
  setvbuf(stdout, nullptr, _IONBF, 0);
  cudaSetDevice(0); cudaFree(0); cuInit(0);
  CUdevice dev; cuDeviceGet(&dev, 0);
  size_t fr, tot; cuMemGetInfo(&fr, &tot); printf("free %zu MiB of %zu MiB\n", fr >> 20, tot >> 20);
  CUmemAllocationProp prop; memset(&prop, 0, sizeof prop);
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED; prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE; prop.location.id = 0;
  prop.allocFlags.usage = CU_MEM_CREATE_USAGE_TILE_POOL;
  for (size_t mb : {96, 128, 192, 256, 320, 384, 448, 512, 1024}) {
    CUmemGenericAllocationHandle h = 0; CUresult r = cuMemCreate(&h, mb << 20, &prop, 0);
    printf("tile pool %4zu MiB -> %d\n", mb, (int)r);
    if (r == 0) cuMemRelease(h);
  }
  // several pools at once?
  { CUmemGenericAllocationHandle h[16]; int n = 0; for (; n < 16; ++n) if (cuMemCreate(&h[n], 256u << 20, &prop, 0)) break; printf("simultaneous 256 MiB pools: %d\n", n); for (int i = 0; i < n; ++i) cuMemRelease(h[i]); }
  // can a tile pool be mapped linearly?
  { CUmemGenericAllocationHandle h = 0; if (TRY(cuMemCreate(&h, 64u << 20, &prop, 0)) == 0) { CUdeviceptr va = 0; TRY(cuMemAddressReserve(&va, 64u << 20, 0, 0, 0)); TRY(cuMemMap(va, 64u << 20, 0, h, 0));
      CUmemAccessDesc acc; memset(&acc, 0, sizeof acc); acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE; acc.location.id = 0; acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE; TRY(cuMemSetAccess(va, 64u << 20, &acc, 1)); TRY(cuMemsetD8(va, 1, 64u << 20)); TRY(cuCtxSynchronize()); } }
  struct { const char *name; unsigned flags; int depth; int ch; CUarray_format f; } v[] = {
    {"sparse layered RG16", CUDA_ARRAY3D_LAYERED | CUDA_ARRAY3D_SPARSE, 512, 2, CU_AD_FORMAT_UNSIGNED_INT16},
    {"sparse layered R32F", CUDA_ARRAY3D_LAYERED | CUDA_ARRAY3D_SPARSE, 512, 1, CU_AD_FORMAT_FLOAT},
    {"sparse layered RG16 16 layers", CUDA_ARRAY3D_LAYERED | CUDA_ARRAY3D_SPARSE, 16, 2, CU_AD_FORMAT_UNSIGNED_INT16},
    {"sparse 3D RG16", CUDA_ARRAY3D_SPARSE, 512, 2, CU_AD_FORMAT_UNSIGNED_INT16},
    {"sparse 3D R16", CUDA_ARRAY3D_SPARSE, 512, 1, CU_AD_FORMAT_UNSIGNED_INT16},
    {"sparse 2D RG16", CUDA_ARRAY3D_SPARSE, 0, 2, CU_AD_FORMAT_UNSIGNED_INT16},
    {"deferred 3D R16", CUDA_ARRAY3D_DEFERRED_MAPPING, 512, 1, CU_AD_FORMAT_UNSIGNED_INT16},
    {"deferred layered RG16 64 layers", CUDA_ARRAY3D_LAYERED | CUDA_ARRAY3D_DEFERRED_MAPPING, 64, 2, CU_AD_FORMAT_UNSIGNED_INT16},
  };
  for (auto &c : v) {
    CUDA_ARRAY3D_DESCRIPTOR ad; memset(&ad, 0, sizeof ad);
    ad.Width = 512; ad.Height = 512; ad.Depth = c.depth; ad.Format = c.f; ad.NumChannels = c.ch; ad.Flags = c.flags;
    CUarray a = nullptr; CUresult r = cuArray3DCreate(&a, &ad);
    printf("%s: create -> %d\n", c.name, (int)r);
    if (r) continue;
    if (c.flags & CUDA_ARRAY3D_SPARSE) { CUDA_ARRAY_SPARSE_PROPERTIES sp; memset(&sp, 0, sizeof sp); r = cuArrayGetSparseProperties(&sp, a); printf("   sparse props -> %d tile %u x %u x %u miptail %llu\n", (int)r, sp.tileExtent.width, sp.tileExtent.height, sp.tileExtent.depth, sp.miptailSize); }
    else { CUDA_ARRAY_MEMORY_REQUIREMENTS rq; memset(&rq, 0, sizeof rq); r = cuArrayGetMemoryRequirements(&rq, a, dev); printf("   mem req -> %d size %zu align %zu\n", (int)r, rq.size, rq.alignment); }
    cuArrayDestroy(a);
  }
  // mipmapped sparse
  { CUDA_ARRAY3D_DESCRIPTOR ad; memset(&ad, 0, sizeof ad); ad.Width = 512; ad.Height = 512; ad.Depth = 512; ad.Format = CU_AD_FORMAT_UNSIGNED_INT16; ad.NumChannels = 2; ad.Flags = CUDA_ARRAY3D_LAYERED | CUDA_ARRAY3D_SPARSE;
    CUmipmappedArray m = nullptr; TRY(cuMipmappedArrayCreate(&m, &ad, 1)); }
  return 0;
}
