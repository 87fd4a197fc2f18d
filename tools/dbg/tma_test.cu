#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cstdlib>
struct alignas(64) Maps { unsigned char m[16][128]; uint32_t valid; };
constexpr int kPP = 144, kPR = 38;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ Maps maps, const void* gdesc, int boxw, int lvl, int gx0, int gy0, int b, uint8_t* out) {
  __shared__ __align__(128) uint8_t s_px[kPR][kPP];
  __shared__ __align__(8) uint64_t s_bar;
  const int tid = threadIdx.x;
  if (tid == 0) {
    const uint32_t bar = smem_u32(&s_bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (boxw & 1) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (boxw & 2) __syncthreads();
  boxw &= ~3;
  if (tid == 0) {
    const uint32_t bar = smem_u32(&s_bar);
    if (gx0 < 0) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); } else {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kPR * boxw) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(&s_px[0][0])), "l"(gdesc ? gdesc : reinterpret_cast<const void*>(&maps.m[lvl][0])), "r"(gx0), "r"(gy0), "r"(b),
        "r"(bar)
        : "memory");
  } }
  __syncthreads();
  {
    const uint32_t bar = smem_u32(&s_bar);
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}"
          : "=r"(done) : "r"(bar) : "memory");
    }
  }
  for (int i = tid; i < kPR * kPP; i += blockDim.x) out[i] = (&s_px[0][0])[i];
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 0;
  const int boxw = variant == 1 ? 128 : kPP;
  const int w = 640, h = 480, pitch = 640, frames = 4;
  std::vector<uint8_t> img((size_t)pitch * h * frames);
  for (size_t i = 0; i < img.size(); ++i) img[i] = (uint8_t)(i * 7 + (i >> 9));
  uint8_t *d, *o;
  cudaMalloc(&d, img.size()); cudaMemcpy(d, img.data(), img.size(), cudaMemcpyHostToDevice);
  cudaMalloc(&o, kPR * kPP);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  printf("entry %d %d %p\n", (int)e, (int)q, fn);
  Maps maps{}; 
  const cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)frames};
  const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)pitch * h};
  const cuuint32_t box[3] = {(cuuint32_t)boxw, kPR, 1u};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  for (int lvl = 0; lvl < 2; ++lvl) {
  CUresult r = ((EncodeTiledFn)fn)((CUtensorMap*)&maps.m[lvl][0], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, strides, box, estr,
      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode %d\n", (int)r);
  }
  for (int t = 0; t < 3; ++t) {
    int lvl = t & 1, gx0 = variant == 3 ? -1 : t == 2 ? 608 : 8, gy0 = t == 2 ? 460 : 15, b = t;
    void* gd = nullptr; if (variant == 2) { cudaMalloc(&gd, 128); cudaMemcpy(gd, &maps.m[lvl][0], 128, cudaMemcpyHostToDevice); }
    k<<<1, 256>>>(maps, gd, boxw | (variant == 4 ? 1 : variant == 5 ? 2 : 0), lvl, gx0, gy0, b, o);
    e = cudaDeviceSynchronize();
    printf("kernel %d: %s\n", t, cudaGetErrorString(e));
    if (e) return 1;
    std::vector<uint8_t> got(kPR * kPP);
    cudaMemcpy(got.data(), o, got.size(), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r = 0; r < kPR; ++r) for (int c = 0; c < kPP; ++c) {
      if (c >= boxw) continue; int x = gx0 + c, y = gy0 + r;
      uint8_t exp = (x < w && y < h) ? img[((size_t)b * h + y) * pitch + x] : 0;
      bad += got[r * boxw + c] != exp;
    }
    printf("  mismatches %d\n", bad);
  }
  return 0;
}
