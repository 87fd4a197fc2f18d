#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap map, const uint8_t* src, int mode, uint8_t* out) {
  __shared__ __align__(128) uint8_t s_px[64 * 128];
  __shared__ __align__(8) uint64_t s_bar;
  const int tid = threadIdx.x;
  const uint32_t bar = smem_u32(&s_bar);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(mode == 5 ? 144 * 38 : 64 * 128) : "memory");
    if (mode == 0) {
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_u32(s_px)), "l"(src), "r"(64 * 128), "r"(bar) : "memory");
    } else if (mode == 3) {
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                   ::"r"(smem_u32(s_px)), "l"(&map), "r"(0), "r"(0), "r"(1), "r"(bar) : "memory");
    } else if (mode == 7) {
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                   ::"r"(smem_u32(s_px)), "l"(&map), "r"(8), "r"(15), "r"(1), "r"(bar) : "memory");
    } else if (mode == 8) {
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                   ::"r"(smem_u32(s_px)), "l"(&map), "r"(16), "r"(15), "r"(1), "r"(bar) : "memory");
    } else if (mode == 9) {
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                   ::"r"(smem_u32(s_px)), "l"((const void*)(src + 65536 + (out ? 0 : 128))), "r"(0), "r"(0), "r"(1), "r"(bar) : "memory");
    } else if (mode == 6) {
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                   ::"r"(smem_u32(s_px)), "l"(&map), "r"(0), "r"(0), "r"(bar) : "memory");
    } else {
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                   ::"r"(smem_u32(s_px)), "l"(&map), "r"(0), "r"(0), "r"(bar) : "memory");
    }
  }
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(bar) : "memory");
  }
  for (int i = tid; i < 64 * 128; i += blockDim.x) out[i] = s_px[i];
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 0;
  const int w = 1024, h = 256;
  std::vector<uint8_t> img((size_t)w * h);
  for (size_t i = 0; i < img.size(); ++i) img[i] = (uint8_t)(i * 7 + (i >> 9));
  uint8_t *d, *o;
  cudaMalloc(&d, img.size()); cudaMemcpy(d, img.data(), img.size(), cudaMemcpyHostToDevice);
  cudaMalloc(&o, 64 * 128);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  alignas(64) CUtensorMap map;
  const cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)(mode == 3 || mode >= 7 ? h / 4 : h), 4};
  const cuuint64_t strides[2] = {(cuuint64_t)w, (cuuint64_t)w * (h / 4)};
  const cuuint32_t box[3] = {mode == 5 ? 144u : 128u, mode == 5 ? 38u : 64u, 1};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  CUresult r = ((EncodeTiledFn)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, mode == 3 || mode >= 7 ? 3 : 2, d, dims, strides, box, estr,
      CU_TENSOR_MAP_INTERLEAVE_NONE, mode == 2 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, mode == 4 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode %d\n", (int)r);
  if (mode == 9) cudaMemcpy(d + 65536, &map, 128, cudaMemcpyHostToDevice);
  k<<<1, 256>>>(map, d, mode, o);
  cudaError_t e = cudaDeviceSynchronize();
  printf("mode %d kernel: %s\n", mode, cudaGetErrorString(e));
  if (e) return 1;
  std::vector<uint8_t> got(64 * 128);
  cudaMemcpy(got.data(), o, got.size(), cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int r2 = 0; r2 < 64; ++r2) for (int c = 0; c < 128; ++c) {
    uint8_t exp = mode == 0 ? img[r2 * 128 + c] : img[(size_t)r2 * w + c];
    bad += got[r2 * 128 + c] != exp;
  }
  printf("  mismatches %d\n", bad);
  return 0;
}
