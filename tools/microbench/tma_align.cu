// Does a tiled TMA tensor copy accept an inner-dimension start coordinate that is not 16-byte aligned?
// (float32 tensor (T, rows); box (100, 16); x = 0, 4, 2, 50, 1.)  One CTA, one copy per launch, result compared.
//   nvcc -std=c++17 -O2 -gencode arch=compute_100a,code=sm_100a -o tma_align tma_align.cu && ./tma_align
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

__global__ void k(const __grid_constant__ CUtensorMap map, int x, int y, float* out)
{
  extern __shared__ __align__(128) float buf[];
  __shared__ uint64_t bar;
  const uint32_t bar_a = static_cast<uint32_t>(__cvta_generic_to_shared(&bar));
  const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(buf));
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(100 * 16 * 4) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(&map)), "r"(x), "r"(y), "r"(bar_a)
                 : "memory");
  }
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done)
                 : "r"(bar_a)
                 : "memory");
  for (int i = threadIdx.x; i < 1600; i += blockDim.x) out[i] = buf[i];
}

int main()
{
  const int T = 4000, rows = 40;
  std::vector<float> h(T * rows);
  for (int i = 0; i < T * rows; ++i) h[i] = static_cast<float>(i);
  float *d, *out;
  cudaMalloc(&d, h.size() * 4);
  cudaMalloc(&out, 1600 * 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  auto encode = reinterpret_cast<CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                              const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                              CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill)>(p);
  CUtensorMap map;
  cuuint64_t gdim[2] = {T, rows}, gstride[1] = {T * 4};
  cuuint32_t box[2] = {100, 16}, es[2] = {1, 1};
  CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, gdim, gstride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode: %d\n", (int)r);
  const int xs[] = {0, 4, 2, 50, 1, 3950};
  for (int x : xs) {
    k<<<1, 128, 1600 * 4>>>(map, x, 3, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("x=%d: %s\n", x, cudaGetErrorString(e));
      return 0;
    }
    std::vector<float> o(1600);
    cudaMemcpy(o.data(), out, 6400, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int rr = 0; rr < 16; ++rr)
      for (int i = 0; i < 100; ++i) {
        const float want = (x + i < T) ? h[(3 + rr) * T + x + i] : 0.f;
        if (o[rr * 100 + i] != want) ++bad;
      }
    printf("x=%d: ok, mismatches=%d\n", x, bad);
  }
  return 0;
}
