// Upper bound for the 500 ms kernels: the arithmetic core alone (bandpower.cuh), 16 warps per SM at 128 registers, windows
// read from shared memory that is filled once -- no TMA, no staging, no tile duties.
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -o fft_core fft_core.cu && ./fft_core
#include <cstdio>
#include <cuda_runtime.h>
#include "../../eeg2video_b200/csrc/bandpower.cuh"

using namespace eegfe;

template <int VEC>
__global__ void __launch_bounds__(512, 1) core_kernel(float* out, int reps)
{
  extern __shared__ __align__(128) float smem[];
  for (int i = threadIdx.x; i < 16 * 404 * 4; i += blockDim.x) smem[i] = 30.0f * __sinf(0.37f * i) + 5.0f;
  __syncthreads();
  const int unit = threadIdx.x % 112;
  const int row = unit / 7, w = unit % 7;
  const float* win = smem + (threadIdx.x / 128) * 16 * 404 + row * 404 + (VEC == 4 ? (w & ~1) * 50 : w * 50);
  float acc = 0.f;
  for (int r = 0; r < reps; ++r) {
    float e[5];
    window_band_energy<4, kHannHalfSec, VEC>(win + (r & 1) * 4, e);
    acc += e[0] + e[1] + e[2] + e[3] + e[4];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int VEC>
static void run(const char* name)
{
  float* out;
  cudaMalloc(&out, 148 * 512 * sizeof(float));
  const int smem = 16 * 404 * 4 * 4 + 64;
  cudaFuncSetAttribute(core_kernel<VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int reps = 200;
  core_kernel<VEC><<<148, 512, smem>>>(out, 10);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  cudaEventRecord(a);
  core_kernel<VEC><<<148, 512, smem>>>(out, reps);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  const double windows = 148.0 * 512 * reps;
  printf("%s: %.3f ms, %.2f G windows/s (%s)\n", name, ms, windows / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

int main()
{
  run<2>("arithmetic core only, LDS.64 ");
  run<4>("arithmetic core only, LDS.128");
  return 0;
}
