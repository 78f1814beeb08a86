// Pipe-throughput microbenchmark for sm_100a (B200): scalar vs packed-f32x2 FP32 ops, shuffles,
// shared-memory loads, MUFU.  Used once to pick the arithmetic form of the DE/PSD kernel
// (see DESIGN.md "Pipe measurements").  Not part of the product library.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define DEV __device__ __forceinline__
DEV u64 fma2(u64 a, u64 b, u64 c){ u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
DEV u64 add2(u64 a, u64 b){ u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
DEV u64 mul2(u64 a, u64 b){ u64 d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
DEV u64 swp(u64 a){ u64 d; asm volatile("{.reg .b32 lo,hi; mov.b64 {lo,hi}, %1; mov.b64 %0, {hi,lo};}" : "=l"(d) : "l"(a)); return d; }
DEV u64 pk(float x, float y){ u64 d; asm volatile("mov.b64 %0, {%1,%2};" : "=l"(d) : "f"(x), "f"(y)); return d; }
DEV float ffma(float a, float b, float c){ float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
DEV float fadd(float a, float b){ float d; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
DEV float fmul(float a, float b){ float d; asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }

constexpr int NCH = 8;      // independent chains per thread
constexpr int UNR = 8;      // unroll
enum Op { FFMA_RRR, FFMA_IMM, FADD_RR, FMUL_RR, FADD_IMM, FFMA2_RRR, FFMA2_IMM, FADD2_RR, FMUL2_IMM, FFMA2_SWAP, FFMA2_BCAST,
          MIX_FFMA_FADD, MIX_FFMA2_FADD2, MIX_FFMA2_IADD, MIX_FFMA2_FFMA, MIX_FFMA2_LDS, SHFL, LDS32, LDS64, LDS128, MUFU_LG2, MIX_FFMA_IADD, NOPS };
const char* names[] = {"ffma rrr","ffma imm","fadd rr","fmul rr","fadd imm","ffma2 rrr","ffma2 imm-bcast","fadd2 rr","fmul2 imm","ffma2 swap(LO_HI)","ffma2 scalar-bcast",
          "mix ffma+fadd (1:1)","mix ffma2+fadd2 (1:1)","mix ffma2+iadd (1:1)","mix ffma2+ffma (1:1)","mix ffma2+lds64 (4:1)","shfl.bfly","lds.32","lds.64","lds.128","mufu.lg2","mix ffma+iadd (1:1)"};

template<int OP>
__global__ void __launch_bounds__(256) bench(float* out, const float* in, int iters, long long* cyc)
{
  extern __shared__ __align__(16) float sm[];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = in[i & 255];
  __syncthreads();
  float f[NCH]; u64 p[NCH]; int n[NCH];
  float b = in[threadIdx.x & 63], c = in[64 + (threadIdx.x & 63)];
  u64 pb = pk(b, c), pc = pk(c, b);
#pragma unroll
  for (int j = 0; j < NCH; j++) { f[j] = in[j] + threadIdx.x; p[j] = pk(f[j], b + j); n[j] = threadIdx.x + j; }
  int sidx = (threadIdx.x * 4) & 4095;
  long long t0 = clock64(); unsigned long long g0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < UNR; u++) {
#pragma unroll
      for (int j = 0; j < NCH; j++) {
        if (OP == FFMA_RRR) f[j] = ffma(f[j], b, c);
        if (OP == FFMA_IMM) f[j] = ffma(f[j], 1.0001f, c);
        if (OP == FADD_RR)  f[j] = fadd(f[j], b);
        if (OP == FADD_IMM) f[j] = fadd(f[j], 1.5f);
        if (OP == FMUL_RR)  f[j] = fmul(f[j], b);
        if (OP == FFMA2_RRR) p[j] = fma2(p[j], pb, pc);
        if (OP == FFMA2_IMM) p[j] = fma2(p[j], pk(1.0001f, 1.0001f), pc);
        if (OP == FADD2_RR) p[j] = add2(p[j], pb);
        if (OP == FMUL2_IMM) p[j] = mul2(p[j], pk(1.0001f, 1.0001f));
        if (OP == FFMA2_SWAP) p[j] = fma2(swp(p[j]), pb, pc);
        if (OP == FFMA2_BCAST) p[j] = fma2(pk(b, b), p[j], pc);
        if (OP == MIX_FFMA_FADD) { if (j & 1) f[j] = ffma(f[j], b, c); else f[j] = fadd(f[j], b); }
        if (OP == MIX_FFMA2_FADD2) { if (j & 1) p[j] = fma2(p[j], pb, pc); else p[j] = add2(p[j], pb); }
        if (OP == MIX_FFMA2_IADD) { if (j & 1) p[j] = fma2(p[j], pb, pc); else n[j] = (n[j] + it) ^ u; }
        if (OP == MIX_FFMA_IADD) { if (j & 1) f[j] = ffma(f[j], b, c); else n[j] = (n[j] + it) ^ u; }
        if (OP == MIX_FFMA2_FFMA) { if (j & 1) p[j] = fma2(p[j], pb, pc); else f[j] = ffma(f[j], b, c); }
        if (OP == MIX_FFMA2_LDS) { if (j < 4) p[j] = fma2(p[j], pb, pc);
              else if (j == 4) { float2 v = *reinterpret_cast<float2*>(&sm[(sidx + 8 * u + 32 * it) & 4095]); f[0] += v.x; f[1] += v.y; } }
        if (OP == SHFL) f[j] = __shfl_xor_sync(0xffffffffu, f[j], 1 + (u & 15));
        if (OP == LDS32) { f[j] += sm[(sidx + 32 * j + u + 4 * it) & 4095]; }
        if (OP == LDS64) { float2 v = *reinterpret_cast<float2*>(&sm[(threadIdx.x * 2 + 64 * j + 512 * u + 2 * it) & 4095]); f[j] += v.x + v.y; }
        if (OP == LDS128) { float4 v = *reinterpret_cast<float4*>(&sm[(threadIdx.x * 4 + 128 * j + 1024 * u + 4 * it) & 4095]); f[j] += v.x + v.w; }
        if (OP == MUFU_LG2) f[j] = __log2f(f[j]);
      }
    }
  }
  long long t1 = clock64(); unsigned long long g1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
  float acc = 0; u64 pacc = 0; int nacc = 0;
#pragma unroll
  for (int j = 0; j < NCH; j++) { acc += f[j]; pacc += p[j]; nacc += n[j]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + (float)pacc + nacc;
  if (threadIdx.x == 0) { cyc[2 * blockIdx.x] = t1 - t0; cyc[2 * blockIdx.x + 1] = (long long)(g1 - g0); }
}

template<int OP> void run(int bps, float* out, float* in, long long* cyc, long long* hcyc)
{
  int nsm = 148, blocks = nsm * bps;
  int iters = (OP == MUFU_LG2 || OP == SHFL) ? 4000 : 16000;
  size_t smem = (size_t)(220 * 1024 / bps) & ~(size_t)1023;     // forces exactly `bps` co-resident blocks per SM
  cudaFuncSetAttribute(bench<OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  bench<OP><<<blocks, 256, smem>>>(out, in, 2000, cyc);   // warm (also ramps clocks)
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  bench<OP><<<blocks, 256, smem>>>(out, in, iters, cyc);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaMemcpy(hcyc, cyc, blocks * 2 * sizeof(long long), cudaMemcpyDeviceToHost);
  double fsum = 0; long long mx = 0; for (int i = 0; i < blocks; i++) { fsum += (double)hcyc[2 * i] / (double)hcyc[2 * i + 1]; if (hcyc[2*i] > mx) mx = hcyc[2*i]; }
  double ghz = fsum / blocks;                                   // SM clock in GHz (cycles per ns)
  double slots = (OP == MIX_FFMA2_LDS) ? 5.0 : (double)NCH;
  double total_warp_instr = (double)iters * UNR * slots * 8.0 * blocks;   // 8 warps per block
  double per_clk_sm = total_warp_instr / nsm / (ms * 1e6 * ghz);
  double per_clk_sm_cyc = total_warp_instr / nsm / (double)mx;
  printf("%-26s warps/SM=%2d  %.3f warp-instr/clk/SM (by max-cyc %.3f)  ms=%.3f  sm_clk=%.0f MHz\n",
         names[OP], bps * 8, per_clk_sm, per_clk_sm_cyc, ms, ghz * 1e3);
  cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) printf("CUDA error %s\n", cudaGetErrorString(e));
}

int main()
{
  float *out, *in; long long* cyc; cudaMalloc(&out, 148 * 8 * 256 * 4); cudaMalloc(&in, 4096 * 4); cudaMalloc(&cyc, 148 * 8 * 16);
  float h[4096]; for (int i = 0; i < 4096; i++) h[i] = 1.0f + 1e-3f * (i % 97);
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  long long* hcyc = (long long*)malloc(148 * 8 * 16);
  for (int bps : {1, 2, 4}) {
    run<FFMA_RRR>(bps, out, in, cyc, hcyc); run<FFMA_IMM>(bps, out, in, cyc, hcyc); run<FADD_RR>(bps, out, in, cyc, hcyc);
    run<FADD_IMM>(bps, out, in, cyc, hcyc); run<FMUL_RR>(bps, out, in, cyc, hcyc);
    run<FFMA2_RRR>(bps, out, in, cyc, hcyc); run<FFMA2_IMM>(bps, out, in, cyc, hcyc); run<FADD2_RR>(bps, out, in, cyc, hcyc);
    run<FMUL2_IMM>(bps, out, in, cyc, hcyc); run<FFMA2_SWAP>(bps, out, in, cyc, hcyc); run<FFMA2_BCAST>(bps, out, in, cyc, hcyc);
    run<MIX_FFMA_FADD>(bps, out, in, cyc, hcyc); run<MIX_FFMA2_FADD2>(bps, out, in, cyc, hcyc); run<MIX_FFMA2_IADD>(bps, out, in, cyc, hcyc);
    run<MIX_FFMA_IADD>(bps, out, in, cyc, hcyc);
    run<MIX_FFMA2_FFMA>(bps, out, in, cyc, hcyc); run<MIX_FFMA2_LDS>(bps, out, in, cyc, hcyc);
    run<SHFL>(bps, out, in, cyc, hcyc); run<LDS32>(bps, out, in, cyc, hcyc); run<LDS64>(bps, out, in, cyc, hcyc); run<LDS128>(bps, out, in, cyc, hcyc);
    run<MUFU_LG2>(bps, out, in, cyc, hcyc);
  }
  return 0;
}
