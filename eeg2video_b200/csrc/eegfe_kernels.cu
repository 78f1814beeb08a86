// libeegfe.so -- fused segmentation + Hann + 200-point FFT + five-band DE/PSD for B200 (sm_100a).
//
// One persistent kernel per analysis mode.  A CTA owns a tile of R rows (one row = one channel of one 2 s clip,
// or one pre-cut window); warp 0 pulls the rows of the NEXT tile from HBM into shared memory with one 1-D TMA
// bulk copy per row (cp.async.bulk ... mbarrier::complete_tx, SASS UBLKCP) while all warps work on the current
// tile.  Rows are addressed by index arithmetic on the raw recording (clip (c, r) of a block starts at sample
// c*2600 + 600 + r*400; window w at +50 w), so neither the clip tensor nor the sliding-window tensor of the
// reference is ever materialised.  Each thread then owns one channel-window: it reads its samples from shared
// memory (LDS.64), runs the register-resident prime-factor FFT of bandpower.cuh with packed f32x2 arithmetic and
// writes 5 PSD + 5 DE values.  No tensor cores (FFT + reduction, no dense contraction), no inter-CTA traffic.
//
// C ABI: include/eegfe.h.  Reference semantics: see the citations in that header and in bandpower.cuh.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/eegfe.h"
#include "bandpower.cuh"

namespace eegfe {

// ---------------------------------------------------------------------------------------------------------------
// job description shared by all feature kernels
// ---------------------------------------------------------------------------------------------------------------
struct Job {
  const float* in;
  float* de;
  float* psd;
  int* status;
  long long total_rows;   // n_units * n_ch
  long long base;         // element offset of unit u: base + (u / d1) * s0 + ((u % d1) / d2) * s1 + (u % d2) * s2
  long long s0, s1, s2;
  long long ch_stride;    // elements between channel rows of one unit
  int d1, d2;
  int n_ch;
  int aligned16;          // every row start is 16-byte aligned -> TMA bulk copies; else cooperative loads
};

// per-mode compile-time geometry
template <int LOAD_, int NWIN_, int HOP_, int NI_, int HANN_, int ROWS_, int NBUF_, int CTAS_>
struct Cfg {
  static constexpr int kLoad = LOAD_;        // samples fetched per row
  static constexpr int kWindows = NWIN_;     // analysis windows per row
  static constexpr int kHop = HOP_;          // samples between window starts
  static constexpr int kNi = NI_;            // live inputs per radix-8 group (4: 100-sample window, 8: 200)
  static constexpr int kHann = HANN_;
  static constexpr int kRows = ROWS_;        // rows per tile
  static constexpr int kBufs = NBUF_;        // shared-memory stages
  static constexpr int kCtasPerSm = CTAS_;
  static constexpr int kUnits = ROWS_ * NWIN_;                 // channel-windows per tile
  static constexpr int kThreads = (kUnits + 31) / 32 * 32;
  static constexpr int kRowBytes = LOAD_ * 4;
  static constexpr int kSmemBytes = NBUF_ * ROWS_ * LOAD_ * 4;
};
//                         LOAD NWIN HOP NI  HANN          ROWS NBUF CTAS
using CfgSliding500 = Cfg<400, 7, 50, 4, kHannHalfSec, 32, 2, 2>;    // 224 threads, 100 KB smem
using CfgOneSec     = Cfg<400, 2, 200, 8, kHannOneSec, 32, 2, 2>;    //  64 threads
using CfgTwoSec     = Cfg<200, 1, 0, 8, kHannTwoSec, 64, 2, 2>;      //  64 threads (only samples 0..199 are read)
using CfgWin100     = Cfg<100, 1, 0, 4, kHannHalfSec, 128, 2, 2>;    // 128 threads, pre-cut 500 ms windows
using CfgWin200     = Cfg<200, 1, 0, 8, kHannOneSec, 64, 2, 2>;      //  64 threads, pre-cut 1 s windows

// ---------------------------------------------------------------------------------------------------------------
// mbarrier / TMA bulk-copy helpers (PTX)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ long long row_offset(const Job& job, long long g)
{
  const long long u = g / job.n_ch;
  const int ch = static_cast<int>(g - u * job.n_ch);
  const long long q = u / job.d1;
  const int rem = static_cast<int>(u - q * job.d1);
  return job.base + q * job.s0 + (rem / job.d2) * job.s1 + (rem % job.d2) * job.s2 + ch * job.ch_stride;
}

// E_b -> psd_b = E_b / count_b (DE_PSD.py:66), de_b = log2(100 psd_b) (:68); counts 4, 5, 7, 18, 69
__device__ __forceinline__ void store_features(const Job& job, long long g, int w, int n_windows, const float (&e)[5])
{
  const long long u = g / job.n_ch;
  const int ch = static_cast<int>(g - u * job.n_ch);
  const long long o = ((u * n_windows + w) * job.n_ch + ch) * 5;
  const float cnt[5] = {4.0f, 5.0f, 7.0f, 18.0f, 69.0f};
  bool zero = false;
#pragma unroll
  for (int b = 0; b < 5; ++b) {
    const float p = __fdiv_rn(e[b], cnt[b]);
    zero |= (p == 0.0f);
    job.psd[o + b] = p;
    job.de[o + b] = log2f(__fmul_rn(100.0f, p));
  }
  if (zero && job.status != nullptr) atomicOr(job.status, EEGFE_STATUS_ZERO_POWER);
}

// ---------------------------------------------------------------------------------------------------------------
// the fused kernel
// ---------------------------------------------------------------------------------------------------------------
template <class C>
__global__ void __launch_bounds__(C::kThreads, C::kCtasPerSm) de_psd_kernel(const Job job)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* const bufs = reinterpret_cast<float*>(smem_raw);
  __shared__ uint64_t full_bar[C::kBufs];

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const bool loader_warp = (tid < 32);
  const long long n_tiles = (job.total_rows + C::kRows - 1) / C::kRows;
  const int row_in_tile = tid / C::kWindows;
  const int w = tid - row_in_tile * C::kWindows;
  const bool has_unit = tid < C::kUnits;

  if (job.aligned16) {
    if (tid == 0) {
#pragma unroll
      for (int b = 0; b < C::kBufs; ++b) mbar_init(&full_bar[b], 1);
      mbar_fence_init();
    }
    __syncthreads();

    auto issue = [&](long long tile, int buf) {          // warp 0: one bulk copy per row of the tile
      const long long row0 = tile * C::kRows;
      const long long left = job.total_rows - row0;
      const int nrows = left < C::kRows ? static_cast<int>(left) : C::kRows;
      if (lane == 0) mbar_arrive_expect_tx(&full_bar[buf], static_cast<uint32_t>(nrows) * C::kRowBytes);
      __syncwarp();
      for (int r = lane; r < nrows; r += 32)
        bulk_copy_g2s(bufs + (buf * C::kRows + r) * C::kLoad, job.in + row_offset(job, row0 + r), C::kRowBytes,
                      &full_bar[buf]);
    };

    long long tile = blockIdx.x;
    if (loader_warp && tile < n_tiles) issue(tile, 0);
    for (int it = 0; tile < n_tiles; tile += gridDim.x, ++it) {
      const int buf = it % C::kBufs;
      const long long next = tile + gridDim.x;
      if (C::kBufs > 1 && loader_warp && next < n_tiles) issue(next, (it + 1) % C::kBufs);
      mbar_wait(&full_bar[buf], (it / C::kBufs) & 1);
      const long long g = tile * C::kRows + row_in_tile;
      if (has_unit && g < job.total_rows) {
        float e[5];
        window_band_energy<C::kNi, C::kHann>(bufs + (buf * C::kRows + row_in_tile) * C::kLoad + w * C::kHop, e);
        store_features(job, g, w, C::kWindows, e);
      }
      __syncthreads();                                   // everyone is done reading `buf`
      if (C::kBufs == 1 && loader_warp && next < n_tiles) issue(next, 0);
    }
  } else {
    // rows not 16-byte aligned (odd block lengths / strides): cooperative 4-byte loads, single stage
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const long long row0 = tile * C::kRows;
      const long long left = job.total_rows - row0;
      const int nrows = left < C::kRows ? static_cast<int>(left) : C::kRows;
      for (int r = tid / 32; r < nrows; r += C::kThreads / 32) {
        const float* src = job.in + row_offset(job, row0 + r);
        for (int i = lane; i < C::kLoad; i += 32) bufs[r * C::kLoad + i] = __ldg(src + i);
      }
      __syncthreads();
      const long long g = row0 + row_in_tile;
      if (has_unit && g < job.total_rows) {
        float e[5];
        window_band_energy<C::kNi, C::kHann>(bufs + row_in_tile * C::kLoad + w * C::kHop, e);
        store_features(job, g, w, C::kWindows, e);
      }
      __syncthreads();
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// byte-exact gathers for the materialising entry points (segment_all_files, seg_sliding_window + save)
// ---------------------------------------------------------------------------------------------------------------
// One "row" = one channel of one clip: `row_elems` contiguous elements copied from src to dst.
// VEC = bytes moved per thread per step (16 when everything is 16-byte aligned, else the element size).
template <int VEC>
__global__ void __launch_bounds__(256) gather_rows_kernel(const unsigned char* __restrict__ src, unsigned char* __restrict__ dst,
                                                           long long n_rows, int row_bytes, int esize, Job geom)
{
  typedef typename std::conditional<VEC == 16, uint4,
          typename std::conditional<VEC == 8, uint2,
          typename std::conditional<VEC == 4, uint32_t, uint16_t>::type>::type>::type vec_t;
  const int per_row = row_bytes / VEC;
  const long long total = n_rows * per_row;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = i / per_row;
    const int v = static_cast<int>(i - row * per_row);
    const unsigned char* s = src + row_offset(geom, row) * esize;
    reinterpret_cast<vec_t*>(dst + row * row_bytes)[v] = reinterpret_cast<const vec_t*>(s)[v];
  }
}

// clips [n_clips][n_ch][400] -> windows [n_clips][7][n_ch][100]; window w = samples [50 w, 50 w + 100)
template <int VEC>
__global__ void __launch_bounds__(256) sliding_windows_kernel(const unsigned char* __restrict__ clips,
                                                               unsigned char* __restrict__ out, long long n_clips,
                                                               int n_ch, int esize)
{
  typedef typename std::conditional<VEC == 8, uint2,
          typename std::conditional<VEC == 4, uint32_t, uint16_t>::type>::type vec_t;
  const int per_win = 100 * esize / VEC;
  const long long total = n_clips * 7 * n_ch * per_win;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % per_win);
    long long t = i / per_win;
    const int ch = static_cast<int>(t % n_ch);
    t /= n_ch;
    const int w = static_cast<int>(t % 7);
    const long long clip = t / 7;
    const unsigned char* s = clips + ((clip * n_ch + ch) * 400 + 50 * w) * esize;
    reinterpret_cast<vec_t*>(out)[i] = reinterpret_cast<const vec_t*>(s)[v];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
static long long g_launches = 0;

static int sm_count()
{
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

template <class C>
static int launch(const Job& job, cudaStream_t stream)
{
  if (job.total_rows == 0) return 0;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(de_psd_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
    if (e != cudaSuccess) return static_cast<int>(e);
    configured = true;
  }
  const long long n_tiles = (job.total_rows + C::kRows - 1) / C::kRows;
  long long grid = static_cast<long long>(sm_count()) * C::kCtasPerSm;
  if (grid > n_tiles) grid = n_tiles;
  de_psd_kernel<C><<<static_cast<unsigned>(grid), C::kThreads, C::kSmemBytes, stream>>>(job);
  ++g_launches;
  return static_cast<int>(cudaGetLastError());
}

static bool is_aligned16(const void* p, std::initializer_list<long long> elem_strides, int esize)
{
  if (reinterpret_cast<uintptr_t>(p) % 16 != 0) return false;
  for (long long s : elem_strides)
    if ((s * esize) % 16 != 0) return false;
  return true;
}

static int esize_of(int dtype)
{
  switch (dtype) {
    case EEGFE_DTYPE_F32: return 4;
    case EEGFE_DTYPE_F64: return 8;
    case EEGFE_DTYPE_F16: return 2;
    case EEGFE_DTYPE_I16: return 2;
    default: return 0;
  }
}

static Job raw_geometry(int64_t n_blocks, int n_ch, int64_t block_stride, int64_t ch_stride, int fs = 200)
{
  Job j{};
  j.total_rows = n_blocks * 200 * n_ch;      // 40 concepts x 5 repetitions per block
  j.base = 3LL * fs;                         // 3 s hint before each concept (segment_raw_signals_200Hz.py:58-62)
  j.s0 = block_stride;
  j.s1 = 13LL * fs;                          // concept stride: 3 s hint + 5 x 2 s
  j.s2 = 2LL * fs;                           // repetition stride: 2 s
  j.d1 = 200;
  j.d2 = 5;
  j.ch_stride = ch_stride;
  j.n_ch = n_ch;
  return j;
}

}  // namespace eegfe

using namespace eegfe;

extern "C" {

int eegfe_abi_version(void) { return EEGFE_ABI_VERSION; }

const char* eegfe_error_string(int code)
{
  switch (code) {
    case 0: return "success";
    case EEGFE_EINVAL: return "invalid argument (mode, shape or null pointer)";
    case EEGFE_ERANGE: return "Segment length mismatch";   // text of the reference's RuntimeError
    case EEGFE_EDTYPE: return "unsupported element type";
    default: return code > 0 ? cudaGetErrorString(static_cast<cudaError_t>(code)) : "unknown error";
  }
}

int eegfe_windows_per_clip(int mode)
{
  switch (mode) {
    case EEGFE_MODE_500MS: return 7;
    case EEGFE_MODE_1S: return 2;
    case EEGFE_MODE_2S: return 1;
    default: return EEGFE_EINVAL;
  }
}

static int dispatch_clip_mode(int mode, Job& job, cudaStream_t stream)
{
  switch (mode) {
    case EEGFE_MODE_500MS: return launch<CfgSliding500>(job, stream);
    case EEGFE_MODE_1S: return launch<CfgOneSec>(job, stream);
    case EEGFE_MODE_2S: return launch<CfgTwoSec>(job, stream);
    default: return EEGFE_EINVAL;
  }
}

int eegfe_de_psd_from_raw(const float* raw, int64_t n_blocks, int n_ch, int64_t block_len, int64_t block_stride,
                          int64_t ch_stride, int mode, float* de, float* psd, int* status, void* stream)
{
  if (n_blocks < 0 || n_ch <= 0 || eegfe_windows_per_clip(mode) < 0) return EEGFE_EINVAL;
  if (n_blocks == 0) return 0;
  if (raw == nullptr || de == nullptr || psd == nullptr) return EEGFE_EINVAL;
  if (block_len < 40 * 2600) return EEGFE_ERANGE;
  if (ch_stride < block_len || block_stride < 0) return EEGFE_EINVAL;
  Job job = raw_geometry(n_blocks, n_ch, block_stride, ch_stride);
  job.in = raw;
  job.de = de;
  job.psd = psd;
  job.status = status;
  job.aligned16 = is_aligned16(raw, {block_stride, ch_stride}, 4);
  return dispatch_clip_mode(mode, job, static_cast<cudaStream_t>(stream));
}

int eegfe_de_psd_from_clips(const float* clips, int64_t n_clips, int n_ch, int mode, float* de, float* psd,
                            int* status, void* stream)
{
  if (n_clips < 0 || n_ch <= 0 || eegfe_windows_per_clip(mode) < 0) return EEGFE_EINVAL;
  if (n_clips == 0) return 0;
  if (clips == nullptr || de == nullptr || psd == nullptr) return EEGFE_EINVAL;
  Job job{};
  job.in = clips;
  job.de = de;
  job.psd = psd;
  job.status = status;
  job.total_rows = n_clips * n_ch;
  job.s0 = static_cast<long long>(n_ch) * 400;
  job.d1 = 1;
  job.d2 = 1;
  job.ch_stride = 400;
  job.n_ch = n_ch;
  job.aligned16 = is_aligned16(clips, {}, 4);
  return dispatch_clip_mode(mode, job, static_cast<cudaStream_t>(stream));
}

int eegfe_de_psd_windows(const float* x, int64_t n_rows, int win_len, int64_t row_stride, float* de, float* psd,
                         int* status, void* stream)
{
  if (n_rows < 0 || (win_len != 100 && win_len != 200 && win_len != 400) || row_stride < win_len) return EEGFE_EINVAL;
  if (n_rows == 0) return 0;
  if (x == nullptr || de == nullptr || psd == nullptr) return EEGFE_EINVAL;
  Job job{};
  job.in = x;
  job.de = de;
  job.psd = psd;
  job.status = status;
  job.total_rows = n_rows;
  job.s0 = row_stride;
  job.d1 = 1;
  job.d2 = 1;
  job.ch_stride = 0;
  job.n_ch = 1;
  job.aligned16 = is_aligned16(x, {row_stride}, 4);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (win_len == 100) return launch<CfgWin100>(job, s);
  if (win_len == 200) return launch<CfgWin200>(job, s);
  return launch<CfgTwoSec>(job, s);          // 400: only samples 0..199 influence the result (DE_PSD.py:58)
}

int eegfe_segment_clips(const void* raw, int dtype, int64_t n_blocks, int n_ch, int64_t block_len,
                        int64_t block_stride, int64_t ch_stride, int fs, void* clips, void* stream)
{
  const int es = esize_of(dtype);
  if (es == 0) return EEGFE_EDTYPE;
  if (n_blocks < 0 || n_ch <= 0 || fs <= 0) return EEGFE_EINVAL;
  if (n_blocks == 0) return 0;
  if (raw == nullptr || clips == nullptr) return EEGFE_EINVAL;
  if (block_len < 40LL * 13 * fs) return EEGFE_ERANGE;
  if (ch_stride < block_len || block_stride < 0) return EEGFE_EINVAL;
  Job geom = raw_geometry(n_blocks, n_ch, block_stride, ch_stride, fs);
  const long long n_rows = geom.total_rows;
  const int row_bytes = 2 * fs * es;
  const bool a16 = is_aligned16(raw, {block_stride, ch_stride, 3LL * fs, 13LL * fs, 2LL * fs}, es) &&
                   reinterpret_cast<uintptr_t>(clips) % 16 == 0;
  const int threads = 256;
  const int grid = sm_count() * 8;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const unsigned char* src = static_cast<const unsigned char*>(raw);
  unsigned char* dst = static_cast<unsigned char*>(clips);
  if (a16) gather_rows_kernel<16><<<grid, threads, 0, s>>>(src, dst, n_rows, row_bytes, es, geom);
  else if (es == 8) gather_rows_kernel<8><<<grid, threads, 0, s>>>(src, dst, n_rows, row_bytes, es, geom);
  else if (es == 4) gather_rows_kernel<4><<<grid, threads, 0, s>>>(src, dst, n_rows, row_bytes, es, geom);
  else gather_rows_kernel<2><<<grid, threads, 0, s>>>(src, dst, n_rows, row_bytes, es, geom);
  ++g_launches;
  return static_cast<int>(cudaGetLastError());
}

int eegfe_sliding_windows(const void* clips, int dtype, int64_t n_clips, int n_ch, void* windows, void* stream)
{
  const int es = esize_of(dtype);
  if (es == 0) return EEGFE_EDTYPE;
  if (n_clips < 0 || n_ch <= 0) return EEGFE_EINVAL;
  if (n_clips == 0) return 0;
  if (clips == nullptr || windows == nullptr) return EEGFE_EINVAL;
  const int threads = 256;
  const int grid = sm_count() * 8;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const unsigned char* src = static_cast<const unsigned char*>(clips);
  unsigned char* dst = static_cast<unsigned char*>(windows);
  // window starts are multiples of 50 elements: 8-byte aligned for 4- and 8-byte types, 4-byte for 2-byte types
  const bool base_ok = reinterpret_cast<uintptr_t>(clips) % 8 == 0 && reinterpret_cast<uintptr_t>(windows) % 8 == 0;
  if (es >= 4 && base_ok) sliding_windows_kernel<8><<<grid, threads, 0, s>>>(src, dst, n_clips, n_ch, es);
  else if (es >= 4 || base_ok) sliding_windows_kernel<4><<<grid, threads, 0, s>>>(src, dst, n_clips, n_ch, es);
  else sliding_windows_kernel<2><<<grid, threads, 0, s>>>(src, dst, n_clips, n_ch, es);
  ++g_launches;
  return static_cast<int>(cudaGetLastError());
}

int eegfe_launch_geometry(int mode, int* grid, int* block, int* smem_bytes, int* rows_per_tile)
{
  int g = 0, b = 0, s = 0, r = 0;
  switch (mode) {
    case EEGFE_MODE_500MS: g = CfgSliding500::kCtasPerSm; b = CfgSliding500::kThreads; s = CfgSliding500::kSmemBytes; r = CfgSliding500::kRows; break;
    case EEGFE_MODE_1S: g = CfgOneSec::kCtasPerSm; b = CfgOneSec::kThreads; s = CfgOneSec::kSmemBytes; r = CfgOneSec::kRows; break;
    case EEGFE_MODE_2S: g = CfgTwoSec::kCtasPerSm; b = CfgTwoSec::kThreads; s = CfgTwoSec::kSmemBytes; r = CfgTwoSec::kRows; break;
    default: return EEGFE_EINVAL;
  }
  if (grid) *grid = g * sm_count();
  if (block) *block = b;
  if (smem_bytes) *smem_bytes = s;
  if (rows_per_tile) *rows_per_tile = r;
  return 0;
}

int64_t eegfe_launch_count(void) { return g_launches; }

}  // extern "C"
