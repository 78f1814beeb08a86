// libeegfe.so -- fused segmentation + Hann + 200-point FFT + five-band DE/PSD for B200 (sm_100a).
//
// Persistent kernels, one CTA per SM.  Rows (one row = one channel of one 2 s clip, or one pre-cut window) are
// addressed by index arithmetic on the raw recording (clip (c, r) of a block starts at sample c*2600 + 600 + r*400;
// window w at +50 w) and pulled into a ring of shared-memory slots with 1-D TMA bulk copies (cp.async.bulk ...
// mbarrier::complete_tx, SASS UBLKCP), so neither the clip tensor nor the sliding-window tensor of the reference is
// ever materialised.  One thread owns one channel-window (or one of its two sweeps): it reads its samples from shared
// memory, runs the register-resident prime-factor FFT of bandpower.cuh in packed f32x2 arithmetic and writes 5 PSD + 5 DE
// values (500 ms kernel: straight from the lane; 1 s / 2 s kernel: through a staged tile that leaves as linear,
// coalesced stores).  No tensor cores (FFT + reduction, no dense contraction), no inter-CTA traffic.
//
//   eegfe_stream.cuh   de_psd_stream_kernel   500 ms windows (sliding over clip rows, or pre-cut): FP32-pipe-bound;
//                                             16 identical warps draw passes from a counter, tile duties fall to the
//                                             last finisher
//   this file          de_psd_kernel          1 s / 2 s windows: HBM-bound; producer warps + worker groups on a ring
//                      (both)                 rows that are only 8- / 4-byte aligned: TMA copies of the 16-byte aligned
//                                             span around each row, read shifted (SHIFT instantiations)
//                      gather / sliding-window / statistics kernels for the materialising and "next row" entry points
//
// C ABI: include/eegfe.h.  Reference semantics: see the citations in that header and in bandpower.cuh.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <atomic>

#include "../../include/eegfe.h"
#include "bandpower.cuh"

namespace eegfe {

// ---------------------------------------------------------------------------------------------------------------
// job description shared by all feature kernels
// ---------------------------------------------------------------------------------------------------------------
struct Job {
  // (time, channel, block) tensor map over the input for the kernels that fetch a tile with ONE TMA tensor copy
  // (valid iff tiles_per_clip > 0 or rows_tma); first member: the TMA unit wants it 64-byte aligned in param space
  CUtensorMap map;
  const float* in;
  float* de;
  float* psd;
  int* status;
  unsigned total_rows;    // n_units * n_ch of THIS launch (< 2^31; the host splits larger jobs)
  long long base;         // element offset of unit u: base + (u / d1) * s0 + ((u % d1) / d2) * s1 + (u % d2) * s2
  long long s0;
  int s1, s2;
  long long ch_stride;    // elements between channel rows of one unit
  unsigned d1, d2;
  unsigned n_ch;
  unsigned n_ch_magic;    // min(floor(2^32 / n_ch), 2^32 - 1): __umulhi(g, magic) is floor(g / n_ch) or one less
  unsigned d1_magic, d2_magic;   // the same for d1 and d2 (set by launch<>; used by the streaming kernel's hot paths)
  long long t_extent;     // valid elements of a channel row (tensor-map bound of the time axis); 0 = unknown
  int row_align;            // bytes every row start is a multiple of: 16 (TMA bulk copies of the rows themselves), else 8
                            // or 4 (TMA copies of the aligned span around each row, read shifted: SHIFT instantiations)
  unsigned tiles_per_clip;  // > 0: tiles never straddle clips (tile = kRows channels of one clip), fetched as tensor boxes
  int rows_tma;             // pre-cut windows in a dense / uniformly strided 2-D array: tile = tensor box of kRows rows
  // optional second product of the 500 ms kernel (GLMNet raw branch): clips_norm[row][400] = (x - mean[ch]) * scale[ch]
  float* norm_out;
  const float* norm_scale;
  const float* norm_mean;
  long long norm_row0;    // global row index of this launch's first row (rows of earlier launches of a split job)
};

// per-mode compile-time geometry
template <int LOAD_, int NWIN_, int HOP_, int NI_, int HANN_, int ROWS_, int GROUPS_, int SLOTS_, int CTAS_, int PAD_,
          int VEC_, int SPLIT_, bool LANEMAP_, int PRODUCERS_>
struct Cfg {
  static constexpr int kLoad = LOAD_;        // samples fetched per row
  static constexpr int kWindows = NWIN_;     // analysis windows per row
  static constexpr int kHop = HOP_;          // samples between window starts
  static constexpr int kNi = NI_;            // live inputs per radix-8 group (4: 100-sample window, 8: 200)
  static constexpr int kHann = HANN_;
  static constexpr int kRows = ROWS_;        // rows per tile
  static constexpr int kGroups = GROUPS_;    // worker groups per CTA; group g owns the CTA's tiles g, g + G, ...
  static constexpr int kSlots = SLOTS_;      // input ring slots per CTA (> kGroups: the surplus is prefetch depth)
  static constexpr int kCtasPerSm = CTAS_;
  static constexpr int kRowStride = LOAD_ + PAD_;              // floats between rows in shared memory (bank skew)
  static constexpr int kVec = VEC_;          // floats per shared-memory load (2: LDS.64, 4: LDS.128)
  static constexpr int kSplit = SPLIT_;      // threads per channel-window (2: even / odd sweep in different warps)
  static constexpr bool kLaneMap = LANEMAP_; // thread -> (row, window) through c_lane_map_500
  static constexpr int kUnits = ROWS_ * NWIN_;                 // channel-windows per tile
  static constexpr int kGroupThreads = kUnits * SPLIT_;        // worker threads per group
  static constexpr int kGroupWarps = kGroupThreads / 32;
  static constexpr int kWorkers = kGroupThreads * GROUPS_;
  static constexpr int kProducers = PRODUCERS_;                // producer warps; warp p loads tiles p, p + P, ...
  static constexpr int kThreads = kWorkers + 32 * PRODUCERS_;
  static constexpr int kRowBytes = LOAD_ * 4;
  static constexpr int kSlotFloats = ROWS_ * kRowStride;
  static constexpr int kOutFloats = kUnits * 5;                // per staging array, laid out [window][row][band]
  static constexpr int kSmemBytes = (SLOTS_ * kSlotFloats + GROUPS_ * 2 * kOutFloats + kWorkers) * 4;
  static_assert(kUnits % 32 == 0, "a tile must fill whole warps");
  static_assert(kRowBytes % 16 == 0 && kRowStride % 4 == 0, "TMA bulk copies need 16-byte aligned rows");
  static_assert(SPLIT_ == 1 || SPLIT_ == 2, "one or two threads per channel-window");
  static_assert(GROUPS_ <= 15, "one named barrier per group");
  // every slot must be refilled by ONE producer warp, in order: with tiles of a slot spread over several producers
  // one of them can wait on the slot's empty barrier two phases ahead, where the parity test reads "free"
  static_assert(SLOTS_ % PRODUCERS_ == 0, "producer p owns the slots congruent to p modulo kProducers");
  static_assert(kSmemBytes <= 227 * 1024, "shared memory per CTA");
};
// Shared-memory bank rules behind PAD / VEC (B200: 32 banks x 4 B; 64-bit loads are served per half-warp,
// 128-bit loads per quarter-warp):
//  * sliding 500 ms: windows start every 50 floats (8-byte aligned) -> LDS.64; stride 404 + the lane map of
//    eegfe_tables.h gives 16 wavefronts per load step (14 ideal, 28 with dense rows and thread = 7 row + window);
//  * 1 s / 2 s / pre-cut: windows start 16-byte aligned -> LDS.128 with lanes = consecutive rows, conflict-free
//    iff (row stride / 4) is odd: 404 -> 101, 204 -> 51, 100 -> 25.
// Group count: the 1 s kernels are where FFT work and HBM traffic are closest to balance: the memory pipeline alone
// (-DEEGFE_NOCOMPUTE) reaches 94 % of the measured HBM peak there, with four worker groups the full kernel 83 %.
// Seven groups (14 worker + 2 producer warps = 16 warps at 128 registers, no spills in the split-sweep form) keep the
// FP32 pipe fed while other groups wait at their store barriers: 5 groups 87 %, 6 groups 97 %, 7 groups 100 % of the
// (copy-measured) HBM peak.  The 2 s kernel is bounded by the memory side instead -- 800-byte requests, 83 % with or
// without the FFT with two producer warps: it was the ISSUE rate of its 800-byte bulk copies (~30 issue cycles each on a
// producer warp).  Four producer warps: 89 %; more groups do not help (they spill at 128 registers).
// Ring sizing: these kernels are HBM-bound at 800 B (400 B) per channel-window and want ~100 KB per SM in flight, hence
// small tiles and as many surplus slots as shared memory holds; a tile is due every ~1 us per SM, so each group writes
// its own tile and the producers only issue copies.  One bulk copy per row costs a producer warp ~30 issue cycles
// (the TMA operands are per-lane, the instruction is uniform): two producer warps.
// CfgSliding500 / CfgWin100 only name the two shapes that run on the streaming kernel (eegfe_stream.cuh).
//                         LOAD NWIN HOP NI  HANN          ROWS GRP SLOT CTAS PAD VEC SPLIT LANEMAP PROD
using CfgSliding500 = Cfg<400, 7, 50, 4, kHannHalfSec, 32, 1, 2, 2, 4, 2, 1, true, 1>;
using CfgOneSec     = Cfg<400, 2, 200, 8, kHannOneSec, 16, 7, 8, 1, 4, 4, 2, false, 2>;  // 7 x 64 + 64 thr, 218 KB
using CfgTwoSec     = Cfg<200, 1, 0, 8, kHannTwoSec, 32, 4, 8, 1, 4, 4, 2, false, 4>;    // 4 x 64 + 128 (samples 0..199)
using CfgWin100     = Cfg<100, 1, 0, 4, kHannHalfSec, 64, 4, 8, 1, 0, 4, 1, false, 2>;
using CfgWin200     = Cfg<200, 1, 0, 8, kHannOneSec, 32, 7, 8, 1, 4, 4, 2, false, 2>;    // 7 x 64 + 64, pre-cut 1 s

__constant__ unsigned char c_lane_map_500[224] = {EEGFE_LANE_MAP_500};

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
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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

// one tiled tensor copy (SASS UTMALDG): box (c0.., c1.., c2) of the tensor map -> dense rows in shared memory.
// The innermost start c0 must sit on a 16-byte boundary (measured: c0 = 2 floats raises "illegal instruction",
// tools/microbench/tma_align.cu), which is why the 50-sample hop of the 500 ms windows cannot ride on it.
__device__ __forceinline__ void tma_load_box(void* dst_smem, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar)
{
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          smem_u32(dst_smem)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}

// row g (32-bit) -> element offset of its first sample, and the index of its first output value
__device__ __forceinline__ long long row_offset(const Job& job, unsigned g, int n_windows, int* out_base)
{
  const unsigned u = g / job.n_ch;
  const unsigned ch = g - u * job.n_ch;
  const unsigned q = u / job.d1;
  const unsigned rem = u - q * job.d1;
  const unsigned c = rem / job.d2;
  const unsigned r = rem - c * job.d2;
  if (out_base) *out_base = static_cast<int>((u * n_windows * job.n_ch + ch) * 5u);
  return job.base + static_cast<long long>(q) * job.s0 + static_cast<long long>(c) * job.s1 +
         static_cast<long long>(r) * job.s2 + static_cast<long long>(ch) * job.ch_stride;
}

// q = floor(g / d), r = g - q d without a hardware division: multiply-high by floor(2^32 / d) is q or q - 1
__device__ __forceinline__ unsigned div_magic(unsigned g, unsigned d, unsigned magic, unsigned& r)
{
  unsigned q = __umulhi(g, magic);
  r = g - q * d;
  if (r >= d) {
    r -= d;
    ++q;
  }
  return q;
}

// row_offset() for the kernels whose Job carries the magic numbers (launch<>): same result, a third of the instructions
__device__ __forceinline__ long long row_offset_fast(const Job& job, unsigned g)
{
  unsigned ch, rem, r;
  const unsigned u = div_magic(g, job.n_ch, job.n_ch_magic, ch);
  const unsigned q = div_magic(u, job.d1, job.d1_magic, rem);
  const unsigned c = div_magic(rem, job.d2, job.d2_magic, r);
  return job.base + static_cast<long long>(q) * job.s0 + static_cast<long long>(c) * job.s1 +
         static_cast<long long>(r) * job.s2 + static_cast<long long>(ch) * job.ch_stride;
}

// E_b -> psd_b = E_b / count_b (DE_PSD.py:66), de_b = log2(100 psd_b) (:68); counts 4, 5, 7, 18, 69.
// The reciprocal multiply and MUFU.LG2 are each good to ~1 ulp (1e-7 relative / 4e-6 absolute at DE ~ 20),
// two orders inside the parity bars.
__device__ __forceinline__ float inv_count(int b)
{
  return b == 0 ? 1.0f / 4.0f : b == 1 ? 1.0f / 5.0f : b == 2 ? 1.0f / 7.0f : b == 3 ? 1.0f / 18.0f : 1.0f / 69.0f;
}
__device__ __forceinline__ bool band_features(const float (&e)[5], float (&psd)[5], float (&de)[5])
{
  bool zero = false;
#pragma unroll
  for (int b = 0; b < 5; ++b) {
    psd[b] = e[b] * inv_count(b);
    zero |= (psd[b] == 0.0f);
    de[b] = __log2f(100.0f * psd[b]);
  }
  return zero;
}

// Write one staged tile to HBM.  The staging arrays are laid out [window][row][band], the features in HBM
// [clip][window][channel][band]: for one window, the rows of the tile that belong to the same clip form ONE
// contiguous run of 5 * rows floats on both sides, so the copy is linear (no per-element index arithmetic) with
// thread `t` of `nthreads` taking consecutive addresses.  A tile spans 1 + (kRows - 1) / n_ch clips at most.
// kSplit == 2: the two sweeps' partials are added and turned into (psd, de) here.
// Returns true if a zero-power band was seen (kSplit == 2 only).
template <class C, int NT>
__device__ __forceinline__ bool store_tile(const Job& job, const float* out_a, const float* out_b, unsigned row0,
                                           int nrows, int t)
{
  constexpr int kIter = (C::kRows * 5 + NT - 1) / NT;      // a run is at most 5 kRows floats
  const int win_stride = static_cast<int>(job.n_ch) * 5;
  bool zero = false;
  int ra = 0;
  while (ra < nrows) {
    const unsigned g = row0 + ra;
    const unsigned u = g / job.n_ch;
    const int ch = static_cast<int>(g - u * job.n_ch);
    // one window per row: [clip][channel][band] is contiguous across clips too -> the whole tile is a single run
    const int seg = C::kWindows == 1 ? nrows : min(nrows - ra, static_cast<int>(job.n_ch) - ch);
    const int n = seg * 5;
    const int obase = static_cast<int>((u * C::kWindows * job.n_ch + ch) * 5u);
    const float* sa = out_a + ra * 5 + t;
    const float* sb = out_b + ra * 5 + t;
    float* gde = job.de + obase + t;
    float* gpsd = job.psd + obase + t;
#pragma unroll 1
    for (int w = 0; w < C::kWindows; ++w) {
      // all loads first, then all stores: the copy is latency-bound on one warp otherwise
      float va[kIter], vb[kIter];
#pragma unroll
      for (int k = 0; k < kIter; ++k)
        if (t + k * NT < n) {
          va[k] = sa[k * NT];
          vb[k] = sb[k * NT];
        }
#pragma unroll
      for (int k = 0; k < kIter; ++k)
        if (t + k * NT < n) {
          if constexpr (C::kSplit == 1) {
            gde[k * NT] = va[k];
            gpsd[k * NT] = vb[k];
          } else {
            const float p = __fadd_rn(va[k], vb[k]) * inv_count((t + k * NT) % 5);
            zero |= (p == 0.0f);
            gpsd[k * NT] = p;
            gde[k * NT] = __log2f(100.0f * p);
          }
        }
      sa += C::kRows * 5;
      sb += C::kRows * 5;
      gde += win_stride;
      gpsd += win_stride;
    }
    ra += seg;
  }
  return zero;
}

__device__ __forceinline__ void group_barrier(int id, int nthreads)
{
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// worker-local unit index -> (row in tile, window)
template <class C>
__device__ __forceinline__ void unit_to_row_window(int unit, int& row, int& w)
{
  if constexpr (C::kLaneMap) {
    const int code = c_lane_map_500[unit];
    row = code >> 3;
    w = code & 7;
  } else if constexpr (C::kWindows == 1) {
    row = unit;
    w = 0;
  } else {                       // lanes run over consecutive rows, the window is warp-uniform
    w = unit / C::kRows;
    row = unit - w * C::kRows;
  }
}

}  // namespace eegfe
#include "eegfe_stream.cuh"
namespace eegfe {

// ---------------------------------------------------------------------------------------------------------------
// the ring kernel (1 s / 2 s modes, pre-cut 1 s windows): warp-specialised, no block-wide barrier in the steady state
//
// A CTA walks its tiles m = 0, 1, 2, ... (global tile blockIdx.x + m * gridDim.x).  Tile m lands in ring slot
// m % kSlots and is processed by worker group m % kGroups.
//
//   producer warp p: for its tiles m = p, p + P, ...   wait empty[slot] -> arm full[slot], publish armed[slot],
//                                                       one TMA bulk copy per row
//   worker group g : for its tiles m = g, g + G, ...   wait armed[slot], full[slot] -> one sweep of one channel-window
//                                                       per thread, register FFT -> arrive empty[slot]
//                                                       group barrier -> staging tile [window][row][band] -> group
//                                                       barrier -> the group writes its tile to HBM (store_tile)
//
// kSplit == 2: even-sweep warps and odd-sweep warps stage their partial band energies; store_tile adds them and does
//              the epilogue (the same additions as the one-thread form, so the results are bit-identical).
// kSplit == 1: a worker runs both sweeps and the epilogue and stages (de, psd).
//
// SHIFT (rows that are only 4- or 8-byte aligned: an odd block length or stride): TMA cannot start a copy there, but
// it can fetch the 16-byte aligned span AROUND the row -- from the row start rounded down to the row end rounded up,
// at most 16 bytes more, which is exactly the slot's row padding -- so the row lands k = 0..3 floats into its
// shared-memory row.  The producer notes k per row, the worker adds it to its window pointer and reads scalar
// (LDS.32; SHIFT = 2, every row 8-byte aligned: LDS.64) instead of LDS.128.  Same HBM traffic and copy count as the
// aligned kernel, no second pass.  The span stays
// inside the caller's allocation as long as that starts and ends on 16-byte boundaries (cudaMalloc, torch: always).
// ---------------------------------------------------------------------------------------------------------------
template <class C, int SHIFT = 0, bool TENSOR = false>
__global__ void __launch_bounds__(C::kThreads, C::kCtasPerSm) de_psd_kernel(const __grid_constant__ Job job)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* const ring = reinterpret_cast<float*>(smem_raw);
  float* const out_stage = ring + C::kSlots * C::kSlotFloats;               // [group][2][kOutFloats]
  // per worker thread: where its window starts inside an input slot, where its 5 results go in the staging tile
  // and its row -- packed into one word that is re-read from shared memory every tile (the compiler otherwise
  // rematerialises the lane-map look-up, an indexed constant load with ~1 us of latency, twice per tile).
  int* const thread_meta = reinterpret_cast<int*>(out_stage + C::kGroups * 2 * C::kOutFloats);   // [kWorkers]
  __shared__ uint64_t full_bar[C::kSlots], empty_bar[C::kSlots];
  // armed[s] = number of tiles whose copies have been ISSUED into slot s.  A parity wait on full[s] only tells
  // "phase k" from "phase k - 2" if the waiter is at most one phase ahead of the barrier; worker groups advance
  // independently, so a group can reach tile m before the slot's previous tile (m - kSlots, another group's unless
  // kSlots % kGroups == 0) has even been requested.  Waiting for armed[s] > k first makes the parity wait unambiguous.
  __shared__ unsigned armed[C::kSlots];
  __shared__ unsigned char row_shift[SHIFT ? C::kSlots : 1][SHIFT ? C::kRows : 1];   // k of every row in flight
  static_assert(!SHIFT || (C::kRowStride >= C::kLoad + 4 && C::kRows <= 32), "a shifted row needs 16 bytes of padding");
  constexpr int kVec = SHIFT == 1 ? 1 : (SHIFT == 2 ? 2 : C::kVec);     // SHIFT 2: every k is 0 or 2 -> LDS.64

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  // Tiles: kRows consecutive rows -- or, with job.tiles_per_clip (tensor-copy producer), kRows consecutive CHANNELS
  // of one clip, so that a tile is a rectangle of the (time, channel, block) tensor; a clip's last tile is short.
  // TENSOR: tiles fetched by one TMA tensor copy each (rows that fit one box, inner extent <= 256) -- an instantiation
  // of its own (measurement option, eegfe_set_tensor_loads), so that the default kernels carry none of its code
  constexpr bool kTensorLoads = TENSOR;
  static_assert(!TENSOR || (C::kLoad == 200 && C::kHann == kHannTwoSec && !SHIFT), "tensor boxes: the 2 s shape only");
  const unsigned tpc = kTensorLoads ? job.tiles_per_clip : 0u;
  const unsigned n_tiles = tpc ? (job.total_rows / job.n_ch) * tpc : (job.total_rows + C::kRows - 1) / C::kRows;
  // tiles of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
  const int n_mine = blockIdx.x < n_tiles ? static_cast<int>((n_tiles - 1 - blockIdx.x) / gridDim.x) + 1 : 0;
  auto tile_row0 = [&](int m) {
    const unsigned gt = blockIdx.x + static_cast<unsigned>(m) * gridDim.x;
    if (tpc == 0) return gt * C::kRows;
    const unsigned clip = gt / tpc;
    return clip * job.n_ch + (gt - clip * tpc) * C::kRows;
  };
  auto tile_nrows = [&](unsigned row0) {
    unsigned left = job.total_rows - row0;
    if (tpc) left = job.n_ch - row0 % job.n_ch;          // rows to the end of the clip
    return static_cast<int>(left < C::kRows ? left : C::kRows);
  };

  if (tid < C::kWorkers) {
    const int gt = tid % C::kGroupThreads;
    int row, w;
    unit_to_row_window<C>(gt % C::kUnits, row, w);
    // bits 0..13: window offset in the slot (floats), 14..24: staging index, 25..30: row in tile
    thread_meta[tid] = (row * C::kRowStride + w * C::kHop) | (((w * C::kRows + row) * 5) << 14) | (row << 25);
  }
  // per-slot state, set up by ALL of warp 0 (lane l: slot l % kSlots; four lanes write the same values): no warp may
  // diverge in front of the block barrier -- ptxas does not always reconverge it there and BAR.SYNC does not either
  // (eegfe_stream.cuh).  `tid < kWorkers` above is warp-uniform (kWorkers is a multiple of 32).
  static_assert(C::kWorkers % 32 == 0 && (C::kSlots & (C::kSlots - 1)) == 0 && C::kSlots <= 32, "set-up by whole warps");
  if (tid < 32) {
    const int sl = tid % C::kSlots;
    mbar_init(&full_bar[sl], 1);
    mbar_init(&empty_bar[sl], C::kGroupWarps);
    armed[sl] = 0;
    mbar_fence_init();
  }
  __syncthreads();

  if (tid >= C::kWorkers) {
    // ------------------------------------------------ producer warps ----------------------------------------------
    const int producer = (tid - C::kWorkers) / 32;
    for (int m = producer; m < n_mine; m += C::kProducers) {
      const int s = m % C::kSlots;
      mbar_wait(&empty_bar[s], ((m / C::kSlots) & 1) ^ 1);
      const unsigned row0 = tile_row0(m);
      const unsigned nrows = tile_nrows(row0);
      if (kTensorLoads && (tpc != 0 || job.rows_tma)) {
        // ONE tensor copy per tile: box (kRowStride samples, kRows rows) -> exactly the slot's padded rows.  Rows past
        // the clip's last channel / the array's last row are out of bounds: zero-filled, never read.
        if (lane == 0) {
          int x = 0, y = static_cast<int>(row0), z = 0;
          if (tpc != 0) {
            const unsigned u = row0 / job.n_ch;
            const unsigned q = u / job.d1;
            const unsigned rem = u - q * job.d1;
            const unsigned c = rem / job.d2;
            x = static_cast<int>(job.base) + static_cast<int>(c) * job.s1 + static_cast<int>(rem - c * job.d2) * job.s2;
            y = static_cast<int>(row0 - u * job.n_ch);
            z = static_cast<int>(q);
          }
          mbar_arrive_expect_tx(&full_bar[s], C::kRows * C::kRowStride * 4);
          st_release_smem(&armed[s], static_cast<unsigned>(m / C::kSlots) + 1u);
          tma_load_box(ring + s * C::kSlotFloats, &job.map, x, y, z, &full_bar[s]);
        }
        __syncwarp();
        continue;
      }
      if constexpr (SHIFT) {
        // the aligned span around every row (kRows <= 32: one row per lane)
        const float* src = nullptr;
        unsigned bytes = 0;
        if (lane < nrows) {
          const float* p = job.in + row_offset(job, row0 + lane, C::kWindows, nullptr);
          const unsigned k = static_cast<unsigned>(reinterpret_cast<uintptr_t>(p) >> 2) & 3u;
          row_shift[s][lane] = static_cast<unsigned char>(k);
          src = p - k;
          bytes = ((k + C::kLoad) * 4u + 15u) & ~15u;
        }
        const unsigned total = __reduce_add_sync(0xffffffffu, bytes);
        __syncwarp();                                              // row_shift[] is written before lane 0 arrives
        if (lane == 0) {
          mbar_arrive_expect_tx(&full_bar[s], total);
          st_release_smem(&armed[s], static_cast<unsigned>(m / C::kSlots) + 1u);
        }
        __syncwarp();
        if (lane < nrows) bulk_copy_g2s(ring + s * C::kSlotFloats + lane * C::kRowStride, src, bytes, &full_bar[s]);
        __syncwarp();
        continue;
      }
      if (lane == 0) {
        mbar_arrive_expect_tx(&full_bar[s], nrows * C::kRowBytes);
        st_release_smem(&armed[s], static_cast<unsigned>(m / C::kSlots) + 1u);
      }
      __syncwarp();
      for (unsigned r = lane; r < nrows; r += 32) {
        const long long off = row_offset(job, row0 + r, C::kWindows, nullptr);
        bulk_copy_g2s(ring + s * C::kSlotFloats + r * C::kRowStride, job.in + off, C::kRowBytes, &full_bar[s]);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------ worker warps ----------------------------------------------
    const int g = tid / C::kGroupThreads;
    const int gt = tid - g * C::kGroupThreads;
    const int sweep = gt / C::kUnits;                  // 0 when kSplit == 1; warp-uniform (kUnits % 32 == 0)
    float* const out_a = out_stage + g * 2 * C::kOutFloats;
    float* const out_b = out_a + C::kOutFloats;
    float* const out_mine = (C::kSplit == 2 && sweep == 1) ? out_b : out_a;
    int j = 0;                                         // this group's j-th tile
    for (int m = g; m < n_mine; m += C::kGroups, ++j) {
      const int s = m % C::kSlots;
      while (ld_acquire_smem(&armed[s]) <= static_cast<unsigned>(m / C::kSlots)) __nanosleep(20);
      mbar_wait(&full_bar[s], (m / C::kSlots) & 1);
      const unsigned row0 = tile_row0(m);
      const int nrows = tile_nrows(row0);
      const int meta = *reinterpret_cast<volatile int*>(thread_meta + tid);
      const bool live = (meta >> 25) < nrows;
      const float* win = ring + s * C::kSlotFloats + (meta & 0x3fff);
      if constexpr (SHIFT) win += live ? row_shift[s][meta >> 25] : 0;
      float va[5], vb[5];
#ifdef EEGFE_NOCOMPUTE      // (measurement builds only: the memory pipeline without the FFT -- one sample per window)
      if (live) {
#pragma unroll
        for (int b = 0; b < 5; ++b) va[b] = vb[b] = win[b * 4];
      }
#else
      if (live) {
        if constexpr (C::kSplit == 1) {
          float e[5];
          window_band_energy<C::kNi, C::kHann, kVec>(win, e);
          if (band_features(e, vb, va) && job.status != nullptr) atomicOr(job.status, EEGFE_STATUS_ZERO_POWER);
        } else {
          const float zero5[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
          sweep_any<C::kNi, C::kHann, kVec>(win, sweep, zero5, va);
        }
      }
#endif
      // (Handing the slot back right after the radix-8 stage, two DFT-25 earlier, measured 4 % SLOWER in 1 s mode.)
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s]);           // this warp no longer reads the input slot
      // staging tile is free again once every thread of the group finished storing the previous tile
      if (j > 0) group_barrier(1 + g, C::kGroupThreads);
      if (live) {
        const int slot_out = (*reinterpret_cast<volatile int*>(thread_meta + tid) >> 14) & 0x7ff;
#pragma unroll
        for (int b = 0; b < 5; ++b) {
          out_mine[slot_out + b] = va[b];
          if constexpr (C::kSplit == 1) out_b[slot_out + b] = vb[b];
        }
      }
      group_barrier(1 + g, C::kGroupThreads);              // all of the group's results are staged
      if (store_tile<C, C::kGroupThreads>(job, out_a, out_b, row0, nrows, gt) && job.status != nullptr)
        atomicOr(job.status, EEGFE_STATUS_ZERO_POWER);
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
    const unsigned char* s = src + row_offset(geom, static_cast<unsigned>(row), 1, nullptr) * esize;
    reinterpret_cast<vec_t*>(dst + row * row_bytes)[v] = reinterpret_cast<const vec_t*>(s)[v];
  }
}

// Rows that are only element-aligned (odd block lengths / strides): one warp per row, consecutive lanes copy
// consecutive elements -- coalesced on both sides and no per-element index arithmetic (the generic kernel above
// divides once per element and manages ~1.5 TB/s on such rows).
template <class T>
__global__ void __launch_bounds__(256) gather_rows_warp_kernel(const T* __restrict__ src, T* __restrict__ dst,
                                                                long long n_rows, int row_elems, Job geom)
{
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long n_warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long row = warp; row < n_rows; row += n_warps) {
    const T* s = src + row_offset(geom, static_cast<unsigned>(row), 1, nullptr);
    T* d = dst + row * row_elems;
    for (int i = lane; i < row_elems; i += 128) {
      T v0 = s[i], v1 = T(), v2 = T(), v3 = T();
      if (i + 32 < row_elems) v1 = s[i + 32];
      if (i + 64 < row_elems) v2 = s[i + 64];
      if (i + 96 < row_elems) v3 = s[i + 96];
      d[i] = v0;
      if (i + 32 < row_elems) d[i + 32] = v1;
      if (i + 64 < row_elems) d[i + 64] = v2;
      if (i + 96 < row_elems) d[i + 96] = v3;
    }
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

// clips [n_clips][n_ch][400] -> [n_clips][n_ch][100][7]: element (clip, ch, t, w) = clip sample 50 w + t -- the layout the
// Seq2Seq trainer stacks inline (EEG2Video_New/Seq2Seq/my_autoregressive_transformer.py:309-314, torch.stack(.., dim=-1)).
// One thread per output element: writes are coalesced, the reads of a warp stay inside one 1600-byte row.
template <class T>
__global__ void __launch_bounds__(256) sliding_windows_last_kernel(const T* __restrict__ clips, T* __restrict__ out,
                                                                    long long n_rows)
{
  const long long total = n_rows * 700;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = i / 700;
    const int e = static_cast<int>(i - row * 700);
    const int t = e / 7;
    const int w = e - t * 7;
    out[i] = clips[row * 400 + 50 * w + t];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// per-channel statistics of the clip samples (training-split normalisation of the GLMNet raw branch)
// ---------------------------------------------------------------------------------------------------------------
// One CTA per (block, channel) row: float64 sum and sum of squares over the 40 x 2000 clip samples of the row
// (hint periods skipped), fixed reduction order -> deterministic.
__global__ void __launch_bounds__(256) channel_row_sums_kernel(const float* __restrict__ raw, long long block_stride,
                                                                long long ch_stride, int n_ch, double* __restrict__ partial)
{
  const long long row = blockIdx.x;
  const long long blk = row / n_ch;
  const int ch = static_cast<int>(row - blk * n_ch);
  const float* base = raw + blk * block_stride + ch * ch_stride;
  double s = 0.0, ss = 0.0;
  for (int i = threadIdx.x; i < 40 * 2000; i += 256) {
    const int c = i / 2000;
    const float v = __ldg(base + c * 2600 + 600 + (i - c * 2000));
    s += v;
    ss += static_cast<double>(v) * v;
  }
  __shared__ double sh_s[8], sh_ss[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_down_sync(0xffffffffu, s, o);
    ss += __shfl_down_sync(0xffffffffu, ss, o);
  }
  if ((threadIdx.x & 31) == 0) {
    sh_s[threadIdx.x >> 5] = s;
    sh_ss[threadIdx.x >> 5] = ss;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < 8; ++w) {
      a += sh_s[w];
      b += sh_ss[w];
    }
    partial[2 * row] = a;
    partial[2 * row + 1] = b;
  }
}

// mean / population std per channel over the blocks selected by `mask` (one thread per channel, block order)
__global__ void channel_stats_finish_kernel(const double* __restrict__ partial, const unsigned char* __restrict__ mask,
                                            long long n_blocks, int n_ch, double* __restrict__ mean, double* __restrict__ std)
{
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= n_ch) return;
  double s = 0.0, ss = 0.0, n = 0.0;
  for (long long b = 0; b < n_blocks; ++b) {
    if (mask != nullptr && mask[b] == 0) continue;
    s += partial[2 * (b * n_ch + ch)];
    ss += partial[2 * (b * n_ch + ch) + 1];
    n += 40.0 * 2000.0;
  }
  const double m = n > 0 ? s / n : 0.0;
  const double var = n > 0 ? ss / n - m * m : 0.0;
  mean[ch] = m;
  std[ch] = sqrt(var > 0.0 ? var : 0.0);
}

// ---------------------------------------------------------------------------------------------------------------
// generic DE_PSD: any window length, any sampling rate (DE_PSD.py:33-39, :49-58) -- the slow, general path
// ---------------------------------------------------------------------------------------------------------------
// The reference takes ANY fre / time_window: Hann of L = int(time_window * fre) points, fft(., 200) truncating or
// zero-padding, bins fStartNum - 1 .. fEndNum - 1 with fNum = int(f / fre * 200) (a start of -1 is Python's "last
// element": bin 99).  The fused kernels cover the three shapes the drivers use (L = 100 / 200 / 400 at 200 Hz); every
// other shape runs here: one CTA per row, thread k forms bin k of the 200-point DFT directly from the (<= 200) live
// weighted samples with a sin/cos table in shared memory -- 20 k FMAs per row instead of 2.4 k, no pruning, fp32.
struct BandBins {
  int lo[5], hi[5];     // inclusive bin range per band; lo may be -1 (= bin 99)
  float inv_count[5];
};

__global__ void __launch_bounds__(128) de_psd_generic_kernel(const float* __restrict__ x, long long n_rows, int n_live,
                                                              long long row_stride, const float* __restrict__ hann,
                                                              BandBins bands, float* __restrict__ de,
                                                              float* __restrict__ psd, int* status)
{
  __shared__ float y[200], tc[200], ts[200], power[100];
  const int k = threadIdx.x;
  for (int i = k; i < 200; i += 128) sincospif(static_cast<float>(i) * 0.01f, &ts[i], &tc[i]);   // angle 2 pi i / 200
  for (long long row = blockIdx.x; row < n_rows; row += gridDim.x) {
    __syncthreads();
    for (int i = k; i < 200; i += 128) y[i] = i < n_live ? __fmul_rn(x[row * row_stride + i], hann[i]) : 0.0f;
    __syncthreads();
    if (k < 100) {
      float re = 0.0f, im = 0.0f;
      int idx = 0;
      for (int n = 0; n < n_live; ++n) {
        re = fmaf(y[n], tc[idx], re);
        im = fmaf(y[n], -ts[idx], im);
        idx += k;
        if (idx >= 200) idx -= 200;
      }
      power[k] = fmaf(re, re, im * im);
    }
    __syncthreads();
    if (k < 5) {
      float e = 0.0f;
      for (int b = bands.lo[k]; b <= bands.hi[k]; ++b) e += power[b < 0 ? b + 100 : b];
      const float p = e * bands.inv_count[k];
      psd[row * 5 + k] = p;
      de[row * 5 + k] = __log2f(100.0f * p);
      if (p == 0.0f && status != nullptr) atomicOr(status, EEGFE_STATUS_ZERO_POWER);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// de = log2(100 psd), elementwise: the SAME device expression the feature kernels use (band_features / store_tile),
// so a rank that received only PSD over NVLink rebuilds DE bit for bit (cohort gather moves half the bytes).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) de_from_psd_kernel(const float* __restrict__ psd, float* __restrict__ de,
                                                           long long n, int* status)
{
  const long long n4 = n >> 2;
  bool zero = false;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long t0 = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if ((reinterpret_cast<uintptr_t>(psd) | reinterpret_cast<uintptr_t>(de)) % 16 == 0) {
    for (long long i = t0; i < n4; i += stride) {
      const float4 p = reinterpret_cast<const float4*>(psd)[i];
      float4 d;
      d.x = __log2f(100.0f * p.x);
      d.y = __log2f(100.0f * p.y);
      d.z = __log2f(100.0f * p.z);
      d.w = __log2f(100.0f * p.w);
      zero |= (p.x == 0.0f) | (p.y == 0.0f) | (p.z == 0.0f) | (p.w == 0.0f);
      reinterpret_cast<float4*>(de)[i] = d;
    }
    for (long long i = 4 * n4 + t0; i < n; i += stride) {
      zero |= (psd[i] == 0.0f);
      de[i] = __log2f(100.0f * psd[i]);
    }
  } else {
    for (long long i = t0; i < n; i += stride) {
      zero |= (psd[i] == 0.0f);
      de[i] = __log2f(100.0f * psd[i]);
    }
  }
  if (zero && status != nullptr) atomicOr(status, EEGFE_STATUS_ZERO_POWER);
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
static std::atomic<long long> g_launches{0};
static std::atomic<long long> g_tma_launches{0};

// ---- per-device state: one process may drive several GPUs (cudaFuncSetAttribute and the SM count are per device) ----
constexpr int kMaxDevices = 64;

static int device_ordinal()
{
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
  return dev % kMaxDevices;
}

// Optional cap on the CTAs of the persistent feature kernels (0 = one per SM).  The kernels occupy an SM completely (all
// registers, ~207 KB of shared memory), so a concurrent NCCL send / receive kernel cannot start until some CTA retires:
// during a cohort gather the ranks leave a few SMs free (cohort.run_cohort) and the transfers overlap the FFTs.
static std::atomic<int> g_cta_limit{0};

static unsigned persistent_grid(unsigned wanted)
{
  const int cap = g_cta_limit.load(std::memory_order_relaxed);
  return (cap > 0 && wanted > static_cast<unsigned>(cap)) ? static_cast<unsigned>(cap) : wanted;
}

static int sm_count()
{
  static std::atomic<int> n[kMaxDevices];
  const int dev = device_ordinal();
  int v = n[dev].load(std::memory_order_relaxed);
  if (v == 0) {
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    n[dev].store(v, std::memory_order_relaxed);
  }
  return v;
}

// Opt one kernel in to `bytes` of dynamic shared memory on the current device, once per (kernel, device).
// `mask` is the kernel's own bit set of devices already configured (racing threads both set the attribute: harmless).
template <class K>
static int configure_smem(std::atomic<unsigned long long>& mask, K kernel, int bytes)
{
  const unsigned long long bit = 1ull << device_ordinal();
  if (mask.load(std::memory_order_acquire) & bit) return 0;
  const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return static_cast<int>(e);
  mask.fetch_or(bit, std::memory_order_release);
  return 0;
}

// 500 ms mode, 16-byte aligned rows: the streaming kernel (eegfe_stream.cuh), one CTA of 16 warps per SM.
template <class SC>
static int launch_stream(const Job& job, cudaStream_t stream)
{
  if (job.total_rows == 0) return 0;
  const unsigned n_tiles = (job.total_rows + SC::kRows - 1) / SC::kRows;
  unsigned grid = persistent_grid(static_cast<unsigned>(sm_count()));
  if (grid > n_tiles) grid = n_tiles;
  constexpr bool kCanNorm = SC::kLoad == 400 && SC::kWindows == 7;
  static std::atomic<unsigned long long> configured{0}, configured_norm{0};
  if (job.norm_out != nullptr) {
    if constexpr (kCanNorm) {
      const int rc = configure_smem(configured_norm, de_psd_stream_kernel<SC, true>, SC::kSmemBytes);
      if (rc != 0) return rc;
      de_psd_stream_kernel<SC, true><<<grid, SC::kThreads, SC::kSmemBytes, stream>>>(job);
    } else {
      return EEGFE_EINVAL;
    }
  } else if (job.row_align < 16) {
    // rows TMA cannot start a copy at -- kernels of their own, the default kernel carries none of their code.
    // Sliding 500 ms windows (rows padded in shared memory): TMA copies of the aligned span around each row, read
    // shifted (LDS.64 when every row is 8-byte aligned, else LDS.32); pre-cut windows (dense rows, nothing to shift
    // into): cp.async.
    if constexpr (kCanNorm) {
      if (job.row_align == 8) {
        static std::atomic<unsigned long long> configured_shift2{0};
        const int rc = configure_smem(configured_shift2, de_psd_stream_kernel<SC, false, false, 2>, SC::kSmemBytes);
        if (rc != 0) return rc;
        de_psd_stream_kernel<SC, false, false, 2><<<grid, SC::kThreads, SC::kSmemBytes, stream>>>(job);
      } else {
        static std::atomic<unsigned long long> configured_shift1{0};
        const int rc = configure_smem(configured_shift1, de_psd_stream_kernel<SC, false, false, 1>, SC::kSmemBytes);
        if (rc != 0) return rc;
        de_psd_stream_kernel<SC, false, false, 1><<<grid, SC::kThreads, SC::kSmemBytes, stream>>>(job);
      }
    } else {
      static std::atomic<unsigned long long> configured_small{0};
      const int rc = configure_smem(configured_small, de_psd_stream_kernel<SC, false, true>, SC::kSmemBytes);
      if (rc != 0) return rc;
      de_psd_stream_kernel<SC, false, true><<<grid, SC::kThreads, SC::kSmemBytes, stream>>>(job);
    }
  } else {
    const int rc = configure_smem(configured, de_psd_stream_kernel<SC, false>, SC::kSmemBytes);
    if (rc != 0) return rc;
    de_psd_stream_kernel<SC, false><<<grid, SC::kThreads, SC::kSmemBytes, stream>>>(job);
  }
  ++g_launches;
  return static_cast<int>(cudaGetLastError());
}

// ---- tensor maps for the ring kernel's optional tensor-copy producer ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn()
{
  // libcuda is not linked: the entry point comes from the runtime (nullptr on a driver without tensor maps)
  static const EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// Tensor-copy producer of the ring kernel: OFF unless switched on (eegfe_set_tensor_loads / -DEEGFE_TENSOR_LOADS=1).
// Measured on B200 (tools/kbench.py, 24 subjects, 2 s mode): 6.0 G channel-windows/s with one tensor copy per
// clip-aligned tile against 6.85 G with one bulk copy per row, whatever the L2 promotion -- so it is kept for
// measurements (DRAM traffic with sector-granular fetches), not as the default.
#ifndef EEGFE_TENSOR_LOADS
#define EEGFE_TENSOR_LOADS 0
#endif
static std::atomic<int> g_tensor_loads{EEGFE_TENSOR_LOADS};

#ifndef EEGFE_TMA_L2_PROMOTION
#define EEGFE_TMA_L2_PROMOTION CU_TENSOR_MAP_L2_PROMOTION_NONE
#endif
// Tensor map for the ring kernel's tensor-copy producer (200-sample rows: 2 s mode, pre-cut 1 s / 2 s windows).
// Fills job.map and job.tiles_per_clip / job.rows_tma; returns false when the job has to stay on per-row bulk copies
// (few channels, aliased blocks, no driver support).  L2 promotion NONE: a 2 s row is 800 B out of every 1600 B, and
// with the 128-byte promotion of the 1-D bulk copies DRAM read 896 B of it (profiles/ncu_traffic.json, round 1).
template <class C>
static bool attach_tensor_map(Job& job)
{
  job.tiles_per_clip = 0;
  job.rows_tma = 0;
  if (g_tensor_loads.load(std::memory_order_relaxed) == 0) return false;
  const EncodeTiledFn encode = encode_tiled_fn();
  if (encode == nullptr) return false;
  const bool rows_mode = (job.n_ch == 1 && job.d1 == 1 && job.ch_stride == 0);
  cuuint64_t gdim[3], gstride[2];
  cuuint32_t box[3] = {static_cast<cuuint32_t>(C::kRowStride), static_cast<cuuint32_t>(C::kRows), 1}, estride[3] = {1, 1, 1};
  unsigned tpc = 0;
  if (rows_mode) {
    if (job.s0 < C::kLoad || job.base != 0) return false;
    gdim[0] = C::kLoad;                      // samples past the window are out of bounds: zero-filled, never read
    gdim[1] = job.total_rows;
    gdim[2] = 1;
    gstride[0] = static_cast<cuuint64_t>(job.s0) * 4;
    gstride[1] = gstride[0] * job.total_rows;
  } else {
    if (job.n_ch < 24 || job.t_extent <= 0 || job.base < 0) return false;
    const unsigned n_clips = job.total_rows / job.n_ch;
    const unsigned n_blocks = (n_clips + job.d1 - 1) / job.d1;
    if (n_blocks > 1 && job.s0 <= 0) return false;          // aliased blocks (stride 0): no tensor view
    tpc = (job.n_ch + C::kRows - 1) / C::kRows;
    gdim[0] = static_cast<cuuint64_t>(job.t_extent);
    gdim[1] = job.n_ch;
    gdim[2] = n_blocks;
    gstride[0] = static_cast<cuuint64_t>(job.ch_stride) * 4;
    gstride[1] = n_blocks > 1 ? static_cast<cuuint64_t>(job.s0) * 4 : gstride[0] * job.n_ch;
  }
  if (gstride[0] == 0 || gstride[0] % 16 != 0 || gstride[1] % 16 != 0 || gstride[0] >= (1ull << 40) ||
      gstride[1] >= (1ull << 40) || gdim[0] > (1ull << 32) || gdim[1] > (1ull << 32) || gdim[2] > (1ull << 32))
    return false;
  if (encode(&job.map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(job.in), gdim, gstride, box, estride,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, EEGFE_TMA_L2_PROMOTION,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return false;
  job.tiles_per_clip = tpc;
  job.rows_tma = rows_mode ? 1 : 0;
  return true;
}

// One launch over `job.total_rows` rows (< 2^31, output indices < 2^31: guaranteed by run_units()).
template <class C>
static int launch(const Job& job_in, bool aligned16, cudaStream_t stream)
{
  if (job_in.total_rows == 0) return 0;
  Job job = job_in;
  job.tiles_per_clip = 0;
  job.rows_tma = 0;
  job.row_align = 16;
  {
    auto magic = [](unsigned d) {
      const unsigned long long m = (1ull << 32) / (d ? d : 1u);
      return m > 0xffffffffull ? 0xffffffffu : static_cast<unsigned>(m);
    };
    job.n_ch_magic = magic(job.n_ch);
    job.d1_magic = magic(job.d1);
    job.d2_magic = magic(job.d2);
  }
  unsigned n_tiles = (job.total_rows + C::kRows - 1) / C::kRows;
  if (!aligned16 && job.norm_out == nullptr) {
    // rows that are only 8- / 4-byte aligned: TMA copies of the 16-byte aligned span around each row, read shifted
    // (SHIFT instantiations of both kernels)
    const bool even = reinterpret_cast<uintptr_t>(job.in) % 8 == 0 && job.base % 2 == 0 && job.s0 % 2 == 0 &&
                      job.s1 % 2 == 0 && job.s2 % 2 == 0 && job.ch_stride % 2 == 0;
    job.row_align = even ? 8 : 4;
    aligned16 = true;                        // same kernels as aligned rows, loader switched by job.row_align
  }
  if (aligned16) {
    if constexpr (C::kLoad == 400 && C::kWindows == 7) {
      return launch_stream<StreamCfg>(job, stream);        // 500 ms sliding windows: the streaming kernel
    } else if constexpr (C::kLoad == 100 && C::kWindows == 1) {
      return launch_stream<StreamCfgWin100>(job, stream);  // pre-cut 500 ms windows: the streaming kernel, dense rows
    } else {
      if (job.norm_out != nullptr) return EEGFE_EINVAL;
      if constexpr (C::kLoad == 200 && C::kHann == kHannTwoSec) {
        // 200-sample rows at a 204-float pitch are one tensor box per tile
        if (job.row_align == 16 && attach_tensor_map<C>(job)) {
          ++g_tma_launches;
          if (job.tiles_per_clip) n_tiles = (job.total_rows / job.n_ch) * job.tiles_per_clip;
        }
      }
      unsigned grid = persistent_grid(static_cast<unsigned>(sm_count()) * C::kCtasPerSm);
      if (grid > n_tiles) grid = n_tiles;
      static std::atomic<unsigned long long> configured{0};
      if (job.tiles_per_clip != 0 || job.rows_tma != 0) {
        if constexpr (C::kLoad == 200 && C::kHann == kHannTwoSec) {
          static std::atomic<unsigned long long> configured_tensor{0};
          const int rc = configure_smem(configured_tensor, de_psd_kernel<C, 0, true>, C::kSmemBytes);
          if (rc != 0) return rc;
          de_psd_kernel<C, 0, true><<<grid, C::kThreads, C::kSmemBytes, stream>>>(job);
        }
      } else if (job.row_align == 8) {
        static std::atomic<unsigned long long> configured_shift2{0};
        const int rc = configure_smem(configured_shift2, de_psd_kernel<C, 2>, C::kSmemBytes);
        if (rc != 0) return rc;
        de_psd_kernel<C, 2><<<grid, C::kThreads, C::kSmemBytes, stream>>>(job);
      } else if (job.row_align < 16) {
        static std::atomic<unsigned long long> configured_shift1{0};
        const int rc = configure_smem(configured_shift1, de_psd_kernel<C, 1>, C::kSmemBytes);
        if (rc != 0) return rc;
        de_psd_kernel<C, 1><<<grid, C::kThreads, C::kSmemBytes, stream>>>(job);
      } else {
        const int rc = configure_smem(configured, de_psd_kernel<C>, C::kSmemBytes);
        if (rc != 0) return rc;
        de_psd_kernel<C><<<grid, C::kThreads, C::kSmemBytes, stream>>>(job);
      }
    }
  } else {
    return EEGFE_EINVAL;                     // the normalised-clip product needs 16-byte aligned rows
  }
  ++g_launches;
  return static_cast<int>(cudaGetLastError());
}

// Split a job of `n_units` units (each n_ch rows, n_windows windows per row) into launches whose row and output
// indices fit 32 bits.  Units are consumed in multiples of `unit_quantum` (d1, so that the affine unit -> offset
// map restarts cleanly); `in_step` is the input offset (elements) per quantum.
template <class C>
static int run_units(Job job, long long n_units, long long unit_quantum, long long in_step, bool aligned16,
                     cudaStream_t stream)
{
  const long long out_per_unit = static_cast<long long>(job.n_ch) * C::kWindows * 5;
  long long max_units = ((1LL << 31) - 1) / out_per_unit;
  max_units -= max_units % unit_quantum;
  if (max_units < unit_quantum) return EEGFE_EINVAL;
  const float* in0 = job.in;
  float* de0 = job.de;
  float* psd0 = job.psd;
  for (long long done = 0; done < n_units; done += max_units) {
    const long long n = (n_units - done < max_units) ? (n_units - done) : max_units;
    job.in = in0 + (done / unit_quantum) * in_step;
    job.de = de0 + done * out_per_unit;
    job.psd = psd0 + done * out_per_unit;
    job.total_rows = static_cast<unsigned>(n * job.n_ch);
    job.norm_row0 = done * job.n_ch;
    const int rc = launch<C>(job, aligned16, stream);
    if (rc != 0) return rc;
  }
  return 0;
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

// units = clips; clip (block, concept c, repetition r) starts at block * block_stride + c * 13 fs + 3 fs + r * 2 fs
static Job raw_geometry(int n_ch, int64_t block_stride, int64_t ch_stride, int fs = 200)
{
  Job j{};
  j.base = 3LL * fs;                         // 3 s hint before each concept (segment_raw_signals_200Hz.py:58-62)
  j.s0 = block_stride;
  j.s1 = 13 * fs;                            // concept stride: 3 s hint + 5 x 2 s
  j.s2 = 2 * fs;                             // repetition stride: 2 s
  j.d1 = 200;                                // 40 concepts x 5 repetitions per block
  j.d2 = 5;
  j.ch_stride = ch_stride;
  j.n_ch = static_cast<unsigned>(n_ch);
  return j;
}

}  // namespace eegfe

#include "eegfe_consumers.cuh"

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

static int dispatch_clip_mode(int mode, const Job& job, long long n_units, long long quantum, long long in_step,
                              bool aligned16, cudaStream_t stream)
{
  switch (mode) {
    case EEGFE_MODE_500MS: return run_units<CfgSliding500>(job, n_units, quantum, in_step, aligned16, stream);
    case EEGFE_MODE_1S: return run_units<CfgOneSec>(job, n_units, quantum, in_step, aligned16, stream);
    case EEGFE_MODE_2S: return run_units<CfgTwoSec>(job, n_units, quantum, in_step, aligned16, stream);
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
  Job job = raw_geometry(n_ch, block_stride, ch_stride);
  job.in = raw;
  job.de = de;
  job.psd = psd;
  job.status = status;
  job.t_extent = block_len;
  return dispatch_clip_mode(mode, job, n_blocks * 200, 200, block_stride,
                            is_aligned16(raw, {block_stride, ch_stride}, 4), static_cast<cudaStream_t>(stream));
}

int eegfe_de_psd_from_concepts(const float* x, int64_t n_blocks, int n_ch, int64_t block_stride, int64_t ch_stride,
                               int64_t concept_stride, int64_t first_offset, int mode, float* de, float* psd,
                               int* status, void* stream)
{
  if (n_blocks < 0 || n_ch <= 0 || eegfe_windows_per_clip(mode) < 0) return EEGFE_EINVAL;
  if (n_blocks == 0) return 0;
  if (x == nullptr || de == nullptr || psd == nullptr) return EEGFE_EINVAL;
  if (concept_stride < 2000 || first_offset < 0 || concept_stride > 0x7fffffff) return EEGFE_EINVAL;
  if (ch_stride < first_offset + 39 * concept_stride + 2000 || block_stride < 0) return EEGFE_ERANGE;
  Job job = raw_geometry(n_ch, block_stride, ch_stride);
  job.base = first_offset;
  job.s1 = static_cast<int>(concept_stride);
  job.t_extent = ch_stride;
  job.in = x;
  job.de = de;
  job.psd = psd;
  job.status = status;
  return dispatch_clip_mode(mode, job, n_blocks * 200, 200, block_stride,
                            is_aligned16(x, {block_stride, ch_stride, concept_stride, first_offset}, 4),
                            static_cast<cudaStream_t>(stream));
}

int eegfe_glmnet_inputs_from_raw(const float* raw, int64_t n_blocks, int n_ch, int64_t block_len, int64_t block_stride,
                                 int64_t ch_stride, const float* ch_scale, const float* ch_mean, float* clips_norm,
                                 float* de, float* psd, int* status, void* stream)
{
  if (n_blocks < 0 || n_ch <= 0) return EEGFE_EINVAL;
  if (n_blocks == 0) return 0;
  if (raw == nullptr || de == nullptr || psd == nullptr || clips_norm == nullptr || ch_scale == nullptr ||
      ch_mean == nullptr)
    return EEGFE_EINVAL;
  if (block_len < 40 * 2600) return EEGFE_ERANGE;
  if (ch_stride < block_len || block_stride < 0) return EEGFE_EINVAL;
  if (!is_aligned16(raw, {block_stride, ch_stride}, 4) || reinterpret_cast<uintptr_t>(clips_norm) % 16 != 0)
    return EEGFE_EINVAL;
  Job job = raw_geometry(n_ch, block_stride, ch_stride);
  job.in = raw;
  job.de = de;
  job.psd = psd;
  job.status = status;
  job.norm_out = clips_norm;
  job.norm_scale = ch_scale;
  job.norm_mean = ch_mean;
  return run_units<CfgSliding500>(job, n_blocks * 200, 200, block_stride, true, static_cast<cudaStream_t>(stream));
}

int eegfe_channel_stats(const float* raw, int64_t n_blocks, int n_ch, int64_t block_len, int64_t block_stride,
                        int64_t ch_stride, const unsigned char* block_mask, double* workspace, double* mean,
                        double* std, void* stream)
{
  if (n_blocks < 0 || n_ch <= 0) return EEGFE_EINVAL;
  if (raw == nullptr || workspace == nullptr || mean == nullptr || std == nullptr) return EEGFE_EINVAL;
  if (block_len < 40 * 2600) return EEGFE_ERANGE;
  if (ch_stride < block_len || block_stride < 0 || n_blocks * n_ch > 0x7fffffff) return EEGFE_EINVAL;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (n_blocks > 0) {
    channel_row_sums_kernel<<<static_cast<unsigned>(n_blocks * n_ch), 256, 0, s>>>(raw, block_stride, ch_stride, n_ch, workspace);
    ++g_launches;
  }
  channel_stats_finish_kernel<<<(n_ch + 63) / 64, 64, 0, s>>>(workspace, block_mask, n_blocks, n_ch, mean, std);
  ++g_launches;
  return static_cast<int>(cudaGetLastError());
}

int eegfe_de_from_psd(const float* psd, int64_t n, float* de, int* status, void* stream)
{
  if (n < 0) return EEGFE_EINVAL;
  if (n == 0) return 0;
  if (psd == nullptr || de == nullptr) return EEGFE_EINVAL;
  long long blocks = (n / 4 + 255) / 256;
  const long long cap = static_cast<long long>(sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  de_from_psd_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(psd, de, n, status);
  ++g_launches;
  return static_cast<int>(cudaGetLastError());
}

int eegfe_copy2d_async(void* dst, int64_t dst_pitch, const void* src, int64_t src_pitch, int64_t width,
                       int64_t height, int kind, void* stream)
{
  if (dst == nullptr || src == nullptr || width < 0 || height < 0 || (kind != 1 && kind != 2)) return EEGFE_EINVAL;
  if (width == 0 || height == 0) return 0;
  return static_cast<int>(cudaMemcpy2DAsync(dst, static_cast<size_t>(dst_pitch), src, static_cast<size_t>(src_pitch),
                                            static_cast<size_t>(width), static_cast<size_t>(height),
                                            kind == 1 ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost,
                                            static_cast<cudaStream_t>(stream)));
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
  job.s0 = static_cast<long long>(n_ch) * 400;
  job.d1 = 1;
  job.d2 = 1;
  job.ch_stride = 400;
  job.t_extent = 400;
  job.n_ch = static_cast<unsigned>(n_ch);
  return dispatch_clip_mode(mode, job, n_clips, 1, job.s0, is_aligned16(clips, {}, 4),
                            static_cast<cudaStream_t>(stream));
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
  job.s0 = row_stride;
  job.d1 = 1;
  job.d2 = 1;
  job.ch_stride = 0;
  job.n_ch = 1;
  const bool a16 = is_aligned16(x, {row_stride}, 4);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (win_len == 100) return run_units<CfgWin100>(job, n_rows, 1, row_stride, a16, s);
  if (win_len == 200) return run_units<CfgWin200>(job, n_rows, 1, row_stride, a16, s);
  return run_units<CfgTwoSec>(job, n_rows, 1, row_stride, a16, s);   // 400: only samples 0..199 count (DE_PSD.py:58)
}

int eegfe_de_psd_generic(const float* x, int64_t n_rows, int n_live, int64_t row_stride, const float* hann,
                         const int* band_lo, const int* band_hi, float* de, float* psd, int* status, void* stream)
{
  if (n_rows < 0 || n_live < 1 || n_live > 200 || row_stride < n_live || band_lo == nullptr || band_hi == nullptr)
    return EEGFE_EINVAL;
  BandBins bands;
  for (int b = 0; b < 5; ++b) {
    if (band_lo[b] < -1 || band_hi[b] > 99 || band_hi[b] < band_lo[b]) return EEGFE_EINVAL;
    bands.lo[b] = band_lo[b];
    bands.hi[b] = band_hi[b];
    bands.inv_count[b] = static_cast<float>(1.0 / static_cast<double>(band_hi[b] - band_lo[b] + 1));
  }
  if (n_rows == 0) return 0;
  if (x == nullptr || hann == nullptr || de == nullptr || psd == nullptr) return EEGFE_EINVAL;
  long long grid = static_cast<long long>(sm_count()) * 16;
  if (grid > n_rows) grid = n_rows;
  de_psd_generic_kernel<<<static_cast<unsigned>(grid), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      x, n_rows, n_live, row_stride, hann, bands, de, psd, status);
  ++g_launches;
  return static_cast<int>(cudaGetLastError());
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
  Job geom = raw_geometry(n_ch, block_stride, ch_stride, fs);
  const int row_bytes = 2 * fs * es;
  const bool a16 = is_aligned16(raw, {block_stride, ch_stride, 3LL * fs, 13LL * fs, 2LL * fs}, es) &&
                   reinterpret_cast<uintptr_t>(clips) % 16 == 0;
  const int threads = 256;
  const int grid = sm_count() * 8;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // rows are indexed with 32 bits inside the kernel: go block-group by block-group
  const long long rows_per_block = 200LL * n_ch;
  long long max_blocks = ((1LL << 31) - 1) / rows_per_block;
  if (max_blocks < 1) return EEGFE_EINVAL;
  for (long long b0 = 0; b0 < n_blocks; b0 += max_blocks) {
    const long long nb = (n_blocks - b0 < max_blocks) ? (n_blocks - b0) : max_blocks;
    const long long n_rows = nb * rows_per_block;
    const unsigned char* src = static_cast<const unsigned char*>(raw) + b0 * block_stride * es;
    unsigned char* dst = static_cast<unsigned char*>(clips) + b0 * rows_per_block * row_bytes;
    if (a16) gather_rows_kernel<16><<<grid, threads, 0, s>>>(src, dst, n_rows, row_bytes, es, geom);
    else if (es == 8) gather_rows_kernel<8><<<grid, threads, 0, s>>>(src, dst, n_rows, row_bytes, es, geom);
    else if (es == 4)
      gather_rows_warp_kernel<uint32_t><<<grid, threads, 0, s>>>(reinterpret_cast<const uint32_t*>(src),
                                                                reinterpret_cast<uint32_t*>(dst), n_rows, 2 * fs, geom);
    else gather_rows_kernel<2><<<grid, threads, 0, s>>>(src, dst, n_rows, row_bytes, es, geom);
    ++g_launches;
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  return 0;
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

int eegfe_sliding_windows_layout(const void* clips, int dtype, int64_t n_clips, int n_ch, int layout, void* windows,
                                 void* stream)
{
  if (layout == EEGFE_WINDOWS_WINDOW_MAJOR) return eegfe_sliding_windows(clips, dtype, n_clips, n_ch, windows, stream);
  if (layout != EEGFE_WINDOWS_LAST) return EEGFE_EINVAL;
  const int es = esize_of(dtype);
  if (es == 0) return EEGFE_EDTYPE;
  if (n_clips < 0 || n_ch <= 0) return EEGFE_EINVAL;
  if (n_clips == 0) return 0;
  if (clips == nullptr || windows == nullptr) return EEGFE_EINVAL;
  const long long n_rows = n_clips * n_ch;
  const int grid = sm_count() * 8;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (es == 8)
    sliding_windows_last_kernel<uint64_t><<<grid, 256, 0, s>>>(static_cast<const uint64_t*>(clips), static_cast<uint64_t*>(windows), n_rows);
  else if (es == 4)
    sliding_windows_last_kernel<uint32_t><<<grid, 256, 0, s>>>(static_cast<const uint32_t*>(clips), static_cast<uint32_t*>(windows), n_rows);
  else
    sliding_windows_last_kernel<uint16_t><<<grid, 256, 0, s>>>(static_cast<const uint16_t*>(clips), static_cast<uint16_t*>(windows), n_rows);
  ++g_launches;
  return static_cast<int>(cudaGetLastError());
}

// threads per block of the column-parallel consumer kernels: the columns of a row in ONE step where possible (310
// columns -> 320 threads: 97 % of the lanes busy; with 256 threads a second step runs at 21 %)
static int column_threads(long long n_cols)
{
  long long t = (n_cols + 31) / 32 * 32;
  if (t < 64) t = 64;
  if (t > 512) t = 512;
  return static_cast<int>(t);
}

int eegfe_select_units(const float* feat, int64_t n_units_in, int n_windows, int n_cols, const int* src_index,
                       int64_t n_out, int reduce_windows, float* out, void* stream)
{
  if (n_units_in < 0 || n_out < 0 || n_windows <= 0 || n_cols <= 0) return EEGFE_EINVAL;
  if (n_out == 0) return 0;
  if (feat == nullptr || src_index == nullptr || out == nullptr) return EEGFE_EINVAL;
  long long blocks = n_out;
  const long long cap = static_cast<long long>(sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  const int threads = column_threads(reduce_windows ? n_cols : static_cast<long long>(n_windows) * n_cols);
  select_units_kernel<<<static_cast<unsigned>(blocks), threads, 0, static_cast<cudaStream_t>(stream)>>>(
      feat, src_index, n_out, n_windows, n_cols, reduce_windows ? 1 : 0, out);
  ++g_launches;
  return static_cast<int>(cudaGetLastError());
}

int64_t eegfe_column_stats_workspace(int64_t n_groups, int64_t n_rows, int n_cols)
{
  if (n_groups < 0 || n_rows < 0 || n_cols <= 0) return EEGFE_EINVAL;
  const int64_t chunks = (n_rows + kStatRowsPerBlock - 1) / kStatRowsPerBlock;
  return 2 * (n_groups > 0 ? n_groups : 1) * (chunks > 0 ? chunks : 1) * n_cols;
}

int eegfe_column_stats(const float* x, int64_t n_groups, int64_t n_rows, int n_cols, int64_t row_stride,
                       int64_t group_stride, double* workspace, double* mean, double* var, double* scale, void* stream)
{
  if (n_groups < 0 || n_rows < 0 || n_cols <= 0 || row_stride < n_cols || group_stride < 0) return EEGFE_EINVAL;
  if (n_groups == 0) return 0;
  if (workspace == nullptr || mean == nullptr || var == nullptr || scale == nullptr) return EEGFE_EINVAL;
  if (n_rows > 0 && x == nullptr) return EEGFE_EINVAL;
  const int64_t chunks64 = (n_rows + kStatRowsPerBlock - 1) / kStatRowsPerBlock;
  if (chunks64 > 0x7fffffff || n_groups > 65535) return EEGFE_EINVAL;
  const int chunks = static_cast<int>(chunks64);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const dim3 part(static_cast<unsigned>(chunks), static_cast<unsigned>(n_groups));
  const dim3 fin(static_cast<unsigned>((n_cols + 127) / 128), static_cast<unsigned>(n_groups));
  if (chunks > 0) {
    column_partial_kernel<0><<<part, column_threads(n_cols), 0, s>>>(x, n_rows, n_cols, row_stride, group_stride, nullptr, workspace);
    ++g_launches;
  }
  column_finish_kernel<0><<<fin, 128, 0, s>>>(workspace, chunks, n_rows, n_cols, nullptr, mean, nullptr);
  ++g_launches;
  if (chunks > 0) {
    column_partial_kernel<1><<<part, column_threads(n_cols), 0, s>>>(x, n_rows, n_cols, row_stride, group_stride, mean, workspace);
    ++g_launches;
  }
  column_finish_kernel<1><<<fin, 128, 0, s>>>(workspace, chunks, n_rows, n_cols, mean, var, scale);
  ++g_launches;
  return static_cast<int>(cudaGetLastError());
}

int eegfe_standardize(const float* x, int64_t n_groups, int64_t n_rows, int n_cols, int64_t row_stride,
                      int64_t group_stride, const double* mean, const double* scale, float* out, void* stream)
{
  if (n_groups < 0 || n_rows < 0 || n_cols <= 0 || row_stride < n_cols || group_stride < 0) return EEGFE_EINVAL;
  if (n_rows == 0 || n_groups == 0) return 0;
  if (x == nullptr || mean == nullptr || scale == nullptr || out == nullptr || n_groups > 65535) return EEGFE_EINVAL;
  // a few thousand blocks in all (each walks several rows): 28 800 one-row blocks spend their time being scheduled
  long long blocks = n_rows;
  long long cap = static_cast<long long>(sm_count()) * 16 / n_groups;
  if (cap < 1) cap = 1;
  if (blocks > cap) blocks = cap;
  const dim3 grid(static_cast<unsigned>(blocks), static_cast<unsigned>(n_groups));
  standardize_kernel<<<grid, column_threads(n_cols), 0, static_cast<cudaStream_t>(stream)>>>(x, n_rows, n_cols, row_stride, group_stride,
                                                                          mean, scale, out);
  ++g_launches;
  return static_cast<int>(cudaGetLastError());
}

int eegfe_launch_geometry(int mode, int* grid, int* block, int* smem_bytes, int* rows_per_tile)
{
  int g = 0, b = 0, s = 0, r = 0;
  switch (mode) {
    case EEGFE_MODE_500MS: g = 1; b = StreamCfg::kThreads; s = StreamCfg::kSmemBytes; r = StreamCfg::kRows; break;
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

int64_t eegfe_launch_count(void) { return g_launches.load(); }

int64_t eegfe_tma_launch_count(void) { return g_tma_launches.load(); }

int eegfe_set_cta_limit(int max_ctas)
{
  return g_cta_limit.exchange(max_ctas > 0 ? max_ctas : 0);
}

int eegfe_set_tensor_loads(int on)
{
  return g_tensor_loads.exchange(on ? 1 : 0);
}

}  // extern "C"
