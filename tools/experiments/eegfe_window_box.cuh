// Window-box form of the fused kernels: ONE TMA tensor copy per (clip, window) tile, dense rows, LDS.128.
//
// Idea.  The sliding-window tensor of the reference, (.., window, channel, 100) (segment_sliding_window.py:6-21),
// is never written to HBM -- but it IS what shared memory holds here: a tile is the box [n_ch rows][LOAD samples]
// that `cp.async.bulk.tensor` (SASS UTMALDG) cuts out of the raw recording at the element coordinate
//     (c * 2600 + 600 + r * 400 + HOP * w,  channel 0,  block)          (segment_raw_signals_200Hz.py:58-65)
// Every window lands 128-byte aligned with dense 400-byte rows.  A tiled tensor copy must start on a 16-byte
// boundary of the innermost axis (measured: x = 2 floats raises "illegal instruction", tools/microbench/tma_align.cu),
// which the even windows do (x = 0, 100, 200, 300 floats into a clip); the odd ones (x = 50, 150, 250) sit 8 bytes
// off and are fetched by the refilling warp with 8-byte cp.async (LDGSTS.64) instead -- same mbarrier, same layout:
//   * LDS.128 with lanes over consecutive rows is conflict-free (row stride / 16 B = 25 is odd); no lane map;
//   * the 50 % overlap of neighbouring windows is served by L2 (the seven boxes of a clip are fetched back to back
//     by the same CTA), so DRAM still sees 1600 B per channel-clip while shared memory receives 2800 B;
//   * one elected thread issues ONE instruction per tile (the 1-D form needs a bulk copy and per-lane address
//     arithmetic per row);
//   * in the order (clip, window, channel) the units of a CTA are consecutive both in the smem ring and in the
//     feature tensors [clip][window][channel][band], so a warp's 32 results are one contiguous run of 160 floats
//     per array: each warp transposes its own results through 1.25 KB of private shared memory and stores them
//     itself -- no cross-warp staging, no store duty, no "drained" handshake.
//
// Structure.  A CTA owns a contiguous range of clips (tile t = local clip * W + window, W boxes per clip) and
// therefore a contiguous range of units u = t * R + row (R = rows per tile = n_ch).  Its warps draw PASSES of 32
// consecutive units from a shared counter; a pass touches 1 + 31 / R (+1) tiles.
//
//   per pass:   wait armed[slot] / full[slot] of every tile touched      (mbarrier transaction count)
//               one channel-window per lane, register FFT (bandpower.cuh), LDS.128
//               consumed[slot] += lanes of that tile    -> the warp that completes a tile resets the count, re-arms
//                                                          full[slot] and issues the copy of tile t + S
//               results -> own scratch -> 5 + 5 coalesced STG.32 per lane
//
// The same kernel serves pre-cut windows (rows of a dense (n_rows, LOAD) array, tile = 64 rows, W = 1).
#pragma once
#include <cuda.h>

namespace eegfe {

#ifndef EEGFE_TMA_WARPS
#define EEGFE_TMA_WARPS 16
#endif
#ifndef EEGFE_TMA_SLOTS
#define EEGFE_TMA_SLOTS 8
#endif

template <int LOAD_, int WIN_, int WINDOWS_, int HOP_, int NI_, int HANN_, int SLOTS_, int WARPS_, int MAXROWS_>
struct TmaCfgT {
  static constexpr int kLoad = LOAD_;          // samples per row of a box (= floats between rows in shared memory)
  static constexpr int kWin = WIN_;            // samples of a window that are read (kLoad - kWin = bank-skew padding)
  static constexpr int kWindows = WINDOWS_;    // boxes per clip
  static constexpr int kHop = HOP_;
  static constexpr int kNi = NI_;
  static constexpr int kHann = HANN_;
  static constexpr int kSlots = SLOTS_;
  static constexpr int kWarps = WARPS_;
  static constexpr int kThreads = WARPS_ * 32;
  static constexpr int kMaxRows = MAXROWS_;    // rows per tile the ring is dimensioned for
  static constexpr int kScratchFloats = 2 * 160;   // per warp: 32 x 5 DE + 32 x 5 PSD
  static_assert((SLOTS_ & (SLOTS_ - 1)) == 0, "slot = tile & (kSlots - 1)");
  static_assert((LOAD_ * 4) % 16 == 0 && (LOAD_ / 4) % 2 == 1, "dense rows: LDS.128 conflict-free iff LOAD / 4 is odd");
  static constexpr int slot_floats(int rows) { return (rows * LOAD_ + 31) & ~31; }     // 128-byte granules
  static constexpr int smem_bytes(int rows) { return (SLOTS_ * slot_floats(rows) + WARPS_ * kScratchFloats) * 4; }
  static_assert(smem_bytes(kMaxRows) <= 227 * 1024, "shared memory per CTA");
};
#ifndef EEGFE_TMA_WARPS_NI8
#define EEGFE_TMA_WARPS_NI8 12
#endif
//                               LOAD WIN NWIN HOP NI  HANN          SLOTS            WARPS                MAXROWS
using TmaCfg500 = TmaCfgT<100, 100, 7, 50, 4, kHannHalfSec, EEGFE_TMA_SLOTS, EEGFE_TMA_WARPS, 64>;
using TmaCfgWin100 = TmaCfgT<100, 100, 1, 0, 4, kHannHalfSec, EEGFE_TMA_SLOTS, EEGFE_TMA_WARPS, 64>;
// 200-sample windows: rows of 204 floats (204 / 4 = 51 is odd; the 4 surplus samples are fetched, never read)
using TmaCfgOneSec = TmaCfgT<204, 200, 2, 200, 8, kHannOneSec, 4, EEGFE_TMA_WARPS_NI8, 62>;
using TmaCfgTwoSec = TmaCfgT<204, 200, 1, 0, 8, kHannTwoSec, 4, EEGFE_TMA_WARPS_NI8, 62>;
using TmaCfgWin200 = TmaCfgT<204, 200, 1, 0, 8, kHannOneSec, 4, EEGFE_TMA_WARPS_NI8, 62>;

struct TmaJob {
  CUtensorMap map;          // (time, channel, block) over the recording; box = (kLoad, rows_per_tile, 1)
  float* de;
  float* psd;
  int* status;
  unsigned n_groups;        // clips of this launch (rows mode: tiles)
  unsigned rows_per_tile;   // R = n_ch (rows mode: 64)
  unsigned long long n_units;   // valid units of this launch = channel-windows (rows mode: rows; the last tile may be partial)
  // clip g -> first sample: block q = g / d1, concept c = (g % d1) / d2, repetition r = g % d2: base + c s1 + r s2
  int base, s1, s2;
  unsigned d1, d2;
  int rows_mode;            // 1: tile t is rows [t R, t R + R) of a dense (n_rows, kLoad) array
  const float* in;          // the recording again, for the windows TMA cannot fetch (8-byte aligned starts)
  long long block_stride, ch_stride;
};

__device__ __forceinline__ void tma_load_box(void* dst_smem, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar)
{
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          smem_u32(dst_smem)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}

__device__ __forceinline__ void cp_async8(void* dst_smem, const void* src_gmem)
{
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
// this thread's earlier cp.async land -> one arrival on `bar` (.noinc: counted against the barrier's init count)
__device__ __forceinline__ void cp_async_arrive(uint64_t* bar)
{
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <class TC>
__global__ void __launch_bounds__(TC::kThreads, 1) de_psd_tma_kernel(const __grid_constant__ TmaJob job)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* const ring = reinterpret_cast<float*>(smem_raw);
  const unsigned R = job.rows_per_tile;
  const unsigned slot_floats = (R * TC::kLoad + 31u) & ~31u;
  float* const scratch = ring + TC::kSlots * slot_floats + (threadIdx.x >> 5) * TC::kScratchFloats;
  __shared__ uint64_t full_bar[TC::kSlots];
  // armed[s]: tiles whose copy has been issued into slot s (makes the parity wait unambiguous, see eegfe_stream.cuh);
  // consumed[s]: units of the slot's current tile that have been read
  __shared__ unsigned armed[TC::kSlots], consumed[TC::kSlots];
  __shared__ unsigned next_pass;

  const int lane = threadIdx.x & 31;
  // this CTA's contiguous range of clips -> tiles, units, first output element
  const unsigned g_lo = static_cast<unsigned>(static_cast<unsigned long long>(job.n_groups) * blockIdx.x / gridDim.x);
  const unsigned g_hi = static_cast<unsigned>(static_cast<unsigned long long>(job.n_groups) * (blockIdx.x + 1) / gridDim.x);
  const unsigned n_tiles = (g_hi - g_lo) * TC::kWindows;
  const unsigned long long unit_lo = static_cast<unsigned long long>(g_lo) * TC::kWindows * R;
  const unsigned long long units_left = job.n_units > unit_lo ? job.n_units - unit_lo : 0ull;
  const unsigned long long units_full = static_cast<unsigned long long>(n_tiles) * R;
  const unsigned n_units = static_cast<unsigned>(units_left < units_full ? units_left : units_full);
  float* const out_de = job.de + unit_lo * 5;
  float* const out_psd = job.psd + unit_lo * 5;
  const unsigned box_bytes = R * TC::kLoad * 4;

  // whole warp: arm the slot's barrier (32 arrivals per phase) and fetch local tile t -- one tensor copy when the
  // window starts on a 16-byte boundary, else 8-byte cp.async (two per row: 32 + 18 lanes x 8 B = 400 B)
  auto load_tile = [&](unsigned t, unsigned generation) {
    const unsigned s = t & (TC::kSlots - 1);
    int x, y, z;
    if (job.rows_mode) {
      x = 0;
      y = static_cast<int>((g_lo + t) * R);
      z = 0;
    } else {
      const unsigned g = g_lo + t / TC::kWindows;
      const unsigned w = t - (t / TC::kWindows) * TC::kWindows;
      const unsigned q = g / job.d1;
      const unsigned rem = g - q * job.d1;
      const unsigned c = rem / job.d2;
      const unsigned r = rem - c * job.d2;
      x = job.base + static_cast<int>(c) * job.s1 + static_cast<int>(r) * job.s2 + static_cast<int>(w) * TC::kHop;
      y = 0;
      z = static_cast<int>(q);
    }
    float* const dst = ring + s * slot_floats;
    if (lane == 0) st_release_smem(&armed[s], generation + 1u);
    if ((x & 3) == 0) {
      if (lane == 0) {
        fence_proxy_async_smem();            // generic-proxy reads of the slot happen-before the async-proxy refill
        mbar_arrive_expect_tx(&full_bar[s], box_bytes);
        tma_load_box(dst, &job.map, x, y, z, &full_bar[s]);
      } else {
        mbar_arrive(&full_bar[s]);
      }
    } else {
      if constexpr (TC::kLoad == 100) {
        const float* src = job.in + static_cast<long long>(z) * job.block_stride + x + 2 * lane;
        float* d = dst + 2 * lane;
#pragma unroll 2
        for (unsigned row = 0; row < R; ++row) {
          cp_async8(d, src);
          if (lane < 18) cp_async8(d + 64, src + 64);
          src += job.ch_stride;
          d += TC::kLoad;
        }
      }
      cp_async_arrive(&full_bar[s]);
    }
  };

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < TC::kSlots; ++s) {
      mbar_init(&full_bar[s], 32);
      armed[s] = 0;
      consumed[s] = 0;
    }
    next_pass = 0;
    mbar_fence_init();
  }
  __syncthreads();
  {
    const unsigned w = threadIdx.x >> 5;
    if (w < TC::kSlots && w < n_tiles) load_tile(w, 0u);
  }

  for (;;) {
    unsigned pass = 0;
    if (lane == 0) pass = atomicAdd(&next_pass, 1u);
    pass = __shfl_sync(0xffffffffu, pass, 0);
    const unsigned u0 = pass * 32u;
    if (u0 >= n_units) break;
    const unsigned u = u0 + lane;
    const bool valid = u < n_units;
    const unsigned uu = valid ? u : n_units - 1;
    const unsigned t = uu / R;
    const unsigned row = uu - t * R;
    const unsigned t_first = __shfl_sync(0xffffffffu, t, 0);
    const unsigned t_last = __shfl_sync(0xffffffffu, t, 31);
    for (unsigned tk = t_first; tk <= t_last; ++tk) {
      const unsigned s = tk & (TC::kSlots - 1), gen = tk / TC::kSlots;
      while (ld_acquire_smem(&armed[s]) <= gen) __nanosleep(20);
      mbar_wait(&full_bar[s], gen & 1);
    }
    float de[5], psd[5];
    if (valid) {
      float e[5];
      window_band_energy<TC::kNi, TC::kHann, 4>(ring + (t & (TC::kSlots - 1)) * slot_floats + row * TC::kLoad, e);
      if (band_features(e, psd, de) && job.status != nullptr) atomicOr(job.status, EEGFE_STATUS_ZERO_POWER);
    }
    __syncwarp();
    // ---- retire: count this warp's lanes into every tile they came from; whoever completes a tile refills its slot ----
    for (unsigned tk = t_first; tk <= t_last; ++tk) {
      const unsigned cnt = __popc(__ballot_sync(0xffffffffu, valid && t == tk));
      unsigned last = 0;
      if (lane == 0) {
        const unsigned s = tk & (TC::kSlots - 1);
        const unsigned left = n_units - tk * R;
        const unsigned rows_here = left < R ? left : R;
        if (atom_add_acq_rel_smem(&consumed[s], cnt) + cnt == rows_here) {
          consumed[s] = 0;
          last = 1;
        }
      }
      last = __shfl_sync(0xffffffffu, last, 0);
      if (last && tk + TC::kSlots < n_tiles) load_tile(tk + TC::kSlots, tk / TC::kSlots + 1);
    }
    // ---- results: [lane][band] -> scratch -> consecutive lanes store consecutive floats (one run of 160 per array) ----
    if (valid) {
#pragma unroll
      for (int b = 0; b < 5; ++b) {
        scratch[lane * 5 + b] = de[b];
        scratch[160 + lane * 5 + b] = psd[b];
      }
    }
    __syncwarp();
    const unsigned n_out = (n_units - u0 < 32u ? n_units - u0 : 32u) * 5u;
    float* const g_de = out_de + static_cast<unsigned long long>(u0) * 5;
    float* const g_psd = out_psd + static_cast<unsigned long long>(u0) * 5;
    float va[5], vb[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      va[k] = scratch[lane + 32 * k];
      vb[k] = scratch[160 + lane + 32 * k];
    }
#pragma unroll
    for (int k = 0; k < 5; ++k)
      if (lane + 32u * k < n_out) {
        g_de[lane + 32 * k] = va[k];
        g_psd[lane + 32 * k] = vb[k];
      }
    __syncwarp();
  }
}

}  // namespace eegfe
