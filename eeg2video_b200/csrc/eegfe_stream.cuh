// Streaming form of the fused 500 ms kernel: 16 identical warps per SM, no specialised producer / storer warp.
//
// Why.  A warp-specialised ring (de_psd_kernel: worker groups + producer warps, what the 1 s / 2 s modes use) leaves
// the FP32 pipe of the sub-partitions that host the producer warps partly idle: warps land on the four SM
// sub-partitions round-robin, and with 16 warps x 128 registers filling the register file a producer warp displaces a
// worker.  In the first form of this kernel (2 CTAs x (7 workers + 1 producer)) that capped the pipe at 14/16
// (ncu: 74 %).  Here every warp is a worker (16 x 32 x 128 registers = the whole file) and the one serial duty --
// issuing the TMA copies of the next tile -- falls to whichever warp happens to finish a tile last.
//
// Structure.  A CTA owns tiles t = 0, 1, ... (global tile blockIdx.x + t * gridDim.x) of 16 rows; tile t lives in
// ring slot t % S.  A tile holds 16 rows x 7 windows = 112 channel-windows = 7 HALF-PASSES of 16 lanes (the lane
// map of eegfe_tables.h makes each half-warp's LDS.64 pattern conflict-free, the 7th 2-way).  The CTA's work is
// the stream of half-passes h = 0 .. 7 * n_tiles - 1; warps draw PASSES (two consecutive half-passes) from a
// shared counter, so a pass may straddle two tiles and no lane ever idles on a tile boundary.
//
//   per pass:   wait full[slot]                         (TMA bytes landed; mbarrier transaction count)
//               one channel-window per lane, register FFT (bandpower.cuh)
//               5 DE + 5 PSD per lane                   -> straight to HBM, [clip][window][channel][band]
//               consumed[slot] += half-passes           -> the warp that completes the tile re-arms full[slot] and
//                                                          issues the 16 bulk copies of tile t + S into the slot
//
// (EEGFE_DIRECT_STORE=0 builds the round-1 store path for comparison: results staged [window][row][band] per tile,
//  a second counter `staged[slot]`, the warp that completes it copies the tile to HBM and bumps `drained[slot]`, which
//  later stagers of the slot wait for.)
//
// No block-wide or group barrier exists after the prologue.  Progress: the oldest unfinished pass only ever waits
// for copies triggered by strictly older passes.
#pragma once

namespace eegfe {

#ifndef EEGFE_STREAM_WARPS
#define EEGFE_STREAM_WARPS 16
#endif
// 1: every lane stores its own 5 DE + 5 PSD values straight to HBM (no staging tile, no store duty, no `staged` /
//    `drained` handshake; the staging memory becomes an eighth input slot).  0: the round-1 form (results staged
//    [window][row][band] per tile, written out by the warp that finishes the tile last).
#ifndef EEGFE_DIRECT_STORE
#define EEGFE_DIRECT_STORE 1
#endif
template <int ROWS_, int WINDOWS_, int HOP_, int LOAD_, int STRIDE_, int VEC_, int SLOTS_, bool LANEMAP_,
          int WARPS_ = EEGFE_STREAM_WARPS>
struct StreamCfgT {
  static constexpr int kRows = ROWS_;          // rows per tile
  static constexpr int kWindows = WINDOWS_;    // analysis windows per row
  static constexpr int kHop = HOP_;
  static constexpr int kLoad = LOAD_;          // samples fetched per row
  static constexpr int kRowStride = STRIDE_;   // floats between rows in shared memory (bank skew, see Cfg)
  static constexpr int kVec = VEC_;            // floats per shared-memory load
  static constexpr bool kLaneMap = LANEMAP_;   // unit order inside a tile through c_lane_map_500_r16
  static constexpr int kRowBytes = kLoad * 4;
  static constexpr int kSlots = SLOTS_;
  static constexpr int kWarps = WARPS_;
  static constexpr int kThreads = kWarps * 32;
  static constexpr int kUnits = kRows * kWindows;          // channel-windows per tile
  static constexpr int kHalfPasses = kUnits / 16;
  static constexpr int kSlotFloats = kRows * kRowStride;
  static constexpr int kOutFloats = kUnits * 5;            // per staging array
  static constexpr int kSplit = 1;                         // store_tile: staged values are final (de, psd)
  static constexpr bool kDirectStore = EEGFE_DIRECT_STORE != 0;
  static constexpr int kSmemBytes = (kSlots * (kSlotFloats + (kDirectStore ? 0 : 2 * kOutFloats)) + kUnits) * 4;
  static_assert(kUnits % 16 == 0, "a tile is a whole number of half-passes");
  static_assert(kRowBytes % 16 == 0 && kRowStride % 4 == 0, "TMA bulk copies need 16-byte aligned rows");
  static_assert(kSmemBytes <= 227 * 1024, "shared memory per CTA");
};
// 500 ms windows sliding over 2 s clip rows (fused segmentation): 16 rows x 7 windows, LDS.64 + lane map
#ifndef EEGFE_STREAM_SLOTS
#define EEGFE_STREAM_SLOTS (EEGFE_DIRECT_STORE ? 8 : 7)
#endif
using StreamCfg = StreamCfgT<16, 7, 50, 400, 404, 2, EEGFE_STREAM_SLOTS, true>;
// pre-cut 500 ms windows (the reference's own call pattern, DE_PSD on a materialised (.., 100) array): 64 rows of 100
// samples, dense rows (LDS.128 over consecutive rows is conflict-free because 100 / 4 = 25 is odd); a tile of
// contiguous rows arrives with ONE bulk copy.
#ifndef EEGFE_WIN100_WARPS
#define EEGFE_WIN100_WARPS 12   // 168 registers, no spills: 11.9 G cw/s against 10.5 G with 16 warps at 128 registers
#endif
using StreamCfgWin100 = StreamCfgT<64, 1, 0, 100, 100, 4, EEGFE_STREAM_SLOTS, false, EEGFE_WIN100_WARPS>;

__constant__ unsigned char c_lane_map_500_r16[112] = {EEGFE_LANE_MAP_500_R16};

__device__ __forceinline__ unsigned atom_add_acq_rel_smem(unsigned* p, unsigned v)
{
  unsigned old;
  asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(smem_u32(p)), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ unsigned ld_acquire_smem(const unsigned* p)
{
  unsigned v;
  asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_smem(unsigned* p, unsigned v)
{
  asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

#ifndef EEGFE_STREAM_DUTY
#define EEGFE_STREAM_DUTY __forceinline__
#endif
// The serial duties (once per tile).  Inlined: as real calls they cost 4 % (uniform registers holding the twiddle
// constants cannot stay live across a call and are re-materialised every pass); -DEEGFE_STREAM_DUTY=__noinline__
// builds the out-of-line form for comparison.

// cp.async of BYTES (4 or 8) per lane (pre-cut windows whose rows are not 16-byte aligned; everything else that is
// misaligned arrives as a TMA copy of the aligned span around it, see stream_load_tile)
template <int BYTES>
__device__ __forceinline__ void cp_async_small(void* dst_smem, const void* src_gmem)
{
  if constexpr (BYTES == 8)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
  else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
// one arrival on `bar`, delivered when the executing thread's earlier cp.async have landed (.noinc: it counts against
// the barrier's init count -- 32 per phase in the cp.async mode, one per lane of the loading warp)
__device__ __forceinline__ void cp_async_arrive(uint64_t* bar)
{
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Dense rows (pre-cut windows) that are only 8- or 4-byte aligned (job.row_align) cannot ride on TMA bulk copies and
// have no row padding to shift into: the SMALL instantiation of the kernel fetches them with 8- / 4-byte cp.async (LDGSTS), one row per step, lanes side
// by side; every lane's arrival on the slot's barrier (32 per phase in this instantiation) fires when its copies have
// landed.  It is a separate kernel so that the TMA kernel's hot loop carries none of this code.
template <class SC>
__device__ __forceinline__ void stream_load_tile_small(const Job* jobp, float* slot, uint64_t* bar, unsigned* armed,
                                                       unsigned generation, unsigned row0, int nrows)
{
  const Job& job = *jobp;
  const int lane = threadIdx.x & 31;
  if (lane == 0) st_release_smem(armed, generation + 1u);
#pragma unroll 1
  for (int r = 0; r < nrows; ++r) {
    const float* src = job.in + row_offset(job, row0 + r, SC::kWindows, nullptr);
    float* dst = slot + r * SC::kRowStride;
    if (job.row_align == 8) {
#pragma unroll 1
      for (int i = 2 * lane; i < SC::kLoad; i += 64) cp_async_small<8>(dst + i, src + i);
    } else {
#pragma unroll 1
      for (int i = lane; i < SC::kLoad; i += 32) cp_async_small<4>(dst + i, src + i);
    }
  }
  cp_async_arrive(bar);                    // every lane: its arrival fires when its copies are in shared memory
  __syncwarp();
}

// whole warp: re-arm `bar` and issue the bulk copies of the tile starting at global row `row0`: one per row, or a
// single one when the rows sit back to back in HBM exactly as they do in the slot (pre-cut windows, dense array).
// SHIFT != 0 (rows that are only 4- / 8-byte aligned): one copy per row of the 16-byte aligned span AROUND the row --
// at most 16 bytes more than the row, the slot's row padding -- so that the row lands k = 0..3 floats into its
// shared-memory row; k is noted per row in `shift` (published by lane 0's arrival on the barrier).
template <class SC, bool SMALL, int SHIFT>
__device__ EEGFE_STREAM_DUTY void stream_load_tile(const Job* jobp, float* slot, uint64_t* bar, unsigned* armed,
                                                   unsigned generation, unsigned row0, int nrows, unsigned char* shift)
{
  if constexpr (SMALL) {
    stream_load_tile_small<SC>(jobp, slot, bar, armed, generation, row0, nrows);
    return;
  }
  const Job& job = *jobp;
  const int lane = threadIdx.x & 31;
  fence_proxy_async_smem();                  // generic-proxy reads of the slot before the async-proxy refill
  if constexpr (SHIFT != 0) {
    static_assert(SHIFT == 0 || (SC::kRows <= 32 && SC::kRowStride >= SC::kLoad + 4), "a shifted row needs 16 bytes of padding");
    const float* src = nullptr;
    unsigned bytes = 0;
    if (lane < nrows) {
      const float* p = job.in + row_offset_fast(job, row0 + lane);
      const unsigned k = static_cast<unsigned>(reinterpret_cast<uintptr_t>(p) >> 2) & 3u;
      shift[lane] = static_cast<unsigned char>(k);
      src = p - k;
      bytes = ((k + SC::kLoad) * 4u + 15u) & ~15u;
    }
    const unsigned total = __reduce_add_sync(0xffffffffu, bytes);
    __syncwarp();                            // shift[] is written before lane 0 arrives
    if (lane == 0) {
      mbar_arrive_expect_tx(bar, total);
      st_release_smem(armed, generation + 1u);
    }
    __syncwarp();
    if (lane < nrows) bulk_copy_g2s(slot + lane * SC::kRowStride, src, bytes, bar);
    __syncwarp();
    return;
  }
  if (lane == 0) {
    mbar_arrive_expect_tx(bar, nrows * SC::kRowBytes);
    st_release_smem(armed, generation + 1u);            // see `armed` in the kernel
  }
  __syncwarp();
  if (SC::kRowStride == SC::kLoad && job.n_ch == 1 && job.d1 == 1 && job.s0 == SC::kLoad) {
    if (lane == 0)
      bulk_copy_g2s(slot, job.in + job.base + static_cast<long long>(row0) * SC::kLoad, nrows * SC::kRowBytes, bar);
  } else {
    for (int r = lane; r < nrows; r += 32) {
      const long long off = row_offset_fast(job, row0 + r);
      bulk_copy_g2s(slot + r * SC::kRowStride, job.in + off, SC::kRowBytes, bar);
    }
  }
  __syncwarp();
}

// whole warp: staged tile -> HBM
template <class SC>
__device__ EEGFE_STREAM_DUTY void stream_store_tile(const Job* jobp, const float* out_a, unsigned row0, int nrows)
{
  store_tile<SC, 32>(*jobp, out_a, out_a + SC::kOutFloats, row0, nrows, threadIdx.x & 31);
  __syncwarp();
}

// GLMNet raw branch: half-pass q of a tile writes rows q, q + 7, q + 14 of the tile out again as per-channel normalised
// clips ((x - mean[ch]) * scale[ch]), 16 lanes x float4 per row -- spread over all passes instead of one warp per tile
// (as a last-reader duty it cost 40 %: 25.6 KB copied by a single warp per tile).
constexpr int kNormTableChannels = 256;     // per-channel scale / shift cached in shared memory up to this many channels
__device__ __forceinline__ void stream_store_norm_rows(const Job& job, const float* slot, unsigned row0, int nrows, int q,
                                                       int lane16, const float* norm_tab)
{
  for (int r = q; r < nrows; r += StreamCfg::kHalfPasses) {
    const unsigned grow = row0 + r;
    const unsigned ch = grow % job.n_ch;
    // (read from global memory, the two factors cost a long-scoreboard stall per row: 20 % of the kernel's stalls)
    const bool cached = job.n_ch <= kNormTableChannels;
    const float sc = cached ? norm_tab[ch] : __ldg(job.norm_scale + ch);
    const float mu = cached ? norm_tab[kNormTableChannels + ch] : __ldg(job.norm_mean + ch);
    const float4* src = reinterpret_cast<const float4*>(slot + r * StreamCfg::kRowStride);
    float4* dst = reinterpret_cast<float4*>(job.norm_out + (job.norm_row0 + grow) * 400);
    float4 v[7];
#pragma unroll
    for (int i = 0; i < 7; ++i)
      if (lane16 + 16 * i < 100) v[i] = src[lane16 + 16 * i];
#pragma unroll
    for (int i = 0; i < 7; ++i)
      if (lane16 + 16 * i < 100) {
        // (x - mean) * (1 / std): subtract first -- folded into one FMA (x / std - mean / std) a channel with a large
        // DC offset would lose the digits the subtraction cancels
        v[i].x = __fmul_rn(__fsub_rn(v[i].x, mu), sc);
        v[i].y = __fmul_rn(__fsub_rn(v[i].y, mu), sc);
        v[i].z = __fmul_rn(__fsub_rn(v[i].z, mu), sc);
        v[i].w = __fmul_rn(__fsub_rn(v[i].w, mu), sc);
        dst[lane16 + 16 * i] = v[i];
      }
  }
}

// SMALL: rows fetched with cp.async (pre-cut windows that are not 16-byte aligned: their rows have no padding to
//        shift into).
// SHIFT: rows fetched as the aligned span around them and read k floats in -- 2: every k is 0 or 2 (8-byte aligned
//        rows), windows still read with LDS.64; 1: any k, scalar LDS.32.
template <class SC, bool NORM, bool SMALL = false, int SHIFT = 0>
__global__ void __launch_bounds__(SC::kThreads, 1) de_psd_stream_kernel(const __grid_constant__ Job job)
{
  using C = SC;
  static_assert(!(SMALL && SHIFT != 0) && !(NORM && SHIFT != 0), "one loader per instantiation; normalised clips need aligned rows");
  constexpr int kVec = SHIFT == 1 ? 1 : C::kVec;
  __shared__ unsigned char row_shift[SHIFT != 0 ? C::kSlots : 1][SHIFT != 0 ? C::kRows : 1];
  static_assert(!NORM || (C::kLoad == 400 && C::kWindows == 7), "normalised clips ride on the sliding 500 ms form");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* const ring = reinterpret_cast<float*>(smem_raw);                     // [slot][row][kRowStride]
  float* const stage = ring + C::kSlots * C::kSlotFloats;                      // [slot][de | psd][window][row][band]
  int* const unit_meta =
      reinterpret_cast<int*>(stage + (C::kDirectStore ? 0 : C::kSlots * 2 * C::kOutFloats));         // [kUnits]
  __shared__ uint64_t full_bar[C::kSlots];
  __shared__ unsigned consumed[C::kSlots], staged[C::kSlots], drained[C::kSlots];
  // armed[s] = number of tiles whose copies have been ISSUED into slot s.  A parity wait on full[s] cannot tell
  // "generation k landed" from "generation k - 2 landed"; a warp that got far ahead of a straggler could reach
  // generation k of a slot before generation k - 1 has even been requested.  Waiting for armed[s] > k first
  // (the barrier is then in phase k or beyond) makes the parity wait unambiguous.
  __shared__ unsigned armed[C::kSlots];
  __shared__ unsigned next_pass;
  __shared__ float norm_tab[NORM ? 2 * kNormTableChannels : 2];
  if constexpr (NORM) {
    // every thread runs the same number of steps (surplus threads repeat the last channel): no warp diverges in front of
    // the block barrier below -- see the set-up comment
    if (job.n_ch <= kNormTableChannels)
      for (unsigned i0 = 0; i0 < job.n_ch; i0 += blockDim.x) {
        const unsigned i = i0 + threadIdx.x < job.n_ch ? i0 + threadIdx.x : job.n_ch - 1;
        norm_tab[i] = job.norm_scale[i];
        norm_tab[kNormTableChannels + i] = job.norm_mean[i];
      }
  }

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const unsigned n_tiles = (job.total_rows + C::kRows - 1) / C::kRows;
  const unsigned n_mine = blockIdx.x < n_tiles ? (n_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
  const unsigned n_half = n_mine * C::kHalfPasses;
  auto tile_row0 = [&](unsigned t) { return (blockIdx.x + t * gridDim.x) * C::kRows; };
  auto tile_nrows = [&](unsigned row0) {
    const unsigned left = job.total_rows - row0;
    return static_cast<int>(left < C::kRows ? left : C::kRows);
  };

  // ---- set-up, written so that NO warp diverges before the block barrier.
  // Seen twice on B200 with this kernel (an out-of-line loader call; scalar shared-memory reads): for `if (tid < 112)` /
  // `if (tid == 0)` blocks ptxas emits, in some instantiations, no BSSY / BSYNC pair around the branch -- the lane groups
  // of warp 3 / warp 0 then reach BAR.SYNC separately, are NOT reconverged by it, and run the code behind it (which
  // keeps per-warp state in uniform registers) as independent groups: the kernel hangs on every input.  A
  // `__syncwarp()` in front of the barrier does not help (ptxas drops it as redundant).  All eight SASS listings that
  // hung lack exactly that BSYNC, all that worked have it (tests/test_abi.py keeps watch).  So: whole
  // warps only -- the unit table is filled by four full warps (the surplus lanes repeat its last entry) and the
  // per-slot state by all of warp 0 (four lanes write the same values to each slot).
  constexpr int kMetaThreads = (C::kUnits + 31) / 32 * 32;
  static_assert(kMetaThreads <= C::kThreads && C::kSlots <= 32, "set-up by whole warps");
  if (tid < kMetaThreads) {                                    // warp-uniform
    const int u = tid < C::kUnits ? tid : C::kUnits - 1;
    int row = u, w = 0;
    if constexpr (C::kLaneMap) {
      static_assert(!C::kLaneMap || C::kUnits == 112, "lane map of 16-row x 7-window tiles");
      const int code = c_lane_map_500_r16[u];
      row = code >> 3;
      w = code & 7;
    }
    // bits 0..13: window offset in the slot (floats), 14..24: staging index, 25..30: row in tile
    unit_meta[u] = (row * C::kRowStride + w * C::kHop) | (((w * C::kRows + row) * 5) << 14) | (row << 25);
  }
  if (tid < 32) {                                              // warp-uniform; lane l sets up slot l % kSlots
    static_assert((C::kSlots & (C::kSlots - 1)) == 0, "every lane of warp 0 repeats the set-up of one slot");
    const int sl = tid % C::kSlots;                            // (the same values from four lanes: no lane is idle)
    mbar_init(&full_bar[sl], SMALL ? 32 : 1);
    consumed[sl] = 0;
    staged[sl] = 0;
    drained[sl] = 0;
    armed[sl] = 0;
    next_pass = 0;
    mbar_fence_init();
  }
  __syncthreads();
  {
    const unsigned w = tid >> 5;
    if (w < C::kSlots && w < n_mine) {
      const unsigned row0 = tile_row0(w);
      stream_load_tile<C, SMALL, SHIFT>(&job, ring + w * C::kSlotFloats, &full_bar[w], &armed[w], 0u, row0,
                                        tile_nrows(row0), row_shift[SHIFT != 0 ? w : 0]);
    }
  }

  for (;;) {
    unsigned pass = 0;
    if (lane == 0) pass = atomicAdd(&next_pass, 1u);
    pass = __shfl_sync(0xffffffffu, pass, 0);
    if (2 * pass >= n_half) break;
    // lane -> (tile, half-pass in tile, slot, generation, unit geometry); evaluated twice per pass (before the FFT
    // and again after it) so that nothing but `pass` and the ten results stays live across the FFT
    unsigned t, gen;
    int s, meta;
    bool valid;
    auto locate = [&]() {
      const unsigned h = 2 * pass + (lane >> 4);
      valid = h < n_half;                                // the very last pass of a CTA may be half empty
      const unsigned hh = valid ? h : n_half - 1;
      t = hh / C::kHalfPasses;
      s = static_cast<int>(t % C::kSlots);
      gen = t / C::kSlots;
      meta = *reinterpret_cast<volatile int*>(unit_meta + static_cast<int>(hh - t * C::kHalfPasses) * 16 + (lane & 15));
    };
    float de[5], psd[5];
    bool live;
    {
      locate();
      const unsigned t0 = __shfl_sync(0xffffffffu, t, 0), t1 = __shfl_sync(0xffffffffu, t, 16);
      while (ld_acquire_smem(&armed[t0 % C::kSlots]) <= t0 / C::kSlots) __nanosleep(20);
      mbar_wait(&full_bar[t0 % C::kSlots], (t0 / C::kSlots) & 1);
      if (t1 != t0) {
        while (ld_acquire_smem(&armed[t1 % C::kSlots]) <= t1 / C::kSlots) __nanosleep(20);
        mbar_wait(&full_bar[t1 % C::kSlots], (t1 / C::kSlots) & 1);
      }
      live = valid && (meta >> 25) < tile_nrows(tile_row0(t));
      if constexpr (NORM) {
        if (valid) {
          const unsigned r0 = tile_row0(t);
          stream_store_norm_rows(job, ring + s * C::kSlotFloats, r0, tile_nrows(r0),
                                 static_cast<int>((2 * pass + (lane >> 4)) - t * C::kHalfPasses), lane & 15, norm_tab);
        }
        __syncwarp();
        asm volatile("" : "+r"(pass));
        locate();
      }
      if (live) {
        float e[5];
        const float* win = ring + s * C::kSlotFloats + (meta & 0x3fff);
        if constexpr (SHIFT != 0) win += row_shift[s][meta >> 25];
        window_band_energy<4, kHannHalfSec, kVec>(win, e);
        if (band_features(e, psd, de) && job.status != nullptr) atomicOr(job.status, EEGFE_STATUS_ZERO_POWER);
      }
    }
    asm volatile("" : "+r"(pass));                        // everything below is recomputed from `pass` alone
    if constexpr (C::kDirectStore) {
      if (live) {
        // features [clip][window][channel][band]: this lane's ten values go straight out.  The 16 lanes of a half-pass
        // cover runs of 8 (lane-mapped tiles) or 16 consecutive channels of one window, i.e. 160 / 320 contiguous bytes
        // per array; the five 4-byte stores of a lane hit the same sectors back to back and merge in L2.
        // (Only the tile and the lane's (row, window) are re-derived here -- `live` implies a valid half-pass -- and
        // g / n_ch is a multiply-high with one correction step: every instruction costs an issue cycle)
        const unsigned h = 2 * pass + (lane >> 4);
        const unsigned tt = h / C::kHalfPasses;
        const int mm = *reinterpret_cast<volatile int*>(unit_meta + static_cast<int>(h - tt * C::kHalfPasses) * 16 + (lane & 15));
        const unsigned g = tile_row0(tt) + static_cast<unsigned>(mm >> 25);
        unsigned ch;
        const unsigned u = div_magic(g, job.n_ch, job.n_ch_magic, ch);
        const unsigned w = C::kWindows == 1 ? 0u : static_cast<unsigned>((mm >> 14) & 0x7ff) / (5u * C::kRows);
        const long long o = (static_cast<long long>(u * C::kWindows + w) * job.n_ch + ch) * 5;
#pragma unroll
        for (int b = 0; b < 5; ++b) {
          job.de[o + b] = de[b];
          job.psd[o + b] = psd[b];
        }
      }
    } else {
      locate();
      if (live) {
        // staging rows of the slot's previous tile must have been written out (true long before, in practice)
        while (ld_acquire_smem(&drained[s]) < gen) __nanosleep(32);
        float* const sd = stage + s * 2 * C::kOutFloats + ((meta >> 14) & 0x7ff);
#pragma unroll
        for (int b = 0; b < 5; ++b) {
          sd[b] = de[b];
          sd[C::kOutFloats + b] = psd[b];
        }
      }
    }
    __syncwarp();
    // ---- retire: per tile touched by this pass (one or two), count this warp's half-passes in; whoever completes
    //      the count refills the input slot (and, with staged results, whoever completes the staging count writes the
    //      tile out) ----
    const unsigned h0 = 2 * pass;
    const unsigned t0 = h0 / C::kHalfPasses;
    const unsigned t1 = (h0 + 1 < n_half) ? (h0 + 1) / C::kHalfPasses : t0;
    const unsigned halves0 = (t1 != t0) ? 1u : ((h0 + 1 < n_half) ? 2u : 1u);   // valid half-passes in tile t0
#pragma unroll 1
    for (unsigned tk = t0; tk <= t1; ++tk) {
      const unsigned cnt = tk == t0 ? halves0 : 1u;
      const int sk = static_cast<int>(tk % C::kSlots);
      const unsigned done = C::kHalfPasses * (tk / C::kSlots + 1);
      unsigned last = 0;
      if (lane == 0) {
        last = (atom_add_acq_rel_smem(&consumed[sk], cnt) + cnt == done) ? 1u : 0u;
        if constexpr (!C::kDirectStore) last |= (atom_add_acq_rel_smem(&staged[sk], cnt) + cnt == done) ? 2u : 0u;
      }
      last = __shfl_sync(0xffffffffu, last, 0);
      if (last & 1u) {
        if (tk + C::kSlots < n_mine) {
          const unsigned r0 = tile_row0(tk + C::kSlots);
          stream_load_tile<C, SMALL, SHIFT>(&job, ring + sk * C::kSlotFloats, &full_bar[sk], &armed[sk],
                                            tk / C::kSlots + 1, r0, tile_nrows(r0), row_shift[SHIFT != 0 ? sk : 0]);
        }
      }
      if constexpr (!C::kDirectStore) {
        if (last & 2u) {
          const unsigned r0 = tile_row0(tk);
          stream_store_tile<C>(&job, stage + sk * 2 * C::kOutFloats, r0, tile_nrows(r0));
          if (lane == 0) st_release_smem(&drained[sk], tk / C::kSlots + 1);
        }
      }
    }
  }
}

}  // namespace eegfe
