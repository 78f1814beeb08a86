// One analysis window -> five band energies.  This is the arithmetic core of the fused kernel.
//
// Reference semantics (EEG_preprocessing/DE_PSD.py:49-68): y = x * hann_L ; X = FFT_200(y) with y truncated or
// zero-padded to 200 samples ; E_b = sum_{k in band b} |X[k]|^2 over the inclusive bin ranges
// [0,3] [3,7] [7,13] [13,30] [30,98].  (psd_b = E_b / count_b and de_b = log2(100 psd_b) happen in the caller.)
//
// Algorithm.  200 = 8 x 25 with gcd(8, 25) = 1, so the 200-point DFT factors by the prime-factor (Good-Thomas)
// map  n = (25 n1 + 8 n2) mod 200,  k == k1 (mod 8), k == k2 (mod 25)  into a radix-8 stage over n1 and a
// 25-point stage over n2 with NO twiddle factors in between:
//
//     X[k1, k2] = sum_{n2} w25^{n2 k2} * B_{k1}[n2],        B_{k1}[n2] = sum_{n1} w8^{n1 k1} y[25 n1 + 8 n2]
//
// * The input is real, so X[200-k] = conj X[k]: only k1 = 0..4 are computed.  B_0 and B_4 are real sequences and
//   share ONE complex 25-point DFT (of B_0 + i B_4), separated afterwards; B_1, B_2, B_3 take one each.
//   Four complex DFT-25 per window in total; each is 5 x 5 Cooley-Tukey (10 radix-5 butterflies, 16 twiddles).
// * 500 ms windows have 100 samples: of the eight n1 of a group only four consecutive ones (mod 8) fall on real
//   samples, the rest is zero padding and is never touched (NI = 4).  1 s / 2 s windows use all eight (NI = 8).
// * The Hann weights are compile-time immediates; bins 99 and 100 are never formed.
//
// Cost per window (fp32 lane-operations): ~2.45 k for NI = 4, ~2.6 k for NI = 8, against ~7.6 k for a textbook
// 200-point complex FFT.  Everything is unrolled at compile time; all table look-ups
// below are constant expressions.
#pragma once
#include <type_traits>
#include "cplx.cuh"
#include "eegfe_tables.h"

#if defined(__CUDACC__)
#define EEGFE_OPAQUE_ZERO() (blockIdx.z)     // grids are 1-D, so this is 0 -- but not to the compiler
#else
#define EEGFE_OPAQUE_ZERO() (0)
#endif

namespace eegfe {

template <int I, int N, class F>
EEGFE_FN void static_for(F&& f)
{
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, N>(f);
  }
}

enum HannId { kHannHalfSec = 0, kHannOneSec = 1, kHannTwoSec = 2 };

template <int HANN, int N>
constexpr float hann_at()
{
  if constexpr (HANN == kHannHalfSec) return tab::kHann100[N];
  else if constexpr (HANN == kHannOneSec) return tab::kHann200[N];
  else return tab::kHann400Head[N];
}

// ---- compile-time index maps ---------------------------------------------------------------------------------
constexpr int n2_of_rho(int rho) { return (22 * rho) % 25; }                 // 8 * n2 == rho (mod 25)
constexpr int rot_of_rho(int rho) { return (8 - ((8 * n2_of_rho(rho)) / 25) % 8) % 8; }   // n1 of sample rho
constexpr int bin_of(int k1, int k2) { return (25 * k1 + 176 * k2) % 200; }  // CRT: k%8 = k1, k%25 = k2
constexpr int fold_bin(int k) { return k <= 100 ? k : 200 - k; }
constexpr int kBandLo[5] = {0, 3, 7, 13, 30};     // DE_PSD.py:35-39,:63 at fs = 200 (inclusive)
constexpr int kBandHi[5] = {3, 7, 13, 30, 98};
constexpr bool in_band(int b, int k) { return k >= kBandLo[b] && k <= kBandHi[b]; }
constexpr int dft25_slot(int k2) { return 5 * (k2 % 5) + k2 / 5; }           // where dft25() leaves output k2

// w8^d = exp(-2 pi i d / 8) as codes: 0 -> 0, +-1 -> +-1, +-2 -> +-sqrt(1/2)
constexpr int w8_re_code(int d) { constexpr int t[8] = {1, 2, 0, -2, -1, -2, 0, 2}; return t[d & 7]; }
constexpr int w8_im_code(int d) { constexpr int t[8] = {0, -2, -1, -2, 0, 2, 1, 2}; return t[d & 7]; }

// ---- window samples ------------------------------------------------------------------------------------------
// Samples are fetched as aligned vectors (identical addresses are merged by the compiler):
//   VEC = 2: float2 / LDS.64 -- sliding 500 ms windows start at multiples of 50 floats, i.e. only 8-byte aligned;
//   VEC = 4: float4 / LDS.128 -- 1 s / 2 s / pre-cut windows start 16-byte aligned;
//   VEC = 1: scalar loads -- rows that sit 4, 8 or 12 bytes into their shared-memory row (ring kernel, SHIFT form).
template <int N, int VEC>
EEGFE_FN float sample(const float* win)
{
  if constexpr (VEC == 1) {
    return win[N];
  } else if constexpr (VEC == 4) {
    const float4 p = *reinterpret_cast<const float4*>(win + (N & ~3));
    return (N & 3) == 0 ? p.x : (N & 3) == 1 ? p.y : (N & 3) == 2 ? p.z : p.w;
  } else {
    const float2 p = *reinterpret_cast<const float2*>(win + (N & ~1));
    return (N & 1) ? p.y : p.x;
  }
}

// ---- radix-5 butterfly (forward, w5 = exp(-2 pi i / 5)), in place ---------------------------------------------
EEGFE_FN void radix5(cf& x0, cf& x1, cf& x2, cf& x3, cf& x4)
{
  const cf t1 = c_add(x1, x4), t2 = c_add(x2, x3), t3 = c_sub(x1, x4), t4 = c_sub(x2, x3);
  const cf m1 = c_fma_s(t2, tab::kC2, c_fma_s(t1, tab::kC1, x0));
  const cf m2 = c_fma_s(t2, tab::kC1, c_fma_s(t1, tab::kC2, x0));
  const cf n1 = c_fma_s(t4, tab::kS2, c_mul_s(t3, tab::kS1));
  const cf n2 = c_fma_s(t4, -tab::kS1, c_mul_s(t3, tab::kS2));
  x0 = c_add(c_add(x0, t1), t2);
  x1 = c_add_mi(m1, n1);
  x4 = c_add_pi(m1, n1);
  x2 = c_add_mi(m2, n2);
  x3 = c_add_pi(m2, n2);
}

// ---- 25-point complex DFT, in place; output k2 ends up in v[dft25_slot(k2)] ------------------------------------
EEGFE_FN void dft25(cf (&v)[25])
{
  static_for<0, 5>([&](auto b_) {              // stage 1: over n_a for each n_b   (n2 = 5 n_a + n_b)
    constexpr int b = decltype(b_)::value;
    radix5(v[b], v[5 + b], v[10 + b], v[15 + b], v[20 + b]);      // -> T[b][k_a] at v[5 k_a + b]
    static_for<1, 5>([&](auto ka_) {
      constexpr int ka = decltype(ka_)::value;
      if constexpr (b > 0) {
        constexpr float wr = tab::kW25Re[b * ka], wi = tab::kW25Im[b * ka];
        v[5 * ka + b] = c_mul_w(v[5 * ka + b], wr, wi);
      }
    });
  });
  static_for<0, 5>([&](auto ka_) {             // stage 2: over n_b for each k_a -> V[k_a + 5 k_b] at v[5 k_a + k_b]
    constexpr int ka = decltype(ka_)::value;
    radix5(v[5 * ka], v[5 * ka + 1], v[5 * ka + 2], v[5 * ka + 3], v[5 * ka + 4]);
  });
}

// ---- band accumulation -----------------------------------------------------------------------------------------
// A band accumulator is a complex register holding (sum re^2, sum im^2): one packed FMA per bin and band.
template <int BIN>
EEGFE_FN void add_power_to_bands(cf (&acc)[5], cf v)
{
  static_for<0, 5>([&](auto b_) {
    constexpr int b = decltype(b_)::value;
    if constexpr (in_band(b, BIN)) acc[b] = c_fma_sq(v, acc[b]);
  });
}

// all 25 outputs of harmonic K1 in {1,2,3}: bins K (< 100) or their mirrors 200 - K
template <int K1>
EEGFE_FN void accumulate_complex(const cf (&v)[25], cf (&acc)[5])
{
  static_for<0, 25>([&](auto k2_) {
    constexpr int k2 = decltype(k2_)::value;
    constexpr int bin = fold_bin(bin_of(K1, k2));
    if constexpr (bin <= 98) add_power_to_bands<bin>(acc, v[dft25_slot(k2)]);
  });
}

// Z = DFT25(B_0 + i B_4):  2 A[k2] = Z[k2] + conj Z[-k2],  2i G[k2] = Z[k2] - conj Z[-k2].
// Powers are accumulated unscaled (the caller applies the exact factor 1/4).
EEGFE_FN void accumulate_real_pair(const cf (&v)[25], cf (&acc4)[5])
{
  static_for<0, 13>([&](auto k2_) {
    constexpr int k2 = decltype(k2_)::value;
    const cf p = v[dft25_slot(k2)], q = v[dft25_slot((25 - k2) % 25)];
    constexpr int bin_a = fold_bin(bin_of(0, k2));
    constexpr int bin_g = fold_bin(bin_of(4, k2));
    if constexpr (bin_a <= 98) add_power_to_bands<bin_a>(acc4, c_add_conj(p, q));
    if constexpr (bin_g <= 98) add_power_to_bands<bin_g>(acc4, c_sub_conj(p, q));
  });
}

// ---- radix-8 stage, 100-sample windows (four live inputs per group) ---------------------------------------------
// Group RHO holds samples RHO + 25 i, i = 0..3, sitting at n1 = (J + i) mod 8 with J = rot_of_rho(RHO).
// With y_i = x_i h_i (Hann weight h_i a compile-time immediate) every pair sum / difference is one multiply and
// one FMA:  y_p + y_q = fma(x_q, h_q, x_p h_p),  y_p - y_q = fma(x_q, -h_q, x_p h_p).
template <int HANN, int RHO, int VEC>
struct Group4 {
  static constexpr int J = rot_of_rho(RHO);
  static constexpr int N2 = n2_of_rho(RHO);
  // signs that make B_2 = (-i)^J (d02 - i d13) come out without negations: u = SU (y0 - y2), v = SV (y1 - y3)
  static constexpr int SU = (J % 4 == 0 || J % 4 == 3) ? 1 : -1;
  static constexpr int SV = (J % 4 == 2 || J % 4 == 3) ? 1 : -1;
  static constexpr float H0 = hann_at<HANN, RHO>(), H1 = hann_at<HANN, RHO + 25>();
  static constexpr float H2 = hann_at<HANN, RHO + 50>(), H3 = hann_at<HANN, RHO + 75>();
  float x0, x1, x2, x3;

  EEGFE_FN explicit Group4(const float* win)
      : x0(sample<RHO, VEC>(win)), x1(sample<RHO + 25, VEC>(win)), x2(sample<RHO + 50, VEC>(win)),
        x3(sample<RHO + 75, VEC>(win)) {}

  // (B_0, B_4) and B_2: the even sweep
  EEGFE_FN void even(cf& b04, cf& b2) const
  {
    constexpr float h0 = H0, h1 = H1, h2 = H2, h3 = H3;
    const float t0 = f_mul(x0, h0), t1 = f_mul(x1, h1);
    const float s02 = f_fma(x2, h2, t0), s13 = f_fma(x3, h3, t1);
    const float u = SU > 0 ? f_fma(x2, -h2, t0) : f_fma(x2, h2, -t0);
    const float v = SV > 0 ? f_fma(x3, -h3, t1) : f_fma(x3, h3, -t1);
    b04 = c_make(f_add(s02, s13), (J % 2 == 0) ? f_sub(s02, s13) : f_sub(s13, s02));
    b2 = (J % 2 == 0) ? c_make(u, v) : c_make(v, u);
  }

  // one component (IM = 0: real part, 1: imaginary part) of B_K1, K1 in {1, 3}:
  //   +-y_axis + (+-sqrt(1/2)) * (sum or signed difference of the diagonal pair)
  template <int K1, int IM>
  EEGFE_FN float odd_component(float sp, float ep, float ya0, float ya1) const
  {
    constexpr int code0 = IM ? w8_im_code((J + 0) * K1) : w8_re_code((J + 0) * K1);
    constexpr int code1 = IM ? w8_im_code((J + 1) * K1) : w8_re_code((J + 1) * K1);
    constexpr int code2 = IM ? w8_im_code((J + 2) * K1) : w8_re_code((J + 2) * K1);
    constexpr int code3 = IM ? w8_im_code((J + 3) * K1) : w8_re_code((J + 3) * K1);
    constexpr bool diag_is_02 = (J % 2 != 0);            // odd J: inputs 0, 2 sit on the diagonals
    constexpr int dq = diag_is_02 ? code0 : code1, dq2 = diag_is_02 ? code2 : code3;      // +-2 each
    constexpr int ax = diag_is_02 ? code1 : code0, ax2 = diag_is_02 ? code3 : code2;      // one is +-1, other 0
    static_assert((dq == 2 || dq == -2) && (dq2 == 2 || dq2 == -2), "diagonal pair");
    static_assert((ax == 0) != (ax2 == 0), "axis pair");
    const float ya = (ax != 0) ? ya0 : ya1;              // ya0 / ya1: first / second member of the axis pair
    constexpr int sa = (ax != 0) ? ax : ax2;
    constexpr int se = diag_is_02 ? SU : SV;             // sign carried by ep
    constexpr float kappa = (dq == dq2) ? (dq > 0 ? tab::kRh : -tab::kRh)
                                        : ((dq > 0) == (se > 0) ? tab::kRh : -tab::kRh);
    return f_fma(kappa, (dq == dq2) ? sp : ep, sa > 0 ? ya : -ya);
  }

  // B_1 and B_3: the odd sweep
  EEGFE_FN void odd(cf& b1, cf& b3) const
  {
    constexpr float h0 = H0, h1 = H1, h2 = H2, h3 = H3;
    float sp, ep, ya0, ya1;
    if constexpr (J % 2 != 0) {          // diagonal pair (0, 2), axis pair (1, 3)
      const float t = f_mul(x0, h0);
      sp = f_fma(x2, h2, t);
      ep = SU > 0 ? f_fma(x2, -h2, t) : f_fma(x2, h2, -t);
      ya0 = f_mul(x1, h1);
      ya1 = f_mul(x3, h3);
    } else {                             // diagonal pair (1, 3), axis pair (0, 2)
      const float t = f_mul(x1, h1);
      sp = f_fma(x3, h3, t);
      ep = SV > 0 ? f_fma(x3, -h3, t) : f_fma(x3, h3, -t);
      ya0 = f_mul(x0, h0);
      ya1 = f_mul(x2, h2);
    }
    b1 = c_make(odd_component<1, 0>(sp, ep, ya0, ya1), odd_component<1, 1>(sp, ep, ya0, ya1));
    b3 = c_make(odd_component<3, 0>(sp, ep, ya0, ya1), odd_component<3, 1>(sp, ep, ya0, ya1));
  }
};

// ---- radix-8 stage, 200-sample windows (all eight inputs) --------------------------------------------------------
// z_n = windowed sample sitting at n1 = n.  The even sweep only needs the sums z_n + z_{n+4}, the odd sweep only
// the differences z_n - z_{n+4}: one multiply and one FMA each.
template <int HANN, int RHO, int VEC>
struct Group8 {
  static constexpr int J = rot_of_rho(RHO);
  static constexpr int N2 = n2_of_rho(RHO);
  const float* win;
  EEGFE_FN explicit Group8(const float* w) : win(w) {}

  template <int N1>
  static constexpr int input_of() { return (N1 - J + 8) % 8; }           // which of the 8 samples sits at n1 = N1
  template <int N1, int SIGN>
  EEGFE_FN float pair() const                                             // z_{N1} + SIGN * z_{N1 + 4}
  {
    constexpr int ia = input_of<N1>(), ib = input_of<N1 + 4>();
    constexpr float ha = hann_at<HANN, RHO + 25 * ia>(), hb = hann_at<HANN, RHO + 25 * ib>();
    const float t = f_mul(sample<RHO + 25 * ia, VEC>(win), ha);
    return f_fma(sample<RHO + 25 * ib, VEC>(win), SIGN > 0 ? hb : -hb, t);
  }
  EEGFE_FN void even(cf& b04, cf& b2) const
  {
    const float sa = pair<0, 1>(), sb = pair<1, 1>(), sc = pair<2, 1>(), sd = pair<3, 1>();
    const float e = f_add(sa, sc), o = f_add(sb, sd);
    b04 = c_make(f_add(e, o), f_sub(e, o));
    b2 = c_make(f_sub(sa, sc), f_sub(sd, sb));
  }
  EEGFE_FN void odd(cf& b1, cf& b3) const
  {
    const float da = pair<0, -1>(), db = pair<1, -1>(), dc = pair<2, -1>(), dd = pair<3, -1>();
    const float p = f_sub(db, dd), q = f_add(db, dd);
    b1 = c_make(f_fma(tab::kRh, p, da), f_fma(-tab::kRh, q, -dc));
    b3 = c_make(f_fma(-tab::kRh, p, da), f_fma(-tab::kRh, q, dc));
  }
};

template <int NI, int HANN, int RHO, int VEC>
using Group = std::conditional_t<NI == 4, Group4<HANN, RHO, VEC>, Group8<HANN, RHO, VEC>>;

// ---- one window -> unnormalised band energies E_b = sum_{k in band b} |X[k]|^2 ------------------------------------
// NI = 4: `win` holds 100 samples (zero-padded transform);  NI = 8: `win` holds 200 samples.
// The work is two independent sweeps over the window, each forming two radix-8 output sequences and running two
// DFT-25 (100 registers of work set instead of 200):
//   sweep 0 (even): harmonics k1 = 0, 4 (one shared DFT) and k1 = 2   -> part_even[b]
//   sweep 1 (odd) : harmonics k1 = 1 and k1 = 3                        -> part_odd[b]
//   E_b = part_even[b] + part_odd[b]
// A thread may run both (500 ms kernel) or the two sweeps may run in different warps (1 s / 2 s kernels); the
// partial sums and their final addition are identical either way, so all paths agree bit for bit.
//
// CODE SIZE.  The sweep is selected at RUN time (a warp-uniform branch) so that both sweeps share the machine code
// of the two DFT-25: fully unrolled, the four DFTs are 13.6 KB of SASS out of a 30 KB loop body, and B200's L1.5
// instruction cache holds 32 KB -- above that every warp streams its instructions from L2 and the producer warp
// starves on instruction fetch (ncu: stall_no_instruction 23 %).  Sharing the DFT code brings the whole kernel
// to ~25 KB.  It also keeps the compiler from merging the two sweeps' loads and radix-8 partial sums (which cost
// ~100 extra live registers).
template <int NI, int HANN, int VEC>
EEGFE_FN void sweep_any(const float* win, int sweep, const float (&carry)[5], float (&part)[5])
{
  cf acc[5];
  static_for<0, 5>([&](auto b_) { acc[decltype(b_)::value] = c_make(carry[decltype(b_)::value], 0.f); });
  cf p[25], q[25];
  if (sweep == 0) {
    static_for<0, 25>([&](auto rho_) {
      constexpr int rho = decltype(rho_)::value;
      const Group<NI, HANN, rho, VEC> g(win);
      g.even(p[g.N2], q[g.N2]);                       // p = B_0 + i B_4, q = B_2
    });
  } else {
    static_for<0, 25>([&](auto rho_) {
      constexpr int rho = decltype(rho_)::value;
      const Group<NI, HANN, rho, VEC> g(win);
      g.odd(p[g.N2], q[g.N2]);                        // p = B_1, q = B_3
    });
  }
  dft25(p);
  if (sweep == 0) {
    accumulate_real_pair(p, acc);
    static_for<0, 5>([&](auto b_) { acc[decltype(b_)::value] = c_mul_s(acc[decltype(b_)::value], 0.25f); });
  } else {
    accumulate_complex<1>(p, acc);
  }
  dft25(q);
  if (sweep == 0) accumulate_complex<2>(q, acc);
  else accumulate_complex<3>(q, acc);
  static_for<0, 5>([&](auto b_) {
    constexpr int b = decltype(b_)::value;
    part[b] = f_add(c_re(acc[b]), c_im(acc[b]));
  });
}

// Both sweeps in one thread.
//  * NI = 8 (1 s / 2 s windows): E_b = part_even[b] + part_odd[b], the same additions the split kernels make, so the
//    one-thread and two-warp paths agree bit for bit.
//  * NI = 4 (500 ms windows; never split): the odd sweep's accumulators START from the even sweep's partial sums
//    (sweep 0 is always the even one), so nothing but five floats crosses the loop back-edge -- holding part_even
//    in registers through the second sweep cost five spills whose reloads sat exposed at the end of every window.
template <int NI, int HANN, int VEC>
EEGFE_FN void window_band_energy(const float* win, float (&energy)[5])
{
  float carry[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  float pe[5];
#if defined(__CUDACC__)
#pragma unroll 1
#endif
  for (int sweep = EEGFE_OPAQUE_ZERO(); sweep < 2; ++sweep) {
    float part[5];
    sweep_any<NI, HANN, VEC>(win, sweep, carry, part);
    static_for<0, 5>([&](auto b_) {
      constexpr int b = decltype(b_)::value;
      if constexpr (NI == 4) {
        carry[b] = part[b];
        energy[b] = part[b];
      } else {
        if (sweep == 0) pe[b] = part[b];
        else energy[b] = f_add(pe[b], part[b]);
      }
    });
  }
}

}  // namespace eegfe
