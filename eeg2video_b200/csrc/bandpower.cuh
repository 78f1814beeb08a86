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
// Cost per window (scalar-equivalent fp32 operations): ~2.45 k for NI = 4, ~2.8 k for NI = 8, against
// ~7.6 k for a textbook 200-point complex FFT.  Everything is unrolled at compile time; all table look-ups
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
// The window start is 8-byte aligned (offsets are multiples of 50 floats); samples are fetched as float2 pairs
// (LDS.64 on the device; identical addresses are merged by the compiler).
template <int N>
EEGFE_FN float sample(const float* win)
{
  const float2 p = *reinterpret_cast<const float2*>(win + (N & ~1));
  return (N & 1) ? p.y : p.x;
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
template <int BIN>
EEGFE_FN void add_to_bands(float (&e)[5], float p)
{
  static_for<0, 5>([&](auto b_) {
    constexpr int b = decltype(b_)::value;
    if constexpr (in_band(b, BIN)) e[b] = f_add(e[b], p);
  });
}

// all 25 outputs of harmonic K1 in {1,2,3}: bins K (< 100) or their mirrors 200 - K
template <int K1>
EEGFE_FN void accumulate_complex(const cf (&v)[25], float (&e)[5])
{
  static_for<0, 25>([&](auto k2_) {
    constexpr int k2 = decltype(k2_)::value;
    constexpr int bin = fold_bin(bin_of(K1, k2));
    if constexpr (bin <= 98) add_to_bands<bin>(e, c_norm2(v[dft25_slot(k2)]));
  });
}

// Z = DFT25(B_0 + i B_4):  2 A[k2] = Z[k2] + conj Z[-k2],  2i G[k2] = Z[k2] - conj Z[-k2].
// Energies are accumulated unscaled into e4[] (the caller applies the exact factor 1/4).
EEGFE_FN void accumulate_real_pair(const cf (&v)[25], float (&e4)[5])
{
  static_for<0, 13>([&](auto k2_) {
    constexpr int k2 = decltype(k2_)::value;
    const cf p = v[dft25_slot(k2)], q = v[dft25_slot((25 - k2) % 25)];
    constexpr int bin_a = fold_bin(bin_of(0, k2));
    constexpr int bin_g = fold_bin(bin_of(4, k2));
    if constexpr (bin_a <= 98) add_to_bands<bin_a>(e4, c_norm2(c_add_conj(p, q)));
    if constexpr (bin_g <= 98) add_to_bands<bin_g>(e4, c_norm2(c_sub_conj(p, q)));
  });
}

// ---- radix-8 stage, 100-sample windows (four live inputs per group) ---------------------------------------------
// Group RHO holds samples RHO + 25 i, i = 0..3, sitting at n1 = (J + i) mod 8 with J = rot_of_rho(RHO).
template <int HANN, int RHO>
struct Group4 {
  static constexpr int J = rot_of_rho(RHO);
  static constexpr int N2 = n2_of_rho(RHO);
  // signs that make B_2 = (-i)^J (d02 - i d13) come out without negations: u = SU (y0 - y2), v = SV (y1 - y3)
  static constexpr int SU = (J % 4 == 0 || J % 4 == 3) ? 1 : -1;
  static constexpr int SV = (J % 4 == 2 || J % 4 == 3) ? 1 : -1;
  float y0, y1, y2, y3;

  EEGFE_FN explicit Group4(const float* win)
  {
    constexpr float h0 = hann_at<HANN, RHO>(), h1 = hann_at<HANN, RHO + 25>();
    constexpr float h2 = hann_at<HANN, RHO + 50>(), h3 = hann_at<HANN, RHO + 75>();
    y0 = f_mul(sample<RHO>(win), h0);
    y1 = f_mul(sample<RHO + 25>(win), h1);
    y2 = f_mul(sample<RHO + 50>(win), h2);
    y3 = f_mul(sample<RHO + 75>(win), h3);
  }
  EEGFE_FN float s02() const { return f_add(y0, y2); }
  EEGFE_FN float s13() const { return f_add(y1, y3); }
  EEGFE_FN float u() const { return SU > 0 ? f_sub(y0, y2) : f_sub(y2, y0); }
  EEGFE_FN float v() const { return SV > 0 ? f_sub(y1, y3) : f_sub(y3, y1); }

  // (B_0, B_4) packed as one complex number
  EEGFE_FN cf even_pair() const
  {
    const float a = s02(), b = s13();
    return c_make(f_add(a, b), (J % 2 == 0) ? f_sub(a, b) : f_sub(b, a));
  }
  EEGFE_FN cf harmonic2() const { return (J % 2 == 0) ? c_make(u(), v()) : c_make(v(), u()); }

  // one component (IM = 0: real part, 1: imaginary part) of B_K1, K1 in {1, 3}:
  //   +-y_axis + (+-sqrt(1/2)) * (sum or signed difference of the diagonal pair)
  template <int K1, int IM>
  EEGFE_FN float odd_component(float sp, float ep) const   // sp / ep: sum / signed difference of the diagonal pair
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
    const float ya = diag_is_02 ? (ax != 0 ? y1 : y3) : (ax != 0 ? y0 : y2);
    constexpr int sa = (ax != 0) ? ax : ax2;
    constexpr int se = diag_is_02 ? SU : SV;                                              // sign carried by ep
    constexpr float kappa = (dq == dq2) ? (dq > 0 ? tab::kRh : -tab::kRh)
                                        : ((dq > 0) == (se > 0) ? tab::kRh : -tab::kRh);
    return f_fma(kappa, (dq == dq2) ? sp : ep, sa > 0 ? ya : -ya);
  }
  template <int K1>
  EEGFE_FN cf harmonic_odd() const
  {
    const float sp = (J % 2 != 0) ? s02() : s13();
    const float ep = (J % 2 != 0) ? u() : v();
    return c_make(odd_component<K1, 0>(sp, ep), odd_component<K1, 1>(sp, ep));
  }
};

// ---- radix-8 stage, 200-sample windows (all eight inputs) --------------------------------------------------------
template <int HANN, int RHO>
struct Group8 {
  static constexpr int J = rot_of_rho(RHO);
  static constexpr int N2 = n2_of_rho(RHO);
  float sa, da, sb, db, sc, dc, sd, dd;       // sums / differences of (z_n, z_{n+4}), z indexed by n1

  template <int N1>
  EEGFE_FN static float z(const float* win)   // windowed sample sitting at n1 = N1
  {
    constexpr int i = (N1 - J + 8) % 8;
    constexpr float h = hann_at<HANN, RHO + 25 * i>();
    return f_mul(sample<RHO + 25 * i>(win), h);
  }
  EEGFE_FN explicit Group8(const float* win)
  {
    const float z0 = z<0>(win), z1 = z<1>(win), z2 = z<2>(win), z3 = z<3>(win);
    const float z4 = z<4>(win), z5 = z<5>(win), z6 = z<6>(win), z7 = z<7>(win);
    sa = f_add(z0, z4); da = f_sub(z0, z4);
    sb = f_add(z1, z5); db = f_sub(z1, z5);
    sc = f_add(z2, z6); dc = f_sub(z2, z6);
    sd = f_add(z3, z7); dd = f_sub(z3, z7);
  }
  EEGFE_FN cf even_pair() const
  {
    const float e = f_add(sa, sc), o = f_add(sb, sd);
    return c_make(f_add(e, o), f_sub(e, o));
  }
  EEGFE_FN cf harmonic2() const { return c_make(f_sub(sa, sc), f_sub(sd, sb)); }
  template <int K1>
  EEGFE_FN cf harmonic_odd() const
  {
    const float p = f_sub(db, dd), q = f_add(db, dd);
    if constexpr (K1 == 1) return c_make(f_fma(tab::kRh, p, da), f_fma(-tab::kRh, q, -dc));
    else return c_make(f_fma(-tab::kRh, p, da), f_fma(-tab::kRh, q, dc));
  }
};

template <int NI, int HANN, int RHO>
using Group = std::conditional_t<NI == 4, Group4<HANN, RHO>, Group8<HANN, RHO>>;

// ---- one window -> unnormalised band energies E_b = sum_{k in band b} |X[k]|^2 ------------------------------------
// NI = 4: `win` holds 100 samples (zero-padded transform);  NI = 8: `win` holds 200 samples.
// Two sweeps over the window keep two DFT-25 work sets (100 registers) live at a time instead of four.
template <int NI, int HANN>
EEGFE_FN void window_band_energy(const float* win, float (&energy)[5])
{
  float e[5] = {0.f, 0.f, 0.f, 0.f, 0.f}, e4[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  {
    cf b04[25], b2[25];
    static_for<0, 25>([&](auto rho_) {
      constexpr int rho = decltype(rho_)::value;
      const Group<NI, HANN, rho> g(win);
      b04[g.N2] = g.even_pair();
      b2[g.N2] = g.harmonic2();
    });
    dft25(b04);
    accumulate_real_pair(b04, e4);
    dft25(b2);
    accumulate_complex<2>(b2, e);
  }
  // Without this the compiler (nvcc AND ptxas) merges the sample loads and the radix-8 partial sums of the two
  // sweeps and keeps ~100 extra values live across the first pair of DFTs -- exactly what the two sweeps are
  // there to avoid.  The second sweep therefore reads through a pointer offset by a run-time zero.
  asm volatile("" ::: "memory");
  win += EEGFE_OPAQUE_ZERO();
  {
    cf b1[25], b3[25];
    static_for<0, 25>([&](auto rho_) {
      constexpr int rho = decltype(rho_)::value;
      const Group<NI, HANN, rho> g(win);
      b1[g.N2] = g.template harmonic_odd<1>();
      b3[g.N2] = g.template harmonic_odd<3>();
    });
    dft25(b1);
    accumulate_complex<1>(b1, e);
    dft25(b3);
    accumulate_complex<3>(b3, e);
  }
  static_for<0, 5>([&](auto b_) {
    constexpr int b = decltype(b_)::value;
    energy[b] = f_fma(0.25f, e4[b], e[b]);
  });
}

}  // namespace eegfe
