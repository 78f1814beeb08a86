// Complex-number primitives for the DE/PSD kernel, with two interchangeable backends:
//
//  * packed  (device, EEGFE_PACKED=1): a complex value lives in one 64-bit register pair and every
//    operation is ONE Blackwell f32x2 instruction (FADD2 / FMUL2 / FFMA2).  ptxas folds the half swaps,
//    scalar broadcasts and per-half sign flips used below into operand modifiers (.LO_HI, .F32, .NP), so a
//    complex add costs 1 issue slot, a multiply by a constant twiddle 2, a radix-5 butterfly 18.
//    Measured on B200 (tools/microbench/pipes.cu): FFMA2 issues at 0.5/clk/SMSP = the same FP32 lane rate
//    as scalar FFMA with half the instructions and register operands; it holds the issue port for both cycles
//    (FFMA2 + IADD3 pairs issue every 3.1 cycles), so a lane-operation costs one issue cycle either way.
//  * scalar  (device with EEGFE_PACKED=0, and the host build used ONLY by tests/hostemu): the same operations,
//    in the same order, with explicitly rounded fp32 add / mul / fma, so both backends -- and the CPU
//    emulation -- produce bit-identical band energies.
#pragma once

#if defined(__CUDACC__)
#define EEGFE_FN __device__ __forceinline__
#else
#include <cmath>
#define EEGFE_FN inline
#endif

#ifndef EEGFE_PACKED
#define EEGFE_PACKED 1
#endif

namespace eegfe {

// ---- explicitly rounded scalar fp32 (never contracted, never re-associated) -------------------------------
#if defined(__CUDACC__)
EEGFE_FN float f_add(float a, float b) { return __fadd_rn(a, b); }
EEGFE_FN float f_sub(float a, float b) { return __fsub_rn(a, b); }
EEGFE_FN float f_mul(float a, float b) { return __fmul_rn(a, b); }
EEGFE_FN float f_fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
#else
EEGFE_FN float f_add(float a, float b) { return a + b; }     // host build uses -ffp-contract=off
EEGFE_FN float f_sub(float a, float b) { return a - b; }
EEGFE_FN float f_mul(float a, float b) { return a * b; }
EEGFE_FN float f_fma(float a, float b, float c) { return std::fmaf(a, b, c); }
#endif

#if defined(__CUDACC__) && EEGFE_PACKED
// ---- packed backend: (re, im) = (lo, hi) halves of one .b64 register ---------------------------------------
struct cf { unsigned long long v; };
typedef unsigned long long u64_;
EEGFE_FN u64_ pk_(float lo, float hi) { u64_ r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
EEGFE_FN u64_ swp_(u64_ a) { u64_ r; asm("{\n\t.reg .b32 lo, hi;\n\tmov.b64 {lo, hi}, %1;\n\tmov.b64 %0, {hi, lo};\n\t}" : "=l"(r) : "l"(a)); return r; }
EEGFE_FN u64_ fma2_(u64_ a, u64_ b, u64_ c) { u64_ r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
EEGFE_FN u64_ add2_(u64_ a, u64_ b) { u64_ r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
EEGFE_FN u64_ sub2_(u64_ a, u64_ b) { u64_ r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
EEGFE_FN u64_ mul2_(u64_ a, u64_ b) { u64_ r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

EEGFE_FN cf c_make(float re, float im) { return cf{pk_(re, im)}; }
EEGFE_FN float c_re(cf a) { return __uint_as_float(static_cast<unsigned>(a.v)); }
EEGFE_FN float c_im(cf a) { return __uint_as_float(static_cast<unsigned>(a.v >> 32)); }
EEGFE_FN cf c_add(cf a, cf b) { return cf{add2_(a.v, b.v)}; }
EEGFE_FN cf c_sub(cf a, cf b) { return cf{sub2_(a.v, b.v)}; }
EEGFE_FN cf c_mul_s(cf a, float s) { return cf{mul2_(a.v, pk_(s, s))}; }                     // a * s
EEGFE_FN cf c_fma_s(cf a, float s, cf c) { return cf{fma2_(a.v, pk_(s, s), c.v)}; }          // a * s + c
EEGFE_FN cf c_add_mi(cf m, cf n) { return cf{fma2_(swp_(n.v), pk_(1.0f, -1.0f), m.v)}; }    // m - i n
EEGFE_FN cf c_add_pi(cf m, cf n) { return cf{fma2_(swp_(n.v), pk_(-1.0f, 1.0f), m.v)}; }    // m + i n
EEGFE_FN cf c_add_conj(cf p, cf q) { return cf{fma2_(q.v, pk_(1.0f, -1.0f), p.v)}; }        // p + conj(q)
EEGFE_FN cf c_sub_conj(cf p, cf q) { return cf{fma2_(q.v, pk_(-1.0f, 1.0f), p.v)}; }        // p - conj(q)
EEGFE_FN cf c_mul_w(cf a, float wr, float wi)                                                 // a * (wr + i wi)
{
  u64_ t = mul2_(a.v, pk_(wr, wr));
  return cf{fma2_(swp_(a.v), pk_(-wi, wi), t)};
}
EEGFE_FN cf c_fma_sq(cf a, cf c) { return cf{fma2_(a.v, a.v, c.v)}; }                         // (re^2, im^2) + c
#else
// ---- scalar backend ------------------------------------------------------------------------------------------
struct cf { float re, im; };
EEGFE_FN cf c_make(float re, float im) { return cf{re, im}; }
EEGFE_FN float c_re(cf a) { return a.re; }
EEGFE_FN float c_im(cf a) { return a.im; }
EEGFE_FN cf c_add(cf a, cf b) { return cf{f_add(a.re, b.re), f_add(a.im, b.im)}; }
EEGFE_FN cf c_sub(cf a, cf b) { return cf{f_sub(a.re, b.re), f_sub(a.im, b.im)}; }
EEGFE_FN cf c_mul_s(cf a, float s) { return cf{f_mul(a.re, s), f_mul(a.im, s)}; }
EEGFE_FN cf c_fma_s(cf a, float s, cf c) { return cf{f_fma(a.re, s, c.re), f_fma(a.im, s, c.im)}; }
EEGFE_FN cf c_add_mi(cf m, cf n) { return cf{f_fma(n.im, 1.0f, m.re), f_fma(n.re, -1.0f, m.im)}; }
EEGFE_FN cf c_add_pi(cf m, cf n) { return cf{f_fma(n.im, -1.0f, m.re), f_fma(n.re, 1.0f, m.im)}; }
EEGFE_FN cf c_add_conj(cf p, cf q) { return cf{f_fma(q.re, 1.0f, p.re), f_fma(q.im, -1.0f, p.im)}; }
EEGFE_FN cf c_sub_conj(cf p, cf q) { return cf{f_fma(q.re, -1.0f, p.re), f_fma(q.im, 1.0f, p.im)}; }
EEGFE_FN cf c_mul_w(cf a, float wr, float wi)
{
  float tr = f_mul(a.re, wr), ti = f_mul(a.im, wr);
  return cf{f_fma(a.im, -wi, tr), f_fma(a.re, wi, ti)};
}
EEGFE_FN cf c_fma_sq(cf a, cf c) { return cf{f_fma(a.re, a.re, c.re), f_fma(a.im, a.im, c.im)}; }
#endif

}  // namespace eegfe
