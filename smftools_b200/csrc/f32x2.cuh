// Packed fp32x2 arithmetic (sm_100: FFMA2 / FMUL2 / FADD2, PTX fma.rn.f32x2 ...).
//
// One warp instruction does two IEEE fp32 FMAs on a 64-bit register pair.  The FMA pipe still needs two passes for
// it, but it takes ONE issue slot -- and the recursions here are issue-bound, not FMA-lane-bound
// (profiles/micro/ffma2_bench.cu: 16 FFMA + 8 LOP3 per iteration 28.9 cycles, 8 FFMA2 + 8 LOP3 21.7).
// ptxas folds the operand shapes the state-vector products need into the instruction itself:
//   a scalar broadcast to both lanes   (R.F32),
//   the two lanes swapped              (R.F32x2.LO_HI),
// so `pack(s, s)` and `pack(hi, lo)` cost no extra instruction.  Each lane rounds exactly like fmaf / * / +, so a
// packed rewrite that keeps the per-lane operation order gives bit-identical results to the scalar code.
#pragma once

namespace smf {

typedef unsigned long long f2_t;

__device__ __forceinline__ f2_t f2_pack(float lo, float hi) {
  f2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(f2_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f2_t f2_fma(f2_t a, f2_t b, f2_t c) {
  f2_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f2_t f2_mul(f2_t a, f2_t b) {
  f2_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f2_t f2_add(f2_t a, f2_t b) {
  f2_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

// (d0, d1) = (a0 * b0 + c0, a1 * b1 + c1) on plain floats: the pairs are formed where they are used, ptxas keeps
// values that are always used together in adjacent registers
__device__ __forceinline__ void fma2f(float& d0, float& d1, float a0, float a1, float b0, float b1, float c0, float c1) {
  f2_unpack(f2_fma(f2_pack(a0, a1), f2_pack(b0, b1), f2_pack(c0, c1)), d0, d1);
}
__device__ __forceinline__ void mul2f(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  f2_unpack(f2_mul(f2_pack(a0, a1), f2_pack(b0, b1)), d0, d1);
}
__device__ __forceinline__ void add2f(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  f2_unpack(f2_add(f2_pack(a0, a1), f2_pack(b0, b1)), d0, d1);
}

}  // namespace smf
