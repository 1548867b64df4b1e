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

// ---------------------------------------------------------------------------------------------------------
// State-vector helpers of the recursions (K = 2, 3, 4 packed, anything else -- or PK = false -- scalar).  Pairing: states (0,1) and
// (2,3) for K = 2, 4; states (1,2) + scalar state 0 for K = 3 (the emission ratios c_1, c_2 are a natural pair,
// c_0 == 1).  Every lane keeps the operation order of the scalar loops next to it: bit-identical results.
// ---------------------------------------------------------------------------------------------------------
// y_j = x_j + sum_{i != j} x_i A_ij  (i ascending; A has a unit diagonal)
template <int K, bool PK = true>
__device__ __forceinline__ void vec_mat_unit(const float (&x)[K], const float (&A)[K][K], float (&y)[K]) {
  if constexpr (PK && K == 2) {
    fma2f(y[0], y[1], x[1], x[0], A[1][0], A[0][1], x[0], x[1]);
  } else if constexpr (PK && K == 3) {
    float y0 = fmaf(x[1], A[1][0], x[0]);
    y0 = fmaf(x[2], A[2][0], y0);
    float y1, y2;
    fma2f(y1, y2, x[0], x[0], A[0][1], A[0][2], x[1], x[2]);
    fma2f(y1, y2, x[2], x[1], A[2][1], A[1][2], y1, y2);
    y[0] = y0;
    y[1] = y1;
    y[2] = y2;
  } else if constexpr (PK && K == 4) {
    float y0, y1, y2, y3;
    fma2f(y0, y1, x[1], x[0], A[1][0], A[0][1], x[0], x[1]);
    fma2f(y0, y1, x[2], x[2], A[2][0], A[2][1], y0, y1);
    fma2f(y0, y1, x[3], x[3], A[3][0], A[3][1], y0, y1);
    fma2f(y2, y3, x[0], x[0], A[0][2], A[0][3], x[2], x[3]);
    fma2f(y2, y3, x[1], x[1], A[1][2], A[1][3], y2, y3);
    fma2f(y2, y3, x[3], x[2], A[3][2], A[2][3], y2, y3);
    y[0] = y0;
    y[1] = y1;
    y[2] = y2;
    y[3] = y3;
  } else {
#pragma unroll
    for (int j = 0; j < K; ++j) {
      float s = x[j];
#pragma unroll
      for (int i = 0; i < K; ++i)
        if (i != j) s = fmaf(x[i], A[i][j], s);
      y[j] = s;
    }
  }
}
// y_i = x_i + sum_{j != i} A_ij x_j  (j ascending)
template <int K, bool PK = true>
__device__ __forceinline__ void mat_vec_unit(const float (&A)[K][K], const float (&x)[K], float (&y)[K]) {
  if constexpr (PK && K == 2) {
    fma2f(y[0], y[1], x[1], x[0], A[0][1], A[1][0], x[0], x[1]);
  } else if constexpr (PK && K == 3) {
    float y0 = fmaf(A[0][1], x[1], x[0]);
    y0 = fmaf(A[0][2], x[2], y0);
    float y1, y2;
    fma2f(y1, y2, x[0], x[0], A[1][0], A[2][0], x[1], x[2]);
    fma2f(y1, y2, x[2], x[1], A[1][2], A[2][1], y1, y2);
    y[0] = y0;
    y[1] = y1;
    y[2] = y2;
  } else if constexpr (PK && K == 4) {
    float y0, y1, y2, y3;
    fma2f(y0, y1, x[1], x[0], A[0][1], A[1][0], x[0], x[1]);
    fma2f(y0, y1, x[2], x[2], A[0][2], A[1][2], y0, y1);
    fma2f(y0, y1, x[3], x[3], A[0][3], A[1][3], y0, y1);
    fma2f(y2, y3, x[0], x[0], A[2][0], A[3][0], x[2], x[3]);
    fma2f(y2, y3, x[1], x[1], A[2][1], A[3][1], y2, y3);
    fma2f(y2, y3, x[3], x[2], A[2][3], A[3][2], y2, y3);
    y[0] = y0;
    y[1] = y1;
    y[2] = y2;
    y[3] = y3;
  } else {
#pragma unroll
    for (int i = 0; i < K; ++i) {
      float s = x[i];
#pragma unroll
      for (int j = 0; j < K; ++j)
        if (j != i) s = fmaf(A[i][j], x[j], s);
      y[i] = s;
    }
  }
}
// first state of the pairs: K = 3 starts at state 1
template <int K>
struct VecPairs {
  static constexpr bool packed = (K == 2 || K == 3 || K == 4);
  static constexpr int first = (K == 3) ? 1 : 0;
};
// d = a .* b
template <int K, bool PK = true>
__device__ __forceinline__ void vec_mul(float (&d)[K], const float (&a)[K], const float (&b)[K]) {
  if constexpr (PK && VecPairs<K>::packed) {
    if (K == 3) d[0] = a[0] * b[0];
#pragma unroll
    for (int k = VecPairs<K>::first; k + 1 < K; k += 2) mul2f(d[k], d[k + 1], a[k], a[k + 1], b[k], b[k + 1]);
  } else {
#pragma unroll
    for (int k = 0; k < K; ++k) d[k] = a[k] * b[k];
  }
}
// d_0 = a_0, d_k = a_k c_k (k >= 1): the emission-ratio product (c_0 == 1 is never read)
template <int K, bool PK = true>
__device__ __forceinline__ void vec_mul_tail(float (&d)[K], const float (&a)[K], const float (&c)[K]) {
  d[0] = a[0];
  if constexpr (PK && K == 3) {
    mul2f(d[1], d[2], a[1], a[2], c[1], c[2]);
  } else if constexpr (PK && K == 4) {
    d[1] = a[1] * c[1];
    mul2f(d[2], d[3], a[2], a[3], c[2], c[3]);
  } else {
#pragma unroll
    for (int k = 1; k < K; ++k) d[k] = a[k] * c[k];
  }
}
// v <- v * s
template <int K, bool PK = true>
__device__ __forceinline__ void vec_scale(float (&v)[K], float s) {
  if constexpr (PK && VecPairs<K>::packed) {
    if (K == 3) v[0] *= s;
#pragma unroll
    for (int k = VecPairs<K>::first; k + 1 < K; k += 2) mul2f(v[k], v[k + 1], v[k], v[k + 1], s, s);
  } else {
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] *= s;
  }
}
// acc_k <- s * b_k + acc_k
template <int K, bool PK = true>
__device__ __forceinline__ void vec_fma_s(float (&acc)[K], float s, const float (&b)[K]) {
  if constexpr (PK && VecPairs<K>::packed) {
    if (K == 3) acc[0] = fmaf(s, b[0], acc[0]);
#pragma unroll
    for (int k = VecPairs<K>::first; k + 1 < K; k += 2) fma2f(acc[k], acc[k + 1], s, s, b[k], b[k + 1], acc[k], acc[k + 1]);
  } else {
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = fmaf(s, b[k], acc[k]);
  }
}
// acc_k <- a_k * s + acc_k   (same products as vec_fma_s with the operands the other way round)
template <int K, bool PK = true>
__device__ __forceinline__ void vec_fma_vs(float (&acc)[K], const float (&a)[K], float s) {
  if constexpr (PK && VecPairs<K>::packed) {
    if (K == 3) acc[0] = fmaf(a[0], s, acc[0]);
#pragma unroll
    for (int k = VecPairs<K>::first; k + 1 < K; k += 2) fma2f(acc[k], acc[k + 1], a[k], a[k + 1], s, s, acc[k], acc[k + 1]);
  } else {
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = fmaf(a[k], s, acc[k]);
  }
}
// acc <- acc + a
template <int K, bool PK = true>
__device__ __forceinline__ void vec_add(float (&acc)[K], const float (&a)[K]) {
  if constexpr (PK && VecPairs<K>::packed) {
    if (K == 3) acc[0] += a[0];
#pragma unroll
    for (int k = VecPairs<K>::first; k + 1 < K; k += 2) add2f(acc[k], acc[k + 1], acc[k], acc[k + 1], a[k], a[k + 1]);
  } else {
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] += a[k];
  }
}

}  // namespace smf
