// Shared device helpers for the sm_100a HMM kernels (posterior / E-step / Viterbi).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "smf_hmm.h"
#include "f32x2.cuh"

namespace smf {

constexpr int kMaxK = SMF_MAX_STATES;
constexpr int kMaxC = SMF_MAX_CHANNELS;
constexpr int kMaxB = SMF_MAX_BINS;

// Linear-domain (fp32) tables for the scaled forward-backward kernels, passed by value in
// the kernel parameter space.  A = exp(log_trans) = trans + eps, pi = start + eps
// (HMM.py:594-595): NOT re-normalised, exactly like the reference's log-space recursion.
template <int K>
struct FbTables {
  float pi[K];
  float A[kMaxB][K][K];
  // emission relative to state 0, log2 domain:  b_k(o)/b_0(o) = 2^(rl0[k][c] + o*rdl[k][c])
  float rl0[K][kMaxC];
  float rdl[K][kMaxC];
  // state-0 emission in natural log (fp64) for the log-likelihood:  log b_0(o) = l00[c] + o*dl0[c]
  double l00[kMaxC];
  double dl0[kMaxC];
  double log_pi_sum;  // log(sum_k pi[k])
  // Fast path (single channel, one transition matrix): column-normalised transitions
  //   A = Ap diag(s),  s_j = A[j][j],  Ap[j][j] = 1
  // so a step costs K(K-1) FMAs; s_j/s_0 is folded into the cached emission ratio
  //   c_j(t) = s_j b_j(o_t) / (s_0 b_0(o_t)) = 2^(rl0p[j] + o*rdl[j][0])   (observed)
  //                                           = 2^(rmiss[j])                (missing)
  // and into the start vector pi2[j] = pi[j] s_0 / s_j.  Every step drops the scalar
  // s_0 b_0(o_t); the log-likelihood adds it back ((L-1) log s_0 + state-0 emission mass).
  float Ap[K][K];
  float pi2[K];
  float rl0p[K];
  float rmiss[K];
  double log_s0;
  double log_pi2_sum;
  double log_s0b[kMaxB];  // distance-binned fast path: log A_b[0][0] of every bin
  int nb;
  int C;
};

// fp64 log-domain tables for Viterbi (HMM.py:647-649).
template <int K>
struct VitTables {
  double log_pi[K];
  double logA[kMaxB][K][K];
  double l1[K][kMaxC];
  double l0[K][kMaxC];
  int nb;
  int C;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Exact power-of-two rescale factor: returns s = 2^-k with m*s in [0.5, 1) and adds k to e.
__device__ __forceinline__ float pow2_rescale(float m, int& e) {
  const int eb = (__float_as_int(m) >> 23) & 0xff;
  const int k = eb - 126;
  e += k;
  return __int_as_float((127 - k) << 23);
}
__device__ __forceinline__ float pow2_rescale_noexp(float m) {
  const int eb = (__float_as_int(m) >> 23) & 0xff;
  return __int_as_float((253 - eb) << 23);
}

// Observation element -> float with NaN = missing (HMM.py:707-708).
__device__ __forceinline__ float load_obs(const void* base, int dtype, long long idx) {
  if (dtype == SMF_OBS_F32) return __ldg(reinterpret_cast<const float*>(base) + idx);
  if (dtype == SMF_OBS_F64) return static_cast<float>(__ldg(reinterpret_cast<const double*>(base) + idx));
  const unsigned v = __ldg(reinterpret_cast<const unsigned char*>(base) + idx);
  return v == 0u ? 0.0f : (v == 1u ? 1.0f : __int_as_float(0x7fc00000));
}
__device__ __forceinline__ double load_obs_f64(const void* base, int dtype, long long idx) {
  if (dtype == SMF_OBS_F32) return static_cast<double>(__ldg(reinterpret_cast<const float*>(base) + idx));
  if (dtype == SMF_OBS_F64) return __ldg(reinterpret_cast<const double*>(base) + idx);
  const unsigned v = __ldg(reinterpret_cast<const unsigned char*>(base) + idx);
  return v == 0u ? 0.0 : (v == 1u ? 1.0 : __longlong_as_double(0x7ff8000000000000LL));
}

// Named barrier over `nthreads` threads (multiple of 32) with id in [1, 15].
__device__ __forceinline__ void group_sync(int id, int nthreads) {
  if (nthreads == 32) {
    __syncwarp();
  } else {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
  }
}

}  // namespace smf
