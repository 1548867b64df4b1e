// Fused forward-backward kernel: posterior decode (K1) and Baum-Welch E-step (K3).
//
// One GROUP of T threads (T = 32*G) owns one read at a time; a CTA holds `groups` groups and
// walks the reads with a static stride (deterministic read -> CTA map).  Every thread owns a
// chunk of `chunk` consecutive positions (odd, so that all per-thread strided shared-memory
// accesses are bank-conflict free).
//
// Per read (M_t = A_{bin(t-1)} diag(b_t), M_0 = diag(b_0); alpha_t = pi M_0..M_t,
// beta_t = M_{t+1}..M_{L-1} 1; gamma_t ~ alpha_t * beta_t; HMM.py:572-629):
//   load    observations -> shared memory (coalesced, converted to fp32, NaN = missing)
//   phase 1 every thread multiplies the transfer matrices of its chunk:  P_i = prod M_t
//   phase 2 warp-shuffle associative scan of the P_i (inclusive prefix and suffix), one more
//           shuffle scan over the warp totals, giving each thread its exclusive prefix vector
//           v_i = pi P_0..P_{i-1} and suffix vector w_i = P_{i+1}..P_{T-1} 1
//   phase 3 thread-local forward sweep from v_i (alpha -> shared staging), backward sweep from
//           w_i: gamma (overwrites alpha in the staging buffer), arg-max state, and -- E-step
//           only -- xi / emission / start sufficient statistics in registers
//   store   staging -> global, coalesced
//
// All recursions run in scaled probability space in fp32 with exact power-of-two rescaling
// (the reference's log-space values are recovered up to rounding: max |d gamma| ~1e-6 against
// its fp64); log-likelihood pieces and all E-step sums are accumulated in fp64.
#pragma once
#include "hmm_common.cuh"

namespace smf {

constexpr int kMaxSeg = 64;  // segments per read the generic kernel can chain

struct FbArgs {
  const void* obs;
  int obs_dtype;
  long long N;
  int L;
  int C;
  const int8_t* bins;  // device [L-1] or nullptr
  float* gamma;        // [N,L,K] / [N,L] / nullptr
  int gamma_state;     // <0: all states
  int8_t* states;      // [N,L] / nullptr
  double* loglik;      // [N] / nullptr
  double* partials;    // STATS: [gridDim.x][S]
  int S;               // stats length
  int chunk;           // positions per thread (odd)
  int T;               // threads per group (multiple of 32)
  int groups;          // reads in flight per CTA
  int flush_every;     // STATS: reads between float->double flushes
  int resc_mask;       // rescale when (step & resc_mask) == resc_mask: 3 = every 4th step, 1, 0 = every step
  int nseg;            // segments per read (1 = the whole read is staged on chip at once)
  // dynamic shared memory layout (byte offsets)
  int off_bins;        // CTA-wide int8 [L]
  int off_sA;          // CTA-wide float [nb*K*K]
  int off_wacc;        // CTA-wide double [warps][S]
  int off_group0;      // first group's region
  int group_bytes;     // bytes per group
  int goff_in, goff_stage, goff_st, goff_scan;  // offsets inside a group region
};

template <int K>
__device__ __forceinline__ void mat_set_identity(float (&P)[K][K]) {
#pragma unroll
  for (int i = 0; i < K; ++i)
#pragma unroll
    for (int j = 0; j < K; ++j) P[i][j] = (i == j) ? 1.0f : 0.0f;
}

template <int K>
__device__ __forceinline__ float mat_max(const float (&P)[K][K]) {
  float m = P[0][0];
#pragma unroll
  for (int i = 0; i < K; ++i)
#pragma unroll
    for (int j = 0; j < K; ++j) m = fmaxf(m, P[i][j]);
  return m;
}

template <int K, bool PK = true>
__device__ __forceinline__ void mat_scale(float (&P)[K][K], float s) {
#pragma unroll
  for (int i = 0; i < K; ++i) vec_scale<K, PK>(P[i], s);  // packed fp32x2 rows (f32x2.cuh)
}

// C = A * B: row i of C = sum_l A[i][l] * (row l of B), l ascending -- packed rows, the scalar code's order per entry
template <int K, bool PK = true>
__device__ __forceinline__ void mat_mul(const float (&A)[K][K], const float (&B)[K][K],
                                        float (&C)[K][K]) {
#pragma unroll
  for (int i = 0; i < K; ++i) {
    float r[K];
#pragma unroll
    for (int j = 0; j < K; ++j) r[j] = B[0][j];
    vec_scale<K, PK>(r, A[i][0]);
#pragma unroll
    for (int l = 1; l < K; ++l) vec_fma_s<K, PK>(r, A[i][l], B[l]);
#pragma unroll
    for (int j = 0; j < K; ++j) C[i][j] = r[j];
  }
}

template <int K>
__device__ __forceinline__ void mat_rescale_if_small(float (&P)[K][K]) {
  const float m = mat_max<K>(P);
  if (m < 9.5367431640625e-07f) mat_scale<K>(P, pow2_rescale_noexp(m));  // 2^-20
}

template <int K>
__device__ __forceinline__ float vec_max(const float (&v)[K]) {
  float m = v[0];
#pragma unroll
  for (int k = 1; k < K; ++k) m = fmaxf(m, v[k]);
  return m;
}

// Relative emission b_k(o_t)/b_0(o_t) (b[0] == 1) and "any channel observed".
// HMM.py:1490-1501 (single) / 1766-1782 (multi): masked channels contribute log-prob 0.
template <int K, bool MULTI>
__device__ __forceinline__ bool emission_rel(const FbTables<K>& tab, const float* __restrict__ in,
                                             int t, int C, float (&b)[K]) {
  b[0] = 1.0f;
  if (!MULTI) {
    const float o = in[t];
    const bool m = (o == o);
#pragma unroll
    for (int k = 1; k < K; ++k)
      b[k] = ex2_approx(m ? fmaf(o, tab.rdl[k][0], tab.rl0[k][0]) : 0.0f);
    return m;
  } else {
    float x[K];
#pragma unroll
    for (int k = 0; k < K; ++k) x[k] = 0.0f;
    bool any = false;
    for (int c = 0; c < C; ++c) {
      const float o = in[t * C + c];
      const bool m = (o == o);
      any |= m;
#pragma unroll
      for (int k = 1; k < K; ++k) x[k] += m ? fmaf(o, tab.rdl[k][c], tab.rl0[k][c]) : 0.0f;
    }
#pragma unroll
    for (int k = 1; k < K; ++k) b[k] = ex2_approx(x[k]);
    return any;
  }
}

template <bool MULTI>
__device__ __forceinline__ bool any_observed(const float* __restrict__ in, int t, int C) {
  if (!MULTI) {
    const float o = in[t];
    return o == o;
  } else {
    bool any = false;
    for (int c = 0; c < C; ++c) {
      const float o = in[t * C + c];
      any |= (o == o);
    }
    return any;
  }
}

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
  return v;
}
__device__ __forceinline__ float warp_sum_f32(float v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
  return v;
}

template <int K, bool MULTI, bool BINNED, bool STATS, int MAXT>
__global__ void __launch_bounds__(MAXT) fb_kernel(const __grid_constant__ FbTables<K> tab,
                                                  const __grid_constant__ FbArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr unsigned kFull = 0xffffffffu;
  constexpr int KK = K * K;
  // E-step register accumulators (fp32, flushed to fp64 every `flush_every` reads).
  constexpr int NXI = STATS ? (BINNED ? kMaxB * KK : KK) : 1;
  constexpr int NEM = STATS ? (MULTI ? K * kMaxC : K) : 1;

  const int tid = threadIdx.x;
  const int T = a.T;
  const int g = tid / T;          // group (read slot) in the CTA
  const int gt = tid - g * T;     // thread in group
  const int lane = tid & 31;
  const int gw = gt >> 5;         // warp in group
  const int G = T >> 5;           // warps per group
  const int wcta = tid >> 5;      // warp in CTA
  const int L = a.L;
  const int C = a.C;
  const int chunk = a.chunk;
  const int bar_id = 1 + g;

  const int8_t* sbins = reinterpret_cast<const int8_t*>(smem + a.off_bins);
  const float* sA = reinterpret_cast<const float*>(smem + a.off_sA);
  double* wacc = reinterpret_cast<double*>(smem + a.off_wacc);
  unsigned char* gbase = smem + a.off_group0 + (size_t)g * a.group_bytes;
  float* s_in = reinterpret_cast<float*>(gbase + a.goff_in);
  float* s_stage = reinterpret_cast<float*>(gbase + a.goff_stage);
  int8_t* s_st = reinterpret_cast<int8_t*>(gbase + a.goff_st);
  float* s_wtot = reinterpret_cast<float*>(gbase + a.goff_scan);  // [32][KK]
  float* s_u = s_wtot + 32 * KK;                                  // [32][K]
  float* s_z = s_u + 32 * K;                                      // [32][K]
  double* s_ll = reinterpret_cast<double*>(s_z + 32 * K);         // [32]
  // long reads: per-segment transfer matrices and boundary vectors
  float* s_segT = reinterpret_cast<float*>(s_ll + 32);            // [kMaxSeg][KK]
  float* s_useg = s_segT + kMaxSeg * KK;                          // [kMaxSeg][K]
  float* s_zseg = s_useg + kMaxSeg * K;                           // [kMaxSeg][K]
  const int nseg = a.nseg;
  const int Lseg = T * chunk;
  if (nseg > 1) sbins = a.bins;  // too long for a shared copy: read the bin indices from global

  if (BINNED && nseg == 1) {
    int8_t* wb = reinterpret_cast<int8_t*>(smem + a.off_bins);
    for (int i = tid; i < L - 1; i += blockDim.x) wb[i] = a.bins[i];
  }
  if (BINNED) {
    float* wA = reinterpret_cast<float*>(smem + a.off_sA);
    for (int i = tid; i < tab.nb * KK; i += blockDim.x) wA[i] = (&tab.A[0][0][0])[i];
  }
  if (STATS) {
    const int nw = blockDim.x >> 5;
    for (int i = tid; i < nw * a.S; i += blockDim.x) wacc[i] = 0.0;
  }
  __syncthreads();

  float acc_xi[NXI];
  float acc_num[NEM];
  float acc_den[NEM];
  float acc_start[K];
  double acc_ll = 0.0;
  if (STATS) {
#pragma unroll
    for (int i = 0; i < NXI; ++i) acc_xi[i] = 0.0f;
#pragma unroll
    for (int i = 0; i < NEM; ++i) {
      acc_num[i] = 0.0f;
      acc_den[i] = 0.0f;
    }
#pragma unroll
    for (int k = 0; k < K; ++k) acc_start[k] = 0.0f;
  }

  const int t0 = gt * chunk;  // segment-local first position of this thread
  const bool want_ll = STATS || (a.loglik != nullptr);

  // Flush fp32 register statistics into this warp's fp64 shared accumulators.
  auto flush_stats = [&]() {
    if (!STATS) return;
    double* my = wacc + (size_t)wcta * a.S;
    const int nxi = BINNED ? tab.nb * KK : KK;
    const int nem = MULTI ? K * C : K;
    int s = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const float r = warp_sum_f32(acc_start[k]);
      if (lane == 0) my[s + k] += (double)r;
      acc_start[k] = 0.0f;
    }
    s += K;
    if (BINNED) {
      for (int i = 0; i < nxi; ++i) {
        const float r = warp_sum_f32(acc_xi[i]);
        if (lane == 0) my[s + i] += (double)r;
        acc_xi[i] = 0.0f;
      }
    } else {
#pragma unroll
      for (int i = 0; i < KK; ++i) {
        const float r = warp_sum_f32(acc_xi[i]);
        if (lane == 0) my[s + i] += (double)r;
        acc_xi[i] = 0.0f;
      }
    }
    s += nxi;
    if (MULTI) {
      // register arrays are laid out [k][kMaxC]; stats are [k][C]
      for (int k = 0; k < K; ++k)
        for (int c = 0; c < C; ++c) {
          const float rn = warp_sum_f32(acc_num[k * kMaxC + c]);
          const float rd = warp_sum_f32(acc_den[k * kMaxC + c]);
          if (lane == 0) {
            my[s + k * C + c] += (double)rn;
            my[s + nem + k * C + c] += (double)rd;
          }
          acc_num[k * kMaxC + c] = 0.0f;
          acc_den[k * kMaxC + c] = 0.0f;
        }
    } else {
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const float rn = warp_sum_f32(acc_num[k]);
        const float rd = warp_sum_f32(acc_den[k]);
        if (lane == 0) {
          my[s + k] += (double)rn;
          my[s + nem + k] += (double)rd;
        }
        acc_num[k] = 0.0f;
        acc_den[k] = 0.0f;
      }
    }
  };

  const long long read_stride = (long long)gridDim.x * a.groups;
  int iter = 0;
  for (long long n = (long long)blockIdx.x * a.groups + g; n < a.N; n += read_stride, ++iter) {
   double ll_read = 0.0;  // thread 0: log-likelihood of the read accumulated over segments
   // Reads longer than the on-chip staging are cut into `nseg` segments and walked twice:
   // pass 0 only forms every segment's transfer matrix (phases 1-2), a short chain turns those
   // into the vectors entering / leaving each segment, pass 1 does the full work per segment.
   for (int pass = (nseg > 1 ? 0 : 1); pass < 2; ++pass) {
    for (int seg = 0; seg < nseg; ++seg) {
    const int seg0 = seg * Lseg;
    const int Ls = min(Lseg, L - seg0);
    const int t1 = min(t0 + chunk, Ls);
    const bool active = t0 < Ls;
    float u_init[K], z_init[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      u_init[k] = (nseg > 1 && pass == 1) ? s_useg[seg * K + k] : tab.pi[k];
      z_init[k] = (nseg > 1 && pass == 1) ? s_zseg[seg * K + k] : 1.0f;
    }
    // ---------------- load: global -> shared (coalesced, converted to fp32) ----------------
    {
      const long long base = (n * (long long)L + seg0) * C;
      const int tot = Ls * C;
      // the element type is decided once per segment, and four independent loads are in flight per thread (the
      // element-wise loop with the type switch inside took 45 % of this kernel's stall samples)
      auto fill = [&](auto ld) {
        int i = gt;
        for (; i + 3 * T < tot; i += 4 * T) {
          const float v0 = ld(base + i), v1 = ld(base + i + T), v2 = ld(base + i + 2 * T), v3 = ld(base + i + 3 * T);
          s_in[i] = v0;
          s_in[i + T] = v1;
          s_in[i + 2 * T] = v2;
          s_in[i + 3 * T] = v3;
        }
        for (; i < tot; i += T) s_in[i] = ld(base + i);
      };
      if (a.obs_dtype == SMF_OBS_F32) {
        const float* p = reinterpret_cast<const float*>(a.obs);
        fill([&](long long idx) { return __ldg(p + idx); });
      } else if (a.obs_dtype == SMF_OBS_F64) {
        const double* p = reinterpret_cast<const double*>(a.obs);
        fill([&](long long idx) { return static_cast<float>(__ldg(p + idx)); });
      } else {
        const unsigned char* p = reinterpret_cast<const unsigned char*>(a.obs);
        fill([&](long long idx) {
          const unsigned v = __ldg(p + idx);
          return v == 0u ? 0.0f : (v == 1u ? 1.0f : __int_as_float(0x7fc00000));
        });
      }
    }
    bool prevseg_obs = false;  // is the position just before this segment observed? (E-step xi)
    if (STATS && seg0 > 0 && gt == 0) {
      for (int c = 0; c < C; ++c) {
        const float o = load_obs(a.obs, a.obs_dtype, (n * (long long)L + seg0 - 1) * C + c);
        prevseg_obs |= (o == o);
      }
    }
    group_sync(bar_id, T);

    // ---------------- phase 1: chunk transfer-matrix product ----------------
    float P[K][K];
    mat_set_identity<K>(P);
    float ll_sumo[MULTI ? kMaxC : 1];
    float ll_cnt[MULTI ? kMaxC : 1];
#pragma unroll
    for (int c = 0; c < (MULTI ? kMaxC : 1); ++c) {
      ll_sumo[c] = 0.0f;
      ll_cnt[c] = 0.0f;
    }
    if (active) {
      int t = t0;
      float b[K];
      if (seg0 + t == 0) {
        emission_rel<K, MULTI>(tab, s_in, 0, C, b);
#pragma unroll
        for (int j = 0; j < K; ++j) P[j][j] = b[j];
        ++t;
      }
#pragma unroll 2
      for (; t < t1; ++t) {
        emission_rel<K, MULTI>(tab, s_in, t, C, b);
        float Q[K][K];
        if (BINNED) {
          const float* A = sA + (int)sbins[seg0 + t - 1] * KK;
#pragma unroll
          for (int i = 0; i < K; ++i)
#pragma unroll
            for (int j = 0; j < K; ++j) {
              float s = P[i][0] * A[j];
#pragma unroll
              for (int l = 1; l < K; ++l) s = fmaf(P[i][l], A[l * K + j], s);
              Q[i][j] = s * b[j];
            }
        } else {
#pragma unroll
          for (int i = 0; i < K; ++i)
#pragma unroll
            for (int j = 0; j < K; ++j) {
              float s = P[i][0] * tab.A[0][0][j];
#pragma unroll
              for (int l = 1; l < K; ++l) s = fmaf(P[i][l], tab.A[0][l][j], s);
              Q[i][j] = s * b[j];
            }
        }
#pragma unroll
        for (int i = 0; i < K; ++i)
#pragma unroll
          for (int j = 0; j < K; ++j) P[i][j] = Q[i][j];
        if (((t - t0) & a.resc_mask) == a.resc_mask) mat_scale<K>(P, pow2_rescale_noexp(mat_max<K>(P)));
      }
      mat_scale<K>(P, pow2_rescale_noexp(mat_max<K>(P)));
      if (want_ll) {
        // state-0 emission mass of this chunk: sum over observed (l00[c] + o*dl0[c]); kept as
        // (count, sum o) so the fp64 combination below is exact up to the fp32 sum of <= chunk
        // values in [0,1].
        for (int tt = t0; tt < t1; ++tt) {
          if (!MULTI) {
            const float o = s_in[tt];
            if (o == o) {
              ll_sumo[0] += o;
              ll_cnt[0] += 1.0f;
            }
          } else {
#pragma unroll
            for (int c = 0; c < kMaxC; ++c)
              if (c < C) {
                const float o = s_in[tt * C + c];
                if (o == o) {
                  ll_sumo[c] += o;
                  ll_cnt[c] += 1.0f;
                }
              }
          }
        }
      }
    }

    // ---------------- phase 2: warp-shuffle scans of the chunk matrices ----------------
    float pre[K][K], suf[K][K];
#pragma unroll
    for (int i = 0; i < K; ++i)
#pragma unroll
      for (int j = 0; j < K; ++j) {
        pre[i][j] = P[i][j];
        suf[i][j] = P[i][j];
      }
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      float q[K][K], r[K][K];
#pragma unroll
      for (int i = 0; i < K; ++i)
#pragma unroll
        for (int j = 0; j < K; ++j) q[i][j] = __shfl_up_sync(kFull, pre[i][j], d);
      if (lane >= d) {
        mat_mul<K>(q, pre, r);
#pragma unroll
        for (int i = 0; i < K; ++i)
#pragma unroll
          for (int j = 0; j < K; ++j) pre[i][j] = r[i][j];
        mat_rescale_if_small<K>(pre);
      }
#pragma unroll
      for (int i = 0; i < K; ++i)
#pragma unroll
        for (int j = 0; j < K; ++j) q[i][j] = __shfl_down_sync(kFull, suf[i][j], d);
      if (lane + d < 32) {
        mat_mul<K>(suf, q, r);
#pragma unroll
        for (int i = 0; i < K; ++i)
#pragma unroll
          for (int j = 0; j < K; ++j) suf[i][j] = r[i][j];
        mat_rescale_if_small<K>(suf);
      }
    }

    float u[K], z[K];  // prefix row vector entering this warp, suffix column vector leaving it
    if (G > 1) {
      if (lane == 31) {
#pragma unroll
        for (int i = 0; i < K; ++i)
#pragma unroll
          for (int j = 0; j < K; ++j) s_wtot[gw * KK + i * K + j] = pre[i][j];
      }
      group_sync(bar_id, T);
      if (gw == 0) {
        float wp[K][K], ws[K][K];
        if (lane < G) {
#pragma unroll
          for (int i = 0; i < K; ++i)
#pragma unroll
            for (int j = 0; j < K; ++j) wp[i][j] = s_wtot[lane * KK + i * K + j];
        } else {
          mat_set_identity<K>(wp);
        }
#pragma unroll
        for (int i = 0; i < K; ++i)
#pragma unroll
          for (int j = 0; j < K; ++j) ws[i][j] = wp[i][j];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          float q[K][K], r[K][K];
#pragma unroll
          for (int i = 0; i < K; ++i)
#pragma unroll
            for (int j = 0; j < K; ++j) q[i][j] = __shfl_up_sync(kFull, wp[i][j], d);
          if (lane >= d) {
            mat_mul<K>(q, wp, r);
#pragma unroll
            for (int i = 0; i < K; ++i)
#pragma unroll
              for (int j = 0; j < K; ++j) wp[i][j] = r[i][j];
            mat_rescale_if_small<K>(wp);
          }
#pragma unroll
          for (int i = 0; i < K; ++i)
#pragma unroll
            for (int j = 0; j < K; ++j) q[i][j] = __shfl_down_sync(kFull, ws[i][j], d);
          if (lane + d < 32) {
            mat_mul<K>(ws, q, r);
#pragma unroll
            for (int i = 0; i < K; ++i)
#pragma unroll
              for (int j = 0; j < K; ++j) ws[i][j] = r[i][j];
            mat_rescale_if_small<K>(ws);
          }
        }
        // exclusive versions: prefix of lane-1, suffix of lane+1
        if (pass == 0 && lane == G - 1) {
          // transfer matrix of the whole segment (normalised)
          const float sc = pow2_rescale_noexp(mat_max<K>(wp));
#pragma unroll
          for (int i = 0; i < K; ++i)
#pragma unroll
            for (int j = 0; j < K; ++j) s_segT[seg * KK + i * K + j] = wp[i][j] * sc;
        }
        float uo[K], zo[K];
#pragma unroll
        for (int j = 0; j < K; ++j) {
          float s = u_init[0] * wp[0][j];
#pragma unroll
          for (int i = 1; i < K; ++i) s = fmaf(u_init[i], wp[i][j], s);
          uo[j] = s;  // u_init * inclusive_prefix(lane)
        }
#pragma unroll
        for (int i = 0; i < K; ++i) {
          float s = ws[i][0] * z_init[0];
#pragma unroll
          for (int j = 1; j < K; ++j) s = fmaf(ws[i][j], z_init[j], s);
          zo[i] = s;  // inclusive_suffix(lane) * z_init
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
          float up = __shfl_up_sync(kFull, uo[k], 1);
          float zn = __shfl_down_sync(kFull, zo[k], 1);
          if (lane == 0) up = u_init[k];
          if (lane == 31) zn = z_init[k];
          s_u[lane * K + k] = up;
          s_z[lane * K + k] = zn;
        }
      }
      group_sync(bar_id, T);
#pragma unroll
      for (int k = 0; k < K; ++k) {
        u[k] = s_u[gw * K + k];
        z[k] = (gw + 1 < G) ? s_z[gw * K + k] : z_init[k];
      }
    } else {
      if (pass == 0 && lane == 31) {
#pragma unroll
        for (int i = 0; i < K; ++i)
#pragma unroll
          for (int j = 0; j < K; ++j) s_segT[seg * KK + i * K + j] = pre[i][j];
      }
      __syncwarp();
#pragma unroll
      for (int k = 0; k < K; ++k) {
        u[k] = u_init[k];
        z[k] = z_init[k];
      }
    }
    if (pass == 0) continue;  // pass 0 stops after the segment's transfer matrix

    // thread-exclusive boundary vectors
    float v[K], w[K];
    {
      float vo[K], wo[K];
#pragma unroll
      for (int j = 0; j < K; ++j) {
        float s = u[0] * pre[0][j];
#pragma unroll
        for (int i = 1; i < K; ++i) s = fmaf(u[i], pre[i][j], s);
        vo[j] = s;
      }
#pragma unroll
      for (int i = 0; i < K; ++i) {
        float s = suf[i][0] * z[0];
#pragma unroll
        for (int j = 1; j < K; ++j) s = fmaf(suf[i][j], z[j], s);
        wo[i] = s;
      }
#pragma unroll
      for (int k = 0; k < K; ++k) {
        v[k] = __shfl_up_sync(kFull, vo[k], 1);
        w[k] = __shfl_down_sync(kFull, wo[k], 1);
        if (lane == 0) v[k] = u[k];
        if (lane == 31) w[k] = z[k];
      }
      float sv = v[0], sw = w[0];
#pragma unroll
      for (int k = 1; k < K; ++k) {
        sv += v[k];
        sw += w[k];
      }
      const float rv = 1.0f / sv, rw = 1.0f / sw;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        v[k] *= rv;
        w[k] *= rw;
      }
    }

    // ---------------- phase 3: local forward / backward sweeps ----------------
    double ll_thread = 0.0;
    if (active) {
      float al[K];
      int e = 0;
#pragma unroll
      for (int k = 0; k < K; ++k) al[k] = v[k];
      {
        int t = t0;
        float b[K];
        if (seg0 + t == 0) {
          emission_rel<K, MULTI>(tab, s_in, 0, C, b);
#pragma unroll
          for (int k = 0; k < K; ++k) {
            al[k] *= b[k];
            s_stage[k] = al[k];
          }
          ++t;
        }
#pragma unroll 2
        for (; t < t1; ++t) {
          emission_rel<K, MULTI>(tab, s_in, t, C, b);
          float nx[K];
          if (BINNED) {
            const float* A = sA + (int)sbins[seg0 + t - 1] * KK;
#pragma unroll
            for (int j = 0; j < K; ++j) {
              float s = al[0] * A[j];
#pragma unroll
              for (int i = 1; i < K; ++i) s = fmaf(al[i], A[i * K + j], s);
              nx[j] = s * b[j];
            }
          } else {
#pragma unroll
            for (int j = 0; j < K; ++j) {
              float s = al[0] * tab.A[0][0][j];
#pragma unroll
              for (int i = 1; i < K; ++i) s = fmaf(al[i], tab.A[0][i][j], s);
              nx[j] = s * b[j];
            }
          }
          if (((t - t0) & a.resc_mask) == a.resc_mask) {
            const float sc = pow2_rescale(vec_max<K>(nx), e);
#pragma unroll
            for (int k = 0; k < K; ++k) nx[k] *= sc;
          }
#pragma unroll
          for (int k = 0; k < K; ++k) {
            al[k] = nx[k];
            s_stage[t * K + k] = nx[k];
          }
        }
      }
      if (want_ll) {
        float s = al[0];
#pragma unroll
        for (int k = 1; k < K; ++k) s += al[k];
        // log of P(chunk | past) with the state-0 emission factored out, plus that factor
        ll_thread = (double)logf(s) + (double)e * 0.6931471805599453094;
#pragma unroll
        for (int c = 0; c < (MULTI ? kMaxC : 1); ++c)
          if (c < C) ll_thread += (double)ll_cnt[c] * tab.l00[c] + (double)ll_sumo[c] * tab.dl0[c];
        if (seg0 + t0 == 0) ll_thread += tab.log_pi_sum;
      }

      // backward sweep
      float bv[K];
#pragma unroll
      for (int k = 0; k < K; ++k) bv[k] = w[k];
      float alc[K];
#pragma unroll
      for (int k = 0; k < K; ++k) alc[k] = s_stage[(t1 - 1) * K + k];
      const int gs = a.gamma_state;
#pragma unroll 2
      for (int t = t1 - 1; t >= t0; --t) {
        float b[K];
        const bool m_t = emission_rel<K, MULTI>(tab, s_in, t, C, b);
        float gk[K];
        float s = 0.0f;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          gk[k] = alc[k] * bv[k];
          s += gk[k];
        }
        const float r = 1.0f / s;
        int best = 0;
        float gbest;
#pragma unroll
        for (int k = 0; k < K; ++k) gk[k] *= r;
        gbest = gk[0];
#pragma unroll
        for (int k = 1; k < K; ++k)
          if (gk[k] > gbest) {
            gbest = gk[k];
            best = k;
          }
        // previous position's alpha must be fetched before its slot neighbours are overwritten
        float alp[K];
        const bool has_prev = (t > t0);
        if (has_prev) {
#pragma unroll
          for (int k = 0; k < K; ++k) alp[k] = s_stage[(t - 1) * K + k];
        } else {
#pragma unroll
          for (int k = 0; k < K; ++k) alp[k] = v[k];
        }
        if (gs < 0) {
#pragma unroll
          for (int k = 0; k < K; ++k) s_stage[t * K + k] = gk[k];
        } else {
          float gsel = gk[0];
#pragma unroll
          for (int k = 1; k < K; ++k)
            if (gs == k) gsel = gk[k];
          s_stage[t * K] = gsel;
        }
        s_st[t] = (int8_t)best;

        if (STATS) {
          if (seg0 + t == 0) {
#pragma unroll
            for (int k = 0; k < K; ++k) acc_start[k] += gk[k];
          }
          if (!MULTI) {
            if (m_t) {
              const float o = s_in[t];
#pragma unroll
              for (int k = 0; k < K; ++k) {
                acc_num[k] = fmaf(gk[k], o, acc_num[k]);
                acc_den[k] += gk[k];
              }
            }
          } else {
#pragma unroll
            for (int c = 0; c < kMaxC; ++c)
              if (c < C) {
                const float o = s_in[t * C + c];
                if (o == o) {
#pragma unroll
                  for (int k = 0; k < K; ++k) {
                    acc_num[k * kMaxC + c] = fmaf(gk[k], o, acc_num[k * kMaxC + c]);
                    acc_den[k * kMaxC + c] += gk[k];
                  }
                }
              }
          }
        }

        if (has_prev || (STATS && seg0 + t > 0)) {
          float cb[K], nb[K];
#pragma unroll
          for (int j = 0; j < K; ++j) cb[j] = b[j] * bv[j];
          float prod[K][K];
          int bin = 0;
          if (BINNED) {
            bin = (int)sbins[seg0 + t - 1];
            const float* A = sA + bin * KK;
#pragma unroll
            for (int i = 0; i < K; ++i)
#pragma unroll
              for (int j = 0; j < K; ++j) prod[i][j] = A[i * K + j] * cb[j];
          } else {
#pragma unroll
            for (int i = 0; i < K; ++i)
#pragma unroll
              for (int j = 0; j < K; ++j) prod[i][j] = tab.A[0][i][j] * cb[j];
          }
#pragma unroll
          for (int i = 0; i < K; ++i) {
            float sN = prod[i][0];
#pragma unroll
            for (int j = 1; j < K; ++j) sN += prod[i][j];
            nb[i] = sN;
          }
          if (STATS) {
            // xi_{t-1}(i,j) = alpha_{t-1}(i) A_ij b_t(j) beta_t(j) / sum, counted only when both
            // endpoints are observed (HMM.py:1583-1591)
            const bool valid = m_t && (t > 0 ? any_observed<MULTI>(s_in, t - 1, C) : prevseg_obs);
            float s2 = 0.0f;
#pragma unroll
            for (int i = 0; i < K; ++i) s2 = fmaf(alp[i], nb[i], s2);
            const float r2 = valid ? (1.0f / s2) : 0.0f;
            if (BINNED) {
              float* dst = acc_xi + bin * KK;
#pragma unroll
              for (int i = 0; i < K; ++i) {
                const float ai = alp[i] * r2;
#pragma unroll
                for (int j = 0; j < K; ++j) dst[i * K + j] = fmaf(ai, prod[i][j], dst[i * K + j]);
              }
            } else {
#pragma unroll
              for (int i = 0; i < K; ++i) {
                const float ai = alp[i] * r2;
#pragma unroll
                for (int j = 0; j < K; ++j)
                  acc_xi[i * K + j] = fmaf(ai, prod[i][j], acc_xi[i * K + j]);
              }
            }
          }
          if ((((t1 - 1) - t) & a.resc_mask) == a.resc_mask) {
            const float sc = pow2_rescale_noexp(vec_max<K>(nb));
#pragma unroll
            for (int k = 0; k < K; ++k) nb[k] *= sc;
          }
#pragma unroll
          for (int k = 0; k < K; ++k) bv[k] = nb[k];
        }
#pragma unroll
        for (int k = 0; k < K; ++k) alc[k] = alp[k];
      }
    }

    // per-read log-likelihood (posterior API) / running total (E-step)
    if (STATS) acc_ll += ll_thread;
    if (a.loglik != nullptr) {
      const double wsum = warp_sum_f64(ll_thread);
      if (lane == 0) s_ll[gw] = wsum;
    }
    group_sync(bar_id, T);

    // ---------------- store: shared staging -> global (coalesced) ----------------
    if (a.loglik != nullptr && gt == 0) {
      for (int i = 0; i < G; ++i) ll_read += s_ll[i];
      if (seg == nseg - 1) a.loglik[n] = ll_read;
    }
    if (a.gamma != nullptr) {
      if (a.gamma_state < 0) {
        float* dst = a.gamma + (n * (long long)L + seg0) * K;
        const int tot = Ls * K;
        for (int i = gt; i < tot; i += T) dst[i] = s_stage[i];
      } else {
        float* dst = a.gamma + n * (long long)L + seg0;
        for (int i = gt; i < Ls; i += T) dst[i] = s_stage[i * K];
      }
    }
    if (a.states != nullptr) {
      int8_t* dst = a.states + n * (long long)L + seg0;
      for (int i = gt; i < Ls; i += T) dst[i] = s_st[i];
    }
    if (nseg > 1) group_sync(bar_id, T);  // staging drained before the next segment is decoded
    }  // segments
    if (pass == 0) {
      // chain the segment matrices into the vectors entering (u) / leaving (z) every segment
      group_sync(bar_id, T);
      if (gt == 0) {
        float x[K];
#pragma unroll
        for (int k = 0; k < K; ++k) x[k] = tab.pi[k];
        for (int sg = 0; sg < nseg; ++sg) {
#pragma unroll
          for (int k = 0; k < K; ++k) s_useg[sg * K + k] = x[k];
          float y[K];
#pragma unroll
          for (int j = 0; j < K; ++j) {
            float sacc = x[0] * s_segT[sg * KK + j];
#pragma unroll
            for (int i = 1; i < K; ++i) sacc = fmaf(x[i], s_segT[sg * KK + i * K + j], sacc);
            y[j] = sacc;
          }
          const float sc = pow2_rescale_noexp(vec_max<K>(y));
#pragma unroll
          for (int k = 0; k < K; ++k) x[k] = y[k] * sc;
        }
#pragma unroll
        for (int k = 0; k < K; ++k) x[k] = 1.0f;
        for (int sg = nseg - 1; sg >= 0; --sg) {
#pragma unroll
          for (int k = 0; k < K; ++k) s_zseg[sg * K + k] = x[k];
          float y[K];
#pragma unroll
          for (int i = 0; i < K; ++i) {
            float sacc = s_segT[sg * KK + i * K] * x[0];
#pragma unroll
            for (int j = 1; j < K; ++j) sacc = fmaf(s_segT[sg * KK + i * K + j], x[j], sacc);
            y[i] = sacc;
          }
          const float sc = pow2_rescale_noexp(vec_max<K>(y));
#pragma unroll
          for (int k = 0; k < K; ++k) x[k] = y[k] * sc;
        }
      }
      group_sync(bar_id, T);
    }
   }  // passes
    if (STATS && ((iter + 1) % a.flush_every) == 0) flush_stats();
  }

  if (STATS) {
    flush_stats();
    const double wl = warp_sum_f64(acc_ll);
    if (lane == 0) wacc[(size_t)wcta * a.S + (a.S - 1)] += wl;
    __syncthreads();
    const int nw = blockDim.x >> 5;
    for (int s = tid; s < a.S; s += blockDim.x) {
      double tot = 0.0;
      for (int wv = 0; wv < nw; ++wv) tot += wacc[(size_t)wv * a.S + s];
      a.partials[(size_t)blockIdx.x * a.S + s] = tot;
    }
  }
}

}  // namespace smf
