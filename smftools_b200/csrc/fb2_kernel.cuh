// Fast path of the fused forward-backward kernel (single channel, homogeneous transitions).
//
// Same algorithm as fb_kernel.cuh (chunk products -> warp-shuffle associative scan -> local
// forward/backward sweeps), re-shaped for instruction throughput and asynchronous data movement:
//   * the chunk length CH is a compile-time odd constant: every per-position loop is fully
//     unrolled (no loop/branch/address arithmetic), and the emission ratios b_k/b_0 of a thread's
//     chunk live in registers, so the transcendental work is done once instead of three times;
//   * a read's observations arrive with ONE bulk asynchronous copy (TMA engine,
//     cp.async.bulk global->shared, completion on an mbarrier); the next read is prefetched as
//     soon as every thread has pulled its chunk into registers, so the copy overlaps phases 2-3;
//   * posteriors and states leave through bulk asynchronous shared->global stores; the few
//     bytes before/after the 16-byte aligned body are written with ordinary stores;
//   * a read whose length is not a multiple of CH is padded (in shared memory) with missing
//     observations, which leave every posterior of the real positions unchanged (beta of the
//     last position stays uniform up to fp32 rounding, ~1e-7) and contribute no statistics.
#pragma once
#include "fb_kernel.cuh"

namespace smf {

struct Fb2Args {
  const void* obs;
  long long N;
  int L;
  float* gamma;     // [N,L,K] / [N,L] / nullptr
  int gamma_state;  // <0: all states
  int8_t* states;   // [N,L] / nullptr
  double* loglik;   // [N] / nullptr
  double* partials; // STATS: [gridDim.x][S]
  int S;
  int T;            // threads per group (multiple of 32)
  int groups;       // reads in flight per CTA
  int flush_every;
  int nbuf;         // output staging buffers per group (2 = a read's bulk stores drain while the next read computes)
  int obuf_bytes;   // bytes between the two staging sets (stage | out1 | st)
  int off_wacc;     // CTA-wide double [warps][S]
  int off_group0;
  int group_bytes;
  // offsets inside a group region (all 16-byte aligned)
  int goff_bar, goff_in, goff_stage, goff_out1, goff_st, goff_scan;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "MBAR_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra MBAR_DONE;\n"
      "bra MBAR_WAIT;\n"
      "MBAR_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// TMA bulk copy global -> shared, completion signalled on an mbarrier (16-byte granules).
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// TMA bulk copy shared -> global (bulk async-group completion).
__device__ __forceinline__ void bulk_store(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <typename ObsT>
__device__ __forceinline__ float obs_to_float(ObsT v);
template <>
__device__ __forceinline__ float obs_to_float<float>(float v) { return v; }
template <>
__device__ __forceinline__ float obs_to_float<double>(double v) { return static_cast<float>(v); }
template <>
__device__ __forceinline__ float obs_to_float<unsigned char>(unsigned char v) {
  return v == 0 ? 0.0f : (v == 1 ? 1.0f : __int_as_float(0x7fc00000));
}
template <typename ObsT>
__device__ __forceinline__ ObsT obs_missing();
template <>
__device__ __forceinline__ float obs_missing<float>() { return __int_as_float(0x7fc00000); }
template <>
__device__ __forceinline__ double obs_missing<double>() { return __longlong_as_double(0x7ff8000000000000LL); }
template <>
__device__ __forceinline__ unsigned char obs_missing<unsigned char>() { return 255; }

// Store bytes [0, nbytes) of a shared buffer to global: the 16-byte aligned body with one bulk
// async store (issued by the calling thread), head / tail bytes with plain stores.  `src` and
// `dst` must be congruent modulo 16.  Called by ONE thread; UNIT is the element size used for
// the ragged ends.
template <typename UNIT>
__device__ __forceinline__ void store_row_bulk(void* dst, const void* src, uint32_t nbytes) {
  const uint32_t mis = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(dst) & 15u);
  uint32_t head = (16u - mis) & 15u;
  if (head > nbytes) head = nbytes;
  const uint32_t body = (nbytes - head) & ~15u;
  const uint32_t tail = nbytes - head - body;
  if (body) bulk_store(static_cast<char*>(dst) + head, static_cast<const char*>(src) + head, body);
  const UNIT* s = static_cast<const UNIT*>(src);
  UNIT* d = static_cast<UNIT*>(dst);
  for (uint32_t i = 0; i < head / sizeof(UNIT); ++i) d[i] = s[i];
  const uint32_t t0 = (head + body) / sizeof(UNIT);
  for (uint32_t i = 0; i < tail / sizeof(UNIT); ++i) d[t0 + i] = s[t0 + i];
}

template <int K, int CH, bool STATS, typename ObsT, int MAXT>
__global__ void __launch_bounds__(MAXT) fb2_kernel(const __grid_constant__ FbTables<K> tab,
                                                   const __grid_constant__ Fb2Args a) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr unsigned kFull = 0xffffffffu;
  constexpr int KK = K * K;
  constexpr int KR = K - 1;  // cached emission ratios per position

  const int tid = threadIdx.x;
  const int T = a.T;
  const int g = tid / T;
  const int gt = tid - g * T;
  const int lane = tid & 31;
  const int gw = gt >> 5;
  const int G = T >> 5;
  const int wcta = tid >> 5;
  const int L = a.L;
  const int bar_id = 1 + g;

  double* wacc = reinterpret_cast<double*>(smem + a.off_wacc);
  unsigned char* gbase = smem + a.off_group0 + (size_t)g * a.group_bytes;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(gbase + a.goff_bar);
  unsigned char* s_in_raw = gbase + a.goff_in;
  float* s_stage_base = reinterpret_cast<float*>(gbase + a.goff_stage);
  float* s_out1_base = reinterpret_cast<float*>(gbase + a.goff_out1);
  int8_t* s_st_base = reinterpret_cast<int8_t*>(gbase + a.goff_st);
  float* s_wtot = reinterpret_cast<float*>(gbase + a.goff_scan);  // [32][KK]
  float* s_u = s_wtot + 32 * KK;
  float* s_z = s_u + 32 * K;
  double* s_ll = reinterpret_cast<double*>(s_z + 32 * K);  // [32]

  if (gt == 0) mbar_init(mbar, 1);
  if (STATS) {
    const int nw = blockDim.x >> 5;
    for (int i = tid; i < nw * a.S; i += blockDim.x) wacc[i] = 0.0;
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();

  constexpr int NXI = STATS ? KK : 1;
  constexpr int NEM = STATS ? K : 1;
  float acc_xi[NXI], acc_num[NEM], acc_den[NEM], acc_start[NEM];
  double acc_ll = 0.0;
#pragma unroll
  for (int i = 0; i < NXI; ++i) acc_xi[i] = 0.0f;
#pragma unroll
  for (int i = 0; i < NEM; ++i) {
    acc_num[i] = 0.0f;
    acc_den[i] = 0.0f;
    acc_start[i] = 0.0f;
  }

  auto flush_stats = [&]() {
    if (!STATS) return;
    double* my = wacc + (size_t)wcta * a.S;
#pragma unroll
    for (int k = 0; k < NEM; ++k) {
      const float r0 = warp_sum_f32(acc_start[k]);
      const float r1 = warp_sum_f32(acc_num[k]);
      const float r2 = warp_sum_f32(acc_den[k]);
      if (lane == 0) {
        my[k] += (double)r0;
        my[K + KK + k] += (double)r1;
        my[K + KK + K + k] += (double)r2;
      }
      acc_start[k] = 0.0f;
      acc_num[k] = 0.0f;
      acc_den[k] = 0.0f;
    }
#pragma unroll
    for (int i = 0; i < NXI; ++i) {
      const float r = warp_sum_f32(acc_xi[i]);
      if (lane == 0) my[K + i] += (double)r;
      acc_xi[i] = 0.0f;
    }
  };

  const int t0 = gt * CH;
  const bool active = t0 < L;
  const bool first = (t0 == 0);
  const bool want_ll = STATS || (a.loglik != nullptr);
  const int gs = a.gamma_state;
  const bool w_gamma_all = !STATS && a.gamma != nullptr && gs < 0;
  const bool w_gamma_one = !STATS && a.gamma != nullptr && gs >= 0;
  const bool w_states = !STATS && a.states != nullptr;
  const uint32_t row_bytes = (uint32_t)L * (uint32_t)sizeof(ObsT);
  const char* obs_base = static_cast<const char*>(a.obs);

  auto issue_load = [&](long long n) {
    const char* src = obs_base + (size_t)n * row_bytes;
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15u);
    const uint32_t bytes = (mis + row_bytes + 15u) & ~15u;
    mbar_expect_tx(mbar, bytes);
    bulk_load(s_in_raw, src - mis, bytes, mbar);
  };

  const long long read_stride = (long long)gridDim.x * a.groups;
  long long n = (long long)blockIdx.x * a.groups + g;
  if (gt == 0 && n < a.N) issue_load(n);

  float A[K][K];
#pragma unroll
  for (int i = 0; i < K; ++i)
#pragma unroll
    for (int j = 0; j < K; ++j) A[i][j] = tab.A[0][i][j];

  int iter = 0;
  for (; n < a.N; n += read_stride, ++iter) {
    // where this read's rows sit relative to 16-byte granules (uniform per group)
    const uint32_t in_mis =
        (uint32_t)(reinterpret_cast<uintptr_t>(obs_base + (size_t)n * row_bytes) & 15u) / (uint32_t)sizeof(ObsT);
    ObsT* s_in = reinterpret_cast<ObsT*>(s_in_raw) + in_mis;
    const size_t obuf = (a.nbuf > 1 && (iter & 1)) ? (size_t)a.obuf_bytes : 0;
    float* s_stage = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(s_stage_base) + obuf);
    float* s_out1 = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(s_out1_base) + obuf);
    int8_t* s_st = s_st_base + obuf;
    if (w_gamma_all) s_stage += (reinterpret_cast<uintptr_t>(a.gamma + (size_t)n * L * K) & 15u) >> 2;
    if (w_gamma_one) s_out1 += (reinterpret_cast<uintptr_t>(a.gamma + (size_t)n * L) & 15u) >> 2;
    if (w_states) s_st += reinterpret_cast<uintptr_t>(a.states + (size_t)n * L) & 15u;

    mbar_wait(mbar, (uint32_t)(iter & 1));

    // ---------------- phase 1: chunk -> registers, transfer-matrix product ----------------
    float rt[CH][KR];        // emission ratios b_k / b_0, k = 1..K-1
    float ov[STATS ? CH : 1];  // observation values (E-step emission statistics)
    unsigned obsmask = 0;    // bit j: position t0+j observed
    float P[K][K];
    mat_set_identity<K>(P);
    float ll_sumo = 0.0f;
    bool prev_obs0 = false;  // observation just before this chunk present? (E-step xi validity)
    if (active) {
      if (STATS && !first) {
        const float op = obs_to_float<ObsT>(s_in[t0 - 1]);
        prev_obs0 = (op == op);
      }
      if (t0 + CH > L) {
        // pad the ragged end with missing observations (this thread is the only reader)
        for (int t = L; t < t0 + CH; ++t) s_in[t] = obs_missing<ObsT>();
      }
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        const float o = obs_to_float<ObsT>(s_in[t0 + j]);
        const bool m = (o == o);
        obsmask |= m ? (1u << j) : 0u;
        if (STATS) ov[j] = m ? o : 0.0f;
        if (want_ll) ll_sumo += m ? o : 0.0f;
        float b[K];
        b[0] = 1.0f;
#pragma unroll
        for (int k = 1; k < K; ++k) {
          b[k] = ex2_approx(m ? fmaf(o, tab.rdl[k][0], tab.rl0[k][0]) : 0.0f);
          rt[j][k - 1] = b[k];
        }
        if (j == 0) {
          // P = I so far: P <- A diag(b), or diag(b) for the first position of the read
#pragma unroll
          for (int i = 0; i < K; ++i)
#pragma unroll
            for (int jj = 0; jj < K; ++jj) {
              const float aij = first ? (i == jj ? 1.0f : 0.0f) : A[i][jj];
              P[i][jj] = (jj == 0) ? aij : aij * b[jj];
            }
        } else {
          float Q[K][K];
#pragma unroll
          for (int i = 0; i < K; ++i)
#pragma unroll
            for (int jj = 0; jj < K; ++jj) {
              float s = P[i][0] * A[0][jj];
#pragma unroll
              for (int l = 1; l < K; ++l) s = fmaf(P[i][l], A[l][jj], s);
              Q[i][jj] = (jj == 0) ? s : s * b[jj];
            }
#pragma unroll
          for (int i = 0; i < K; ++i)
#pragma unroll
            for (int jj = 0; jj < K; ++jj) P[i][jj] = Q[i][jj];
        }
        if ((j & 3) == 3 || j == CH - 1) mat_scale<K>(P, pow2_rescale_noexp(mat_max<K>(P)));
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

    float u[K], z[K];
    if (G > 1) {
      if (lane == 31) {
#pragma unroll
        for (int i = 0; i < K; ++i)
#pragma unroll
          for (int j = 0; j < K; ++j) s_wtot[gw * KK + i * K + j] = pre[i][j];
      }
      group_sync(bar_id, T);  // B1: chunk registers loaded by everyone, warp totals visible
    } else {
      __syncwarp();
    }
    if (gt == 0) {
      // the observation buffer is free: prefetch the next read of this group
      const long long nn = n + read_stride;
      if (nn < a.N) issue_load(nn);
      // the bulk stores that last used this iteration's staging set must have finished READING it
      if (a.nbuf > 1) bulk_wait_read_1(); else bulk_wait_read_all();
    }
    if (G > 1) {
      if (gw < 2) {
        // warp 0: exclusive prefix vectors u_w = pi T_0..T_{w-1};  warp 1: exclusive suffix vectors
        // z_w = T_{w+1}..T_{G-1} 1   (T_w = product of warp w's chunk matrices)
        float wm[K][K];
        if (lane < G) {
#pragma unroll
          for (int i = 0; i < K; ++i)
#pragma unroll
            for (int j = 0; j < K; ++j) wm[i][j] = s_wtot[lane * KK + i * K + j];
        } else {
          mat_set_identity<K>(wm);
        }
        if (gw == 0) {
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            float q[K][K], r[K][K];
#pragma unroll
            for (int i = 0; i < K; ++i)
#pragma unroll
              for (int j = 0; j < K; ++j) q[i][j] = __shfl_up_sync(kFull, wm[i][j], d);
            if (lane >= d) {
              mat_mul<K>(q, wm, r);
#pragma unroll
              for (int i = 0; i < K; ++i)
#pragma unroll
                for (int j = 0; j < K; ++j) wm[i][j] = r[i][j];
              mat_rescale_if_small<K>(wm);
            }
          }
#pragma unroll
          for (int j = 0; j < K; ++j) {
            float sacc = tab.pi[0] * wm[0][j];
#pragma unroll
            for (int i = 1; i < K; ++i) sacc = fmaf(tab.pi[i], wm[i][j], sacc);
            float up = __shfl_up_sync(kFull, sacc, 1);
            if (lane == 0) up = tab.pi[j];
            s_u[lane * K + j] = up;
          }
        } else {
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            float q[K][K], r[K][K];
#pragma unroll
            for (int i = 0; i < K; ++i)
#pragma unroll
              for (int j = 0; j < K; ++j) q[i][j] = __shfl_down_sync(kFull, wm[i][j], d);
            if (lane + d < 32) {
              mat_mul<K>(wm, q, r);
#pragma unroll
              for (int i = 0; i < K; ++i)
#pragma unroll
                for (int j = 0; j < K; ++j) wm[i][j] = r[i][j];
              mat_rescale_if_small<K>(wm);
            }
          }
#pragma unroll
          for (int i = 0; i < K; ++i) {
            float sacc = wm[i][0];
#pragma unroll
            for (int j = 1; j < K; ++j) sacc += wm[i][j];
            float zn = __shfl_down_sync(kFull, sacc, 1);
            if (lane == 31) zn = 1.0f;
            s_z[lane * K + i] = zn;
          }
        }
      }
      group_sync(bar_id, T);  // B2: boundary vectors visible; staging buffers reusable
#pragma unroll
      for (int k = 0; k < K; ++k) {
        u[k] = s_u[gw * K + k];
        z[k] = s_z[gw * K + k];
      }
    } else {
      __syncwarp();
#pragma unroll
      for (int k = 0; k < K; ++k) {
        u[k] = tab.pi[k];
        z[k] = 1.0f;
      }
    }

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

    // ---------------- phase 3: local forward / backward sweeps (fully unrolled) ----------------
    double ll_thread = 0.0;
    if (active) {
      float al[K];
      int e = 0;
#pragma unroll
      for (int k = 0; k < K; ++k) al[k] = v[k];
      float* stg = s_stage + (size_t)t0 * K;
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        float nx[K];
        if (j == 0) {
          // first position of the read: alpha_0 = pi * b_0 (no transition)
#pragma unroll
          for (int jj = 0; jj < K; ++jj) {
            float s = al[0] * A[0][jj];
#pragma unroll
            for (int i = 1; i < K; ++i) s = fmaf(al[i], A[i][jj], s);
            s = first ? al[jj] : s;
            nx[jj] = (jj == 0) ? s : s * rt[0][jj - 1];
          }
        } else {
#pragma unroll
          for (int jj = 0; jj < K; ++jj) {
            float s = al[0] * A[0][jj];
#pragma unroll
            for (int i = 1; i < K; ++i) s = fmaf(al[i], A[i][jj], s);
            nx[jj] = (jj == 0) ? s : s * rt[j][jj - 1];
          }
        }
        if ((j & 3) == 3) {
          const float sc = want_ll ? pow2_rescale(vec_max<K>(nx), e) : pow2_rescale_noexp(vec_max<K>(nx));
#pragma unroll
          for (int k = 0; k < K; ++k) nx[k] *= sc;
        }
#pragma unroll
        for (int k = 0; k < K; ++k) al[k] = nx[k];
        if (K == 2) {
          *reinterpret_cast<float2*>(stg + j * 2) = make_float2(nx[0], nx[1]);
        } else if (K == 4) {
          *reinterpret_cast<float4*>(stg + j * 4) = make_float4(nx[0], nx[1], nx[2 % K], nx[3 % K]);
        } else {
#pragma unroll
          for (int k = 0; k < K; ++k) stg[j * K + k] = nx[k];
        }
      }
      if (want_ll) {
        float s = al[0];
#pragma unroll
        for (int k = 1; k < K; ++k) s += al[k];
        ll_thread = (double)logf(s) + (double)e * 0.6931471805599453094 +
                    (double)__popc(obsmask) * tab.l00[0] + (double)ll_sumo * tab.dl0[0];
        if (first) ll_thread += tab.log_pi_sum;
      }

      float bv[K];
#pragma unroll
      for (int k = 0; k < K; ++k) bv[k] = w[k];
      float alc[K];
#pragma unroll
      for (int k = 0; k < K; ++k) alc[k] = al[k];  // alpha of the chunk's last position
#pragma unroll
      for (int j = CH - 1; j >= 0; --j) {
        float gk[K];
        float s = 0.0f;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          gk[k] = alc[k] * bv[k];
          s += gk[k];
        }
        const float r = rcp_approx(s);
#pragma unroll
        for (int k = 0; k < K; ++k) gk[k] *= r;
        int best = 0;
        float gbest = gk[0];
#pragma unroll
        for (int k = 1; k < K; ++k)
          if (gk[k] > gbest) {
            gbest = gk[k];
            best = k;
          }
        // alpha of the previous position (this thread's chunk, or the incoming prefix vector)
        float alp[K];
        if (j > 0) {
          if (K == 2) {
            const float2 t2 = *reinterpret_cast<const float2*>(stg + (j - 1) * 2);
            alp[0] = t2.x;
            alp[1 % K] = t2.y;
          } else if (K == 4) {
            const float4 t4 = *reinterpret_cast<const float4*>(stg + (j - 1) * 4);
            alp[0] = t4.x;
            alp[1 % K] = t4.y;
            alp[2 % K] = t4.z;
            alp[3 % K] = t4.w;
          } else {
#pragma unroll
            for (int k = 0; k < K; ++k) alp[k] = stg[(j - 1) * K + k];
          }
        } else {
#pragma unroll
          for (int k = 0; k < K; ++k) alp[k] = v[k];
        }
        if (!STATS) {
          if (w_gamma_all) {
            if (K == 2) {
              *reinterpret_cast<float2*>(stg + j * 2) = make_float2(gk[0], gk[1 % K]);
            } else if (K == 4) {
              *reinterpret_cast<float4*>(stg + j * 4) = make_float4(gk[0], gk[1 % K], gk[2 % K], gk[3 % K]);
            } else {
#pragma unroll
              for (int k = 0; k < K; ++k) stg[j * K + k] = gk[k];
            }
          } else if (w_gamma_one) {
            float gsel = gk[0];
#pragma unroll
            for (int k = 1; k < K; ++k)
              if (gs == k) gsel = gk[k];
            s_out1[t0 + j] = gsel;
          }
          if (w_states) s_st[t0 + j] = (int8_t)best;
        }

        if (STATS) {
          if (j == 0 && first) {
#pragma unroll
            for (int k = 0; k < K; ++k) acc_start[k] += gk[k];
          }
          if ((obsmask >> j) & 1u) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
              acc_num[k] = fmaf(gk[k], ov[j], acc_num[k]);
              acc_den[k] += gk[k];
            }
          }
        }

        const bool need_step = (j > 0) || STATS;
        if (need_step) {
          float cb[K], nb[K];
          cb[0] = bv[0];
#pragma unroll
          for (int k = 1; k < K; ++k) cb[k] = rt[j][k - 1] * bv[k];
          if (STATS) {
            float prod[K][K];
#pragma unroll
            for (int i = 0; i < K; ++i) {
              float sN = 0.0f;
#pragma unroll
              for (int jj = 0; jj < K; ++jj) {
                prod[i][jj] = A[i][jj] * cb[jj];
                sN += prod[i][jj];
              }
              nb[i] = sN;
            }
            // xi_{t-1}(i,jj), t = t0+j > 0, both endpoints observed (HMM.py:1583-1591)
            const bool prev_obs = (j > 0) ? (((obsmask >> (j > 0 ? j - 1 : 0)) & 1u) != 0u) : prev_obs0;
            const bool valid = ((obsmask >> j) & 1u) && prev_obs;
            float s2 = 0.0f;
#pragma unroll
            for (int i = 0; i < K; ++i) s2 = fmaf(alp[i], nb[i], s2);
            const float r2 = valid ? rcp_approx(s2) : 0.0f;
#pragma unroll
            for (int i = 0; i < K; ++i) {
              const float ai = alp[i] * r2;
#pragma unroll
              for (int jj = 0; jj < K; ++jj) acc_xi[i * K + jj] = fmaf(ai, prod[i][jj], acc_xi[i * K + jj]);
            }
          } else {
#pragma unroll
            for (int i = 0; i < K; ++i) {
              float sN = A[i][0] * cb[0];
#pragma unroll
              for (int jj = 1; jj < K; ++jj) sN = fmaf(A[i][jj], cb[jj], sN);
              nb[i] = sN;
            }
          }
          if ((((CH - 1) - j) & 3) == 3) {
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

    if (STATS) acc_ll += ll_thread;
    if (a.loglik != nullptr) {
      const double wsum = warp_sum_f64(ll_thread);
      if (lane == 0) s_ll[gw] = wsum;
    }
    if (!STATS) fence_proxy_async();  // make the staging writes visible to the bulk-copy engine
    group_sync(bar_id, T);            // B3: staging complete

    // ---------------- store: bulk async shared -> global ----------------
    if (gt == 0) {
      if (a.loglik != nullptr) {
        double s = 0.0;
        for (int i = 0; i < G; ++i) s += s_ll[i];
        a.loglik[n] = s;
      }
      if (!STATS) {
        if (w_gamma_all) store_row_bulk<float>(a.gamma + (size_t)n * L * K, s_stage, (uint32_t)L * K * 4u);
        if (w_gamma_one) store_row_bulk<float>(a.gamma + (size_t)n * L, s_out1, (uint32_t)L * 4u);
        if (w_states) store_row_bulk<int8_t>(a.states + (size_t)n * L, s_st, (uint32_t)L);
        bulk_commit();
      }
    }
    if (STATS && ((iter + 1) % a.flush_every) == 0) flush_stats();
  }

  if (gt == 0) bulk_wait_all();  // all bulk stores complete before the CTA retires
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
