// Sequential (one thread per read) Baum-Welch E-step for large batches: three recursions per position.
//
// The chunk-parallel E-step kernels (fb2_kernel<STATS>, walk_kernel + estep_sweep_kernel) buy their
// intra-read parallelism with work: the two-kernel path runs FOUR recursions per position (forward and
// backward boundary walks, then a forward and a backward sweep inside every chunk) plus K^2-wide boundary
// traffic, 294 thread-instructions per position at K = 4.  With >= ~10^5 reads in a launch the reads
// alone fill the machine, and the classic checkpointed forward-backward needs only THREE:
//
//   seq_forward_kernel   one thread walks its read forward (K(K-1) FMAs per position, scaled fp32, the
//                        column-normalised tables of fb2_kernel), stores alpha at every 16th position
//                        (checkpoints, [chunk][read][K]: coalesced across the reads of a warp) and the
//                        read's log-likelihood;
//   seq_backward_kernel  the same thread layout walks the read backward chunk by chunk: alpha inside the
//                        chunk is recomputed from its checkpoint into a shared-memory stack
//                        ([position][thread] float4: conflict-free), then ONE backward recursion carries
//                        beta through the whole read while gamma / xi statistics accumulate in fp32
//                        registers (flushed every few chunks into per-warp fp64 accumulators; per-CTA
//                        partials are reduced in a fixed order by reduce_partials_kernel: deterministic).
//
// Observations are read straight from global memory, 16 positions per thread and load (uint4 for the
// uint8 code, 4 x float4 for floats): every sector fetched is fully used, nothing is staged or transposed.
// For the uint8 code (0 / 1 / missing) the per-state emission ratios come from a 3-row shared-memory table
// (one LDS.128) instead of K - 1 ex2 evaluations.
// HBM traffic per position: 2 observation reads + 2 * 4K / 16 bytes of checkpoints (K = 4, uint8: 4 B).
#pragma once
#include "fb2_kernel.cuh"

namespace smf {

constexpr int kSeqCH = 16;        // positions per chunk (checkpoint distance)
constexpr int kSeqThreads = 128;  // reads per CTA

struct SeqArgs {
  const void* obs;
  long long N;
  int L;
  int Tact;          // chunks per read = ceil(L / 16)
  float* ckpt;       // [Tact][N][KP] alpha entering chunk c (slot 0 unused); KP = 4 for K >= 3, 2 for K = 2
  double* ll;        // [N] per-read log-likelihood (written by the forward kernel, summed by the backward one)
  double* partials;  // [gridDim.x][S]
  int S;
  int flush_chunks;  // fp32 register statistics are flushed every this many chunks
};

template <int K>
struct SeqVec {
  static constexpr int KP = (K == 2) ? 2 : 4;
};

// ---- 16 observations of one read in registers -------------------------------------------------------
template <typename ObsT>
struct SeqChunk;

template <>
struct SeqChunk<unsigned char> {
  uint4 w;
  __device__ __forceinline__ void load(const unsigned char* p, int len) {
    (void)len;  // rows of the uint8 code are whole 16-byte groups (L % 16 == 0)
    w = __ldg(reinterpret_cast<const uint4*>(p));
  }
  __device__ __forceinline__ unsigned code(int j) const {
    const unsigned x = (j < 4) ? w.x : (j < 8 ? w.y : (j < 12 ? w.z : w.w));
    return (x >> (8 * (j & 3))) & 0xffu;
  }
  __device__ __forceinline__ bool observed(int j) const { return code(j) < 2u; }
};

template <>
struct SeqChunk<float> {
  float4 q[4];
  __device__ __forceinline__ void load(const float* p, int len) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      if (4 * g < len)
        q[g] = __ldg(reinterpret_cast<const float4*>(p) + g);
      else
        q[g] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  __device__ __forceinline__ float value(int j) const {
    const float4 v = q[j >> 2];
    return (j & 3) == 0 ? v.x : ((j & 3) == 1 ? v.y : ((j & 3) == 2 ? v.z : v.w));
  }
  __device__ __forceinline__ bool observed(int j) const {
    const float o = value(j);
    return o == o;
  }
};

// Emission ratios c_k = s_k b_k(o) / (s_0 b_0(o)) (c_0 = 1), "observed" flag and the observed value.
template <int K, typename ObsT>
struct SeqEmit;

template <int K>
struct SeqEmit<K, unsigned char> {
  const float4* tab4;  // shared: rows for code 0, 1, missing
  __device__ __forceinline__ void get(const SeqChunk<unsigned char>& ch, int j, float (&c)[K], bool& m, float& om) const {
    const unsigned code = ch.code(j);
    const float4 r = tab4[min(code, 2u)];
    c[0] = 1.0f;
    c[1] = r.y;
    if (K > 2) c[2 % K] = r.z;
    if (K > 3) c[3 % K] = r.w;
    m = code < 2u;
    om = code == 1u ? 1.0f : 0.0f;
  }
};

template <int K>
struct SeqEmit<K, float> {
  float rdl[K], rl0p[K], rmiss[K];
  __device__ __forceinline__ void get(const SeqChunk<float>& ch, int j, float (&c)[K], bool& m, float& om) const {
    const float o = ch.value(j);
    m = (o == o);
    om = m ? o : 0.0f;
    c[0] = 1.0f;
#pragma unroll
    for (int k = 1; k < K; ++k) c[k] = ex2_approx(m ? fmaf(o, rdl[k], rl0p[k]) : rmiss[k]);
  }
};

template <int K, typename ObsT>
__device__ __forceinline__ void seq_emit_init(const FbTables<K>& tab, SeqEmit<K, ObsT>& em, float4* s_tab) {
  if constexpr (sizeof(ObsT) == 1) {
    if (threadIdx.x < 3) {
      float r[4] = {1.0f, 1.0f, 1.0f, 1.0f};
      for (int k = 1; k < K; ++k) {
        const float x = threadIdx.x == 0 ? tab.rl0p[k] : (threadIdx.x == 1 ? tab.rl0p[k] + tab.rdl[k][0] : tab.rmiss[k]);
        r[k] = exp2f(x);
      }
      s_tab[threadIdx.x] = make_float4(r[0], r[1], r[2], r[3]);
    }
    em.tab4 = s_tab;
    __syncthreads();
  } else {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      em.rdl[k] = tab.rdl[k][0];
      em.rl0p[k] = tab.rl0p[k];
      em.rmiss[k] = tab.rmiss[k];
    }
  }
}

template <int K>
__device__ __forceinline__ void seq_store_vec(float* p, const float (&y)[K]) {
  if (K == 2) {
    *reinterpret_cast<float2*>(p) = make_float2(y[0], y[1 % K]);
  } else {
    *reinterpret_cast<float4*>(p) = make_float4(y[0], y[1 % K], y[2 % K], K > 3 ? y[3 % K] : 0.0f);
  }
}
template <int K>
__device__ __forceinline__ void seq_load_vec(const float* p, float (&y)[K]) {
  if (K == 2) {
    const float2 t = *reinterpret_cast<const float2*>(p);
    y[0] = t.x;
    y[1 % K] = t.y;
  } else {
    const float4 t = *reinterpret_cast<const float4*>(p);
    y[0] = t.x;
    y[1 % K] = t.y;
    y[2 % K] = t.z;
    if (K > 3) y[3 % K] = t.w;
  }
}

// one forward step: al <- (al A') .* c   (no transition into the very first position of the read)
template <int K>
__device__ __forceinline__ void seq_fwd_step(float (&al)[K], const float (&A)[K][K], const float (&c)[K], bool no_trans) {
  float y[K];
#pragma unroll
  for (int jj = 0; jj < K; ++jj) {
    float s = al[jj];
#pragma unroll
    for (int i = 0; i < K; ++i)
      if (i != jj) s = fmaf(al[i], A[i][jj], s);
    y[jj] = no_trans ? al[jj] : s;
  }
#pragma unroll
  for (int k = 1; k < K; ++k) y[k] *= c[k];
#pragma unroll
  for (int k = 0; k < K; ++k) al[k] = y[k];
}

// -----------------------------------------------------------------------------------------------------
// forward: checkpoints + log-likelihood
// -----------------------------------------------------------------------------------------------------
template <int K, typename ObsT>
__global__ void __launch_bounds__(kSeqThreads) seq_forward_kernel(const __grid_constant__ FbTables<K> tab,
                                                                  const __grid_constant__ SeqArgs a) {
  constexpr int KP = SeqVec<K>::KP;
  __shared__ float4 s_tab[3];
  SeqEmit<K, ObsT> em;
  seq_emit_init<K, ObsT>(tab, em, s_tab);
  float A[K][K];
#pragma unroll
  for (int i = 0; i < K; ++i)
#pragma unroll
    for (int j = 0; j < K; ++j) A[i][j] = tab.Ap[i][j];
  const int L = a.L;
  const size_t cstride = (size_t)a.N * KP;

  for (long long n = (long long)blockIdx.x * kSeqThreads + threadIdx.x; n < a.N; n += (long long)gridDim.x * kSeqThreads) {
    const ObsT* row = reinterpret_cast<const ObsT*>(a.obs) + n * (long long)L;
    float al[K];
    {
      float s = 0.0f;
#pragma unroll
      for (int k = 0; k < K; ++k) s += tab.pi2[k];
      const float r = 1.0f / s;
#pragma unroll
      for (int k = 0; k < K; ++k) al[k] = tab.pi2[k] * r;
    }
    int e = 0, cnt = 0;
    double sumo = 0.0;
    float* dst = a.ckpt + cstride + (size_t)n * KP;
    SeqChunk<ObsT> cur, nxt;
    cur.load(row, min(kSeqCH, L));
    for (int c = 0; c < a.Tact; ++c) {
      const int t0 = c * kSeqCH;
      const int len = min(kSeqCH, L - t0);
      if (c + 1 < a.Tact) nxt.load(row + t0 + kSeqCH, min(kSeqCH, L - t0 - kSeqCH));
      if (c > 0) {
        seq_store_vec<K>(dst, al);
        dst += cstride;
      }
      float tsum = 0.0f;
#pragma unroll
      for (int j = 0; j < kSeqCH; ++j) {
        if (j < len) {
          float cf[K];
          bool m;
          float om;
          em.get(cur, j, cf, m, om);
          cnt += m ? 1 : 0;
          tsum += om;
          seq_fwd_step<K>(al, A, cf, c == 0 && j == 0);
          if ((j & 3) == 3) {
            const float sc = pow2_rescale(vec_max<K>(al), e);
#pragma unroll
            for (int k = 0; k < K; ++k) al[k] *= sc;
          }
        }
      }
      sumo += (double)tsum;
      cur = nxt;
    }
    float s = al[0];
#pragma unroll
    for (int k = 1; k < K; ++k) s += al[k];
    a.ll[n] = (double)logf(s) + (double)e * 0.6931471805599453094 + tab.log_pi2_sum + (double)cnt * tab.l00[0] +
              sumo * tab.dl0[0] + (double)(L - 1) * tab.log_s0;
  }
}

// -----------------------------------------------------------------------------------------------------
// backward: alpha recomputed per chunk, one beta recursion, statistics
// -----------------------------------------------------------------------------------------------------
template <int K, typename ObsT>
__global__ void __launch_bounds__(kSeqThreads, (K == 4 ? 4 : 5)) seq_backward_kernel(const __grid_constant__ FbTables<K> tab,
                                                                                      const __grid_constant__ SeqArgs a) {
  constexpr int KP = SeqVec<K>::KP;
  constexpr int KK = K * K;
  extern __shared__ __align__(16) unsigned char smem[];
  // [warps][S] fp64 accumulators | [16][128][KP] alpha stack | emission table
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kWarps = kSeqThreads / 32;
  double* wacc = reinterpret_cast<double*>(smem) + (size_t)warp * a.S;
  float* stack = reinterpret_cast<float*>(smem + (((size_t)kWarps * a.S * 8 + 15) & ~(size_t)15));
  float4* s_tab = reinterpret_cast<float4*>(stack + (size_t)kSeqCH * kSeqThreads * KP);
  for (int i = lane; i < a.S; i += 32) wacc[i] = 0.0;
  SeqEmit<K, ObsT> em;
  seq_emit_init<K, ObsT>(tab, em, s_tab);  // contains a __syncthreads for the table
  float* my = stack + (size_t)threadIdx.x * KP;  // + j * kSeqThreads * KP
  constexpr int jstride = kSeqThreads * KP;

  float A[K][K];
#pragma unroll
  for (int i = 0; i < K; ++i)
#pragma unroll
    for (int j = 0; j < K; ++j) A[i][j] = tab.Ap[i][j];

  float acc_xi[KK], acc_num[K], acc_den[K], acc_start[K];
  double acc_ll = 0.0;
#pragma unroll
  for (int i = 0; i < KK; ++i) acc_xi[i] = 0.0f;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    acc_num[k] = 0.0f;
    acc_den[k] = 0.0f;
    acc_start[k] = 0.0f;
  }
  auto flush = [&]() {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const float r0 = warp_sum_f32(acc_start[k]);
      const float r1 = warp_sum_f32(acc_num[k]);
      const float r2 = warp_sum_f32(acc_den[k]);
      if (lane == 0) {
        wacc[k] += (double)r0;
        wacc[K + KK + k] += (double)r1;
        wacc[K + KK + K + k] += (double)r2;
      }
      acc_start[k] = 0.0f;
      acc_num[k] = 0.0f;
      acc_den[k] = 0.0f;
    }
#pragma unroll
    for (int i = 0; i < KK; ++i) {
      const float r = warp_sum_f32(acc_xi[i]);
      // the constant factor Ap[i][j] of xi is applied here, once per flush
      if (lane == 0) wacc[K + i] += (double)r * (double)tab.Ap[i / K][i % K];
      acc_xi[i] = 0.0f;
    }
  };

  const int L = a.L;
  const size_t cstride = (size_t)a.N * KP;
  const long long nblk = (a.N + kSeqThreads - 1) / kSeqThreads;

  for (long long blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
    const long long n = blk * kSeqThreads + threadIdx.x;
    const bool live = n < a.N;
    const long long nn = live ? n : a.N - 1;  // idle threads shadow the last read (results discarded)
    const ObsT* row = reinterpret_cast<const ObsT*>(a.obs) + nn * (long long)L;
    float bv[K];
#pragma unroll
    for (int k = 0; k < K; ++k) bv[k] = 1.0f / K;
    SeqChunk<ObsT> cur, prv;
    {
      const int t0 = (a.Tact - 1) * kSeqCH;
      cur.load(row + t0, L - t0);
    }
    int since_flush = 0;
    for (int c = a.Tact - 1; c >= 0; --c) {
      const int t0 = c * kSeqCH;
      const int len = min(kSeqCH, L - t0);
      const bool first = (c == 0);
      float v[K];
      if (!first) {
        prv.load(row + t0 - kSeqCH, kSeqCH);
        seq_load_vec<K>(a.ckpt + (size_t)c * cstride + (size_t)nn * KP, v);
      } else {
#pragma unroll
        for (int k = 0; k < K; ++k) v[k] = tab.pi2[k];
      }
      {
        float sv = v[0];
#pragma unroll
        for (int k = 1; k < K; ++k) sv += v[k];
        const float rv = rcp_approx(sv);
#pragma unroll
        for (int k = 0; k < K; ++k) v[k] *= rv;
      }
      const bool prev_obs0 = first ? false : prv.observed(kSeqCH - 1);

      // ---- alpha of the chunk, recomputed from its checkpoint ----
      float al[K];
#pragma unroll
      for (int k = 0; k < K; ++k) al[k] = v[k];
#pragma unroll
      for (int j = 0; j < kSeqCH; ++j) {
        if (j < len) {
          float cf[K];
          bool m;
          float om;
          em.get(cur, j, cf, m, om);
          seq_fwd_step<K>(al, A, cf, first && j == 0);
          if ((j & 3) == 3) {
            const float sc = pow2_rescale_noexp(vec_max<K>(al));
#pragma unroll
            for (int k = 0; k < K; ++k) al[k] *= sc;
          }
          seq_store_vec<K>(my + j * jstride, al);
        }
      }

      // ---- backward over the chunk: gamma, xi, statistics ----
      float alc[K];
#pragma unroll
      for (int k = 0; k < K; ++k) alc[k] = al[k];
#pragma unroll
      for (int j = kSeqCH - 1; j >= 0; --j) {
        if (j < len) {
          float cr[K];
          bool m;
          float om;
          em.get(cur, j, cr, m, om);
          float gk[K];
          float gs = 0.0f;
#pragma unroll
          for (int k = 0; k < K; ++k) {
            gk[k] = alc[k] * bv[k];
            gs += gk[k];
          }
          const float g0 = rcp_approx(gs);
          const float gr = m ? g0 : 0.0f;
#pragma unroll
          for (int k = 0; k < K; ++k) {
            const float g = gk[k] * gr;
            acc_num[k] = fmaf(g, om, acc_num[k]);
            acc_den[k] += g;
          }
          float alp[K];
          if (j > 0) {
            seq_load_vec<K>(my + (j - 1) * jstride, alp);
          } else {
#pragma unroll
            for (int k = 0; k < K; ++k) alp[k] = v[k];
          }
          const bool at_start = first && j == 0;
          if (at_start) {
            // gamma_0 of the read (start statistics, counted whether or not position 0 is observed)
#pragma unroll
            for (int k = 0; k < K; ++k) acc_start[k] += gk[k] * g0;
          }
          float cb[K], nb[K];
          cb[0] = bv[0];
#pragma unroll
          for (int k = 1; k < K; ++k) cb[k] = cr[k] * bv[k];
#pragma unroll
          for (int i = 0; i < K; ++i) {
            float acc = cb[i];
#pragma unroll
            for (int jj = 0; jj < K; ++jj)
              if (jj != i) acc = fmaf(A[i][jj], cb[jj], acc);
            nb[i] = acc;
          }
          const bool prev_obs = (j > 0) ? cur.observed(j > 0 ? j - 1 : 0) : prev_obs0;
          const bool valid = m && prev_obs && !at_start;
          float s2 = 0.0f;
#pragma unroll
          for (int i = 0; i < K; ++i) s2 = fmaf(alp[i], nb[i], s2);
          const float r2 = valid ? rcp_approx(s2) : 0.0f;
#pragma unroll
          for (int i = 0; i < K; ++i) {
            const float ai = alp[i] * r2;
#pragma unroll
            for (int jj = 0; jj < K; ++jj) acc_xi[i * K + jj] = fmaf(ai, cb[jj], acc_xi[i * K + jj]);
          }
          if ((j & 3) == 0) {
            const float sc = pow2_rescale_noexp(vec_max<K>(nb));
#pragma unroll
            for (int k = 0; k < K; ++k) nb[k] *= sc;
          }
#pragma unroll
          for (int k = 0; k < K; ++k) {
            bv[k] = nb[k];
            alc[k] = alp[k];
          }
        }
      }
      cur = prv;
      if (++since_flush == a.flush_chunks) {
        if (!live) {
#pragma unroll
          for (int i = 0; i < KK; ++i) acc_xi[i] = 0.0f;
#pragma unroll
          for (int k = 0; k < K; ++k) acc_num[k] = acc_den[k] = acc_start[k] = 0.0f;
        }
        flush();
        since_flush = 0;
      }
    }
    if (!live) {
#pragma unroll
      for (int i = 0; i < KK; ++i) acc_xi[i] = 0.0f;
#pragma unroll
      for (int k = 0; k < K; ++k) acc_num[k] = acc_den[k] = acc_start[k] = 0.0f;
    }
    flush();
    if (live) acc_ll += a.ll[n];
  }
  // per-warp log-likelihood, then the CTA's partial vector
  {
    double v = acc_ll;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
    if (lane == 0) wacc[a.S - 1] += v;
  }
  __syncthreads();
  double* all = reinterpret_cast<double*>(smem);
  for (int sidx = threadIdx.x; sidx < a.S; sidx += blockDim.x) {
    double tot = 0.0;
    for (int wv = 0; wv < kWarps; ++wv) tot += all[(size_t)wv * a.S + sidx];
    a.partials[(size_t)blockIdx.x * a.S + sidx] = tot;
  }
}

}  // namespace smf
