// E-step sweeps of the two-kernel path (K >= 3): Baum-Welch statistics from the chunk boundary
// vectors walk_kernel left in HBM.
//
// With the boundary vectors known, the chunks of a read are independent, so nothing here is
// CTA-wide: the unit of work is a SLICE -- 32 consecutive chunks of one read, one chunk per lane --
// and every warp runs its own pipeline over the slices it owns:
//   * the slice's observations (32*CH elements) arrive with one bulk asynchronous copy into the
//     warp's private double buffer (mbarrier completion); the copy of the next slice is issued
//     before the current one is processed;
//   * each lane sweeps its chunk forward, stacking the scaled alpha vectors in a private
//     shared-memory column, then backward, forming gamma / xi on the fly and accumulating the
//     sufficient statistics in registers (fp32), flushed every few slices into per-warp fp64
//     accumulators; per-CTA partials are reduced in a fixed order by reduce_partials_kernel.
// No CTA barrier, no padding, no limit on the read length; the position loops are rolled, so the
// chunk length is a run-time (odd) number and the register footprint stays below the spill line
// that the fully unrolled fb2_kernel E-step variants cross for K = 4.
#pragma once
#include "fb2_kernel.cuh"

namespace smf {

constexpr int kSweepWarps = 8;  // warps per CTA

struct SweepArgs {
  const void* obs;
  long long N;
  int L;
  int CH;        // chunk length (odd)
  int Tact;      // chunks per read
  int slices;    // slices per read = ceil(Tact / 32)
  const float* bnd_v;  // [Tact][N][K]
  const float* bnd_w;
  const double* ll_in;  // [N]
  double* partials;     // [gridDim.x][S]
  int S;
  int flush_every;
  int in_bytes;     // bytes of one observation buffer (16-byte multiple)
  int warp_bytes;   // bytes of one warp's region: 16 (two mbarriers) + 2 * in_bytes + stack
  int off_warp0;    // first warp region (after the CTA's fp64 accumulators)
};

template <int K, typename ObsT>
__global__ void __launch_bounds__(kSweepWarps * 32, 2) estep_sweep_kernel(const __grid_constant__ FbTables<K> tab,
                                                                         const __grid_constant__ SweepArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int KK = K * K;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int L = a.L;
  const int CH = a.CH;

  double* wacc = reinterpret_cast<double*>(smem) + (size_t)warp * a.S;
  unsigned char* wbase = smem + a.off_warp0 + (size_t)warp * a.warp_bytes;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(wbase);  // [2]
  unsigned char* in_raw = wbase + 16;
  float* stack = reinterpret_cast<float*>(wbase + 16 + 2 * (size_t)a.in_bytes) + (size_t)lane * CH * K;

  for (int i = lane; i < a.S; i += 32) wacc[i] = 0.0;
  if (lane == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();

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
  auto ratios = [&](float o, float (&c)[K]) {
    const bool m = (o == o);
#pragma unroll
    for (int k = 1; k < K; ++k) c[k] = ex2_approx(m ? fmaf(o, tab.rdl[k][0], tab.rl0p[k]) : tab.rmiss[k]);
    return m;
  };

  const long long total = a.N * (long long)a.slices;
  const long long wstride = (long long)gridDim.x * kSweepWarps;
  const int slice_pos = 32 * CH;
  const char* obs_base = static_cast<const char*>(a.obs);

  auto issue_load = [&](long long item, int buf) {
    const long long n = item / a.slices;
    const int s = (int)(item - n * a.slices);
    const int p0 = s * slice_pos;
    const int npos = min(slice_pos, L - p0);
    const char* src = obs_base + ((size_t)n * L + p0) * sizeof(ObsT);
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15u);
    const uint32_t bytes = (mis + (uint32_t)npos * (uint32_t)sizeof(ObsT) + 15u) & ~15u;
    mbar_expect_tx(&mbar[buf], bytes);
    bulk_load(in_raw + (size_t)buf * a.in_bytes, src - mis, bytes, &mbar[buf]);
  };

  long long item = (long long)blockIdx.x * kSweepWarps + warp;
  if (lane == 0 && item < total) issue_load(item, 0);

  int it = 0;
  for (; item < total; item += wstride, ++it) {
    const int buf = it & 1;
    if (lane == 0 && item + wstride < total) issue_load(item + wstride, buf ^ 1);

    const long long n = item / a.slices;
    const int s = (int)(item - n * a.slices);
    const int p0 = s * slice_pos;
    const int c = s * 32 + lane;
    const int t0 = c * CH;
    const bool active = c < a.Tact;
    const bool first = (c == 0);
    const int len = active ? min(CH, L - t0) : 0;

    // boundary vectors (issued before waiting for the observations)
    float v[K], w[K];
    {
      const bool has_v = active && !first;
      const bool has_w = active && c < a.Tact - 1;
      const size_t boff = ((size_t)c * (size_t)a.N + (size_t)n) * K;
      float sv = 0.0f, sw = 0.0f;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        v[k] = has_v ? __ldg(a.bnd_v + boff + k) : tab.pi2[k];
        w[k] = has_w ? __ldg(a.bnd_w + boff + k) : 1.0f;
        sv += v[k];
        sw += w[k];
      }
      const float rv = rcp_approx(sv), rw = rcp_approx(sw);
#pragma unroll
      for (int k = 0; k < K; ++k) {
        v[k] *= rv;
        w[k] *= rw;
      }
    }
    // observation just before the chunk (xi needs both endpoints observed, HMM.py:1583-1591)
    const char* row = obs_base + (size_t)n * L * sizeof(ObsT);
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(row + (size_t)p0 * sizeof(ObsT)) & 15u);
    const ObsT* s_in =
        reinterpret_cast<const ObsT*>(in_raw + (size_t)buf * a.in_bytes + mis) + (size_t)lane * CH;
    bool prev_obs0 = false;
    if (active && !first && lane == 0) {
      const float op = obs_to_float<ObsT>(__ldg(reinterpret_cast<const ObsT*>(row) + t0 - 1));
      prev_obs0 = (op == op);
    }

    mbar_wait(&mbar[buf], (uint32_t)((it >> 1) & 1));

    if (active && !first && lane != 0) {
      const float op = obs_to_float<ObsT>(s_in[-1]);
      prev_obs0 = (op == op);
    }

    // ---------------- forward sweep: alpha column ----------------
    float al[K];
#pragma unroll
    for (int k = 0; k < K; ++k) al[k] = v[k];
#pragma unroll 2
    for (int j = 0; j < len; ++j) {
      const float o = obs_to_float<ObsT>(s_in[j]);
      float cf[K];
      ratios(o, cf);
      float y[K];
#pragma unroll
      for (int jj = 0; jj < K; ++jj) {
        float acc = al[jj];
#pragma unroll
        for (int i = 0; i < K; ++i)
          if (i != jj) acc = fmaf(al[i], A[i][jj], acc);
        y[jj] = (first && j == 0) ? al[jj] : acc;  // alpha_0 = pi2 * c: no transition
      }
#pragma unroll
      for (int k = 1; k < K; ++k) y[k] *= cf[k];
      if ((j & 3) == 3) {
        const float sc = pow2_rescale_noexp(vec_max<K>(y));
#pragma unroll
        for (int k = 0; k < K; ++k) y[k] *= sc;
      }
#pragma unroll
      for (int k = 0; k < K; ++k) al[k] = y[k];
      if (K == 4) {
        *reinterpret_cast<float4*>(stack + j * 4) = make_float4(y[0], y[1 % K], y[2 % K], y[3 % K]);
      } else if (K == 2) {
        *reinterpret_cast<float2*>(stack + j * 2) = make_float2(y[0], y[1 % K]);
      } else {
#pragma unroll
        for (int k = 0; k < K; ++k) stack[j * K + k] = y[k];
      }
    }

    // ---------------- backward sweep: gamma, xi, statistics ----------------
    float bv[K], alc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      bv[k] = w[k];
      alc[k] = al[k];
    }
#pragma unroll 2
    for (int j = len - 1; j >= 0; --j) {
      // gamma only feeds the emission statistics here: fold the "observed" mask into its normaliser
      const float o = obs_to_float<ObsT>(s_in[j]);
      const bool m = (o == o);
      const float om = m ? o : 0.0f;
      float gk[K];
      float gs = 0.0f;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        gk[k] = alc[k] * bv[k];
        gs += gk[k];
      }
      const float gr = m ? rcp_approx(gs) : 0.0f;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const float g = gk[k] * gr;
        acc_num[k] = fmaf(g, om, acc_num[k]);
        acc_den[k] += g;
      }
      float alp[K];
      if (j > 0) {
        if (K == 4) {
          const float4 t4 = *reinterpret_cast<const float4*>(stack + (j - 1) * 4);
          alp[0] = t4.x;
          alp[1 % K] = t4.y;
          alp[2 % K] = t4.z;
          alp[3 % K] = t4.w;
        } else if (K == 2) {
          const float2 t2 = *reinterpret_cast<const float2*>(stack + (j - 1) * 2);
          alp[0] = t2.x;
          alp[1 % K] = t2.y;
        } else {
#pragma unroll
          for (int k = 0; k < K; ++k) alp[k] = stack[(j - 1) * K + k];
        }
      } else {
#pragma unroll
        for (int k = 0; k < K; ++k) alp[k] = v[k];
      }
      const bool at_start = first && j == 0;
      if (at_start) {
        // gamma_0 of the read (start statistics, counted whether or not position 0 is observed)
        const float g0 = rcp_approx(gs);
#pragma unroll
        for (int k = 0; k < K; ++k) acc_start[k] += gk[k] * g0;
      }
      float cr[K], cb[K], nb[K];
      ratios(o, cr);
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
      bool prev_obs = prev_obs0;
      if (j > 0) {
        const float op = obs_to_float<ObsT>(s_in[j - 1]);
        prev_obs = (op == op);
      }
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
      if (((len - 1 - j) & 3) == 3) {
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

    if (s == 0 && lane == 0) acc_ll += a.ll_in[n];
    if (((it + 1) % a.flush_every) == 0) flush();
    __syncwarp();  // every lane is done with this observation buffer before it is refilled
  }

  flush();
  if (lane == 0) wacc[a.S - 1] += acc_ll;
  __syncthreads();
  double* all = reinterpret_cast<double*>(smem);
  for (int sidx = threadIdx.x; sidx < a.S; sidx += blockDim.x) {
    double tot = 0.0;
    for (int wv = 0; wv < kSweepWarps; ++wv) tot += all[(size_t)wv * a.S + sidx];
    a.partials[(size_t)blockIdx.x * a.S + sidx] = tot;
  }
}

}  // namespace smf
