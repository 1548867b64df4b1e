// Posterior decode of SHORT reads: several reads per warp.
//
// fb2_kernel gives every read whole warps; a read of 170 positions (the reference's amplicon case)
// then occupies 10 of 32 lanes.  Here a warp owns a BUNDLE of R = floor(32 / Tg) consecutive reads,
// read r on lanes [r*Tg, (r+1)*Tg) with Tg = ceil(L / CH) chunks of CH positions, and everything is
// warp-local (no CTA barrier):
//   * consecutive reads are contiguous in HBM, so a bundle's observations arrive with ONE bulk
//     asynchronous copy into the warp's private double buffer (mbarrier completion, next bundle
//     prefetched), and its posteriors / states leave with ONE bulk store each from a staging tile laid
//     out exactly like the global rows;
//   * phase 1 (chunk transfer-matrix product), phase 2 (Kogge-Stone prefix + suffix scans, SEGMENTED:
//     a lane only combines with lanes of its own read) and phase 3 (forward sweep stacking alpha in
//     a private shared-memory column, backward sweep writing gamma / arg-max to the staging tile) use
//     the same scaled fp32 arithmetic and column-normalised tables as fb2_kernel;
//   * position loops are rolled: the chunk length is a run-time odd number chosen by the planner to
//     keep the lanes busy (e.g. L = 170: CH = 17, Tg = 10, R = 3 -> 94 % of the lanes), and the ragged
//     last chunk of a read simply runs fewer iterations (no padding, nothing to clobber).
#pragma once
#include "fb2_kernel.cuh"

namespace smf {

constexpr int kShortWarps = 4;

struct ShortArgs {
  const void* obs;
  long long N;
  int L;
  int CH;   // chunk length (odd)
  int Tg;   // lanes (chunks) per read
  int R;    // reads per warp
  float* gamma;     // [N,L,K] / [N,L] / nullptr
  int gamma_state;  // < 0: all states
  int8_t* states;   // [N,L] / nullptr
  double* loglik;   // [N] / nullptr
  double* partials; // STATS: [gridDim.x][S] per-CTA partial statistics
  int S;            // STATS: statistics length
  int flush_every;  // STATS: bundles between fp32 -> fp64 flushes
  int wacc_bytes;   // STATS: CTA-wide fp64 accumulators [warps][S] in front of the warp regions
  int in_bytes;     // one observation buffer
  int stack_off;    // byte offsets inside a warp's region
  int gam_off, st_off;
  int warp_bytes;
};

// CHT > 0: the chunk length is the compile-time constant CHT -- the position loops are fully unrolled
// (no loop / address / predicate arithmetic, emission ratios cached in registers) and a lane simply
// skips the copies beyond its chunk; CHT == 0: run-time chunk length a.CH with rolled, software-
// pipelined loops (any odd length, any observation type).
// STATS = true (compile-time chunk only): Baum-Welch E-step -- no outputs; the backward sweep
// accumulates start / transition / emission statistics and the log-likelihood exactly like
// fb2_kernel<STATS> (fp32 registers -> per-warp fp64 shared accumulators -> per-CTA partials).
template <int K, typename ObsT, int CHT, bool STATS = false>
__global__ void __launch_bounds__(kShortWarps * 32, 4) short_kernel(const __grid_constant__ FbTables<K> tab,
                                                                   const __grid_constant__ ShortArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr unsigned kFull = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int L = a.L, CH = (CHT > 0) ? CHT : a.CH, Tg = a.Tg, R = a.R;
  constexpr int NRT = (CHT > 0) ? CHT : 1;
  const int seg = lane / Tg;         // read of the bundle this lane works for
  const int li = lane - seg * Tg;    // chunk index inside the read
  const int t0 = li * CH;
  const bool first = (li == 0);
  const bool last = (li == Tg - 1);
  const int len = (seg < R) ? max(0, min(CH, L - t0)) : 0;

  static_assert(!STATS || CHT > 0, "the E-step variant needs a compile-time chunk length");
  constexpr int KK = K * K;
  unsigned char* wbase = smem + (STATS ? a.wacc_bytes : 0) + (size_t)warp * a.warp_bytes;
  double* wacc = reinterpret_cast<double*>(smem) + (size_t)warp * (STATS ? a.S : 0);
  if (STATS) {
    for (int i = lane; i < a.S; i += 32) wacc[i] = 0.0;
  }
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
#pragma unroll
    for (int k = 0; k < NEM; ++k) {
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
    for (int i = 0; i < NXI; ++i) {
      const float r = warp_sum_f32(acc_xi[i]);
      // the constant factor Ap[i][j] of xi is applied here, once per flush
      if (lane == 0) wacc[K + i] += (double)r * (double)tab.Ap[(i / K) % K][i % K];
      acc_xi[i] = 0.0f;
    }
  };
  uint64_t* mbar = reinterpret_cast<uint64_t*>(wbase);  // [2]
  unsigned char* in_raw = wbase + 16;
  float* stack = reinterpret_cast<float*>(wbase + a.stack_off) + (size_t)lane * CH * K;
  float* s_gam_base = reinterpret_cast<float*>(wbase + a.gam_off);
  int8_t* s_st_base = reinterpret_cast<int8_t*>(wbase + a.st_off);

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
  // packed fp32x2 (f32x2.cuh) for K >= 3; K = 2 measured 2-5 % slower with it here (profiles/r02_policy_audit4.txt)
  auto vecA = [&](const float (&x)[K], float (&y)[K]) {
    if constexpr (K >= 3) {
      vec_mat_unit<K>(x, A, y);
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
  };
  auto tailAvec = [&](const float (&bvv)[K], const float (&c)[K], float (&cb)[K], float (&nb)[K]) {
    if constexpr (K >= 3) {
      vec_mul_tail<K>(cb, bvv, c);
      mat_vec_unit<K>(A, cb, nb);
    } else {
      cb[0] = bvv[0];
#pragma unroll
      for (int k = 1; k < K; ++k) cb[k] = c[k] * bvv[k];
#pragma unroll
      for (int i = 0; i < K; ++i) {
        float s = cb[i];
#pragma unroll
        for (int jj = 0; jj < K; ++jj)
          if (jj != i) s = fmaf(A[i][jj], cb[jj], s);
        nb[i] = s;
      }
    }
  };
  auto ratios = [&](float o, float (&c)[K]) {
    const bool m = (o == o);
    c[0] = 1.0f;
#pragma unroll
    for (int k = 1; k < K; ++k) c[k] = ex2_approx(m ? fmaf(o, tab.rdl[k][0], tab.rl0p[k]) : tab.rmiss[k]);
    return m;
  };

  const bool want_ll = STATS || (a.loglik != nullptr);
  const int gs = a.gamma_state;
  const bool w_gamma_all = !STATS && (a.gamma != nullptr && gs < 0);
  const bool w_gamma_one = !STATS && (a.gamma != nullptr && gs >= 0);
  const bool w_states = !STATS && (a.states != nullptr);
  const char* obs_base = static_cast<const char*>(a.obs);
  const long long nbundles = (a.N + R - 1) / R;
  const long long wstride = (long long)gridDim.x * kShortWarps;

  auto issue_load = [&](long long b, int buf) {
    const long long n0 = b * R;
    const int nr = (int)min((long long)R, a.N - n0);
    const char* src = obs_base + (size_t)n0 * L * sizeof(ObsT);
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15u);
    const uint32_t bytes = (mis + (uint32_t)(nr * L) * (uint32_t)sizeof(ObsT) + 15u) & ~15u;
    mbar_expect_tx(&mbar[buf], bytes);
    bulk_load(in_raw + (size_t)buf * a.in_bytes, src - mis, bytes, &mbar[buf]);
  };

  long long b = (long long)blockIdx.x * kShortWarps + warp;
  if (lane == 0 && b < nbundles) issue_load(b, 0);

  int it = 0;
  for (; b < nbundles; b += wstride, ++it) {
    const int buf = it & 1;
    if (lane == 0 && b + wstride < nbundles) issue_load(b + wstride, buf ^ 1);
    const long long n0 = b * R;
    const int nr = (int)min((long long)R, a.N - n0);
    const bool active = (seg < nr) && (len > 0);
    const int mylen = active ? len : 0;
    const uint32_t mis =
        (uint32_t)(reinterpret_cast<uintptr_t>(obs_base + (size_t)n0 * L * sizeof(ObsT)) & 15u);
    const ObsT* s_in = reinterpret_cast<const ObsT*>(in_raw + (size_t)buf * a.in_bytes + mis) + (size_t)seg * L + t0;

    mbar_wait(&mbar[buf], (uint32_t)((it >> 1) & 1));

    // ---------------- phase 1: chunk transfer matrix ----------------
    float P[K][K];
    mat_set_identity<K>(P);
    float ll_sumo = 0.0f;
    int ll_cnt = 0;
    float rt[NRT][K];  // cached emission ratios (compile-time chunk only)
    using mask_t = typename std::conditional<(CHT > 32), unsigned long long, unsigned>::type;
    mask_t obsmask = 0;      // bit j: position t0 + j observed (E-step)
    bool prev_obs0 = false;  // position just before the chunk observed (xi needs both ends, HMM.py:1583-1591)
    if (STATS && active && !first) {
      const float op = obs_to_float<ObsT>(s_in[-1]);  // same read, previous lane's last position
      prev_obs0 = (op == op);
    }
    if constexpr (CHT > 0) {
      // Positions past the end of the chunk (ragged last chunk of a read, idle lanes) are treated as
      // missing observations: they leave the posteriors of the real positions unchanged (beta of the
      // last real position stays uniform), so nothing but the loads and the output stores is predicated.
      const float kMissing = __int_as_float(0x7fc00000);
#pragma unroll
      for (int j = 0; j < CHT; ++j) {
        const float o = (j < mylen) ? obs_to_float<ObsT>(s_in[j]) : kMissing;
        float c[K];
        const bool m = ratios(o, c);
        if (STATS) obsmask |= m ? ((mask_t)1 << j) : (mask_t)0;
#pragma unroll
        for (int k = 0; k < K; ++k) rt[j][k] = c[k];
        if (want_ll && m) {
          ll_sumo += o;
          ++ll_cnt;
        }
#pragma unroll
        for (int i = 0; i < K; ++i) {
          float q[K];
          vecA(P[i], q);
#pragma unroll
          for (int jj = 0; jj < K; ++jj) {
            const float x = (j == 0 && first) ? P[i][jj] : q[jj];  // P <- diag(c) at the start of a read
            P[i][jj] = (jj == 0) ? x : x * c[jj];
          }
        }
        if ((j & 3) == 3 || j == CHT - 1) mat_scale<K>(P, pow2_rescale_noexp(mat_max<K>(P)));
      }
    } else {
    // The emission ratios of position j+1 are formed while position j is folded in: the loop-carried
      // dependency is then only the K(K-1) FMA chain, not load -> ex2 -> multiply.
      float cn[K];
      float on = (mylen > 0) ? obs_to_float<ObsT>(s_in[0]) : 0.0f;
      bool mn = ratios(on, cn);
      for (int j = 0; j < mylen; ++j) {
        float c[K];
#pragma unroll
        for (int k = 0; k < K; ++k) c[k] = cn[k];
        const float o = on;
        const bool m = mn;
        if (j + 1 < mylen) {
          on = obs_to_float<ObsT>(s_in[j + 1]);
          mn = ratios(on, cn);
        }
        if (want_ll && m) {
          ll_sumo += o;
          ++ll_cnt;
        }
        const bool no_trans = first && j == 0;  // P <- diag(c) for the first position of a read
#pragma unroll
        for (int i = 0; i < K; ++i) {
          float q[K];
          vecA(P[i], q);
#pragma unroll
          for (int jj = 0; jj < K; ++jj) {
            const float x = no_trans ? P[i][jj] : q[jj];
            P[i][jj] = (jj == 0) ? x : x * c[jj];
          }
        }
        if ((j & 3) == 3 || j == mylen - 1) mat_scale<K>(P, pow2_rescale_noexp(mat_max<K>(P)));
      }
    }

    // ---------------- phase 2: segmented warp-shuffle scans ----------------
    float pre[K][K], suf[K][K];
#pragma unroll
    for (int i = 0; i < K; ++i)
#pragma unroll
      for (int j = 0; j < K; ++j) {
        pre[i][j] = P[i][j];
        suf[i][j] = P[i][j];
      }
    int lvl = 0;
    for (int d = 1; d < Tg; d <<= 1, ++lvl) {
      float q[K][K], r[K][K];
#pragma unroll
      for (int i = 0; i < K; ++i)
#pragma unroll
        for (int j = 0; j < K; ++j) q[i][j] = __shfl_up_sync(kFull, pre[i][j], d);
      mat_mul<K>(q, pre, r);
      const bool up_ok = li >= d;
#pragma unroll
      for (int i = 0; i < K; ++i)
#pragma unroll
        for (int j = 0; j < K; ++j) pre[i][j] = up_ok ? r[i][j] : pre[i][j];
#pragma unroll
      for (int i = 0; i < K; ++i)
#pragma unroll
        for (int j = 0; j < K; ++j) q[i][j] = __shfl_down_sync(kFull, suf[i][j], d);
      mat_mul<K>(suf, q, r);
      const bool dn_ok = li + d < Tg;
#pragma unroll
      for (int i = 0; i < K; ++i)
#pragma unroll
        for (int j = 0; j < K; ++j) suf[i][j] = dn_ok ? r[i][j] : suf[i][j];
      if (lvl & 1) {
        mat_scale<K>(pre, pow2_rescale_noexp(mat_max<K>(pre)));
        mat_scale<K>(suf, pow2_rescale_noexp(mat_max<K>(suf)));
      }
    }
    float v[K], w[K];
    {
      float vo[K], wo[K];
#pragma unroll
      for (int j = 0; j < K; ++j) {
        float s = tab.pi2[0] * pre[0][j];
#pragma unroll
        for (int i = 1; i < K; ++i) s = fmaf(tab.pi2[i], pre[i][j], s);
        vo[j] = s;
      }
#pragma unroll
      for (int i = 0; i < K; ++i) {
        float s = suf[i][0];
#pragma unroll
        for (int j = 1; j < K; ++j) s += suf[i][j];
        wo[i] = s;
      }
      float sv = 0.0f, sw = 0.0f;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        v[k] = __shfl_up_sync(kFull, vo[k], 1);
        w[k] = __shfl_down_sync(kFull, wo[k], 1);
        if (first) v[k] = tab.pi2[k];
        if (last) w[k] = 1.0f;
        sv += v[k];
        sw += w[k];
      }
      const float rv = want_ll ? (1.0f / sv) : rcp_approx(sv);
      const float rw = rcp_approx(sw);
#pragma unroll
      for (int k = 0; k < K; ++k) {
        v[k] *= rv;
        w[k] *= rw;
      }
    }

    // the previous bundle's bulk stores have finished reading the staging tile
    if (lane == 0) bulk_wait_read_all();
    __syncwarp();

    // where the bundle's rows sit relative to 16-byte granules (keeps smem and global congruent)
    float* s_gam = s_gam_base;
    int8_t* s_st = s_st_base;
    if (w_gamma_all) s_gam += (reinterpret_cast<uintptr_t>(a.gamma + (size_t)n0 * L * K) & 15u) >> 2;
    if (w_gamma_one) s_gam += (reinterpret_cast<uintptr_t>(a.gamma + (size_t)n0 * L) & 15u) >> 2;
    if (w_states) s_st += reinterpret_cast<uintptr_t>(a.states + (size_t)n0 * L) & 15u;
    const int pos0 = seg * L + t0;  // my first position inside the bundle

    // ---------------- phase 3: forward sweep (alpha column), backward sweep (gamma, states) ----------------
    double ll_thread = 0.0;
    if constexpr (CHT > 0) {
      float al[K];
      int e = 0;
#pragma unroll
      for (int k = 0; k < K; ++k) al[k] = v[k];
      float* gout = s_gam + (size_t)pos0 * (w_gamma_all ? K : 1);
      int8_t* sout = s_st + pos0;
#pragma unroll
      for (int j = 0; j < CHT; ++j) {
        float y[K];
        vecA(al, y);
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const float x = (j == 0 && first) ? al[k] : y[k];
          y[k] = (k == 0) ? x : x * rt[j][k];
        }
        if ((j & 3) == 3) {
          const float sc = want_ll ? pow2_rescale(vec_max<K>(y), e) : pow2_rescale_noexp(vec_max<K>(y));
#pragma unroll
          for (int k = 0; k < K; ++k) y[k] *= sc;
        }
#pragma unroll
        for (int k = 0; k < K; ++k) al[k] = y[k];
        if (K == 2) {
          *reinterpret_cast<float2*>(stack + j * 2) = make_float2(y[0], y[1 % K]);
        } else if (K == 4) {
          *reinterpret_cast<float4*>(stack + j * 4) = make_float4(y[0], y[1 % K], y[2 % K], y[3 % K]);
        } else {
#pragma unroll
          for (int k = 0; k < K; ++k) stack[j * K + k] = y[k];
        }
      }
      if (want_ll && active) {
        float s = al[0];
#pragma unroll
        for (int k = 1; k < K; ++k) s += al[k];
        // every one of the CHT steps (pads included) dropped the scalar s_0
        ll_thread = (double)logf(s) + (double)e * 0.6931471805599453094 + (double)ll_cnt * tab.l00[0] +
                    (double)ll_sumo * tab.dl0[0] + (double)(first ? CHT - 1 : CHT) * tab.log_s0;
        if (first) ll_thread += tab.log_pi2_sum;
      }
      float bv[K], alc[K];
#pragma unroll
      for (int k = 0; k < K; ++k) {
        bv[k] = w[k];
        alc[k] = al[k];
      }
#pragma unroll
      for (int j = CHT - 1; j >= 0; --j) {
        const bool real = j < mylen;
        float gk[K];
        float gsum = 0.0f;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          gk[k] = alc[k] * bv[k];
          gsum += gk[k];
        }
        const float gr = rcp_approx(gsum);
#pragma unroll
        for (int k = 0; k < K; ++k) gk[k] *= gr;
        if (STATS) {
          const bool mj = ((obsmask >> j) & (mask_t)1) != 0;  // pads are never observed
          if (j == 0 && first && active) {
#pragma unroll
            for (int k = 0; k < NEM; ++k) acc_start[k] += gk[k];
          }
          const float oj = mj ? obs_to_float<ObsT>(s_in[j]) : 0.0f;
#pragma unroll
          for (int k = 0; k < NEM; ++k) {
            const float g = mj ? gk[k] : 0.0f;
            acc_num[k] = fmaf(g, oj, acc_num[k]);
            acc_den[k] += g;
          }
        }
        if (real) {
          if (w_gamma_all) {
            if (K == 2) {
              *reinterpret_cast<float2*>(gout + j * 2) = make_float2(gk[0], gk[1 % K]);  // 8-byte aligned (fb2_aligned)
            } else {
#pragma unroll
              for (int k = 0; k < K; ++k) gout[j * K + k] = gk[k];
            }
          } else if (w_gamma_one) {
            float gsel = gk[0];
#pragma unroll
            for (int k = 1; k < K; ++k)
              if (gs == k) gsel = gk[k];
            gout[j] = gsel;
          }
          if (w_states) {
            int best = 0;
            float gbest = gk[0];
#pragma unroll
            for (int k = 1; k < K; ++k)
              if (gk[k] > gbest) {
                gbest = gk[k];
                best = k;
              }
            sout[j] = (int8_t)best;
          }
        }
        if (j > 0 || STATS) {
          float cb[K], nb[K];
          tailAvec(bv, rt[j], cb, nb);
          if (STATS) {
            // xi_{t-1}(i,jj) ~ alpha_{t-1}(i) Ap[i][jj] cb[jj], t = t0 + j > 0, both ends observed
            float alp[K];
            if (j > 0) {
              if (K == 2) {
                const float2 t2 = *reinterpret_cast<const float2*>(stack + (j > 0 ? j - 1 : 0) * 2);
                alp[0] = t2.x;
                alp[1 % K] = t2.y;
              } else {
#pragma unroll
                for (int k = 0; k < K; ++k) alp[k] = stack[(j > 0 ? j - 1 : 0) * K + k];
              }
            } else {
#pragma unroll
              for (int k = 0; k < K; ++k) alp[k] = v[k];
            }
            const bool prev_obs = (j > 0) ? (((obsmask >> (j > 0 ? j - 1 : 0)) & (mask_t)1) != 0) : prev_obs0;
            const bool valid = (((obsmask >> j) & (mask_t)1) != 0) && prev_obs && !(j == 0 && first);
            float s2 = 0.0f;
#pragma unroll
            for (int i = 0; i < K; ++i) s2 = fmaf(alp[i], nb[i], s2);
            const float r2 = valid ? rcp_approx(s2) : 0.0f;
#pragma unroll
            for (int i = 0; i < K; ++i) {
              const float ai = alp[i] * r2;
#pragma unroll
              for (int jj = 0; jj < K; ++jj) acc_xi[(i * K + jj) % NXI] = fmaf(ai, cb[jj], acc_xi[(i * K + jj) % NXI]);
            }
          }
          if (j > 0 && ((CHT - 1 - j) & 3) == 3) {
            const float sc = pow2_rescale_noexp(vec_max<K>(nb));
#pragma unroll
            for (int k = 0; k < K; ++k) nb[k] *= sc;
          }
#pragma unroll
          for (int k = 0; k < K; ++k) bv[k] = nb[k];
          if (j == 0) {
            // nothing left to do in this chunk
          } else if (K == 2) {
            const float2 t2 = *reinterpret_cast<const float2*>(stack + (j - 1) * 2);
            alc[0] = t2.x;
            alc[1 % K] = t2.y;
          } else if (K == 4) {
            const float4 t4 = *reinterpret_cast<const float4*>(stack + (j - 1) * 4);
            alc[0] = t4.x;
            alc[1 % K] = t4.y;
            alc[2 % K] = t4.z;
            alc[3 % K] = t4.w;
          } else {
#pragma unroll
            for (int k = 0; k < K; ++k) alc[k] = stack[(j - 1) * K + k];
          }
        }
      }
    } else
    {
      float al[K];
      int e = 0;
#pragma unroll
      for (int k = 0; k < K; ++k) al[k] = v[k];
      float fn[K];
      ratios((mylen > 0) ? obs_to_float<ObsT>(s_in[0]) : 0.0f, fn);
      for (int j = 0; j < mylen; ++j) {
        float cf[K], y[K];
#pragma unroll
        for (int k = 0; k < K; ++k) cf[k] = fn[k];
        if (j + 1 < mylen) ratios(obs_to_float<ObsT>(s_in[j + 1]), fn);
        vecA(al, y);
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const float x = (first && j == 0) ? al[k] : y[k];
          y[k] = (k == 0) ? x : x * cf[k];
        }
        if ((j & 3) == 3) {
          const float sc = want_ll ? pow2_rescale(vec_max<K>(y), e) : pow2_rescale_noexp(vec_max<K>(y));
#pragma unroll
          for (int k = 0; k < K; ++k) y[k] *= sc;
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
          al[k] = y[k];
          stack[j * K + k] = y[k];
        }
      }
      if (want_ll && active) {
        float s = al[0];
#pragma unroll
        for (int k = 1; k < K; ++k) s += al[k];
        ll_thread = (double)logf(s) + (double)e * 0.6931471805599453094 + (double)ll_cnt * tab.l00[0] +
                    (double)ll_sumo * tab.dl0[0] + (double)(first ? mylen - 1 : mylen) * tab.log_s0;
        if (first) ll_thread += tab.log_pi2_sum;
      }
      float bv[K], alc[K];
#pragma unroll
      for (int k = 0; k < K; ++k) {
        bv[k] = w[k];
        alc[k] = al[k];
      }
      float rn[K], an[K];  // ratios of position j, alpha of position j-1: fetched one step ahead
      ratios((mylen > 0) ? obs_to_float<ObsT>(s_in[mylen - 1]) : 0.0f, rn);
#pragma unroll
      for (int k = 0; k < K; ++k) an[k] = (mylen > 1) ? stack[(mylen - 2) * K + k] : 0.0f;
      for (int j = mylen - 1; j >= 0; --j) {
        float cr[K], alp[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
          cr[k] = rn[k];
          alp[k] = an[k];
        }
        if (j > 0) {
          ratios(obs_to_float<ObsT>(s_in[j - 1]), rn);
          if (j > 1) {
#pragma unroll
            for (int k = 0; k < K; ++k) an[k] = stack[(j - 2) * K + k];
          }
        }
        float gk[K];
        float gsum = 0.0f;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          gk[k] = alc[k] * bv[k];
          gsum += gk[k];
        }
        const float gr = rcp_approx(gsum);
#pragma unroll
        for (int k = 0; k < K; ++k) gk[k] *= gr;
        if (w_gamma_all) {
#pragma unroll
          for (int k = 0; k < K; ++k) s_gam[(size_t)(pos0 + j) * K + k] = gk[k];
        } else if (w_gamma_one) {
          float gsel = gk[0];
#pragma unroll
          for (int k = 1; k < K; ++k)
            if (gs == k) gsel = gk[k];
          s_gam[pos0 + j] = gsel;
        }
        if (w_states) {
          int best = 0;
          float gbest = gk[0];
#pragma unroll
          for (int k = 1; k < K; ++k)
            if (gk[k] > gbest) {
              gbest = gk[k];
              best = k;
            }
          s_st[pos0 + j] = (int8_t)best;
        }
        if (j > 0) {
          float cb[K], nb[K];
          tailAvec(bv, cr, cb, nb);
          if (((mylen - 1 - j) & 3) == 3) {
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
    }

    if (STATS) {
      acc_ll += ll_thread;
      if (((it + 1) % a.flush_every) == 0) flush_stats();
    }
    // ---------------- store: one bulk store per output for the whole bundle ----------------
    fence_proxy_async();
    __syncwarp();
    if (!STATS) {
      const uint32_t npos = (uint32_t)(nr * L);
      if (w_gamma_all) store_row_bulk<float>(a.gamma + (size_t)n0 * L * K, s_gam, npos * K * 4u, lane);
      if (w_gamma_one) store_row_bulk<float>(a.gamma + (size_t)n0 * L, s_gam, npos * 4u, lane);
      if (w_states) store_row_bulk<int8_t>(a.states + (size_t)n0 * L, s_st, npos, lane);
      if (lane == 0) bulk_commit();
    }
    if (!STATS && want_ll) {
      // segmented sum over the lanes of a read
      double s = ll_thread;
      for (int d = 1; d < Tg; d <<= 1) {
        const double o = __shfl_down_sync(kFull, s, d);
        if (li + d < Tg) s += o;
      }
      // after a power-of-two ladder lane li = 0 holds the sum of lanes [0, 2^ceil(log2 Tg)) restricted by the guard
      if (first && seg < nr) a.loglik[n0 + seg] = s;
    }
    __syncwarp();  // every lane is done with this observation buffer before it is refilled
  }
  if (lane == 0) bulk_wait_all();
  if (STATS) {
    flush_stats();
    const double wl = warp_sum_f64(acc_ll);
    if (lane == 0) wacc[a.S - 1] += wl;
    __syncthreads();
    const double* all = reinterpret_cast<const double*>(smem);
    for (int sidx = threadIdx.x; sidx < a.S; sidx += blockDim.x) {
      double tot = 0.0;
      for (int wv = 0; wv < kShortWarps; ++wv) tot += all[(size_t)wv * a.S + sidx];
      a.partials[(size_t)blockIdx.x * a.S + sidx] = tot;
    }
  }
}

}  // namespace smf
