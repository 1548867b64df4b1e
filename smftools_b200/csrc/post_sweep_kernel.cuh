// Posterior sweeps of the two-kernel path for reads of ANY length.
//
// fb2_kernel (with or without the boundary walk) keeps a whole read on chip, which bounds the read
// length (~21 k positions for K = 2, ~11 k for K = 4).  With the chunk boundary vectors from
// walk_kernel the chunks of a read are independent, so -- exactly like estep_sweep_kernel -- the unit
// of work can be a SLICE (32 consecutive chunks of one read, one per lane) owned by one warp:
// private double-buffered bulk load of the slice's observations, forward sweep stacking alpha in
// the warp's staging tile, backward sweep overwriting it in place with gamma (the tile is laid out
// like the global rows: position-major, so lane columns and output rows coincide), one bulk store
// per output and slice.  No CTA barrier, no read-length limit.  Position loops are rolled and
// software-pipelined (run-time odd chunk length).
#pragma once
#include "fb2_kernel.cuh"

namespace smf {

constexpr int kPostWarps = 4;  // warps per CTA (small CTAs: shared memory limits residency)

struct PostSweepArgs {
  const void* obs;
  long long N;
  int L;
  int CH;      // chunk length (odd)
  int Tact;    // chunks per read
  int slices;  // slices per read = ceil(Tact / 32)
  const float* bnd_v;  // [Tact][N][K]
  const float* bnd_w;
  float* gamma;     // [N,L,K] / [N,L] / nullptr
  int gamma_state;  // < 0: all states
  int8_t* states;   // [N,L] / nullptr
  int in_bytes;     // one observation buffer
  int stage_off, one_off, st_off;  // byte offsets inside a warp's region
  int warp_bytes;
};

// CHT > 0: compile-time chunk length -- fully unrolled sweeps, emission ratios cached in registers,
// positions past the end of a ragged chunk treated as missing observations (only the loads and the
// output stores are predicated); CHT == 0: run-time chunk length, rolled software-pipelined loops.
template <int K, typename ObsT, int CHT = 0>
__global__ void __launch_bounds__(kPostWarps * 32, 4) post_sweep_kernel(const __grid_constant__ FbTables<K> tab,
                                                                       const __grid_constant__ PostSweepArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int L = a.L;
  const int CH = (CHT > 0) ? CHT : a.CH;
  constexpr int NRT = (CHT > 0) ? CHT : 1;

  unsigned char* wbase = smem + (size_t)warp * a.warp_bytes;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(wbase);  // [2]
  unsigned char* in_raw = wbase + 16;
  float* s_stage_base = reinterpret_cast<float*>(wbase + a.stage_off);  // alpha, then gamma (all states) in place
  float* s_one_base = reinterpret_cast<float*>(wbase + a.one_off);      // single-state gamma
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
  auto ratios = [&](float o, float (&c)[K]) {
    const bool m = (o == o);
    c[0] = 1.0f;
#pragma unroll
    for (int k = 1; k < K; ++k) c[k] = ex2_approx(m ? fmaf(o, tab.rdl[k][0], tab.rl0p[k]) : tab.rmiss[k]);
  };

  const int gs = a.gamma_state;
  const bool w_gamma_all = (a.gamma != nullptr && gs < 0);
  const bool w_gamma_one = (a.gamma != nullptr && gs >= 0);
  const bool w_states = (a.states != nullptr);
  const long long total = a.N * (long long)a.slices;
  const long long wstride = (long long)gridDim.x * kPostWarps;
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

  long long item = (long long)blockIdx.x * kPostWarps + warp;
  if (lane == 0 && item < total) issue_load(item, 0);

  int it = 0;
  for (; item < total; item += wstride, ++it) {
    const int buf = it & 1;
    if (lane == 0 && item + wstride < total) issue_load(item + wstride, buf ^ 1);

    const long long n = item / a.slices;
    const int s = (int)(item - n * a.slices);
    const int p0 = s * slice_pos;
    const int npos = min(slice_pos, L - p0);
    const int c = s * 32 + lane;
    const int t0 = c * CH;
    const bool active = c < a.Tact;
    const bool first = (c == 0);
    const int len = active ? min(CH, L - t0) : 0;

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
    const uint32_t mis =
        (uint32_t)(reinterpret_cast<uintptr_t>(obs_base + ((size_t)n * L + p0) * sizeof(ObsT)) & 15u);
    const ObsT* s_in =
        reinterpret_cast<const ObsT*>(in_raw + (size_t)buf * a.in_bytes + mis) + (size_t)lane * CH;

    // staging tile congruent (mod 16 bytes) with the global rows it is stored to
    const size_t gpos = (size_t)n * L + p0;  // first position of the slice in the [N, L] grid
    float* s_stage = s_stage_base;
    float* s_one = s_one_base;
    int8_t* s_st = s_st_base;
    if (w_gamma_all) s_stage += (reinterpret_cast<uintptr_t>(a.gamma + gpos * K) & 15u) >> 2;
    if (w_gamma_one) s_one += (reinterpret_cast<uintptr_t>(a.gamma + gpos) & 15u) >> 2;
    if (w_states) s_st += reinterpret_cast<uintptr_t>(a.states + gpos) & 15u;
    float* stg = s_stage + (size_t)lane * CH * K;  // my column == my rows of the slice

    // the previous slice's bulk stores have finished reading the tile
    if (lane == 0) bulk_wait_read_all();
    __syncwarp();
    mbar_wait(&mbar[buf], (uint32_t)((it >> 1) & 1));

    if constexpr (CHT > 0) {
      // ---------------- unrolled sweeps ----------------
      const float kMissing = __int_as_float(0x7fc00000);
      float rt[NRT][K];
      float al[K];
#pragma unroll
      for (int k = 0; k < K; ++k) al[k] = v[k];
#pragma unroll
      for (int j = 0; j < CHT; ++j) {
        const float o = (j < len) ? obs_to_float<ObsT>(s_in[j]) : kMissing;
        float cf[K], y[K];
        ratios(o, cf);
#pragma unroll
        for (int k = 0; k < K; ++k) rt[j][k] = cf[k];
#pragma unroll
        for (int jj = 0; jj < K; ++jj) {
          float acc = al[jj];
#pragma unroll
          for (int i = 0; i < K; ++i)
            if (i != jj) acc = fmaf(al[i], A[i][jj], acc);
          const float x = (j == 0 && first) ? al[jj] : acc;
          y[jj] = (jj == 0) ? x : x * cf[jj];
        }
        if ((j & 3) == 3) {
          const float sc = pow2_rescale_noexp(vec_max<K>(y));
#pragma unroll
          for (int k = 0; k < K; ++k) y[k] *= sc;
        }
#pragma unroll
        for (int k = 0; k < K; ++k) al[k] = y[k];
        if (active) {  // idle lanes of the last slice own no column
          if (K == 4) {
            *reinterpret_cast<float4*>(stg + j * 4) = make_float4(y[0], y[1 % K], y[2 % K], y[3 % K]);
          } else {
#pragma unroll
            for (int k = 0; k < K; ++k) stg[j * K + k] = y[k];
          }
        }
      }
      float bv[K], alc[K];
#pragma unroll
      for (int k = 0; k < K; ++k) {
        bv[k] = w[k];
        alc[k] = al[k];
      }
#pragma unroll
      for (int j = CHT - 1; j >= 0; --j) {
        const bool real = j < len;
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
        float alp[K];
        if (j > 0) {
          if (K == 4) {
            const float4 t4 = *reinterpret_cast<const float4*>(stg + (j - 1) * 4);
            alp[0] = t4.x;
            alp[1 % K] = t4.y;
            alp[2 % K] = t4.z;
            alp[3 % K] = t4.w;
          } else {
#pragma unroll
            for (int k = 0; k < K; ++k) alp[k] = stg[(j - 1) * K + k];
          }
        }
        if (real) {
          if (w_gamma_all) {
            if (K == 4) {
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
            s_one[lane * CHT + j] = gsel;
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
            s_st[lane * CHT + j] = (int8_t)best;
          }
        }
        if (j > 0) {
          float cb[K], nb[K];
          cb[0] = bv[0];
#pragma unroll
          for (int k = 1; k < K; ++k) cb[k] = rt[j][k] * bv[k];
#pragma unroll
          for (int i = 0; i < K; ++i) {
            float acc = cb[i];
#pragma unroll
            for (int jj = 0; jj < K; ++jj)
              if (jj != i) acc = fmaf(A[i][jj], cb[jj], acc);
            nb[i] = acc;
          }
          if (((CHT - 1 - j) & 3) == 3) {
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
    } else {
      // ---------------- forward sweep (ratios of position j + 1 formed while j is folded in) ----------------
      float al[K];
#pragma unroll
      for (int k = 0; k < K; ++k) al[k] = v[k];
      {
        float fn[K];
        ratios((len > 0) ? obs_to_float<ObsT>(s_in[0]) : 0.0f, fn);
        for (int j = 0; j < len; ++j) {
          float cf[K], y[K];
#pragma unroll
          for (int k = 0; k < K; ++k) cf[k] = fn[k];
          if (j + 1 < len) ratios(obs_to_float<ObsT>(s_in[j + 1]), fn);
#pragma unroll
          for (int jj = 0; jj < K; ++jj) {
            float acc = al[jj];
#pragma unroll
            for (int i = 0; i < K; ++i)
              if (i != jj) acc = fmaf(al[i], A[i][jj], acc);
            const float x = (first && j == 0) ? al[jj] : acc;  // alpha_0 = pi2 * c: no transition
            y[jj] = (jj == 0) ? x : x * cf[jj];
          }
          if ((j & 3) == 3) {
            const float sc = pow2_rescale_noexp(vec_max<K>(y));
#pragma unroll
            for (int k = 0; k < K; ++k) y[k] *= sc;
          }
#pragma unroll
          for (int k = 0; k < K; ++k) {
            al[k] = y[k];
            stg[j * K + k] = y[k];
          }
        }
      }
      // ---------------- backward sweep: gamma over alpha in place, arg-max states ----------------
      {
        float bv[K], alc[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
          bv[k] = w[k];
          alc[k] = al[k];
        }
        float rn[K], an[K];
        ratios((len > 0) ? obs_to_float<ObsT>(s_in[len - 1]) : 0.0f, rn);
#pragma unroll
        for (int k = 0; k < K; ++k) an[k] = (len > 1) ? stg[(len - 2) * K + k] : 0.0f;
        for (int j = len - 1; j >= 0; --j) {
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
              for (int k = 0; k < K; ++k) an[k] = stg[(j - 2) * K + k];
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
            for (int k = 0; k < K; ++k) stg[j * K + k] = gk[k];
          } else if (w_gamma_one) {
            float gsel = gk[0];
#pragma unroll
            for (int k = 1; k < K; ++k)
              if (gs == k) gsel = gk[k];
            s_one[lane * CH + j] = gsel;
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
            s_st[lane * CH + j] = (int8_t)best;
          }
          if (j > 0) {
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
        }
      }
    }
    // ---------------- store ----------------
    fence_proxy_async();
    __syncwarp();
    if (w_gamma_all) store_row_bulk<float>(a.gamma + gpos * K, s_stage, (uint32_t)npos * K * 4u, lane);
    if (w_gamma_one) store_row_bulk<float>(a.gamma + gpos, s_one, (uint32_t)npos * 4u, lane);
    if (w_states) store_row_bulk<int8_t>(a.states + gpos, s_st, (uint32_t)npos, lane);
    if (lane == 0) bulk_commit();
    __syncwarp();  // every lane is done with this observation buffer before it is refilled
  }
  if (lane == 0) bulk_wait_all();
}

}  // namespace smf
