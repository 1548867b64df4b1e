// Viterbi decode (K2): max-plus forward sweep with packed back-pointers + back-trace.
//
// One thread owns one read and walks it sequentially in fp64, repeating the reference's
// operation order exactly (HMM.py:647-669):
//     logB = o*log(p+eps) + (1-o)*log1p(-p+eps)      (two products, one sum; 0 when masked)
//     cand = delta[t-1,i] + logA[i,j];  best over i with FIRST maximum;  delta[t,j] = best + logB
// so delta -- and therefore every arg-max decision, including exact ties -- is bit-identical
// to torch's fp64 CPU result when the log tables come from the same torch ops.
//
// Data movement: a CTA of kVitReads reads stages a [reads x kVitTile positions] observation
// tile through shared memory (coalesced row segments in, conflict-free per-thread rows out).
// The K two-bit back-pointers of a position are packed into one int8 in a shared-memory tile,
// which is flushed coalesced to a [N, L] byte scratch in HBM (one read's worth of
// back-pointers is L bytes; 128 resident reads x 10 kb does not fit on chip), and streamed
// back tile by tile, in reverse, for the back-trace.  States leave through the same tile.
#pragma once
#include "hmm_common.cuh"

namespace smf {

constexpr int kVitReads = 128;  // reads (threads) per CTA
constexpr int kVitTile = 32;    // positions per staged tile

struct VitArgs {
  const void* obs;
  int obs_dtype;
  long long N;
  int L;
  int C;
  const int8_t* bins;  // device [L-1] or nullptr
  int8_t* states;      // [N, L]
  uint8_t* bp;         // [N, L] packed back-pointers (scratch)
};

template <int K, bool MULTI, bool BINNED, bool WIDE>
__global__ void __launch_bounds__(kVitReads) viterbi_kernel(const __grid_constant__ VitTables<K> tab,
                                                            const __grid_constant__ VitArgs a) {
  using obs_t = typename std::conditional<WIDE, double, float>::type;
  extern __shared__ __align__(16) unsigned char smem[];
  const int C = a.C;
  const int L = a.L;
  const int row_elems = kVitTile * C;
  const int row_stride = row_elems | 1;  // odd stride (in elements): conflict-free per-thread rows
  obs_t* tile = reinterpret_cast<obs_t*>(smem);
  constexpr int kBpStride = kVitTile + 4;  // bytes; 9 words -> conflict-free byte rows
  uint8_t* bpt = smem + (size_t)kVitReads * row_stride * sizeof(obs_t);
  int8_t* sbins = reinterpret_cast<int8_t*>(bpt + kVitReads * kBpStride);

  const int tid = threadIdx.x;
  const long long n0 = (long long)blockIdx.x * kVitReads;
  const int nrows = (int)min((long long)kVitReads, a.N - n0);
  const bool live = tid < nrows;

  if (BINNED) {
    for (int i = tid; i < L - 1; i += kVitReads) sbins[i] = a.bins[i];
  }

  double delta[K];
#pragma unroll
  for (int k = 0; k < K; ++k) delta[k] = 0.0;

  // ---------------- forward max-plus sweep ----------------
  for (int tile0 = 0; tile0 < L; tile0 += kVitTile) {
    const int tw = min(kVitTile, L - tile0);
    __syncthreads();  // previous tile fully consumed (and sbins visible)
    for (int idx = tid; idx < nrows * tw * C; idx += kVitReads) {
      const int row = idx / (tw * C);
      const int col = idx - row * (tw * C);
      const long long gidx = ((n0 + row) * (long long)L + tile0) * C + col;
      if (WIDE)
        tile[row * row_stride + col] = (obs_t)load_obs_f64(a.obs, a.obs_dtype, gidx);
      else
        tile[row * row_stride + col] = (obs_t)load_obs(a.obs, a.obs_dtype, gidx);
    }
    __syncthreads();
    if (live) {
      const obs_t* my = tile + tid * row_stride;
      for (int j = 0; j < tw; ++j) {
        const int t = tile0 + j;
        double logB[K];
#pragma unroll
        for (int k = 0; k < K; ++k) logB[k] = 0.0;
        if (!MULTI) {
          const double o = (double)my[j];
          if (o == o) {
#pragma unroll
            for (int k = 0; k < K; ++k)
              logB[k] = __dadd_rn(__dmul_rn(o, tab.l1[k][0]),
                                  __dmul_rn(__dsub_rn(1.0, o), tab.l0[k][0]));
          }
        } else {
          for (int c = 0; c < C; ++c) {
            const double o = (double)my[j * C + c];
            if (o == o) {
#pragma unroll
              for (int k = 0; k < K; ++k)
                logB[k] = __dadd_rn(logB[k], __dadd_rn(__dmul_rn(o, tab.l1[k][c]),
                                                       __dmul_rn(__dsub_rn(1.0, o), tab.l0[k][c])));
            }
          }
        }
        unsigned packed = 0;
        if (t == 0) {
#pragma unroll
          for (int k = 0; k < K; ++k) delta[k] = __dadd_rn(tab.log_pi[k], logB[k]);
        } else {
          const int bin = BINNED ? (int)sbins[t - 1] : 0;
          double nd[K];
#pragma unroll
          for (int jn = 0; jn < K; ++jn) {
            double best = __dadd_rn(delta[0], tab.logA[bin][0][jn]);
            unsigned arg = 0;
#pragma unroll
            for (int i = 1; i < K; ++i) {
              const double cnd = __dadd_rn(delta[i], tab.logA[bin][i][jn]);
              if (cnd > best) {
                best = cnd;
                arg = i;
              }
            }
            nd[jn] = __dadd_rn(best, logB[jn]);
            packed |= arg << (2 * jn);
          }
#pragma unroll
          for (int k = 0; k < K; ++k) delta[k] = nd[k];
        }
        bpt[tid * kBpStride + j] = (uint8_t)packed;
      }
    }
    __syncthreads();
    for (int idx = tid; idx < nrows * tw; idx += kVitReads) {
      const int row = idx / tw;
      const int col = idx - row * tw;
      a.bp[(n0 + row) * (long long)L + tile0 + col] = bpt[row * kBpStride + col];
    }
  }

  // last state: first maximum of delta[L-1] (torch.argmax, HMM.py:664)
  int s = 0;
  {
    double best = delta[0];
#pragma unroll
    for (int k = 1; k < K; ++k)
      if (delta[k] > best) {
        best = delta[k];
        s = k;
      }
  }

  // ---------------- back-trace, tiles in reverse ----------------
  const int ntiles = (L + kVitTile - 1) / kVitTile;
  for (int ti = ntiles - 1; ti >= 0; --ti) {
    const int tile0 = ti * kVitTile;
    const int tw = min(kVitTile, L - tile0);
    __syncthreads();  // global bp writes of this CTA visible; previous tile flushed
    for (int idx = tid; idx < nrows * tw; idx += kVitReads) {
      const int row = idx / tw;
      const int col = idx - row * tw;
      bpt[row * kBpStride + col] = a.bp[(n0 + row) * (long long)L + tile0 + col];
    }
    __syncthreads();
    if (live) {
      uint8_t* my = bpt + tid * kBpStride;
      for (int j = tw - 1; j >= 0; --j) {
        const unsigned packed = my[j];
        my[j] = (uint8_t)s;  // state at position tile0 + j
        if (tile0 + j > 0) s = (int)((packed >> (2 * s)) & 3u);  // s[t-1] = psi[t, s[t]]
      }
    }
    __syncthreads();
    for (int idx = tid; idx < nrows * tw; idx += kVitReads) {
      const int row = idx / tw;
      const int col = idx - row * tw;
      a.states[(n0 + row) * (long long)L + tile0 + col] = (int8_t)bpt[row * kBpStride + col];
    }
  }
}

}  // namespace smf
