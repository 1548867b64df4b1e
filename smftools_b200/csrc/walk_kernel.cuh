// Boundary-walk kernel of the two-kernel forward-backward path (K >= 3).
//
// For K >= 3 the chunk-product + matrix-scan phases of fb2_kernel cost ~K^3 FMAs per position and a
// K x K register matrix per thread; the E-step variant additionally carries K^2 + 3K accumulators
// and spills.  With >= ~10^4 reads per launch there is enough read-level parallelism to get the
// chunk boundary vectors the cheap way instead: ONE THREAD PER READ AND DIRECTION walks the read
// sequentially (K(K-1) FMAs per position, scaled fp32, the same column-normalised tables as
// fb2_kernel) and records
//     v_i = alpha just before chunk i            (forward CTAs, blockIdx.y == 0; chunk 0 starts from pi2)
//     w_i = beta at the last position of chunk i (backward CTAs, blockIdx.y == 1; the last chunk ends in 1)
// chunk-major ([i][n][K]: coalesced across the reads of a warp), plus the read's log-likelihood.
// fb2_kernel<..., WALK = true> then only runs its thread-local sweeps from those vectors: no chunk
// matrices, no scans, no cross-warp hand-off.  Observations are staged through shared memory like in
// viterbi_kernel (coalesced 128-byte row segments, one tile prefetched into registers, odd row
// stride).  Cost: one extra forward and one extra backward read of the observations (2 x 4 B per
// position for fp32 input) and 2 * 8K / CH bytes per position of boundary vectors (written + read).
#pragma once
#include "hmm_common.cuh"
#include "viterbi_kernel.cuh"

namespace smf {

constexpr int kWalkTile = 16;                 // positions per staged tile (one register per tile row and thread)
constexpr int kWalkRowsPerLoad = 32 / kWalkTile;  // rows one warp-wide load covers
constexpr int kWalkLoads = 32 / kWalkRowsPerLoad;  // loads per tile and thread

struct WalkArgs {
  const void* obs;
  long long N;
  int L;
  int CH;         // chunk length of the consumer kernel
  int Tact;       // chunks per read = ceil(L / CH)
  float* bnd_v;   // [Tact][N][K]
  float* bnd_w;   // [Tact][N][K]
  double* ll;     // [N] log-likelihood per read (may be nullptr)
};

template <int K, int OBSK>
__global__ void __launch_bounds__(kVitReads, (K >= 4 ? 7 : 8)) walk_kernel(const __grid_constant__ FbTables<K> tab,
                                                            const __grid_constant__ WalkArgs a) {
  using src_t = typename std::conditional<OBSK == SMF_OBS_F64, double,
                                          typename std::conditional<OBSK == SMF_OBS_U8, unsigned char, float>::type>::type;
  extern __shared__ __align__(16) unsigned char smem[];
  const int L = a.L;
  constexpr int row_stride = kWalkTile | 1;
  float* tile = reinterpret_cast<float*>(smem);

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int wbase = (tid >> 5) * 32;
  const long long n0 = (long long)blockIdx.x * kVitReads;
  const int nrows = (int)min((long long)kVitReads, a.N - n0);
  const bool live = tid < nrows;
  const long long n = n0 + tid;
  const size_t bstride = (size_t)a.N * K;  // floats between consecutive chunks
  const bool backward = (blockIdx.y != 0);
  // rows of this warp that exist; rows beyond are clamped to the last valid one (loads stay in bounds,
  // the values are never used)
  const int rows_w = max(0, min(32, nrows - wbase));

  float A[K][K];
#pragma unroll
  for (int i = 0; i < K; ++i)
#pragma unroll
    for (int j = 0; j < K; ++j) A[i][j] = tab.Ap[i][j];

  float nxt[kWalkLoads];
  const src_t* wsrc = reinterpret_cast<const src_t*>(a.obs) + (n0 + wbase) * (long long)L;
  const int lcol = lane % kWalkTile;  // column of the tile this lane fetches
  const int lrow = lane / kWalkTile;  // row offset inside one warp-wide load
  auto fetch_tile = [&](int tile0) {
    if (rows_w == 0) return;
    const int col = min(tile0 + lcol, L - 1);  // clamped: columns past the end are never consumed
    const src_t* p = wsrc + col;
    if (rows_w == 32) {
#pragma unroll
      for (int r = 0; r < kWalkLoads; ++r)
        nxt[r] = obs_to_float<src_t>(__ldg(p + (r * kWalkRowsPerLoad + lrow) * L));  // 32-bit offset: < 32 * 2^24
    } else {
#pragma unroll
      for (int r = 0; r < kWalkLoads; ++r)
        nxt[r] = obs_to_float<src_t>(__ldg(p + min(r * kWalkRowsPerLoad + lrow, rows_w - 1) * L));
    }
  };
  auto commit_tile = [&]() {
#pragma unroll
    for (int r = 0; r < kWalkLoads; ++r) tile[(wbase + r * kWalkRowsPerLoad + lrow) * row_stride + lcol] = nxt[r];
  };
  auto ratios = [&](float o, float (&c)[K]) {
    const bool m = (o == o);
#pragma unroll
    for (int k = 1; k < K; ++k) c[k] = ex2_approx(m ? fmaf(o, tab.rdl[k][0], tab.rl0p[k]) : tab.rmiss[k]);
    return m;
  };
  const int ntiles = (L + kWalkTile - 1) / kWalkTile;
  const float* my = tile + tid * row_stride;

  if (!backward) {
    // ---------------- forward walk ----------------
    float al[K];
    {
      float s = 0.0f;
#pragma unroll
      for (int k = 0; k < K; ++k) s += tab.pi2[k];
      const float r = 1.0f / s;
#pragma unroll
      for (int k = 0; k < K; ++k) al[k] = tab.pi2[k] * r;
    }
    int e = 0;          // power-of-two exponent taken out of alpha so far
    int cnt = 0;        // observed positions
    double sumo = 0.0;  // sum of observed values (state-0 emission mass)
    int next_emit = a.CH - 1;  // v_{chunk_i} = alpha after this position
    int chunk_i = 1;
    float* dstv = a.bnd_v + bstride + (size_t)n * K;
    fetch_tile(0);
    for (int ti = 0; ti < ntiles; ++ti) {
      const int tile0 = ti * kWalkTile;
      const int tw = min(kWalkTile, L - tile0);
      __syncthreads();
      commit_tile();
      __syncthreads();
      if (ti + 1 < ntiles) fetch_tile(tile0 + kWalkTile);
      if (live) {
        float tsum = 0.0f;  // per-tile partial of sumo (<= 32 terms in fp32, then fp64)
        const int ej = next_emit - tile0;  // tile-local position after which v is recorded (chunks are longer than a tile)
        auto fstep = [&](int j, bool first_tile) {
          const float o = my[j];
          float c[K];
          const bool m = ratios(o, c);
          cnt += m ? 1 : 0;
          tsum += m ? o : 0.0f;
          float y[K];
#pragma unroll
          for (int jj = 0; jj < K; ++jj) {
            float s = al[jj];
#pragma unroll
            for (int i = 0; i < K; ++i)
              if (i != jj) s = fmaf(al[i], A[i][jj], s);
            y[jj] = (first_tile && j == 0) ? al[jj] : s;  // alpha_0 = pi2 * c: no transition
          }
#pragma unroll
          for (int k = 1; k < K; ++k) y[k] *= c[k];
          if ((j & 3) == 3) {  // tiles start at multiples of 16: (t & 3) == (j & 3)
            const float sc = pow2_rescale(vec_max<K>(y), e);
#pragma unroll
            for (int k = 0; k < K; ++k) y[k] *= sc;
          }
#pragma unroll
          for (int k = 0; k < K; ++k) al[k] = y[k];
          if (j == ej && chunk_i < a.Tact) {
#pragma unroll
            for (int k = 0; k < K; ++k) dstv[k] = al[k];
          }
        };
        if (ti > 0 && tw == kWalkTile) {
#pragma unroll
          for (int j = 0; j < kWalkTile; ++j) fstep(j, false);
        } else {
#pragma unroll 4
          for (int j = 0; j < tw; ++j) fstep(j, ti == 0);
        }
        if (ej < tw) {
          ++chunk_i;
          dstv += bstride;
          next_emit += a.CH;
        }
        sumo += (double)tsum;
      }
    }
    if (live && a.ll != nullptr) {
      float s = al[0];
#pragma unroll
      for (int k = 1; k < K; ++k) s += al[k];
      a.ll[n] = (double)logf(s) + (double)e * 0.6931471805599453094 + tab.log_pi2_sum + (double)cnt * tab.l00[0] +
                sumo * tab.dl0[0] + (double)(L - 1) * tab.log_s0;
    }
  } else {
    // ---------------- backward walk ----------------
    float bv[K];
#pragma unroll
    for (int k = 0; k < K; ++k) bv[k] = 1.0f / K;
    int wi = a.Tact - 1;       // next boundary to record is w_{wi-1} ...
    int wpos = wi * a.CH - 1;  // ... = beta at the last position of chunk wi-1
    float* dstw = a.bnd_w + (size_t)(wi > 0 ? wi - 1 : 0) * bstride + (size_t)n * K;
    fetch_tile((ntiles - 1) * kWalkTile);
    for (int ti = ntiles - 1; ti >= 0; --ti) {
      const int tile0 = ti * kWalkTile;
      const int tw = min(kWalkTile, L - tile0);
      __syncthreads();
      commit_tile();
      __syncthreads();
      if (ti > 0) fetch_tile(tile0 - kWalkTile);
      if (live) {
        const int wj = wpos - tile0;  // tile-local position whose beta is recorded (at most one per tile)
        auto bstep = [&](int j) {
          if (j == wj) {  // bv == beta_t
#pragma unroll
            for (int k = 0; k < K; ++k) dstw[k] = bv[k];
          }
          const float o = my[j];
          float c[K];
          ratios(o, c);
          float cb[K], nb[K];
          cb[0] = bv[0];
#pragma unroll
          for (int k = 1; k < K; ++k) cb[k] = c[k] * bv[k];
#pragma unroll
          for (int i = 0; i < K; ++i) {
            float s = cb[i];
#pragma unroll
            for (int jj = 0; jj < K; ++jj)
              if (jj != i) s = fmaf(A[i][jj], cb[jj], s);
            nb[i] = s;
          }
          if ((j & 3) == 0) {
            const float sc = pow2_rescale_noexp(vec_max<K>(nb));
#pragma unroll
            for (int k = 0; k < K; ++k) nb[k] *= sc;
          }
#pragma unroll
          for (int k = 0; k < K; ++k) bv[k] = nb[k];
        };
        if (tw == kWalkTile) {
#pragma unroll
          for (int j = kWalkTile - 1; j >= 0; --j) bstep(j);
        } else {
#pragma unroll 4
          for (int j = tw - 1; j >= 0; --j) bstep(j);
        }
        if (wj >= 0 && wj < tw) {
          dstw -= bstride;
          wpos -= a.CH;
        }
      }
    }
  }
}

}  // namespace smf
