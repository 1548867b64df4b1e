// Layer reductions computed from the packed device results: what the reference's
// downstream consumers derive from the dense float64 layers by re-reading them on the host --
//   * per-layer "modified" / "valid" cell counts  (tools/partitioned_hmm.py:519-529, _plot_feature_fractions)
//   * per-column sums / finite counts -> nan-mean + rolling mean  (hmm/call_hmm_peaks.py:151-160)
//   * runs-per-read and run-size histograms       (tools/partitioned_hmm.py:572-630)
// -- here straight from class_out / merge-chain code bytes + the read-span validity, before any dense
// float64 layer exists.  All integer results are exact (64-bit counters).
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "smf_hmm.h"

namespace smf {

constexpr int kColThreads = 256;

struct ColArgs {
  const uint8_t* cls;    // class_out or merge-chain code [N, V]
  const uint8_t* valid;  // [N, V] or nullptr
  int packed;
  long long N;
  int V;
  int n_rows;  // 1 + n_classes output rows: row 0 = any / member, row c + 1 = size class c
  unsigned long long* colsum;    // [n_rows][V]
  unsigned long long* colvalid;  // [V]
  int reads_per_block;
};

// Thread = one grid column over a block of reads; per-column counters live in shared memory
// ([row][thread]: conflict free), flushed once with 64-bit atomics.
__global__ void __launch_bounds__(kColThreads) layer_colsums_kernel(const __grid_constant__ ColArgs a) {
  extern __shared__ unsigned int cnt[];  // [(n_rows + 1)][kColThreads]; last row = valid count
  const int v = blockIdx.x * kColThreads + threadIdx.x;
  for (int r = 0; r <= a.n_rows; ++r) cnt[r * kColThreads + threadIdx.x] = 0u;
  const long long n0 = (long long)blockIdx.y * a.reads_per_block;
  const long long n1 = min(a.N, n0 + a.reads_per_block);
  if (v < a.V) {
    for (long long n = n0; n < n1; ++n) {
      const unsigned c = a.cls[n * a.V + v];
      bool ok = a.valid ? (a.valid[n * a.V + v] != 0) : true;
      unsigned cl, any;
      if (a.packed) {
        ok = ok && (c & 0x80u);
        cl = c & 15u;
        any = (c & 0x40u) ? 1u : 0u;
      } else {
        cl = c;
        any = c > 0u ? 1u : 0u;
      }
      if (ok) {
        cnt[a.n_rows * kColThreads + threadIdx.x] += 1u;
        cnt[threadIdx.x] += any;
        if (cl > 0u && (int)cl < a.n_rows) cnt[cl * kColThreads + threadIdx.x] += 1u;
      }
    }
    for (int r = 0; r < a.n_rows; ++r) {
      const unsigned x = cnt[r * kColThreads + threadIdx.x];
      if (x) atomicAdd(&a.colsum[(size_t)r * a.V + v], (unsigned long long)x);
    }
    const unsigned x = cnt[a.n_rows * kColThreads + threadIdx.x];
    if (x) atomicAdd(&a.colvalid[v], (unsigned long long)x);
  }
}

struct MetricArgs {
  const unsigned long long* colsum;    // [V]
  const unsigned long long* colvalid;  // [V]
  int V;
  int window;
  double* means;   // [V]
  double* metric;  // [V]
};

// nan-mean per column (0 where no read covers the column) and the 'same'-mode rolling mean of np.convolve.
__global__ void column_metric_kernel(const __grid_constant__ MetricArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int w = a.window;
  // np.convolve(..., "same") returns max(V, w) values: the centre of the full convolution, offset (min(V, w) - 1) / 2
  const int out_len = (w > 1 && w > a.V) ? w : a.V;
  if (i >= out_len) return;
  auto mean_at = [&](int j) -> double {
    const unsigned long long d = a.colvalid[j];
    return d ? (double)a.colsum[j] / (double)d : 0.0;
  };
  if (i < a.V) a.means[i] = mean_at(i);
  if (w <= 1) {
    a.metric[i] = mean_at(i);
    return;
  }
  // full[k] = sum_j means[j] * kern[k - j], same = full[off + i]; kern = 1 / w
  const double kv = 1.0 / (double)w;
  const int k = ((w < a.V ? w : a.V) - 1) / 2 + i;
  double s = 0.0;
  // numpy accumulates over the kernel index m = k - j from 0 upwards; keep that order
  for (int m = 0; m < w; ++m) {
    const int j = k - m;
    if (j >= 0 && j < a.V) s += mean_at(j) * kv;
  }
  a.metric[i] = s;
}

struct RunArgs {
  const uint8_t* cls;
  const uint8_t* valid;
  int packed;
  int select;  // -1: any / member, c >= 0: size class c
  long long N;
  int V;
  unsigned long long* count_hist;  // [V / 2 + 2]: reads with k runs
  unsigned long long* size_hist;   // [V + 1]: runs of s columns
};

__device__ __forceinline__ bool run_member(const RunArgs& a, long long idx) {
  const unsigned c = a.cls[idx];
  bool ok = a.valid ? (a.valid[idx] != 0) : true;
  if (a.packed) {
    ok = ok && (c & 0x80u);
    return ok && (a.select < 0 ? (c & 0x40u) != 0 : (int)(c & 15u) == a.select + 1);
  }
  return ok && (a.select < 0 ? c > 0u : (int)c == a.select + 1);
}

// One warp per read: 32 columns per step, membership ballot -> falling edges close runs.
__global__ void __launch_bounds__(256) run_histograms_kernel(const __grid_constant__ RunArgs a) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * 256LL + threadIdx.x) >> 5;
  const long long nwarps = (gridDim.x * 256LL) >> 5;
  for (long long n = warp0; n < a.N; n += nwarps) {
    const long long base = n * a.V;
    int open_start = -1;  // first column of the run that is open at the start of this word
    int runs = 0;
    unsigned cur = 0;
    {
      const int v = lane;
      cur = __ballot_sync(0xffffffffu, v < a.V && run_member(a, base + v));
    }
    for (int w0 = 0; w0 < a.V; w0 += 32) {
      unsigned nxt = 0;
      {
        const int v = w0 + 32 + lane;
        nxt = __ballot_sync(0xffffffffu, v < a.V && run_member(a, base + v));
      }
      // a run ends at bit i when bit i is set and the following column is clear
      const unsigned follow = (cur >> 1) | ((nxt & 1u) << 31);
      const unsigned ends = cur & ~follow;
      if ((ends >> lane) & 1u) {
        // start: one past the nearest clear bit below lane in this word, else the run came in open
        const unsigned below_clear = ~cur & ((lane == 0) ? 0u : (0xffffffffu >> (32 - lane)));
        const int start = below_clear ? (w0 + (31 - __clz(below_clear)) + 1) : (open_start >= 0 ? open_start : w0);
        const int size = w0 + lane + 1 - start;
        atomicAdd(&a.size_hist[size], 1ULL);
      }
      runs += __popc(ends);
      // run open across the boundary into the next word: remember where it began
      if ((cur >> 31) & 1u) {
        if (nxt & 1u) {
          const unsigned clear = ~cur;
          open_start = clear ? (w0 + (31 - __clz(clear)) + 1) : (open_start >= 0 ? open_start : w0);
        } else {
          open_start = -1;
        }
      } else {
        open_start = -1;
      }
      cur = nxt;
    }
    if (lane == 0) atomicAdd(&a.count_hist[runs], 1ULL);
  }
}

}  // namespace smf

using namespace smf;

extern "C" {

int smf_hmm_layer_colsums(const uint8_t* cls_or_code, int packed, const uint8_t* valid, int64_t n_reads,
                          int64_t n_grid, int n_classes, uint64_t* colsum, uint64_t* colvalid, void* stream) {
  if (n_reads < 0 || n_grid < 0 || n_classes < 0 || n_classes > 15) return SMF_ERR_BAD_ARG;
  if (!colsum || !colvalid) return SMF_ERR_BAD_ARG;
  if (n_grid > 0x3fffffff) return SMF_ERR_TOO_LONG;
  cudaStream_t st = (cudaStream_t)stream;
  const int n_rows = n_classes + 1;
  if (cudaMemsetAsync(colsum, 0, (size_t)n_rows * (size_t)n_grid * 8, st) != cudaSuccess) return SMF_ERR_CUDA;
  if (cudaMemsetAsync(colvalid, 0, (size_t)n_grid * 8, st) != cudaSuccess) return SMF_ERR_CUDA;
  if (n_reads == 0 || n_grid == 0) return SMF_OK;
  if (!cls_or_code) return SMF_ERR_BAD_ARG;
  ColArgs a;
  a.cls = cls_or_code;
  a.valid = valid;
  a.packed = packed;
  a.N = n_reads;
  a.V = (int)n_grid;
  a.n_rows = n_rows;
  a.colsum = reinterpret_cast<unsigned long long*>(colsum);
  a.colvalid = reinterpret_cast<unsigned long long*>(colvalid);
  const int col_blocks = (int)((n_grid + kColThreads - 1) / kColThreads);
  // enough CTAs to fill the machine, reads per CTA a multiple of 64
  long long want_y = (148LL * 16 + col_blocks - 1) / col_blocks;
  if (want_y < 1) want_y = 1;
  long long rpb = (n_reads + want_y - 1) / want_y;
  rpb = (rpb + 63) / 64 * 64;
  if (rpb < 64) rpb = 64;
  a.reads_per_block = (int)rpb;
  const long long gy = (n_reads + rpb - 1) / rpb;
  if (gy > 65535) return SMF_ERR_TOO_LONG;
  const size_t smem = (size_t)(n_rows + 1) * kColThreads * sizeof(unsigned int);
  layer_colsums_kernel<<<dim3((unsigned)col_blocks, (unsigned)gy), kColThreads, smem, st>>>(a);
  return cudaGetLastError() == cudaSuccess ? SMF_OK : SMF_ERR_CUDA;
}

int smf_hmm_column_metric(const uint64_t* colsum, const uint64_t* colvalid, int64_t n_grid, int rolling_window,
                          double* means, double* metric, void* stream) {
  if (n_grid < 0 || rolling_window < 1) return SMF_ERR_BAD_ARG;
  if (n_grid == 0) return SMF_OK;
  if (!colsum || !colvalid || !means || !metric) return SMF_ERR_BAD_ARG;
  if (n_grid > 0x3fffffff) return SMF_ERR_TOO_LONG;
  MetricArgs a;
  a.colsum = reinterpret_cast<const unsigned long long*>(colsum);
  a.colvalid = reinterpret_cast<const unsigned long long*>(colvalid);
  a.V = (int)n_grid;
  a.window = rolling_window;
  a.means = means;
  a.metric = metric;
  const int64_t out_len = (rolling_window > 1 && rolling_window > n_grid) ? rolling_window : n_grid;
  column_metric_kernel<<<(int)((out_len + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
  return cudaGetLastError() == cudaSuccess ? SMF_OK : SMF_ERR_CUDA;
}

int smf_hmm_run_histograms(const uint8_t* cls_or_code, int packed, int select_class, const uint8_t* valid,
                           int64_t n_reads, int64_t n_grid, uint64_t* count_hist, uint64_t* size_hist, void* stream) {
  if (n_reads < 0 || n_grid < 0 || select_class < -1 || select_class > 14) return SMF_ERR_BAD_ARG;
  if (!count_hist || !size_hist) return SMF_ERR_BAD_ARG;
  if (n_grid > 0x3fffffff) return SMF_ERR_TOO_LONG;
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(count_hist, 0, (size_t)(n_grid / 2 + 2) * 8, st) != cudaSuccess) return SMF_ERR_CUDA;
  if (cudaMemsetAsync(size_hist, 0, (size_t)(n_grid + 1) * 8, st) != cudaSuccess) return SMF_ERR_CUDA;
  if (n_reads == 0) return SMF_OK;
  if (n_grid == 0) {
    // every read has zero runs
    const unsigned long long n = (unsigned long long)n_reads;
    return cudaMemcpyAsync(count_hist, &n, 8, cudaMemcpyHostToDevice, st) == cudaSuccess ? SMF_OK : SMF_ERR_CUDA;
  }
  if (!cls_or_code) return SMF_ERR_BAD_ARG;
  RunArgs a;
  a.cls = cls_or_code;
  a.valid = valid;
  a.packed = packed;
  a.select = select_class;
  a.N = n_reads;
  a.V = (int)n_grid;
  a.count_hist = reinterpret_cast<unsigned long long*>(count_hist);
  a.size_hist = reinterpret_cast<unsigned long long*>(size_hist);
  long long blocks = (n_reads + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  run_histograms_kernel<<<(int)blocks, 256, 0, st>>>(a);
  return cudaGetLastError() == cudaSuccess ? SMF_OK : SMF_ERR_CUDA;
}

}  // extern "C"
