// Footprint / feature calling kernels (K4 call_features, K5 merge + size classes, K6 span mask).
//
// One CTA owns one read (row).  Runs of a boolean row are enumerated by stream compaction:
// every thread classifies one cell as run-start / run-end, warp ballots + a CTA prefix give
// each boundary its ordinal, and the (start, end) pairs land in shared memory in order.
// Dense layers are then produced with coalesced row writes; each cell finds its run by a
// binary search over the (sorted) run starts held in shared memory.
#pragma once
#include "hmm_common.cuh"

namespace smf {

constexpr int kFeatThreads = 256;
constexpr int kFeatWarps = kFeatThreads / 32;
constexpr int kMaxClasses = 16;

struct ClassTable {
  double lo[kMaxClasses];
  double hi[kMaxClasses];
  int n;
};

// First size class with lo <= len < hi (HMM.py:1335-1338, 1107-1108); -1 if none.
__device__ __forceinline__ int pick_class(const ClassTable& ct, double len) {
  int cls = -1;
  for (int c = ct.n - 1; c >= 0; --c)
    if (ct.lo[c] <= len && len < ct.hi[c]) cls = c;
  return cls;
}

// Enumerate the maximal runs of pred over [0, n) into rs[] / re[] (end exclusive), in order.
// Equivalent to BaseHMM._runs_from_bool (HMM.py:927-938).  s_cnt: int[2*kFeatWarps + 2].
template <class Pred>
__device__ int enumerate_runs(Pred pred, int n, int* rs, int* re, int* s_cnt) {
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  int total_s = 0, total_e = 0;  // uniform across the CTA
  for (int base = 0; base <= n; base += kFeatThreads) {
    const int v = base + tid;
    const bool cur = (v < n) && pred(v);
    const bool prev = (v > 0) && (v <= n) && pred(v - 1);
    const bool is_start = cur && !prev;
    const bool is_end = prev && !cur;  // run ended at v (exclusive)
    const unsigned bs = __ballot_sync(0xffffffffu, is_start);
    const unsigned be = __ballot_sync(0xffffffffu, is_end);
    if (lane == 0) {
      s_cnt[w] = __popc(bs);
      s_cnt[kFeatWarps + w] = __popc(be);
    }
    __syncthreads();
    int off_s = 0, off_e = 0, blk_s = 0, blk_e = 0;
    for (int i = 0; i < kFeatWarps; ++i) {
      const int cs = s_cnt[i], ce = s_cnt[kFeatWarps + i];
      if (i < w) {
        off_s += cs;
        off_e += ce;
      }
      blk_s += cs;
      blk_e += ce;
    }
    const unsigned lt = (1u << lane) - 1u;
    if (is_start) rs[total_s + off_s + __popc(bs & lt)] = v;
    if (is_end) re[total_e + off_e + __popc(be & lt)] = v;
    total_s += blk_s;
    total_e += blk_e;
    __syncthreads();
  }
  return total_s;
}

// Largest r in [0, R) with rs[r] <= v, or -1.
__device__ __forceinline__ int find_run(const int* rs, int R, int v) {
  int lo = 0, hi = R;  // first index with rs > v
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (rs[mid] <= v) lo = mid + 1; else hi = mid;
  }
  return lo - 1;
}

struct FeatArgs {
  const int8_t* states;  // [N, L] or nullptr
  const float* post;     // [N, L] or nullptr
  int target;
  float thr;
  long long N;
  int L;
  const long long* coords;  // [L] site coordinates (strictly increasing)
  const int* grid_left;     // [L] searchsorted(full, coords[t], left)
  const int* grid_right;    // [L] searchsorted(full, coords[t], right)
  int V;
  uint8_t* class_out;  // [N, V]
  int* run_len;        // [N, V]
  int* intervals;      // [max,4] or nullptr
  long long max_intervals;
  unsigned long long* n_intervals;
};

__global__ void __launch_bounds__(kFeatThreads) call_features_kernel(const __grid_constant__ ClassTable ct,
                                                                      const __grid_constant__ FeatArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int maxruns = (a.L + 1) / 2 + 1;
  int* rs = reinterpret_cast<int*>(smem);
  int* re = rs + maxruns;
  int* s_cnt = re + maxruns;  // 2*W+2
  int8_t* rcls = reinterpret_cast<int8_t*>(s_cnt + 2 * kFeatWarps + 2);
  const int tid = threadIdx.x, lane = tid & 31;

  for (long long n = blockIdx.x; n < a.N; n += gridDim.x) {
    const int8_t* srow = a.states ? a.states + n * (long long)a.L : nullptr;
    const float* prow = a.post ? a.post + n * (long long)a.L : nullptr;
    const int target = a.target;
    const float thr = a.thr;
    auto pred = [&](int t) -> bool {
      return prow ? (__ldg(prow + t) >= thr) : ((int)__ldg(srow + t) == target);
    };
    const int R = enumerate_runs(pred, a.L, rs, re, s_cnt);
    // classify runs; rs/re become grid spans
    for (int r0 = 0; r0 < R; r0 += kFeatThreads) {
      const int r = r0 + tid;
      bool keep = false;
      int gs = 0, ge = 0, cls = -1;
      if (r < R) {
        const int s = rs[r], e = re[r];
        const long long glen = __ldg(a.coords + e - 1) - __ldg(a.coords + s) + 1;
        gs = __ldg(a.grid_left + s);
        ge = __ldg(a.grid_right + e - 1);
        if (glen > 0) cls = pick_class(ct, (double)glen);
        keep = (cls >= 0) && (gs < ge);
        if (!keep) cls = -1;
      }
      if (r < R) {  // each thread rewrites only the entries it just read
        rs[r] = gs;
        re[r] = ge;
        rcls[r] = (int8_t)cls;
      }
      if (a.intervals != nullptr) {
        const unsigned bk = __ballot_sync(0xffffffffu, keep);
        unsigned long long basei = 0;
        if (lane == 0 && bk) basei = atomicAdd(a.n_intervals, (unsigned long long)__popc(bk));
        basei = __shfl_sync(0xffffffffu, basei, 0);
        if (keep) {
          const unsigned long long slot = basei + __popc(bk & ((1u << lane) - 1u));
          if (slot < (unsigned long long)a.max_intervals) {
            int* dst = a.intervals + slot * 4;
            dst[0] = (int)n;
            dst[1] = gs;
            dst[2] = ge;
            dst[3] = cls;
          }
        }
      }
    }
    __syncthreads();
    if (a.class_out != nullptr) {  // dense layers are optional: interval records alone are enough downstream
      uint8_t* crow = a.class_out + n * (long long)a.V;
      int* lrow = a.run_len ? a.run_len + n * (long long)a.V : nullptr;
      for (int v = tid; v < a.V; v += kFeatThreads) {
        const int r = find_run(rs, R, v);
        uint8_t c = 0;
        int len = 0;
        if (r >= 0 && v < re[r] && rcls[r] >= 0) {
          c = (uint8_t)(rcls[r] + 1);
          len = re[r] - rs[r];
        }
        crow[v] = c;
        if (lrow) lrow[v] = len;
      }
    }
    __syncthreads();
  }
}

// Interval-only feature calling: one WARP per read, no shared memory and no CTA barriers.
// 32 positions per step: the membership ballot gives run starts / ends as bit masks; the lane
// that sees a run end finds the matching start in the same mask (or the start carried over from
// earlier words), classifies the run and appends one (read, grid_start, grid_end, class) record
// through a warp-aggregated atomic.  Same records as call_features_kernel.
__global__ void __launch_bounds__(kFeatThreads) call_intervals_kernel(const __grid_constant__ ClassTable ct,
                                                                       const __grid_constant__ FeatArgs a) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const unsigned lt = (1u << lane) - 1u;
  for (long long n = warp0; n < a.N; n += nwarps) {
    const int8_t* srow = a.states ? a.states + n * (long long)a.L : nullptr;
    const float* prow = a.post ? a.post + n * (long long)a.L : nullptr;
    unsigned carry = 0;   // membership of the position just before this word
    int open_start = 0;   // start of the run that is open when the word begins
    for (int base = 0; base <= a.L; base += 32) {
      const int t = base + lane;
      bool cur = false;
      if (t < a.L) cur = prow ? (__ldg(prow + t) >= a.thr) : ((int)__ldg(srow + t) == a.target);
      const unsigned m = __ballot_sync(0xffffffffu, cur);
      const unsigned shifted = (m << 1) | carry;
      const unsigned starts = m & ~shifted;
      const unsigned ends = ~m & shifted;  // bit i: a run ended at position base+i (exclusive)
      bool keep = false;
      int gs = 0, ge = 0, cls = -1;
      if ((ends >> lane) & 1u) {
        const unsigned before = starts & lt;
        const int s0 = before ? (base + 31 - __clz(before)) : open_start;
        const int e0 = t;
        const long long glen = __ldg(a.coords + e0 - 1) - __ldg(a.coords + s0) + 1;
        gs = __ldg(a.grid_left + s0);
        ge = __ldg(a.grid_right + e0 - 1);
        if (glen > 0) cls = pick_class(ct, (double)glen);
        keep = (cls >= 0) && (gs < ge);
      }
      const unsigned bk = __ballot_sync(0xffffffffu, keep);
      if (bk) {
        unsigned long long basei = 0;
        if (lane == 0) basei = atomicAdd(a.n_intervals, (unsigned long long)__popc(bk));
        basei = __shfl_sync(0xffffffffu, basei, 0);
        if (keep) {
          const unsigned long long slot = basei + __popc(bk & lt);
          if (slot < (unsigned long long)a.max_intervals) {
            int4 rec = make_int4((int)n, gs, ge, cls);
            *reinterpret_cast<int4*>(a.intervals + slot * 4) = rec;
          }
        }
      }
      if (starts) {
        const int hs = 31 - __clz(starts);
        if (ends == 0u || hs > 31 - __clz(ends)) open_start = base + hs;  // that run is still open
      }
      carry = m >> 31;
    }
  }
}

struct MergeArgs {
  const uint8_t* layer;  // [N, V]
  long long N;
  int V;
  const long long* coords;  // [V]
  int coords_are_ints;
  long long dt;
  uint8_t* merged;     // [N, V]
  int* merged_len;     // [N, V]
  uint8_t* class_out;  // [N, V] or nullptr
  int* len_cls;        // [N, V] or nullptr
};

// merge_intervals_to_new_layer + write_size_class_layers_from_binary (HMM.py:1006-1123).
__global__ void __launch_bounds__(kFeatThreads) merge_runs_kernel(const __grid_constant__ ClassTable ct,
                                                                   const __grid_constant__ MergeArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int maxruns = (a.V + 1) / 2 + 1;
  int* rs = reinterpret_cast<int*>(smem);
  int* re = rs + maxruns;
  int* sgs = re + maxruns;   // segment start (grid) per run
  int* sge = sgs + maxruns;  // segment end per run
  int* s_cnt = sge + maxruns;
  int8_t* rcls = reinterpret_cast<int8_t*>(s_cnt + 2 * kFeatWarps + 2);
  const int tid = threadIdx.x;

  for (long long n = blockIdx.x; n < a.N; n += gridDim.x) {
    const uint8_t* row = a.layer + n * (long long)a.V;
    auto pred = [&](int v) -> bool { return __ldg(row + v) != 0; };
    const int R = enumerate_runs(pred, a.V, rs, re, s_cnt);
    // Sequential chain merge (gap measured against the growing merged run, HMM.py:1045-1057).
    // One warp walks the runs 32 at a time: a run starts a new segment iff its gap > dt.
    if (tid < 32) {
      int seg_first = 0;  // index of the first run of the open segment
      for (int r0 = 0; r0 < R; r0 += 32) {
        const int r = r0 + tid;
        bool starts = false;
        if (r < R) {
          if (r == 0) {
            starts = true;
          } else {
            const long long gap = a.coords_are_ints
                                      ? (__ldg(a.coords + rs[r]) - __ldg(a.coords + re[r - 1] - 1) - 1)
                                      : (long long)(rs[r] - re[r - 1]);
            starts = gap > a.dt;
          }
        }
        const unsigned bsx = __ballot_sync(0xffffffffu, starts);
        if (r < R) {
          // first run of my segment: highest start flag at or below my lane, else carry-in
          const unsigned le = bsx & (0xffffffffu >> (31 - tid));
          const int f = le ? (r0 + 31 - __clz(le)) : seg_first;
          sgs[r] = f;
        }
        if (bsx) seg_first = r0 + 31 - __clz(bsx);
      }
    }
    __syncthreads();
    // sgs[r] = index of first run of r's segment.  Segment end: last run l with sgs[l] == sgs[r];
    // found from the next segment's first run (scan forward over run indices in parallel).
    for (int r = tid; r < R; r += kFeatThreads) {
      // binary search the first run index whose segment-first exceeds mine
      const int f = sgs[r];
      int lo = r + 1, hi = R;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (sgs[mid] > f) hi = mid; else lo = mid + 1;
      }
      sge[r] = re[lo - 1];
    }
    __syncthreads();
    for (int r = tid; r < R; r += kFeatThreads) {
      const int S = rs[sgs[r]], E = sge[r];
      int cls = -1;
      if (a.class_out != nullptr) {
        const double len_bp = a.coords_are_ints
                                  ? (double)(__ldg(a.coords + E - 1) - __ldg(a.coords + S) + 1)
                                  : (double)(E - S);
        cls = pick_class(ct, len_bp);
      }
      rcls[r] = (int8_t)cls;
    }
    __syncthreads();
    for (int r = tid; r < R; r += kFeatThreads) sgs[r] = rs[sgs[r]];  // now the grid start
    __syncthreads();
    uint8_t* mrow = a.merged + n * (long long)a.V;
    int* lrow = a.merged_len + n * (long long)a.V;
    uint8_t* crow = a.class_out ? a.class_out + n * (long long)a.V : nullptr;
    int* clrow = a.len_cls ? a.len_cls + n * (long long)a.V : nullptr;
    for (int v = tid; v < a.V; v += kFeatThreads) {
      const int r = find_run(rs, R, v);
      uint8_t m = 0, c = 0;
      int len = 0;
      if (r >= 0 && v < sge[r]) {  // inside r's merged segment (run cell or joined gap)
        m = 1;
        len = sge[r] - sgs[r];
        if (rcls[r] >= 0) c = (uint8_t)(rcls[r] + 1);
      }
      mrow[v] = m;
      lrow[v] = len;
      if (crow) crow[v] = c;
      if (clrow) clrow[v] = c ? len : 0;
    }
    __syncthreads();
  }
}

struct SpanArgs {
  const uint8_t* covered;  // [N, V] or nullptr
  const long long* coords; // [V]
  const double* start;     // [N]
  const double* end;       // [N]
  long long N;
  int V;
  uint8_t* valid;  // [N, V]
};

// mask_layers_outside_read_span (HMM.py:193-226) as a validity bitmap.
__global__ void __launch_bounds__(256) span_mask_kernel(const __grid_constant__ SpanArgs a) {
  for (long long n = blockIdx.y; n < a.N; n += gridDim.y) {
    bool all_valid = false;
    long long s = 0, e = 0;
    if (a.covered == nullptr) {
      const double ds = a.start[n], de = a.end[n];
      if (!isfinite(ds) || !isfinite(de)) {
        all_valid = true;
      } else {
        s = (long long)ds;  // int(start): truncation toward zero
        e = (long long)de;
      }
    }
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < a.V; v += gridDim.x * blockDim.x) {
      bool ok;
      if (a.covered != nullptr) {
        ok = a.covered[n * (long long)a.V + v] != 0;
      } else {
        const long long c = __ldg(a.coords + v);
        ok = all_valid || !(c < s || c > e);
      }
      a.valid[n * (long long)a.V + v] = ok ? 1 : 0;
    }
  }
}

}  // namespace smf
