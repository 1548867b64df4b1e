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
  // fused chain (smf_hmm_merge_chain): one packed code byte per cell instead of merged/class_out/len_cls
  //   bits 0-3 size class + 1 (0 = none) | 0x40 member of a merged run | 0x80 inside the read span
  uint8_t* packed;              // [N, V] or nullptr
  const uint8_t* covered;       // [N, V] or nullptr: explicit coverage mask (HMM.py:193-204)
  const long long* span_coords; // [V] coordinates compared against the read span (HMM.py:178-191)
  const double* read_start;     // [N]
  const double* read_end;       // [N]
};

constexpr uint8_t kCodeMember = 0x40;
constexpr uint8_t kCodeValid = 0x80;

// ---------------------------------------------------------------------------------------
// merge_intervals_to_new_layer + write_size_class_layers_from_binary (HMM.py:1006-1123), optionally
// fused with mask_layers_outside_read_span (HMM.py:148-231): the chain of tools/partitioned_hmm.py:78-108.
//
// One CTA per read; the row lives on chip as BIT MASKS (V/32 words, a few KB for any realistic grid),
// so the kernel streams: 1 byte in, 1 + 4 (packed) bytes out per cell, no per-run tables.
//   A. membership words   mem[w]  (byte loads + ballot)
//   B. nxt[w] = first non-empty word >= w        (one warp, ballot scan over the words)
//   C. every falling edge (last cell p of a run) finds the next run start s with bit operations and
//      bridges the gap [p+1, s) into mrg[] iff its coordinate width <= the threshold.  The decision
//      only involves the two neighbouring runs, exactly as in the reference's chain merge where the
//      growing merged run always ends with the last run appended (HMM.py:1045-1057).
//   D. pz[w] / nz[w] = nearest word at or below / above w that has a clear bit in mrg
//   E. per cell: start / end of its merged run from clz / ffs on mrg (+ pz / nz across full words),
//      size class from the coordinates of the run's first and last cell, read-span validity; each lane
//      writes 4 consecutive cells (32-bit code store, 128-bit length store) when the row allows it.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ int mc_run_start(const uint32_t* mrg, const int* pz, int v) {
  const int w = v >> 5, b = v & 31;
  const uint32_t zeros = ~mrg[w] & ((1u << b) - 1u);  // clear bits below b
  if (zeros) return (w << 5) + (32 - __clz(zeros));
  const int w2 = (w > 0) ? pz[w - 1] : -1;
  if (w2 < 0) return 0;
  return (w2 << 5) + (32 - __clz(~mrg[w2]));
}
__device__ __forceinline__ int mc_run_end(const uint32_t* mrg, const int* nz, int v, int V, int W) {
  const int w = v >> 5, b = v & 31;
  const uint32_t zeros = ~mrg[w] & ~((2u << b) - 1u);  // clear bits above b (b == 31: none)
  if (zeros) return min((w << 5) + __ffs(zeros) - 1, V);
  const int w2 = (w + 1 < W) ? nz[w + 1] : W;
  if (w2 >= W) return V;
  return min((w2 << 5) + __ffs(~mrg[w2]) - 1, V);
}

__global__ void __launch_bounds__(kFeatThreads) merge_runs_kernel(const __grid_constant__ ClassTable ct,
                                                                   const __grid_constant__ MergeArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int V = a.V;
  const int W = (V + 31) >> 5;
  uint32_t* mem = reinterpret_cast<uint32_t*>(smem);  // [W]
  uint32_t* mrg = mem + W;                            // [W]
  int* nxt = reinterpret_cast<int*>(mrg + W);         // [W]
  int* pz = nxt + W;                                  // [W]
  int* nz = pz + W;                                   // [W]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool want_cls = (a.class_out != nullptr) || (a.packed != nullptr && ct.n > 0);
  const bool vec_ok = ((V & 3) == 0) && ((reinterpret_cast<uintptr_t>(a.merged_len) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(a.packed) & 3) == 0) &&
                      ((reinterpret_cast<uintptr_t>(a.merged) & 3) == 0) &&
                      ((reinterpret_cast<uintptr_t>(a.class_out) & 3) == 0) &&
                      ((reinterpret_cast<uintptr_t>(a.len_cls) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(a.covered) & 3) == 0);

  for (long long n = blockIdx.x; n < a.N; n += gridDim.x) {
    const uint8_t* row = a.layer + n * (long long)V;
    // ---- A: membership bit masks
    for (int w = warp; w < W; w += kFeatWarps) {
      const int v = (w << 5) + lane;
      const bool mbit = (v < V) && (__ldg(row + v) != 0);
      const uint32_t word = __ballot_sync(0xffffffffu, mbit);
      if (lane == 0) {
        mem[w] = word;
        mrg[w] = word;
      }
    }
    __syncthreads();
    // ---- B: next non-empty word (suffix scan by one warp)
    if (warp == 0) {
      int carry = W;
      for (int base = ((W - 1) >> 5) << 5; base >= 0; base -= 32) {
        const int w = base + lane;
        const bool f = (w < W) && (mem[w] != 0u);
        const uint32_t bal = __ballot_sync(0xffffffffu, f);
        const uint32_t up = bal >> lane;
        if (w < W) nxt[w] = up ? (w + __ffs(up) - 1) : carry;
        if (bal) carry = base + __ffs(bal) - 1;
      }
    }
    __syncthreads();
    // ---- C: bridge the gaps that are narrow enough
    for (int w = tid; w < W; w += kFeatThreads) {
      const uint32_t m = mem[w];
      if (m == 0u) continue;
      const uint32_t nb0 = (w + 1 < W) ? (mem[w + 1] & 1u) : 0u;
      uint32_t fall = m & ~((m >> 1) | (nb0 << 31));  // last cell of every run ending in this word
      while (fall) {
        const int b = __ffs(fall) - 1;
        fall &= fall - 1u;
        const int p = (w << 5) + b;
        const uint32_t rest = (b == 31) ? 0u : (m & ~((2u << b) - 1u));
        int s;
        if (rest) {
          s = (w << 5) + __ffs(rest) - 1;
        } else {
          const int w2 = (w + 1 < W) ? nxt[w + 1] : W;
          if (w2 >= W) continue;  // trailing zeros: no later run
          s = (w2 << 5) + __ffs(mem[w2]) - 1;
        }
        const long long gap = a.coords_are_ints ? (__ldg(a.coords + s) - __ldg(a.coords + p) - 1)
                                                : (long long)(s - p - 1);
        if (gap > a.dt) continue;
        // set bits (p, s) in mrg
        const int f0 = p + 1, l0 = s - 1;
        if (f0 > l0) continue;
        const int wf = f0 >> 5, wl = l0 >> 5;
        for (int ww = wf; ww <= wl; ++ww) {
          uint32_t bits = 0xffffffffu;
          if (ww == wf) bits &= ~((1u << (f0 & 31)) - 1u);
          if (ww == wl) bits &= ((l0 & 31) == 31) ? 0xffffffffu : ((2u << (l0 & 31)) - 1u);
          atomicOr(&mrg[ww], bits);
        }
      }
    }
    __syncthreads();
    // ---- D: nearest word with a clear bit, below (pz) and above (nz)
    if (warp == 0) {
      int carry = -1;
      for (int base = 0; base < W; base += 32) {
        const int w = base + lane;
        const bool f = (w < W) && (mrg[w] != 0xffffffffu);
        const uint32_t bal = __ballot_sync(0xffffffffu, f);
        const uint32_t dn = bal & (0xffffffffu >> (31 - lane));
        if (w < W) pz[w] = dn ? (base + 31 - __clz(dn)) : carry;
        if (bal) carry = base + 31 - __clz(bal);
      }
    } else if (warp == 1) {
      int carry = W;
      for (int base = ((W - 1) >> 5) << 5; base >= 0; base -= 32) {
        const int w = base + lane;
        const bool f = (w < W) && (mrg[w] != 0xffffffffu);
        const uint32_t bal = __ballot_sync(0xffffffffu, f);
        const uint32_t up = bal >> lane;
        if (w < W) nz[w] = up ? (w + __ffs(up) - 1) : carry;
        if (bal) carry = base + __ffs(bal) - 1;
      }
    }
    __syncthreads();
    // ---- E: outputs
    bool all_valid = false;
    long long sp = 0, ep = 0;
    if (a.packed != nullptr && a.covered == nullptr) {
      const double ds = a.read_start[n], de = a.read_end[n];
      if (!isfinite(ds) || !isfinite(de)) {
        all_valid = true;
      } else {
        sp = (long long)ds;  // int(start): truncation toward zero (HMM.py:219-220)
        ep = (long long)de;
      }
    }
    const long long roff = n * (long long)V;
    auto cell_valid = [&](int v) -> bool {
      if (a.covered != nullptr) return __ldg(a.covered + roff + v) != 0;
      if (all_valid) return true;
      const long long c = __ldg(a.span_coords + v);
      return !(c < sp || c > ep);
    };
    auto classify = [&](int S, int E) -> int {
      if (!want_cls) return -1;
      const double len_bp = a.coords_are_ints ? (double)(__ldg(a.coords + E - 1) - __ldg(a.coords + S) + 1)
                                              : (double)(E - S);
      return pick_class(ct, len_bp);
    };
    const int step = vec_ok ? 4 : 1;
    for (int v0 = tid * step; v0 < V; v0 += kFeatThreads * step) {
      const uint32_t bits = (mrg[v0 >> 5] >> (v0 & 31)) & (vec_ok ? 0xFu : 0x1u);
      int curE = -1, curLen = 0, curCls = -1;
      uint32_t code4 = 0, mem4 = 0, cls4 = 0;
      int len4[4] = {0, 0, 0, 0};
      int clen4[4] = {0, 0, 0, 0};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (k >= step) break;
        const int v = v0 + k;
        uint32_t code = 0;
        if ((bits >> k) & 1u) {
          if (v >= curE) {
            const int S = mc_run_start(mrg, pz, v);
            curE = mc_run_end(mrg, nz, v, V, W);
            curLen = curE - S;
            curCls = classify(S, curE);
          }
          len4[k] = curLen;
          code = kCodeMember | (curCls >= 0 ? (uint32_t)(curCls + 1) : 0u);
          mem4 |= 1u << (8 * k);
          if (curCls >= 0) {
            cls4 |= (uint32_t)(curCls + 1) << (8 * k);
            clen4[k] = curLen;
          }
        }
        if (a.packed != nullptr && cell_valid(v)) code |= kCodeValid;
        code4 |= code << (8 * k);
      }
      if (vec_ok) {
        *reinterpret_cast<int4*>(a.merged_len + roff + v0) = make_int4(len4[0], len4[1], len4[2], len4[3]);
        if (a.packed != nullptr) {
          *reinterpret_cast<uint32_t*>(a.packed + roff + v0) = code4;
        } else {
          *reinterpret_cast<uint32_t*>(a.merged + roff + v0) = mem4;
          if (a.class_out) *reinterpret_cast<uint32_t*>(a.class_out + roff + v0) = cls4;
          if (a.len_cls)
            *reinterpret_cast<int4*>(a.len_cls + roff + v0) = make_int4(clen4[0], clen4[1], clen4[2], clen4[3]);
        }
      } else {
        a.merged_len[roff + v0] = len4[0];
        if (a.packed != nullptr) {
          a.packed[roff + v0] = (uint8_t)code4;
        } else {
          a.merged[roff + v0] = (uint8_t)mem4;
          if (a.class_out) a.class_out[roff + v0] = (uint8_t)cls4;
          if (a.len_cls) a.len_cls[roff + v0] = clen4[0];
        }
      }
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
