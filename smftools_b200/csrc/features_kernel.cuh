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

// Vectorised variant of call_intervals_kernel for 16-byte aligned rows (L % 16 == 0): a warp still owns
// a read, but walks it in super-blocks of 1024 positions -- each lane loads its 32 positions with
// 128-bit loads and folds them into one membership word with byte-compare / nibble-gather
// arithmetic (1 instruction per position instead of a load + compare + ballot per position).
// A run end at bit i of a lane's word finds its start from the clear bits below i, or -- when the
// run began before the word -- from a warp-level exclusive max-scan of the lanes' last clear
// positions (plus the carry from earlier super-blocks).  Lanes with several ends take them in rounds;
// each round appends its records through one warp-aggregated atomic.  Same records as
// call_intervals_kernel (order within the output is unspecified in both).
template <bool POST>
__global__ void __launch_bounds__(kFeatThreads) call_intervals_vec_kernel(const __grid_constant__ ClassTable ct,
                                                                           const __grid_constant__ FeatArgs a) {
  constexpr unsigned kFull = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const unsigned lt = (1u << lane) - 1u;
  const int L = a.L;
  const uint32_t target4 = (uint32_t)(uint8_t)a.target * 0x01010101u;
  for (long long n = warp0; n < a.N; n += nwarps) {
    const int8_t* srow = POST ? nullptr : a.states + n * (long long)L;
    const float* prow = POST ? a.post + n * (long long)L : nullptr;
    int carry_lz = -1;        // last non-member position before this super-block (-1: none)
    unsigned carry_bit = 0;   // membership of the position just before this super-block
    for (int base = 0; base <= L; base += 1024) {
      const int p0 = base + 32 * lane;
      unsigned m = 0;
      if (!POST) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (p0 + 16 * h < L) {
            const uint4 x = __ldg(reinterpret_cast<const uint4*>(srow + p0 + 16 * h));
            const uint32_t c[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int q = 0; q < 4; ++q)
              m |= ((((__vcmpeq4(c[q], target4)) & 0x01010101u) * 0x01020408u) >> 24) << (16 * h + 4 * q);
          }
        }
      } else {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          if (p0 + 4 * q < L) {
            const float4 x = __ldg(reinterpret_cast<const float4*>(prow + p0 + 4 * q));
            const unsigned nib = (x.x >= a.thr ? 1u : 0u) | (x.y >= a.thr ? 2u : 0u) | (x.z >= a.thr ? 4u : 0u) |
                                 (x.w >= a.thr ? 8u : 0u);
            m |= nib << (4 * q);
          }
        }
      }
      // membership of the position before my word
      unsigned prevbit = __shfl_up_sync(kFull, m >> 31, 1);
      if (lane == 0) prevbit = carry_bit;
      unsigned ends = ~m & ((m << 1) | prevbit);  // bit i: a run ended at position p0 + i (exclusive)
      // last clear position at or before the end of each lane's word: exclusive max-scan over lanes
      const unsigned z = ~m;
      const int lz = z ? (p0 + 31 - __clz(z)) : -1;
      int inc = lz;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(kFull, inc, d);
        if (lane >= d) inc = max(inc, o);
      }
      int before = __shfl_up_sync(kFull, inc, 1);
      if (lane == 0) before = -1;
      before = max(before, carry_lz);
      // run ends, one per lane and round
      while (__any_sync(kFull, ends != 0u)) {
        bool keep = false;
        int gs = 0, ge = 0, cls = -1;
        if (ends) {
          const int i = __ffs(ends) - 1;
          ends &= ends - 1u;
          const int e0 = p0 + i;
          const unsigned zb = z & ((1u << i) - 1u);  // clear bits below the end
          const int s0 = (zb ? (p0 + 31 - __clz(zb)) : before) + 1;
          const long long glen = __ldg(a.coords + e0 - 1) - __ldg(a.coords + s0) + 1;
          gs = __ldg(a.grid_left + s0);
          ge = __ldg(a.grid_right + e0 - 1);
          if (glen > 0) cls = pick_class(ct, (double)glen);
          keep = (cls >= 0) && (gs < ge);
        }
        const unsigned bk = __ballot_sync(kFull, keep);
        if (bk) {
          unsigned long long basei = 0;
          if (lane == 0) basei = atomicAdd(a.n_intervals, (unsigned long long)__popc(bk));
          basei = __shfl_sync(kFull, basei, 0);
          if (keep) {
            const unsigned long long slot = basei + __popc(bk & lt);
            if (slot < (unsigned long long)a.max_intervals)
              *reinterpret_cast<int4*>(a.intervals + slot * 4) = make_int4((int)n, gs, ge, cls);
          }
        }
      }
      carry_lz = max(carry_lz, __shfl_sync(kFull, inc, 31));
      carry_bit = __shfl_sync(kFull, m >> 31, 31);
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
//   A. membership words   mem[w]  (32-bit loads of 4 cells, nibbles OR-reduced over 8 lanes)
//   B. nxt[w] = first non-empty word >= w        (one warp, ballot scan over the words)
//   C. every falling edge (last cell p of a run) finds the next run start s with bit operations and
//      bridges the gap [p+1, s) into mrg[] iff its coordinate width <= the threshold.  The decision
//      only involves the two neighbouring runs, exactly as in the reference's chain merge where the
//      growing merged run always ends with the last run appended (HMM.py:1045-1057).
//   D. pz[w] / nz[w] = nearest word at or below / above w that has a clear bit in mrg
//   E. ONE LANE PER WORD: the lane walks the (few) sub-runs of its 32 cells, gets the start / end of
//      a run that crosses the word boundary from clz / ffs on mrg (+ pz / nz across full words), the
//      size class from the coordinates of the run's first and last cell, the read-span validity as
//      a 32-bit mask, and writes code bytes + run lengths into a per-warp staging tile (padded,
//      conflict free); the warp then copies the tile out with coalesced 128-bit stores.
// ---------------------------------------------------------------------------------------
constexpr int kMcThreads = 128;  // threads per CTA of merge_runs_kernel
constexpr int kMcWarps = kMcThreads / 32;
constexpr int kMcCodeStride = 48;   // bytes per lane in the code tile (32 used)
constexpr int kMcLenStride = 144;   // bytes per lane in the length tile (128 used)
constexpr int kMcTileBytes = 32 * (kMcCodeStride + kMcLenStride);

struct IntClassTable {
  long long lo[kMaxClasses];  // ceil(lo): len >= lo  <=>  len >= ceil(lo) for integer len
  long long hi[kMaxClasses];  // ceil(hi): len <  hi  <=>  len <  ceil(hi)
  int n;
};
__device__ __forceinline__ int pick_class_int(const IntClassTable& ct, long long len) {
  int cls = -1;
  for (int c = ct.n - 1; c >= 0; --c)
    if (ct.lo[c] <= len && len < ct.hi[c]) cls = c;
  return cls;
}

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
// bits [lo, hi) of a 32-bit word, for any integers lo <= hi (clamped to the word)
__device__ __forceinline__ uint32_t mc_range_mask(int lo, int hi) {
  lo = max(lo, 0);
  hi = min(hi, 32);
  if (lo >= hi) return 0u;
  const uint32_t upto_hi = (hi == 32) ? 0xffffffffu : ((1u << hi) - 1u);
  return upto_hi & ~((1u << lo) - 1u);
}
// the four low bits of x spread to the low bit of four bytes
__device__ __forceinline__ uint32_t mc_spread4(uint32_t x) { return ((x & 0xFu) * 0x00204081u) & 0x01010101u; }

__global__ void __launch_bounds__(kMcThreads) merge_runs_kernel(const __grid_constant__ IntClassTable ct,
                                                                   const __grid_constant__ MergeArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int V = a.V;
  const int W = (V + 31) >> 5;
  const int Wpad = (W + 3) & ~3;
  uint32_t* mem = reinterpret_cast<uint32_t*>(smem);  // [Wpad]
  uint32_t* mrg = mem + Wpad;                         // [Wpad]
  int* nxt = reinterpret_cast<int*>(mrg + Wpad);      // [Wpad]
  int* pz = nxt + Wpad;                               // [Wpad]
  int* nz = pz + Wpad;                                // [Wpad]
  __shared__ int s_cnt[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned char* tile = smem + (size_t)Wpad * 20 + (size_t)warp * kMcTileBytes;
  unsigned char* tcode = tile;
  unsigned char* tlen = tile + 32 * kMcCodeStride;
  const bool want_cls = (ct.n > 0) && ((a.class_out != nullptr) || (a.packed != nullptr));
  const bool packed = (a.packed != nullptr);
  const bool in_vec = ((V & 3) == 0) && ((reinterpret_cast<uintptr_t>(a.layer) & 3) == 0);
  const bool out_vec = ((V & 3) == 0) && ((reinterpret_cast<uintptr_t>(a.merged_len) & 15) == 0) &&
                       ((reinterpret_cast<uintptr_t>(a.packed) & 3) == 0) &&
                       ((reinterpret_cast<uintptr_t>(a.merged) & 3) == 0) &&
                       ((reinterpret_cast<uintptr_t>(a.class_out) & 3) == 0) &&
                       ((reinterpret_cast<uintptr_t>(a.len_cls) & 15) == 0);
  const bool cov_vec = ((V & 3) == 0) && ((reinterpret_cast<uintptr_t>(a.covered) & 3) == 0);

  // Non-decreasing span coordinates (the normal case: genomic positions) turn the read-span test
  // of a row into one column range; otherwise every cell is tested against its coordinate.
  bool span_sorted = false;
  if (packed && a.covered == nullptr) {
    bool ok = true;
    for (int v = tid + 1; v < V; v += kMcThreads) ok = ok && (__ldg(a.span_coords + v) >= __ldg(a.span_coords + v - 1));
    span_sorted = __syncthreads_and(ok) != 0;
  }

  for (long long n = blockIdx.x; n < a.N; n += gridDim.x) {
    const long long roff = n * (long long)V;
    const uint8_t* row = a.layer + roff;
    // ---- A: membership bit masks
    if (in_vec) {
      for (int base = warp * 128; base < V; base += kMcWarps * 128) {
        const int v = base + lane * 4;
        const uint32_t x = (v < V) ? __ldg(reinterpret_cast<const uint32_t*>(row + v)) : 0u;
        const uint32_t ne = __vcmpne4(x, 0u);  // 0xff per non-zero byte
        uint32_t word = ((ne & 0x01010101u) * 0x01020408u) >> 24;  // 4 cells -> nibble
        word <<= 4 * (lane & 7);
        word |= __shfl_xor_sync(0xffffffffu, word, 1);
        word |= __shfl_xor_sync(0xffffffffu, word, 2);
        word |= __shfl_xor_sync(0xffffffffu, word, 4);
        const int w = (base >> 5) + (lane >> 3);
        if ((lane & 7) == 0 && w < W) {
          mem[w] = word;
          mrg[w] = word;
        }
      }
    } else {
      for (int w = warp; w < W; w += kMcWarps) {
        const int v = (w << 5) + lane;
        const bool mbit = (v < V) && (__ldg(row + v) != 0);
        const uint32_t word = __ballot_sync(0xffffffffu, mbit);
        if (lane == 0) {
          mem[w] = word;
          mrg[w] = word;
        }
      }
    }
    // read-span column range of this row (sorted coordinates): two-level counting search
    int vlo = 0, vhi = V;
    bool all_valid = false;
    long long sp = 0, ep = 0;
    if (packed && a.covered == nullptr) {
      const double ds = a.read_start[n], de = a.read_end[n];
      if (!isfinite(ds) || !isfinite(de)) {
        all_valid = true;
      } else {
        sp = (long long)ds;  // int(start): truncation toward zero (HMM.py:219-220)
        ep = (long long)de;
      }
      if (span_sorted && !all_valid) {
        // vlo = #cells with c < sp, vhi = #cells with c <= ep
        const int seg = (V + kMcThreads - 1) / kMcThreads;
        const int pos = min(tid * seg, V - 1);
        const long long c0 = __ldg(a.span_coords + pos);
        const int segs_lo = __syncthreads_count(tid * seg < V && c0 < sp);   // segments starting below sp
        const int segs_hi = __syncthreads_count(tid * seg < V && c0 <= ep);
        // the boundary lies inside segment segs-1 (if any): count inside it
        const int blo = max(segs_lo - 1, 0) * seg, bhi = max(segs_hi - 1, 0) * seg;
        int cl = 0, ch = 0;
        for (int i = tid; i < seg; i += kMcThreads) {
          if (blo + i < V && __ldg(a.span_coords + blo + i) < sp) ++cl;
          if (bhi + i < V && __ldg(a.span_coords + bhi + i) <= ep) ++ch;
        }
        if (tid == 0) {
          s_cnt[0] = 0;
          s_cnt[1] = 0;
        }
        __syncthreads();
        if (cl) atomicAdd(&s_cnt[0], cl);
        if (ch) atomicAdd(&s_cnt[1], ch);
        __syncthreads();
        vlo = (segs_lo > 0) ? blo + s_cnt[0] : 0;
        vhi = (segs_hi > 0) ? bhi + s_cnt[1] : 0;
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
    for (int w = tid; w < W; w += kMcThreads) {
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
        const int f0 = p + 1, l0 = s - 1;  // set bits (p, s) in mrg
        if (f0 > l0) continue;
        const int wf = f0 >> 5, wl = l0 >> 5;
        for (int ww = wf; ww <= wl; ++ww)
          atomicOr(&mrg[ww], mc_range_mask(f0 - (ww << 5), l0 + 1 - (ww << 5)));
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
    // ---- E: one lane per word -> staging tile -> coalesced copy-out
    for (int blk = warp; blk * 32 < W; blk += kMcWarps) {
      const int w = blk * 32 + lane;
      const int c0 = w << 5;  // first cell of my word
      uint32_t m = (w < W) ? mrg[w] : 0u;
      // validity bits of my 32 cells
      uint32_t vb = 0u;
      if (packed && w < W) {
        if (a.covered != nullptr) {
          if (cov_vec) {
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              const int v = c0 + 4 * g;
              const uint32_t x = (v < V) ? __ldg(reinterpret_cast<const uint32_t*>(a.covered + roff + v)) : 0u;
              vb |= ((((__vcmpne4(x, 0u)) & 0x01010101u) * 0x01020408u) >> 24) << (4 * g);
            }
          } else {
            for (int i = 0; i < 32 && c0 + i < V; ++i) vb |= (__ldg(a.covered + roff + c0 + i) != 0 ? 1u : 0u) << i;
          }
        } else if (all_valid) {
          vb = 0xffffffffu;
        } else if (span_sorted) {
          vb = mc_range_mask(vlo - c0, vhi - c0);
        } else {
          for (int i = 0; i < 32 && c0 + i < V; ++i) {
            const long long c = __ldg(a.span_coords + c0 + i);
            vb |= (!(c < sp || c > ep) ? 1u : 0u) << i;
          }
        }
      }
      // default fill: no run -> length 0, code = validity bit
      uint4* lrow = reinterpret_cast<uint4*>(tlen + lane * kMcLenStride);
#pragma unroll
      for (int g = 0; g < 8; ++g) lrow[g] = make_uint4(0u, 0u, 0u, 0u);
      uint32_t cw[8];
#pragma unroll
      for (int g = 0; g < 8; ++g) cw[g] = mc_spread4(vb >> (4 * g)) * (uint32_t)kCodeValid;
      uint4* crow = reinterpret_cast<uint4*>(tcode + lane * kMcCodeStride);
      crow[0] = make_uint4(cw[0], cw[1], cw[2], cw[3]);
      crow[1] = make_uint4(cw[4], cw[5], cw[6], cw[7]);
      // sub-runs of the word
      int* lcell = reinterpret_cast<int*>(tlen + lane * kMcLenStride);
      unsigned char* ccell = tcode + lane * kMcCodeStride;
      while (m) {
        const int b0 = __ffs(m) - 1;
        const uint32_t above = ~m & ~((1u << b0) - 1u);  // clear bits at or above b0
        const int b1 = above ? (__ffs(above) - 2) : 31;  // last cell of the sub-run
        m &= ~mc_range_mask(b0, b1 + 1);
        int S = c0 + b0, E = c0 + b1 + 1;
        if (b0 == 0 && w > 0 && (mrg[w - 1] >> 31)) S = mc_run_start(mrg, pz, c0);
        if (b1 == 31 && w + 1 < W && (mrg[w + 1] & 1u)) E = mc_run_end(mrg, nz, c0 + 31, V, W);
        E = min(E, V);
        const int len = E - S;
        int cls = -1;
        if (want_cls) {
          const long long len_bp =
              a.coords_are_ints ? (__ldg(a.coords + E - 1) - __ldg(a.coords + S) + 1) : (long long)(E - S);
          cls = pick_class_int(ct, len_bp);
        }
        const unsigned char cd = (unsigned char)(kCodeMember | (cls >= 0 ? (cls + 1) : 0));
        // fully covered groups of 4 cells get vector stores, the ragged ends cell stores
        const int gs = (b0 + 3) >> 2, ge = ((b1 + 1) >> 2) - 1;
        const int head_end = (gs <= ge) ? 4 * gs : b1 + 1;
        for (int i = b0; i < head_end; ++i) {
          lcell[i] = len;
          ccell[i] = cd | (((vb >> i) & 1u) ? kCodeValid : 0);
        }
        if (gs <= ge) {
          const uint4 len4 = make_uint4((uint32_t)len, (uint32_t)len, (uint32_t)len, (uint32_t)len);
          const uint32_t cd4 = (uint32_t)cd * 0x01010101u;
          for (int g = gs; g <= ge; ++g) {
            lrow[g] = len4;
            *reinterpret_cast<uint32_t*>(ccell + 4 * g) = cd4 | (mc_spread4(vb >> (4 * g)) * (uint32_t)kCodeValid);
          }
          for (int i = 4 * (ge + 1); i <= b1; ++i) {
            lcell[i] = len;
            ccell[i] = cd | (((vb >> i) & 1u) ? kCodeValid : 0);
          }
        }
      }
      __syncwarp();
      // copy-out: 1024 cells of this block, coalesced
      const int cbase = blk * 1024;
      if (out_vec) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int o = 4 * lane + 128 * k;  // cell offset inside the block (4 cells per lane)
          const int v = cbase + o;
          if (v < V) {
            const uint32_t c4 =
                *reinterpret_cast<const uint32_t*>(tcode + (o >> 5) * kMcCodeStride + (o & 31));
            const uint4 l4 = *reinterpret_cast<const uint4*>(tlen + (o >> 5) * kMcLenStride + (o & 31) * 4);
            *reinterpret_cast<uint4*>(a.merged_len + roff + v) = l4;
            if (packed) {
              *reinterpret_cast<uint32_t*>(a.packed + roff + v) = c4;
            } else {
              *reinterpret_cast<uint32_t*>(a.merged + roff + v) = (c4 >> 6) & 0x01010101u;
              const uint32_t k4 = c4 & 0x0f0f0f0fu;
              if (a.class_out) *reinterpret_cast<uint32_t*>(a.class_out + roff + v) = k4;
              if (a.len_cls)
                *reinterpret_cast<uint4*>(a.len_cls + roff + v) =
                    make_uint4((k4 & 0xffu) ? l4.x : 0u, (k4 & 0xff00u) ? l4.y : 0u, (k4 & 0xff0000u) ? l4.z : 0u,
                               (k4 & 0xff000000u) ? l4.w : 0u);
            }
          }
        }
      } else {
        for (int o = lane; o < 1024; o += 32) {
          const int v = cbase + o;
          if (v >= V) break;
          const unsigned char c1 = tcode[(o >> 5) * kMcCodeStride + (o & 31)];
          const int l1 = *reinterpret_cast<const int*>(tlen + (o >> 5) * kMcLenStride + (o & 31) * 4);
          a.merged_len[roff + v] = l1;
          if (packed) {
            a.packed[roff + v] = c1;
          } else {
            a.merged[roff + v] = (c1 >> 6) & 1u;
            if (a.class_out) a.class_out[roff + v] = c1 & 15u;
            if (a.len_cls) a.len_cls[roff + v] = (c1 & 15u) ? l1 : 0;
          }
        }
      }
      __syncwarp();
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
