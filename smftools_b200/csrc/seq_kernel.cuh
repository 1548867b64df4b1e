// Sequential (one thread per read) forward-backward for large batches and for GROUPED launches.
//
// The chunk-parallel kernels (fb2_kernel, walk_kernel + sweeps) buy their intra-read parallelism with work:
// the two-kernel E-step runs FOUR recursions per position (forward and backward boundary walks, then a
// forward and a backward sweep inside every chunk) plus K^2-wide boundary traffic, 294 thread-instructions
// per position at K = 4.  With >= ~10^5 reads in flight the reads alone fill the machine, and the classic
// checkpointed forward-backward needs only THREE recursions:
//
//   seq_forward_kernel   one thread walks its read forward (K(K-1) FMAs per position, scaled fp32, the
//                        column-normalised tables of fb2_kernel), stores alpha at every 16th position
//                        (checkpoints, [chunk][read][KP]: coalesced across the reads of a warp) and the
//                        read's log-likelihood;
//   seq_backward_kernel  the same thread layout walks the read backward chunk by chunk: alpha inside the
//                        chunk is recomputed from its checkpoint into a shared-memory stack
//                        ([position][thread] float4: conflict free), then ONE backward recursion carries
//                        beta through the whole read.  MODE_STATS: gamma / xi statistics accumulate in fp32
//                        registers (flushed every few chunks into per-warp fp64 accumulators, reduced in a
//                        fixed order: deterministic).  MODE_POST: posteriors and arg-max states are written
//                        from registers, 16 bytes per store.
//
// Observations are read straight from global memory, 16 positions per thread and load (uint4 for the uint8
// code, 4 x float4 for floats): every sector fetched is fully used, nothing is staged or transposed.  For
// the uint8 code (0 / 1 / missing) the per-state emission ratios come from a 3-row shared-memory table (one
// LDS.128) instead of K - 1 ex2 evaluations.
//
// GROUPED launches: the work list is a table of groups (own observation matrix, read count, read length,
// parameter slot); a work item is a block of 128 reads of one group, the per-group tables live in device
// memory (written by the host or by mstep_kernel), statistics are kept per item and reduced per group.
// Many small per-reference / per-barcode fits therefore run in ONE launch per EM iteration, without a
// host round trip (the M-step and the convergence test run on the device too).
#pragma once
#include "f32x2.cuh"
#include "fb2_kernel.cuh"

namespace smf {

constexpr int kSeqCH = 16;        // positions per chunk (checkpoint distance)
constexpr int kSeqThreads = 128;  // reads per CTA = reads per work item
// The backward kernel's alpha stack is THREAD-major: a thread's 16 alpha vectors of the current chunk are one row of
// kSeqCH * KP words plus a pad that makes the row an odd number of 16-byte units (K >= 3: 17, K = 2: 9), so the
// 128-bit accesses of a quarter warp fall into distinct bank groups.  Posteriors are written into the same row as a
// dense run laid out like the output (MODE_POST), which a warp then copies out 16 bytes per lane.
template <int K>
struct SeqRowWords {
  static constexpr int value = (K == 2) ? 36 : 68;
};

struct SeqGroup {
  const void* obs;   // [N, row_stride] observations
  long long N;       // reads
  int L;             // positions per read
  int row_stride;    // elements between rows
  int unaligned;     // 0: every row starts on a 16-byte boundary and a 16-element load never leaves the buffer (row_stride a
                     // multiple of 16 for uint8 / 4 for float): vector loads; 1: element-wise loads (E-step of arbitrary rows)
  int Tact;          // chunks per read = ceil(L / 16)
  int item0;         // first work item of the group (grouped launches)
  float* ckpt;       // [Tact][N][KP] alpha entering chunk c (slot 0 unused)
  double* ll;        // [N] per-read log-likelihood
  float* gamma;      // MODE_POST: [N, L, K] or [N, L] (gamma_state >= 0); may be nullptr
  int8_t* states;    // MODE_POST: [N, L]; may be nullptr
  int gamma_state;
  int out_stride;    // MODE_POST: positions between output rows (multiple of 4, >= L)
};

struct SeqArgs {
  SeqGroup g;               // single-group launch
  const SeqGroup* groups;   // grouped launch: device table
  const void* tabs;         // grouped launch: device FbTables<K>[n_groups]
  const int* item_group;    // grouped launch: group of every work item
  const int* active;        // grouped launch: per-group flag, 0 = skip the group (nullptr = all active)
  long long n_items;
  double* partials;         // MODE_STATS: [gridDim.x][S] (single) or [n_items][S] (grouped)
  int S;
  int flush_chunks;         // fp32 register statistics are flushed every this many chunks
};

constexpr int MODE_STATS = 0;
constexpr int MODE_POST = 1;

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
  __device__ __forceinline__ void load(const unsigned char* p, int len, int unaligned = 0) {
    if (!unaligned) {
      w = __ldg(reinterpret_cast<const uint4*>(p));  // rows padded to whole 16-byte groups
      return;
    }
    unsigned v[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};  // beyond the row: missing
#pragma unroll
    for (int j = 0; j < kSeqCH; ++j)
      if (j < len) v[j >> 2] = (v[j >> 2] & ~(0xffu << (8 * (j & 3)))) | ((unsigned)__ldg(p + j) << (8 * (j & 3)));
    w = make_uint4(v[0], v[1], v[2], v[3]);
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
  __device__ __forceinline__ void load(const float* p, int len, int unaligned = 0) {
    if (!unaligned) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        if (4 * g < len)
          q[g] = __ldg(reinterpret_cast<const float4*>(p) + g);
        else
          q[g] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      return;
    }
    float v[kSeqCH];
#pragma unroll
    for (int j = 0; j < kSeqCH; ++j) v[j] = (j < len) ? __ldg(p + j) : 0.0f;
#pragma unroll
    for (int g = 0; g < 4; ++g) q[g] = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
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
    c[0] = r.x;  // == 1, taken from the row so that (c_0, c_1) arrives as a register pair (f32x2.cuh)
    c[1] = r.y;
    if (K > 2) c[2 % K] = r.z;
    if (K > 3) c[3 % K] = r.w;
    m = code < 2u;
    // K <= 3: the row's last slot carries the observation value (0 / 1 / 0 for missing)
    om = (K == 4) ? (code == 1u ? 1.0f : 0.0f) : r.w;
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

// (re)build the emission lookup for the current tables; contains CTA barriers in the uint8 variant
template <int K, typename ObsT>
__device__ __forceinline__ void seq_emit_init(const FbTables<K>& tab, SeqEmit<K, ObsT>& em, float4* s_tab) {
  if constexpr (sizeof(ObsT) == 1) {
    __syncthreads();  // everyone is done with the previous table
    if (threadIdx.x < 3) {
      float r[4] = {1.0f, 1.0f, 1.0f, 1.0f};
      for (int k = 1; k < K; ++k) {
        const float x = threadIdx.x == 0 ? tab.rl0p[k] : (threadIdx.x == 1 ? tab.rl0p[k] + tab.rdl[k][0] : tab.rmiss[k]);
        r[k] = exp2f(x);
      }
      if (K < 4) r[3] = (threadIdx.x == 1) ? 1.0f : 0.0f;  // observation value of the code (SeqEmit::get)
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

// the handful of table entries the kernels keep in registers
template <int K>
struct SeqTab {
  float A[K][K];
  float pi2[K];
};
template <int K>
__device__ __forceinline__ void seq_tab_load(const FbTables<K>& tab, SeqTab<K>& t) {
#pragma unroll
  for (int i = 0; i < K; ++i) {
    t.pi2[i] = tab.pi2[i];
#pragma unroll
    for (int j = 0; j < K; ++j) t.A[i][j] = tab.Ap[i][j];
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
// (vector helpers: f32x2.cuh -- packed fp32x2 for K = 2, 3, 4 in the scalar code's per-lane operation order)
template <int K, bool PK = true>
__device__ __forceinline__ void seq_fwd_step(float (&al)[K], const float (&A)[K][K], const float (&c)[K], bool no_trans) {
  float y[K];
  vec_mat_unit<K, PK>(al, A, y);
#pragma unroll
  for (int k = 0; k < K; ++k) y[k] = no_trans ? al[k] : y[k];
  if constexpr (PK && (K == 4 || K == 2)) {
    // whole pairs, c[0] == 1 (exact): a multiply of ONE lane would make ptxas re-assemble the pair with moves
    vec_mul<K, PK>(al, y, c);
  } else {
    vec_mul_tail<K, PK>(al, y, c);
  }
}

// work-item bookkeeping shared by the kernels: which group / tables / read this thread works on
template <int K, bool GROUPED>
struct SeqItem {
  SeqGroup g;
  long long n;    // read of this thread (may be >= g.N: idle)
  bool skip;
  int group;
  __device__ __forceinline__ void locate(const SeqArgs& a, long long item) {
    if (GROUPED) {
      group = a.item_group[item];
      g = a.groups[group];
      skip = a.active != nullptr && a.active[group] == 0;
      n = (item - g.item0) * kSeqThreads + threadIdx.x;
    } else {
      group = 0;
      g = a.g;
      skip = false;
      n = item * kSeqThreads + threadIdx.x;
    }
  }
};

// -----------------------------------------------------------------------------------------------------
// forward: checkpoints + log-likelihood
// -----------------------------------------------------------------------------------------------------
// (a device function so that the fused fit loop below runs it too; tab0 is read only when !GROUPED)
template <int K, typename ObsT, bool GROUPED>
__device__ __forceinline__ void seq_forward_body(const FbTables<K>& tab0, const SeqArgs& a, float4* s_tab) {
  constexpr int KP = SeqVec<K>::KP;
  SeqEmit<K, ObsT> em;
  SeqTab<K> T;
  double log_pi2_sum = 0.0, l00 = 0.0, dl0 = 0.0, log_s0 = 0.0;
  if (!GROUPED) {
    log_pi2_sum = tab0.log_pi2_sum;
    l00 = tab0.l00[0];
    dl0 = tab0.dl0[0];
    log_s0 = tab0.log_s0;
    seq_emit_init<K, ObsT>(tab0, em, s_tab);
    seq_tab_load<K>(tab0, T);
  }
  int cur_group = -1;

  for (long long item = blockIdx.x; item < a.n_items; item += gridDim.x) {
    SeqItem<K, GROUPED> it;
    it.locate(a, item);
    if (it.skip) continue;  // uniform across the CTA
    if (GROUPED && it.group != cur_group) {
      const FbTables<K>& tg = reinterpret_cast<const FbTables<K>*>(a.tabs)[it.group];
      seq_emit_init<K, ObsT>(tg, em, s_tab);
      seq_tab_load<K>(tg, T);
      log_pi2_sum = tg.log_pi2_sum;
      l00 = tg.l00[0];
      dl0 = tg.dl0[0];
      log_s0 = tg.log_s0;
      cur_group = it.group;
    }
    const SeqGroup& g = it.g;
    if (it.n >= g.N) continue;
    const long long n = it.n;
    const int L = g.L;
    const size_t cstride = (size_t)g.N * KP;
    const ObsT* row = reinterpret_cast<const ObsT*>(g.obs) + n * (long long)g.row_stride;
    float al[K];
    {
      float s = 0.0f;
#pragma unroll
      for (int k = 0; k < K; ++k) s += T.pi2[k];
      const float r = 1.0f / s;
#pragma unroll
      for (int k = 0; k < K; ++k) al[k] = T.pi2[k] * r;
    }
    int e = 0, cnt = 0;
    double sumo = 0.0;
    float* dst = g.ckpt + cstride + (size_t)n * KP;
    SeqChunk<ObsT> cur, nxt;
    cur.load(row, min(kSeqCH, L), g.unaligned);
    for (int c = 0; c < g.Tact; ++c) {
      const int t0 = c * kSeqCH;
      const int len = min(kSeqCH, L - t0);
      if (c + 1 < g.Tact) nxt.load(row + t0 + kSeqCH, min(kSeqCH, L - t0 - kSeqCH), g.unaligned);
      if (c > 0) {
        seq_store_vec<K>(dst, al);
        dst += cstride;
      }
      float tsum = 0.0f;
      auto step = [&](int j) {
        float cf[K];
        bool m;
        float om;
        em.get(cur, j, cf, m, om);
        cnt += m ? 1 : 0;
        tsum += om;
        seq_fwd_step<K>(al, T.A, cf, c == 0 && j == 0);
        if ((j & 3) == 3) {
          const float sc = pow2_rescale(vec_max<K>(al), e);
          vec_scale<K>(al, sc);
        }
      };
      if (len == kSeqCH) {
#pragma unroll
        for (int j = 0; j < kSeqCH; ++j) step(j);
      } else {
#pragma unroll
        for (int j = 0; j < kSeqCH; ++j)
          if (j < len) step(j);
      }
      sumo += (double)tsum;
      cur = nxt;
    }
    float s = al[0];
#pragma unroll
    for (int k = 1; k < K; ++k) s += al[k];
    g.ll[n] = (double)logf(s) + (double)e * 0.6931471805599453094 + log_pi2_sum + (double)cnt * l00 + sumo * dl0 +
              (double)(L - 1) * log_s0;
  }
}

template <int K, typename ObsT, bool GROUPED>
__global__ void __launch_bounds__(kSeqThreads) seq_forward_kernel(const __grid_constant__ FbTables<K> tab0,
                                                                  const __grid_constant__ SeqArgs a) {
  __shared__ float4 s_tab[3];
  seq_forward_body<K, ObsT, GROUPED>(tab0, a, s_tab);
}

// MODE_POST copy-out plan of a lane.  A read's chunk of posteriors is ONE contiguous run of 16 * KO floats in the
// output (KO = K, or 1 for a single state column) and one dense run in the read's stack row; the warp's 32 runs
// are cut into 16-byte pieces, P = 4 * KO per read, piece p = it * 32 + lane of iteration `it`.  The (read, piece)
// pattern repeats every PER iterations (P = 12: 3 iterations = 8 reads, otherwise 1), so a lane keeps at most three
// (row offset, output offset) pairs and every iteration costs one LDS.128, one STG.128 and two adds.
template <int K>
struct SeqCopyPlan {
  int P, PER, RPP;        // pieces per read, iterations per period, reads per period
  int s_off[3], g_off[3]; // word offset in the warp's stack block / float offset from the warp's output base
  int r_of[3], q_of[3];   // read within the period and piece within the read
  int run0;               // first word of the dense run inside a row
  __device__ __forceinline__ void init(int KO, int out_stride, int lane) {
    constexpr int RS = SeqRowWords<K>::value;
    P = 4 * KO;
    PER = (P == 12) ? 3 : 1;
    RPP = 32 * PER / P;
    run0 = (KO == 1) ? ((K == 2) ? 16 : 48) : ((K == 3) ? 16 : 0);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int p = i * 32 + lane;
      const int r = p / P, q = p - r * P;
      r_of[i] = r;
      q_of[i] = q;
      s_off[i] = r * RS + run0 + 4 * q;
      g_off[i] = r * out_stride * KO + 4 * q;
    }
  }
  // wst: the warp's stack block (32 rows); gbase: output address of the warp's first read at position t0
  __device__ __forceinline__ void run(const float* __restrict__ wst, float* __restrict__ gbase, int KO, int out_stride,
                                      long long reads_left, int len) const {
    constexpr int RS = SeqRowWords<K>::value;
    const int nf4 = (len * KO + 3) >> 2;  // pieces of a partial last chunk (rounded up: the excess lands in the row pad)
    const int groups = 32 / RPP;
    const int gstep = RPP * out_stride * KO;
    int sb = 0, gb = 0, rb = 0;
    for (int grp = 0; grp < groups; ++grp) {
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        if (i < PER && q_of[i] < nf4 && rb + r_of[i] < reads_left) {
          const float4 v = *reinterpret_cast<const float4*>(wst + sb + s_off[i]);
          *reinterpret_cast<float4*>(gbase + (size_t)gb + g_off[i]) = v;
        }
      }
      sb += RPP * RS;
      gb += gstep;
      rb += RPP;
    }
  }
};

// resident CTAs per SM the backward kernel is compiled for (register cap = 65536 / (128 * blocks))
constexpr int seq_bwd_min_blocks(int K, size_t obs_size, int mode) {
  // posterior decode of float rows, K <= 3: one CTA fewer (128 registers), so that the chunk-ahead observation
  // loads stay in registers (K = 3, 100,000 x 5 kb: 4.27 -> 3.37 ms); K = 4 measured mixed with 3 CTAs and keeps 4
  if (mode == 1 && obs_size > 1 && K < 4) return 4;
  // K = 4 E-step on uint8-coded rows: the packed fp32x2 code fits 96 registers without spilling, so a fifth CTA
  // is resident (the kernel is latency-bound at 16 warps per SM: 61 % issue, 54 % of the FMA pipe); float rows
  // spill 24-184 bytes at that cap and measured slower with it (262,144 x 8 kb: 11.7 vs 10.8 ms), so they keep four
  if (K == 4) return (mode == 0 && obs_size == 1) ? 5 : 4;
  if (K == 2 && mode == 0 && obs_size == 1) return 6;  // 80 registers without spilling (float rows: 24-32 bytes)
  return 5;
}

// -----------------------------------------------------------------------------------------------------
// backward: alpha recomputed per chunk, one beta recursion; statistics (MODE_STATS) or posteriors (MODE_POST)
// -----------------------------------------------------------------------------------------------------
template <int K, typename ObsT, bool GROUPED, int MODE>
__device__ __forceinline__ void seq_backward_body(const FbTables<K>& tab0, const SeqArgs& a, unsigned char* smem) {
  constexpr int KP = SeqVec<K>::KP;
  constexpr int KK = K * K;
  constexpr bool STATS = (MODE == MODE_STATS);
  // Packed fp32x2 arithmetic in the E-step only: issue-bound at four resident CTAs per SM (configs[3]: 14.0 -> 10.8 ms).
  // Posterior decode measured mixed (float rows of large batches 6-13 % faster, uint8 rows and batches of ~25,000
  // reads 11-13 % slower: the pair moves sit on the dependency chain) and keeps the scalar code.
  constexpr bool PK = STATS;
  // [warps][S] fp64 accumulators | [16][128][KP] alpha stack | emission table
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kWarps = kSeqThreads / 32;
  const int S = STATS ? a.S : 0;
  double* wacc = reinterpret_cast<double*>(smem) + (size_t)warp * S;
  float* stack = reinterpret_cast<float*>(smem + (((size_t)kWarps * S * 8 + 15) & ~(size_t)15));
  constexpr int RS = SeqRowWords<K>::value;
  float4* s_tab = reinterpret_cast<float4*>(stack + (size_t)kSeqThreads * RS);
  if (STATS)
    for (int i = lane; i < S; i += 32) wacc[i] = 0.0;
  SeqEmit<K, ObsT> em;
  SeqTab<K> T;
  float Ap_flush[K][K];  // the constant xi factor, applied once per flush
  if (!GROUPED) {
    seq_emit_init<K, ObsT>(tab0, em, s_tab);
    seq_tab_load<K>(tab0, T);
#pragma unroll
    for (int i = 0; i < K; ++i)
#pragma unroll
      for (int j = 0; j < K; ++j) Ap_flush[i][j] = tab0.Ap[i][j];
  }
  int cur_group = -1;
  float* my = stack + (size_t)threadIdx.x * RS;  // this thread's row: alpha_j at word j * KP
  constexpr int jstride = KP;

  float acc_xi[STATS ? KK : 1], acc_num[STATS ? K : 1], acc_den[STATS ? K : 1], acc_start[STATS ? K : 1];
  double acc_ll = 0.0;
  auto zero_acc = [&]() {
    if (STATS) {
#pragma unroll
      for (int i = 0; i < KK; ++i) acc_xi[STATS ? i : 0] = 0.0f;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        acc_num[STATS ? k : 0] = 0.0f;
        acc_den[STATS ? k : 0] = 0.0f;
        acc_start[STATS ? k : 0] = 0.0f;
      }
    }
  };
  zero_acc();
  auto flush = [&]() {
    if (!STATS) return;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const float r0 = warp_sum_f32(acc_start[STATS ? k : 0]);
      const float r1 = warp_sum_f32(acc_num[STATS ? k : 0]);
      const float r2 = warp_sum_f32(acc_den[STATS ? k : 0]);
      if (lane == 0) {
        wacc[k] += (double)r0;
        wacc[K + KK + k] += (double)r1;
        wacc[K + KK + K + k] += (double)r2;
      }
    }
#pragma unroll
    for (int i = 0; i < KK; ++i) {
      const float r = warp_sum_f32(acc_xi[STATS ? i : 0]);
      if (lane == 0) wacc[K + i] += (double)r * (double)Ap_flush[i / K][i % K];
    }
    zero_acc();
  };
  // CTA partial vector -> global slot (fixed order over the warps), accumulators cleared
  auto emit_partials = [&](long long slot) {
    if (!STATS) return;
    {
      double v = acc_ll;
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
      if (lane == 0) wacc[S - 1] += v;
      acc_ll = 0.0;
    }
    __syncthreads();
    double* all = reinterpret_cast<double*>(smem);
    for (int sidx = threadIdx.x; sidx < S; sidx += kSeqThreads) {
      double tot = 0.0;
      for (int wv = 0; wv < kWarps; ++wv) tot += all[(size_t)wv * S + sidx];
      a.partials[(size_t)slot * S + sidx] = tot;
    }
    __syncthreads();
    for (int i = lane; i < S; i += 32) wacc[i] = 0.0;
    __syncwarp();
  };

  for (long long item = blockIdx.x; item < a.n_items; item += gridDim.x) {
    SeqItem<K, GROUPED> it;
    it.locate(a, item);
    if (it.skip) {
      if (GROUPED && STATS) emit_partials(item);  // zeros: the reduction never reads stale values
      continue;
    }
    if (GROUPED && it.group != cur_group) {
      const FbTables<K>& tg = reinterpret_cast<const FbTables<K>*>(a.tabs)[it.group];
      seq_emit_init<K, ObsT>(tg, em, s_tab);
      seq_tab_load<K>(tg, T);
#pragma unroll
      for (int i = 0; i < K; ++i)
#pragma unroll
        for (int j = 0; j < K; ++j) Ap_flush[i][j] = tg.Ap[i][j];
      cur_group = it.group;
    }
    const SeqGroup& g = it.g;
    const bool live = it.n < g.N;
    const long long nn = live ? it.n : g.N - 1;  // idle threads shadow the last read (results discarded)
    const int L = g.L;
    const size_t cstride = (size_t)g.N * KP;
    const ObsT* row = reinterpret_cast<const ObsT*>(g.obs) + nn * (long long)g.row_stride;
    float* grow = nullptr;
    int8_t* srow = nullptr;
    SeqCopyPlan<K> cp;
    if (!STATS) cp.init(g.gamma_state < 0 ? K : 1, g.out_stride, lane);
    if (!STATS) {
      if (g.gamma != nullptr) grow = g.gamma + (size_t)nn * g.out_stride * (g.gamma_state < 0 ? K : 1);
      if (g.states != nullptr) srow = g.states + (size_t)nn * g.out_stride;
    }
    float bv[K];
#pragma unroll
    for (int k = 0; k < K; ++k) bv[k] = 1.0f / K;
    SeqChunk<ObsT> cur, prv;
    {
      const int t0 = (g.Tact - 1) * kSeqCH;
      cur.load(row + t0, L - t0, g.unaligned);
    }
    int since_flush = 0;
    float vnext[K];
#pragma unroll
    for (int k = 0; k < K; ++k) vnext[k] = 0.0f;
    if (g.Tact > 1) seq_load_vec<K>(g.ckpt + (size_t)(g.Tact - 1) * cstride + (size_t)nn * KP, vnext);
    for (int c = g.Tact - 1; c >= 0; --c) {
      const int t0 = c * kSeqCH;
      const int len = min(kSeqCH, L - t0);
      const bool first = (c == 0);
      float v[K];
      if (sizeof(ObsT) > 1 && c >= 2) {
        // float rows: the register prefetch below holds 16 registers per chunk, which ptxas trades away under the
        // register cap (the loads then sit right before their use); an L2 prefetch two chunks ahead costs none
        asm volatile("prefetch.global.L2 [%0];" ::"l"(row + t0 - 2 * kSeqCH));
      }
      if (!first) {
        prv.load(row + t0 - kSeqCH, kSeqCH, g.unaligned);
#pragma unroll
        for (int k = 0; k < K; ++k) v[k] = vnext[k];
        // checkpoint of the chunk before this one: in flight while this chunk is processed
        if (c > 1) seq_load_vec<K>(g.ckpt + (size_t)(c - 1) * cstride + (size_t)nn * KP, vnext);
      } else {
#pragma unroll
        for (int k = 0; k < K; ++k) v[k] = T.pi2[k];
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
      auto fstep = [&](int j) {
        float cf[K];
        bool m;
        float om;
        em.get(cur, j, cf, m, om);
        seq_fwd_step<K, PK>(al, T.A, cf, first && j == 0);
        if ((j & 3) == 3) {
          const float sc = pow2_rescale_noexp(vec_max<K>(al));
          vec_scale<K, PK>(al, sc);
        }
        seq_store_vec<K>(my + j * jstride, al);
      };
      if (len == kSeqCH) {
#pragma unroll
        for (int j = 0; j < kSeqCH; ++j) fstep(j);
      } else {
#pragma unroll
        for (int j = 0; j < kSeqCH; ++j)
          if (j < len) fstep(j);
      }
      // ---- backward over the chunk ----
      float alc[K];
#pragma unroll
      for (int k = 0; k < K; ++k) alc[k] = al[k];
      unsigned sq[!STATS ? 4 : 1] = {0};  // MODE_POST: the chunk's 16 arg-max states, one byte each
      float gq[!STATS ? (K == 3 ? 12 : 4) : 1];  // MODE_POST: four positions of posteriors waiting for a 16-byte store
      auto bstep = [&](int j) {
        float cr[K];
        bool m;
        float om;
        em.get(cur, j, cr, m, om);
        float gk[K];
        vec_mul<K, PK>(gk, alc, bv);
        float gs = 0.0f;
#pragma unroll
        for (int k = 0; k < K; ++k) gs += gk[k];
        const float g0 = rcp_approx(gs);
        if (STATS) {
          const float gr = m ? g0 : 0.0f;
          if constexpr (STATS) {
            float gg[K];
#pragma unroll
            for (int k = 0; k < K; ++k) gg[k] = gk[k];
            vec_scale<K, PK>(gg, gr);
            vec_fma_vs<K, PK>(acc_num, gg, om);
            vec_add<K, PK>(acc_den, gg);
          }
        } else {
          vec_scale<K, PK>(gk, g0);
          if (srow != nullptr) {
            unsigned best = 0;
            float gbest = gk[0];
#pragma unroll
            for (int k = 1; k < K; ++k)
              if (gk[k] > gbest) {
                gbest = gk[k];
                best = k;
              }
            sq[!STATS ? (j >> 2) : 0] |= best << (8 * (j & 3));
          }
        }
        float alp[K];
        if (j > 0) {
          seq_load_vec<K>(my + (j - 1) * jstride, alp);
        } else {
#pragma unroll
          for (int k = 0; k < K; ++k) alp[k] = v[k];
        }
        // MODE_POST: gamma_j takes the stack slot of alpha_j (already consumed); the chunk leaves from there
        if (!STATS && grow != nullptr) {
          if (g.gamma_state >= 0) {
            float gsel = gk[0];
#pragma unroll
            for (int k = 1; k < K; ++k)
              if (g.gamma_state == k) gsel = gk[k];
            gq[!STATS ? (j & 3) : 0] = gsel;
            if ((j & 3) == 0)
              *reinterpret_cast<float4*>(my + cp.run0 + j) = make_float4(gq[0], gq[!STATS ? 1 : 0], gq[!STATS ? 2 : 0], gq[!STATS ? 3 : 0]);
          } else if (K == 3) {
            // four positions = three 16-byte pieces of the dense run, which sits at the top of the row so that it
            // never reaches the alpha slots still to be read (run word 16 + 3 j  >  alpha_{j-1} word 4 j - 1)
#pragma unroll
            for (int k = 0; k < K; ++k) gq[(!STATS && K == 3) ? (j & 3) * 3 + k : 0] = gk[k];
            if ((j & 3) == 0) {
              float4* d4 = reinterpret_cast<float4*>(my + 16 + 3 * j);
#pragma unroll
              for (int x = 0; x < 3; ++x)
                d4[x] = make_float4(gq[(!STATS && K == 3) ? 4 * x : 0], gq[(!STATS && K == 3) ? 4 * x + 1 : 0],
                                    gq[(!STATS && K == 3) ? 4 * x + 2 : 0], gq[(!STATS && K == 3) ? 4 * x + 3 : 0]);
            }
          } else {
            seq_store_vec<K>(my + j * jstride, gk);  // K = 2, 4: the slot of alpha_j is exactly gamma_j's place
          }
        }
        const bool at_start = first && j == 0;
        if (STATS && at_start) {
          // gamma_0 of the read (start statistics, counted whether or not position 0 is observed)
#pragma unroll
          for (int k = 0; k < K; ++k) acc_start[STATS ? k : 0] += gk[k] * g0;
        }
        float cb[K], nb[K];
        if constexpr (PK && (K == 4 || K == 2)) {
          vec_mul<K, PK>(cb, cr, bv);  // cr[0] == 1: exact
        } else {
          vec_mul_tail<K, PK>(cb, bv, cr);
        }
        mat_vec_unit<K, PK>(T.A, cb, nb);
        if (STATS) {
          const bool prev_obs = (j > 0) ? cur.observed(j > 0 ? j - 1 : 0) : prev_obs0;
          const bool valid = m && prev_obs && !at_start;
          float s2 = 0.0f;
#pragma unroll
          for (int i = 0; i < K; ++i) s2 = fmaf(alp[i], nb[i], s2);
          const float r2 = valid ? rcp_approx(s2) : 0.0f;
          if constexpr (STATS) {
            float ai[K];
#pragma unroll
            for (int i = 0; i < K; ++i) ai[i] = alp[i];
            vec_scale<K, PK>(ai, r2);
#pragma unroll
            for (int i = 0; i < K; ++i) {
              float row[K];
#pragma unroll
              for (int jj = 0; jj < K; ++jj) row[jj] = acc_xi[i * K + jj];
              vec_fma_s<K, PK>(row, ai[i], cb);
#pragma unroll
              for (int jj = 0; jj < K; ++jj) acc_xi[i * K + jj] = row[jj];
            }
          }
        }
        if ((j & 3) == 0) {
          const float sc = pow2_rescale_noexp(vec_max<K>(nb));
          vec_scale<K, PK>(nb, sc);
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
          bv[k] = nb[k];
          alc[k] = alp[k];
        }
      };
      if (len == kSeqCH) {
#pragma unroll
        for (int j = kSeqCH - 1; j >= 0; --j) bstep(j);
      } else {
#pragma unroll
        for (int j = kSeqCH - 1; j >= 0; --j)
          if (j < len) bstep(j);
      }
      if (!STATS) {
        if (srow != nullptr && live) {
          if ((g.out_stride & 15) == 0) {
            *reinterpret_cast<uint4*>(srow + t0) = make_uint4(sq[0], sq[!STATS ? 1 : 0], sq[!STATS ? 2 : 0], sq[!STATS ? 3 : 0]);
          } else {
#pragma unroll
            for (int w = 0; w < 4; ++w)
              if (4 * w < len) *reinterpret_cast<unsigned*>(srow + t0 + 4 * w) = sq[!STATS ? w : 0];
          }
        }
#pragma unroll
        for (int w = 0; w < 4; ++w) sq[!STATS ? w : 0] = 0;
        if (g.gamma != nullptr) {
          // the warp's 32 read-chunks leave together: a read's 16 positions are one contiguous run of the output
          __syncwarp();
          const long long n0 = it.n - lane;  // first read of this warp
          const int KO = g.gamma_state < 0 ? K : 1;
          cp.run(stack + (size_t)(warp * 32) * RS, g.gamma + ((size_t)n0 * g.out_stride + t0) * KO, KO, g.out_stride,
                 g.N - n0, len);
          __syncwarp();
        }
      }
      cur = prv;
      if (STATS && ++since_flush == a.flush_chunks) {
        if (!live) zero_acc();
        flush();
        since_flush = 0;
      }
    }
    if (STATS) {
      if (!live) zero_acc();
      flush();
      if (live) acc_ll += g.ll[it.n];
      if (GROUPED) emit_partials(item);
    }
  }
  if (STATS && !GROUPED) emit_partials(blockIdx.x);
}

template <int K, typename ObsT, bool GROUPED, int MODE>
__global__ void __launch_bounds__(kSeqThreads, seq_bwd_min_blocks(K, sizeof(ObsT), MODE)) seq_backward_kernel(const __grid_constant__ FbTables<K> tab0,
                                                                                      const __grid_constant__ SeqArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  seq_backward_body<K, ObsT, GROUPED, MODE>(tab0, a, smem);
}

// Single-channel, homogeneous-transition tables from the fp64 log tables (the same arithmetic as the host's
// build_fb_tables in launch.h, usable on the device by the grouped M-step).
template <int K>
__host__ __device__ inline void seq_tables_from_logs(const double* log_start, const double* log_trans, const double* log_e1,
                                                     const double* log_e0, FbTables<K>* out) {
  const double log2e = 1.4426950408889634074;
  double pisum = 0.0;
  for (int k = 0; k < K; ++k) {
    out->pi[k] = (float)exp(log_start[k]);
    pisum += (double)out->pi[k];
  }
  out->log_pi_sum = log(pisum);
  for (int i = 0; i < K; ++i)
    for (int j = 0; j < K; ++j) out->A[0][i][j] = (float)exp(log_trans[i * K + j]);
  const double l0_0 = log_e0[0];
  const double d_0 = log_e1[0] - l0_0;
  out->l00[0] = l0_0;
  out->dl0[0] = d_0;
  double sdiag[K];
  for (int j = 0; j < K; ++j) sdiag[j] = exp(log_trans[j * K + j]);
  for (int i = 0; i < K; ++i)
    for (int j = 0; j < K; ++j) out->Ap[i][j] = (i == j) ? 1.0f : (float)(exp(log_trans[i * K + j]) / sdiag[j]);
  double p2sum = 0.0;
  for (int k = 0; k < K; ++k) {
    const double l0 = log_e0[k];
    const double d = log_e1[k] - l0;
    out->rl0[k][0] = (float)((l0 - l0_0) * log2e);
    out->rdl[k][0] = (float)((d - d_0) * log2e);
    const double ratio = sdiag[k] / sdiag[0];
    out->pi2[k] = (float)(exp(log_start[k]) / ratio);
    p2sum += (double)out->pi2[k];
    out->rl0p[k] = (float)((l0 - l0_0) * log2e + log2(ratio));
    out->rmiss[k] = (float)log2(ratio);
  }
  out->log_s0 = log_trans[0];
  out->log_pi2_sum = log(p2sum);
  out->nb = 1;
  out->C = 1;
}

// Dynamic-range guard of the scaled fp32 recursions (the rule of rescale_cadence() in api.cu, restated for
// single-channel homogeneous models so that the device-side M-step can apply it to its own parameters):
// true when a rescale every 4th step keeps every still-relevant entry inside fp32's normal range.
template <int K>
__host__ __device__ inline bool seq_cadence4_ok(const double* lt, const double* l1, const double* l0) {
  double logf = 0.0;
  for (int o = 0; o < 2; ++o) {
    double lo = 1e300, hi = -1e300;
    for (int k = 0; k < K; ++k) {
      const double v = o ? l1[k] : l0[k];
      lo = v < lo ? v : lo;
      hi = v > hi ? v : hi;
    }
    if (lo - hi < logf) logf = lo - hi;
  }
  double logg = 0.0;
  for (int i = 0; i < K; ++i) {
    double lo = 1e300, hi = -1e300;
    for (int j = 0; j < K; ++j) {
      const double v = lt[i * K + j];
      lo = v < lo ? v : lo;
      hi = v > hi ? v : hi;
    }
    if (lo - hi < logg) logg = lo - hi;
  }
  return 2.0 * logg + 5.0 * logf >= -64.0;
}

constexpr int SEQ_MSTEP_ALL = 0;
constexpr int SEQ_MSTEP_REDUCE_ONLY = 1;
constexpr int SEQ_MSTEP_UPDATE_ONLY = 2;

// Per-group EM bookkeeping of a grouped fit, all in device memory.
struct SeqFitState {
  double* params;   // [G][K + K*K + K]: start | trans | emission (fp64, the reference's parameter layout)
  double* stats;    // [G][S]
  double* hist;     // [G][max_iter] log-likelihood before each M-step
  int* active;      // [G] 1 while the group still iterates
  int* n_iter;      // [G] iterations done
  int* status;      // [G] SMF_FIT_* (0 ran all iterations, 1 converged, 2 handed back to the caller)
  void* tabs;       // [G] FbTables<K>
  const SeqGroup* groups;   // [G] device group table
  const double* partials;   // [n_items][S] per-item statistics of the E-step
  int n_groups, S, max_iter, iter;  // iter < 0: build the tables of the initial parameters only
  int what;                         // SEQ_MSTEP_*: reduce + update (one process), or the two halves separately so
                                    // that the caller can all-reduce stats[G][S] over ranks in between
  double eps, tol;
  int update_start, update_trans, update_emission;
};

// Statistics of group g from the per-item partials of the E-step, in a fixed order that does not depend on the
// CTA size: warp w takes statistics w, w + warps, ...; lane l sums items l, l + 32, ...; one xor-shuffle tree.
__device__ __forceinline__ void seq_reduce_group(const SeqFitState& f, int g) {
  const SeqGroup grp = f.groups[g];
  const long long n_items = (grp.N + kSeqThreads - 1) / kSeqThreads;
  const int nw = blockDim.x >> 5, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int sidx = w; sidx < f.S; sidx += nw) {
    double tot = 0.0;
    for (long long i = lane; i < n_items; i += 32) tot += f.partials[(size_t)(grp.item0 + i) * f.S + sidx];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, d);
    if (lane == 0) f.stats[(size_t)g * f.S + sidx] = tot;
  }
}

// M-step + convergence test + table rebuild of group g by ONE thread: the closed forms and the unconditional
// re-normalisation of HMM.py:1598-1614, the stop rule of :1616-1628.  iter < 0: tables of the initial parameters.
template <int K>
__device__ inline void seq_mstep_update(const SeqFitState& f, int g, int iter) {
  constexpr int KK = K * K;
  double* p = f.params + (size_t)g * (K + KK + K);
  double* start = p;
  double* trans = p + K;
  double* em = p + K + KK;
  const double eps = f.eps;
  if (iter < 0) {
    f.active[g] = 1;
    f.n_iter[g] = 0;
    f.status[g] = 0;
  } else {
    if (f.active[g] == 0) return;
    const double* st = f.stats + (size_t)g * f.S;
    const double ll = st[f.S - 1];
    f.hist[(size_t)g * f.max_iter + iter] = ll;
    f.n_iter[g] = iter + 1;
    if (f.update_start) {
      double s[K], tot = 0.0;
      for (int k = 0; k < K; ++k) {
        s[k] = st[k] + eps;
        tot += s[k];
      }
      for (int k = 0; k < K; ++k) start[k] = s[k] / tot;
    }
    if (f.update_trans) {
      for (int i = 0; i < K; ++i) {
        double t[K], tot = 0.0;
        for (int j = 0; j < K; ++j) {
          t[j] = st[K + i * K + j] + eps;
          tot += t[j];
        }
        if (tot == 0.0) tot = 1.0;
        for (int j = 0; j < K; ++j) trans[i * K + j] = t[j] / tot;
      }
    }
    if (f.update_emission) {
      for (int k = 0; k < K; ++k) {
        double e = (st[K + KK + k] + eps) / (st[K + KK + K + k] + 2.0 * eps);
        e = e < eps ? eps : (e > 1.0 - eps ? 1.0 - eps : e);
        em[k] = e;
      }
    }
    // the reference re-normalises every parameter after every update (HMM.py:492-511, 1480-1484)
    {
      double s[K], tot = 0.0;
      for (int k = 0; k < K; ++k) {
        s[k] = start[k] + eps;
        tot += s[k];
      }
      for (int k = 0; k < K; ++k) start[k] = s[k] / tot;
      for (int i = 0; i < K; ++i) {
        double t[K], rt = 0.0;
        for (int j = 0; j < K; ++j) {
          t[j] = trans[i * K + j] + eps;
          rt += t[j];
        }
        if (rt == 0.0) rt = 1.0;
        for (int j = 0; j < K; ++j) trans[i * K + j] = t[j] / rt;
      }
      for (int k = 0; k < K; ++k) em[k] = em[k] < eps ? eps : (em[k] > 1.0 - eps ? 1.0 - eps : em[k]);
    }
    if (iter > 0) {
      const double prev = f.hist[(size_t)g * f.max_iter + iter - 1];
      const double scale = fabs(ll) > 1.0 ? fabs(ll) : 1.0;
      if (fabs(ll - prev) < f.tol * scale) {
        f.active[g] = 0;
        f.status[g] = 1;
      }
    }
  }
  double ls[K], lt[KK], l1[K], l0[K];
  for (int k = 0; k < K; ++k) {
    ls[k] = log(start[k] + eps);
    l1[k] = log(em[k] + eps);
    l0[k] = log1p(-em[k] + eps);
  }
  for (int i = 0; i < KK; ++i) lt[i] = log(trans[i] + eps);
  if (!seq_cadence4_ok<K>(lt, l1, l0)) {
    // parameters outside the dynamic range the scaled fp32 recursion covers with its fixed rescale cadence: stop
    // here and let the caller continue this group on the general path
    if (f.active[g] != 0) f.status[g] = 2;
    f.active[g] = 0;
    return;
  }
  FbTables<K>* out = reinterpret_cast<FbTables<K>*>(f.tabs) + g;
  if (iter < 0) {
    FbTables<K> tab;
    memset(&tab, 0, sizeof(tab));
    seq_tables_from_logs<K>(ls, lt, l1, l0, &tab);
    *out = tab;
  } else {
    seq_tables_from_logs<K>(ls, lt, l1, l0, out);  // the same fields every time: the rest stays as first written
  }
}

// One CTA per group: the group's statistics from the per-item partials, then thread 0 updates (see above).
template <int K>
__global__ void __launch_bounds__(256) seq_mstep_kernel(const __grid_constant__ SeqFitState f) {
  const int g = blockIdx.x;
  if (f.iter >= 0 && f.what != SEQ_MSTEP_UPDATE_ONLY) {
    if (f.active[g] == 0) {
      // a stopped group contributes nothing any more (its slot of an all-reduced statistics matrix must be defined)
      if (f.what == SEQ_MSTEP_REDUCE_ONLY)
        for (int sidx = threadIdx.x; sidx < f.S; sidx += blockDim.x) f.stats[(size_t)g * f.S + sidx] = 0.0;
      return;  // uniform across the CTA
    }
    seq_reduce_group(f, g);
    __syncthreads();
  }
  if (threadIdx.x != 0 || f.what == SEQ_MSTEP_REDUCE_ONLY) return;
  seq_mstep_update<K>(f, g, f.iter);
}

// -----------------------------------------------------------------------------------------------------
// The whole Baum-Welch loop of a SMALL grouped fit in one cooperative launch (every work item has its own resident
// CTA): forward + backward of an item run back to back in the same threads, two grid barriers per iteration
// separate the E-step from the M-step (groups dealt round-robin to the CTAs), and the loop ends as soon as no
// group is active.  Same arithmetic and summation orders as the three-kernel sequence: identical results.
// -----------------------------------------------------------------------------------------------------
__device__ __forceinline__ void seq_grid_barrier(unsigned* counter, unsigned& target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += gridDim.x;
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
    } while ((int)(seen - target) < 0);
    __threadfence();
  }
  __syncthreads();
}

template <int K, typename ObsT>
__global__ void __launch_bounds__(kSeqThreads, 1) seq_fit_loop_kernel(const __grid_constant__ SeqArgs a,
                                                                      const __grid_constant__ SeqFitState f,
                                                                      unsigned* barrier_counter) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ float4 s_tab[3];
  const FbTables<K>& unused = *reinterpret_cast<const FbTables<K>*>(a.tabs);
  unsigned target = 0;
  for (int iter = 0; iter < f.max_iter; ++iter) {
    seq_forward_body<K, ObsT, true>(unused, a, s_tab);
    seq_backward_body<K, ObsT, true, MODE_STATS>(unused, a, smem);
    seq_grid_barrier(barrier_counter, target);
    for (int g = blockIdx.x; g < f.n_groups; g += gridDim.x) {
      if (f.active[g] == 0) continue;  // uniform across the CTA
      seq_reduce_group(f, g);
      __syncthreads();
      if (threadIdx.x == 0) seq_mstep_update<K>(f, g, iter);
      __syncthreads();
    }
    seq_grid_barrier(barrier_counter, target);
    int any = 0;
    for (int g = threadIdx.x; g < f.n_groups; g += kSeqThreads) any |= f.active[g];
    if (__syncthreads_or(any) == 0) break;
  }
}

}  // namespace smf
