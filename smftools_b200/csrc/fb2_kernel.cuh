// Fast path of the fused forward-backward kernel (single channel, homogeneous transitions).
//
// Same algorithm as fb_kernel.cuh (chunk products -> warp-shuffle associative scan -> local
// forward/backward sweeps), re-shaped for instruction throughput and asynchronous data movement:
//   * the chunk length CH is a compile-time odd constant: every per-position loop is fully
//     unrolled (no loop/branch/address arithmetic), and the emission ratios b_k/b_0 of a thread's
//     chunk live in registers, so the transcendental work is done once instead of three times;
//   * a read's observations arrive with ONE bulk asynchronous copy (TMA engine,
//     cp.async.bulk global->shared, completion on an mbarrier); the next read is prefetched as
//     soon as every thread has pulled its chunk into registers, so the copy overlaps phases 2-3;
//   * posteriors and states leave through bulk asynchronous shared->global stores; the few
//     bytes before/after the 16-byte aligned body are written with ordinary stores;
//   * a read whose length is not a multiple of CH is padded (in shared memory) with missing
//     observations, which leave every posterior of the real positions unchanged (beta of the
//     last position stays uniform up to fp32 rounding, ~1e-7) and contribute no statistics.
#pragma once
#include <type_traits>
#include "fb_kernel.cuh"

namespace smf {

struct Fb2Args {
  const void* obs;
  long long N;
  int L;
  float* gamma;     // [N,L,K] / [N,L] / nullptr
  int gamma_state;  // <0: all states
  int8_t* states;   // [N,L] / nullptr
  double* loglik;   // [N] / nullptr
  double* partials; // STATS: [gridDim.x][S]
  int S;
  int T;            // threads per group (multiple of 32)
  int groups;       // reads in flight per CTA
  int flush_every;
  int alias_in;     // 1: the observation buffer shares memory with the staging buffer (no prefetch; the next
                    //    read is loaded once this read's stores have left shared memory)
  int nbuf;         // output staging buffers per group (2 = a read's bulk stores drain while the next read computes)
  int obuf_bytes;   // bytes between the two staging sets (stage | out1 | st)
  int off_wacc;     // CTA-wide double [warps][S]
  int off_group0;
  int group_bytes;
  // offsets inside a group region (all 16-byte aligned)
  int goff_bar, goff_in, goff_stage, goff_out1, goff_st, goff_scan;
  // WALK variant: chunk boundary vectors from walk_kernel ([chunk][read][K]) and per-read log-likelihoods
  const float* bnd_v;
  const float* bnd_w;
  const double* ll_in;
  int Tact;  // chunks per read
  int gc;    // 1: staging tile sized for the compact (OUT == 3) layout
  const int8_t* bins;  // BINNED: device [L-1] distance-bin index of every transition
  int n_bins;
  int off_bins;     // BINNED, CTA-wide: int8 [Lpad] bin of the transition INTO position t at index t (0 for t = 0 / pads)
  int off_btab;     // BINNED, CTA-wide: float [n_bins][kBinRow] column-normalised transitions + log2 scale ratios
  int off_ls0;      // BINNED, CTA-wide: double [n_bins] log A_b[0][0]
  int off_sxi;      // BINNED E-step (K = 2), CTA-wide: float4 [blockDim.x][n_bins | 1] per-thread, per-bin xi accumulators
  int late_load;  // 1: the next read is requested at the end of the iteration (multi-channel E-step: the backward
                  //    sweep re-reads the observations from shared memory for the per-channel emission statistics)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "MBAR_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra MBAR_DONE;\n"
      "bra MBAR_WAIT;\n"
      "MBAR_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// TMA bulk copy global -> shared, completion signalled on an mbarrier (16-byte granules).
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// TMA bulk copy shared -> global (bulk async-group completion).
__device__ __forceinline__ void bulk_store(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <typename ObsT>
__device__ __forceinline__ float obs_to_float(ObsT v);
template <>
__device__ __forceinline__ float obs_to_float<float>(float v) { return v; }
template <>
__device__ __forceinline__ float obs_to_float<double>(double v) { return static_cast<float>(v); }
template <>
__device__ __forceinline__ float obs_to_float<unsigned char>(unsigned char v) {
  return v == 0 ? 0.0f : (v == 1 ? 1.0f : __int_as_float(0x7fc00000));
}
template <typename ObsT>
__device__ __forceinline__ ObsT obs_missing();
template <>
__device__ __forceinline__ float obs_missing<float>() { return __int_as_float(0x7fc00000); }
template <>
__device__ __forceinline__ double obs_missing<double>() { return __longlong_as_double(0x7ff8000000000000LL); }
template <>
__device__ __forceinline__ unsigned char obs_missing<unsigned char>() { return 255; }

// Store bytes [0, nbytes) of a shared buffer to global, called by a whole warp: the 16-byte
// aligned body goes out as one bulk async store issued by lane 0, the ragged head / tail (< 16
// bytes each) are copied by lanes 1..31 with plain stores.  `src` and `dst` must be congruent
// modulo 16; UNIT is the element size (1 or 4) used for the ragged ends.
template <typename UNIT>
__device__ __forceinline__ void store_row_bulk(void* dst, const void* src, uint32_t nbytes, int lane) {
  const uint32_t mis = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(dst) & 15u);
  uint32_t head = (16u - mis) & 15u;
  if (head > nbytes) head = nbytes;
  const uint32_t body = (nbytes - head) & ~15u;
  const uint32_t tail = nbytes - head - body;
  if (lane == 0 && body) bulk_store(static_cast<char*>(dst) + head, static_cast<const char*>(src) + head, body);
  const UNIT* s = static_cast<const UNIT*>(src);
  UNIT* d = static_cast<UNIT*>(dst);
  const uint32_t nh = head / sizeof(UNIT), nt = tail / sizeof(UNIT);
  const uint32_t tb = (head + body) / sizeof(UNIT);
  const uint32_t i = (uint32_t)lane;
  if (i >= 1 && i - 1 < nh) d[i - 1] = s[i - 1];            // lanes 1..15: head
  if (i >= 16 && i - 16 < nt) d[tb + i - 16] = s[tb + i - 16];  // lanes 16..31: tail
}

// OUT selects the output set at compile time: 0 = all-state posterior + states (decode),
// 1 = single-state posterior + states (annotate), 2 = decided at run time from the pointers.
//
// WALK = true: the chunk boundary vectors come from walk_kernel (walk_kernel.cuh) instead of the
// chunk-product + scan phases; the kernel is then just the load / local sweeps / store pipeline.
//
// C > 1: multi-channel observations [N, L, C] (MultiBernoulliHMM, HMM.py:1766-1782): the emission ratio of a
// position is the product over its observed channels, computed once in phase 1 like the single-channel ratio --
// everything after that (chunk products, scans, sweeps, stores) is unchanged.  The E-step keeps K x C emission
// accumulators and re-reads the observations from shared memory in the backward sweep (late_load).
//
// BINNED: distance-binned transitions (DistanceBinnedSingleBernoulliHMM, HMM.py:2127-2188): the transition into
// position t is A_b with b = bins[t-1].  Every bin's matrix is column-normalised on its own (A_b = A'_b diag(s_b));
// a shared-memory table holds, per bin, the off-diagonal entries of A'_b and log2(s_b[k] / s_b[0]), one 16- or
// 32-byte row that a position fetches with its bin index in each of the three phases; the scale ratio is folded
// into the cached emission ratio, the dropped scalar s_b[0] returns in the log-likelihood.  Posterior decode only.
template <int K, int CH, bool STATS, typename ObsT, int MAXT, int MINB, int OUT, bool WALK = false, int C = 1,
          bool BINNED = false>
__global__ void __launch_bounds__(MAXT, MINB) fb2_kernel(const __grid_constant__ FbTables<K> tab,
                                                   const __grid_constant__ Fb2Args a) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr unsigned kFull = 0xffffffffu;
  constexpr int KK = K * K;
  constexpr int KR = K - 1;  // emission ratios per position
  // K = 4 would need 3*CH registers for the cached ratios: keep the observation instead and
  // recompute the ratios (3 ex2 per position and phase; the kernel is FMA-bound there anyway)
  constexpr bool RECOMP = (K >= 4);
  static_assert(C == 1 || (!WALK && OUT == 2 && K <= 3), "multi-channel variant: K <= 3, run-time output set, single kernel");
  static_assert(!BINNED || (C == 1 && !WALK && OUT == 2 && K <= 3 && (!STATS || K == 2)),
                "distance-binned variant: posterior decode for K <= 3, E-step for K = 2");
  constexpr int kBinRow = (K == 2) ? 4 : 8;  // floats per bin: K (K - 1) off-diagonal entries + K - 1 scale ratios (+ pad)

  const int tid = threadIdx.x;
  const int T = a.T;
  const int g = tid / T;
  const int gt = tid - g * T;
  const int lane = tid & 31;
  const int gw = gt >> 5;
  const int G = T >> 5;
  const int wcta = tid >> 5;
  const int L = a.L;
  const int bar_id = 1 + g;

  double* wacc = reinterpret_cast<double*>(smem + a.off_wacc);
  unsigned char* gbase = smem + a.off_group0 + (size_t)g * a.group_bytes;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(gbase + a.goff_bar);
  unsigned char* s_in_raw = gbase + a.goff_in;
  float* s_stage_base = reinterpret_cast<float*>(gbase + a.goff_stage);
  float* s_out1_base = reinterpret_cast<float*>(gbase + a.goff_out1);
  int8_t* s_st_base = reinterpret_cast<int8_t*>(gbase + a.goff_st);
  float* s_wtot = reinterpret_cast<float*>(gbase + a.goff_scan);  // [32][KK]
  float* s_u = s_wtot + 32 * KK;
  float* s_z = s_u + 32 * K;
  double* s_ll = reinterpret_cast<double*>(s_z + 32 * K);  // [32]

  if (gt == 0) mbar_init(mbar, 1);
  const int8_t* s_bins = reinterpret_cast<const int8_t*>(smem + a.off_bins);
  const float* s_btab = reinterpret_cast<const float*>(smem + a.off_btab);
  const double* s_ls0 = reinterpret_cast<const double*>(smem + a.off_ls0);
  if (BINNED) {
    int8_t* wb = reinterpret_cast<int8_t*>(smem + a.off_bins);
    const int lpad = ((L + CH - 1) / CH) * CH;
    for (int i = tid; i < lpad; i += blockDim.x) wb[i] = (i >= 1 && i < L) ? a.bins[i - 1] : (int8_t)0;
    float* wt = reinterpret_cast<float*>(smem + a.off_btab);
    double* wl = reinterpret_cast<double*>(smem + a.off_ls0);
    for (int b = tid; b < a.n_bins; b += blockDim.x) {
      int o = 0;
      for (int i = 0; i < K; ++i)
        for (int j = 0; j < K; ++j)
          if (i != j) wt[b * kBinRow + o++] = tab.A[b][i][j] / tab.A[b][j][j];
      for (int k = 1; k < K; ++k) wt[b * kBinRow + o++] = log2f(tab.A[b][k][k] / tab.A[b][0][0]);
      wl[b] = tab.log_s0b[b];
    }
  }
  if (STATS) {
    const int nw = blockDim.x >> 5;
    for (int i = tid; i < nw * a.S; i += blockDim.x) wacc[i] = 0.0;
  }
  // BINNED E-step: the bin of a position is the same for every read, and a thread owns the same positions of every
  // read it processes -- so the transition statistics are kept per thread and per bin in shared memory (one float4 =
  // the K x K = 4 sums of one bin; an odd number of float4 per thread keeps the 128-bit accesses conflict free)
  const int xi_stride = a.n_bins | 1;
  float4* s_xi = reinterpret_cast<float4*>(smem + a.off_sxi) + (size_t)tid * xi_stride;
  if (BINNED && STATS)
    for (int b = 0; b < a.n_bins; ++b) s_xi[b] = make_float4(0.f, 0.f, 0.f, 0.f);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();

  constexpr int NXI = STATS ? KK : 1;
  constexpr int NEM = STATS ? K : 1;
  constexpr int NEC = STATS ? K * C : 1;  // emission accumulators: [state][channel]
  float acc_xi[NXI], acc_num[NEC], acc_den[NEC], acc_start[NEM];
  double acc_ll = 0.0;
#pragma unroll
  for (int i = 0; i < NXI; ++i) acc_xi[i] = 0.0f;
#pragma unroll
  for (int i = 0; i < NEM; ++i) acc_start[i] = 0.0f;
#pragma unroll
  for (int i = 0; i < NEC; ++i) {
    acc_num[i] = 0.0f;
    acc_den[i] = 0.0f;
  }

  const int off_em = K + (BINNED ? a.n_bins : 1) * KK;  // emission statistics follow the transition block(s)
  auto flush_stats = [&]() {
    if (!STATS) return;
    double* my = wacc + (size_t)wcta * a.S;
#pragma unroll
    for (int k = 0; k < NEM; ++k) {
      const float r0 = warp_sum_f32(acc_start[k]);
      if (lane == 0) my[k] += (double)r0;
      acc_start[k] = 0.0f;
    }
#pragma unroll
    for (int i = 0; i < NEC; ++i) {
      const float r1 = warp_sum_f32(acc_num[i]);
      const float r2 = warp_sum_f32(acc_den[i]);
      if (lane == 0) {
        my[off_em + i] += (double)r1;
        my[off_em + K * C + i] += (double)r2;
      }
      acc_num[i] = 0.0f;
      acc_den[i] = 0.0f;
    }
    if (BINNED) {
      // per-bin sums: the constant factor A'_b[i][j] (unit diagonal) is applied here, once per flush
      for (int b = 0; b < a.n_bins; ++b) {
        const float4 q = s_xi[b];
        s_xi[b] = make_float4(0.f, 0.f, 0.f, 0.f);
        const float r00 = warp_sum_f32(q.x), r01 = warp_sum_f32(q.y), r10 = warp_sum_f32(q.z), r11 = warp_sum_f32(q.w);
        if (lane == 0) {
          const float* row = s_btab + b * kBinRow;  // [A'_01, A'_10, ...]
          double* dst = my + K + b * KK;
          dst[0] += (double)r00;
          dst[1 % KK] += (double)r01 * (double)row[0];
          dst[2 % KK] += (double)r10 * (double)row[1];
          dst[3 % KK] += (double)r11;
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < NXI; ++i) {
        const float r = warp_sum_f32(acc_xi[i]);
        if (lane == 0) my[K + i] += (double)r * (double)tab.Ap[(i / K) % K][i % K];
        acc_xi[i] = 0.0f;
      }
    }
  };

  const int t0 = gt * CH;
  const bool active = t0 < L;
  const bool first = (t0 == 0);
  const bool want_ll = !WALK && (STATS || (a.loglik != nullptr));
  const int gs = a.gamma_state;
  // OUT == 3 (K = 2 only): all-state posterior + states like OUT == 0, but the staging tile holds ONE float
  // per position -- the forward sweep stashes alpha as a signed ratio (the smaller component over the
  // larger, sign = which one), the backward sweep overwrites it with gamma_0, and the warp expands
  // (gamma_0, 1 - gamma_0) on its way out with coalesced 64-bit stores.  9 instead of 13 bytes of
  // shared memory per position: more reads resident per SM.
  constexpr bool GC = (OUT == 3);
  static_assert(!GC || (K == 2 && !STATS), "compact staging is for the K = 2 posterior kernel");
  const bool w_gamma_all = !STATS && (OUT == 0 || GC || (OUT == 2 && a.gamma != nullptr && gs < 0));
  const bool w_gamma_one = !STATS && (OUT == 1 || (OUT == 2 && a.gamma != nullptr && gs >= 0));
  const bool w_states = !STATS && (OUT != 2 || a.states != nullptr);
  const uint32_t row_bytes = (uint32_t)L * (uint32_t)(C * sizeof(ObsT));
  const char* obs_base = static_cast<const char*>(a.obs);
  // this warp's slice of the read: positions [w0, w1)
  const int w0 = gw * 32 * CH;
  const int w1 = min(w0 + 32 * CH, L);
  const int Lpad = ((L + CH - 1) / CH) * CH;

  auto issue_load = [&](long long n) {
    const char* src = obs_base + (size_t)n * row_bytes;
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15u);
    const uint32_t bytes = (mis + row_bytes + 15u) & ~15u;
    mbar_expect_tx(mbar, bytes);
    bulk_load(s_in_raw, src - mis, bytes, mbar);
  };

  const long long read_stride = (long long)gridDim.x * a.groups;
  long long n = (long long)blockIdx.x * a.groups + g;
  if (gt == 0 && n < a.N) issue_load(n);

  // column-normalised transitions: Ap[j][j] == 1 (see FbTables)
  float A[K][K];
#pragma unroll
  for (int i = 0; i < K; ++i)
#pragma unroll
    for (int j = 0; j < K; ++j) A[i][j] = tab.Ap[i][j];

  // row-vector x Ap  and  Ap x column-vector with the unit diagonal folded into the addend
  // uint8-coded observations have three possible emission ratios per state: evaluated once per thread (the same
  // ex2 of the same argument as the per-position expression, so results are bit-identical), selected per position
  constexpr bool CODE_TAB = (sizeof(ObsT) == 1) && C == 1 && !BINNED && !RECOMP;
  float rc0[K], rc1[K], rcm[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    rc0[k] = CODE_TAB ? ex2_approx(fmaf(0.0f, tab.rdl[k][0], tab.rl0p[k])) : 0.0f;
    rc1[k] = CODE_TAB ? ex2_approx(fmaf(1.0f, tab.rdl[k][0], tab.rl0p[k])) : 0.0f;
    rcm[k] = CODE_TAB ? ex2_approx(tab.rmiss[k]) : 0.0f;
  }
  // Packed fp32x2 (f32x2.cuh) for K = 3 only: measured 3-6 % faster there (21 % at 544 positions), while K = 2
  // (1-7 % slower: the pair moves cost what the packing saves) and K = 4 (6-20 % slower) keep the scalar code
  // (profiles/r02_policy_audit4.txt against r02_policy_audit.txt).
  constexpr bool PK = (K == 3);
  auto vecA = [&](const float (&Am)[K][K], const float (&x)[K], float (&y)[K]) { vec_mat_unit<K, PK>(x, Am, y); };
  auto Avec = [&](const float (&Am)[K][K], const float (&x)[K], float (&y)[K]) { mat_vec_unit<K, PK>(Am, x, y); };
  // BINNED: the table row of the transition into position t (off-diagonal entries of A'_b, then log2 scale ratios)
  auto bin_row = [&](int t, float (&Ab)[K][K], float (&rs)[K]) {
    const float* row = s_btab + (int)s_bins[t] * kBinRow;
    float r[kBinRow];
    if (K == 2) {
      const float4 q = *reinterpret_cast<const float4*>(row);
      r[0] = q.x;
      r[1] = q.y;
      r[2] = q.z;
      r[3] = q.w;
    } else {
      const float4 q0 = *reinterpret_cast<const float4*>(row);
      const float4 q1 = *reinterpret_cast<const float4*>(row + 4);
      r[0] = q0.x;
      r[1] = q0.y;
      r[2] = q0.z;
      r[3] = q0.w;
      r[4 % kBinRow] = q1.x;
      r[5 % kBinRow] = q1.y;
      r[6 % kBinRow] = q1.z;
      r[7 % kBinRow] = q1.w;
    }
    int o = 0;
#pragma unroll
    for (int i = 0; i < K; ++i)
#pragma unroll
      for (int j = 0; j < K; ++j) Ab[i][j] = (i == j) ? 1.0f : r[o++];
    rs[0] = 0.0f;
#pragma unroll
    for (int k = 1; k < K; ++k) rs[k] = r[o++];
  };

  int iter = 0;
  for (; n < a.N; n += read_stride, ++iter) {
    // where this read's rows sit relative to 16-byte granules (uniform per group)
    const uint32_t in_mis =
        (uint32_t)(reinterpret_cast<uintptr_t>(obs_base + (size_t)n * row_bytes) & 15u) / (uint32_t)sizeof(ObsT);
    ObsT* s_in = reinterpret_cast<ObsT*>(s_in_raw) + in_mis;
    const size_t obuf = (a.nbuf > 1 && (iter & 1)) ? (size_t)a.obuf_bytes : 0;
    float* s_stage = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(s_stage_base) + obuf);
    float* s_out1 = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(s_out1_base) + obuf);
    int8_t* s_st = s_st_base + obuf;
    if (w_gamma_all && !GC) s_stage += (reinterpret_cast<uintptr_t>(a.gamma + (size_t)n * L * K) & 15u) >> 2;
    if (w_gamma_one) s_out1 += (reinterpret_cast<uintptr_t>(a.gamma + (size_t)n * L) & 15u) >> 2;
    if (w_states) s_st += reinterpret_cast<uintptr_t>(a.states + (size_t)n * L) & 15u;

    float bvv[WALK ? K : 1], bww[WALK ? K : 1];
    if (WALK) {
      // boundary vectors of this thread's chunk (issued before waiting for the observations)
      const bool has_v = active && !first;
      const bool has_w = active && gt < a.Tact - 1;
      const size_t boff = ((size_t)gt * (size_t)a.N + (size_t)n) * K;
#pragma unroll
      for (int k = 0; k < (WALK ? K : 1); ++k) {
        bvv[k] = has_v ? __ldg(a.bnd_v + boff + k) : tab.pi2[k];
        bww[k] = has_w ? __ldg(a.bnd_w + boff + k) : 1.0f;
      }
    }

    mbar_wait(mbar, (uint32_t)(iter & 1));

    // ---------------- phase 1: chunk -> registers, transfer-matrix product ----------------
    float rt[RECOMP ? 1 : CH][KR];          // c_k(t) = s_k b_k / (s_0 b_0), k = 1..K-1
    float ov[((STATS && C == 1) || RECOMP) ? CH : 1];  // observation values (NaN = missing)
    auto ratios = [&](int j, float (&c)[K]) {
      c[0] = 1.0f;
      if (RECOMP) {
        const float o = ov[j];
        const bool m = (o == o);
#pragma unroll
        for (int k = 1; k < K; ++k) c[k] = ex2_approx(m ? fmaf(o, tab.rdl[k][0], tab.rl0p[k]) : tab.rmiss[k]);
      } else {
#pragma unroll
        for (int k = 1; k < K; ++k) c[k] = rt[RECOMP ? 0 : j][k - 1];
      }
    };
    using mask_t = typename std::conditional<(CH > 32), unsigned long long, unsigned>::type;
    mask_t obsmask = 0;        // bit j: position t0+j observed (E-step only)
    float P[K][K];
    mat_set_identity<K>(P);
    float ll_sumo[C];
    int ll_cnt[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      ll_sumo[c] = 0.0f;
      ll_cnt[c] = 0;
    }
    bool prev_obs0 = false;  // observation just before this chunk present? (E-step xi validity)
    unsigned in_dep = 0;
    if (active) {
      if (STATS && !first) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const float op = obs_to_float<ObsT>(s_in[(t0 - 1) * C + c]);
          prev_obs0 = prev_obs0 || (op == op);  // multi-channel: a position counts when any channel is observed
        }
      }
      if (t0 + CH > L) {
        // pad the ragged end with missing observations (this thread is the only reader)
        for (int t = L * C; t < (t0 + CH) * C; ++t) s_in[t] = obs_missing<ObsT>();
      }
      if (want_ll) {
        // state-0 emission mass of the chunk as (count, sum o) per channel: combined in fp64 below
        for (int j = 0; j < CH; ++j) {
#pragma unroll
          for (int c = 0; c < C; ++c) {
            const float o = obs_to_float<ObsT>(s_in[(t0 + j) * C + c]);
            if (o == o) {
              ll_sumo[c] += o;
              ++ll_cnt[c];
            }
          }
        }
      }
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        float c[K];
        c[0] = 1.0f;
        bool m;
        float o;
        float Ab[K][K], rsb[K];
        if (BINNED) bin_row(t0 + j, Ab, rsb);
        const float(&Am)[K][K] = BINNED ? Ab : A;
        if (BINNED) {
          o = obs_to_float<ObsT>(s_in[t0 + j]);
          m = (o == o);
          const bool start = first && j == 0;  // no transition into the first position: no scale ratio either
#pragma unroll
          for (int k = 1; k < K; ++k)
            c[k] = ex2_approx((m ? fmaf(o, tab.rdl[k][0], tab.rl0[k][0]) : 0.0f) + (start ? 0.0f : rsb[k]));
        } else if (CODE_TAB) {
          const unsigned code = (unsigned)s_in[t0 + j];
          m = code < 2u;
          o = m ? (code == 1u ? 1.0f : 0.0f) : __int_as_float(0x7fc00000);
#pragma unroll
          for (int k = 1; k < K; ++k) c[k] = (code == 0u) ? rc0[k] : ((code == 1u) ? rc1[k] : rcm[k]);
        } else if (C == 1) {
          o = obs_to_float<ObsT>(s_in[t0 + j]);
          m = (o == o);
#pragma unroll
          for (int k = 1; k < K; ++k) c[k] = ex2_approx(m ? fmaf(o, tab.rdl[k][0], tab.rl0p[k]) : tab.rmiss[k]);
        } else {
          // log2 of the ratio: the transition-scale term + one term per observed channel
          float x[K];
#pragma unroll
          for (int k = 1; k < K; ++k) x[k] = tab.rmiss[k];
          m = false;
          o = 0.0f;
#pragma unroll
          for (int ch = 0; ch < C; ++ch) {
            const float oc = obs_to_float<ObsT>(s_in[(t0 + j) * C + ch]);
            const bool mc = (oc == oc);
            m = m || mc;
#pragma unroll
            for (int k = 1; k < K; ++k) x[k] += mc ? fmaf(oc, tab.rdl[k][ch], tab.rl0[k][ch]) : 0.0f;
          }
#pragma unroll
          for (int k = 1; k < K; ++k) c[k] = ex2_approx(x[k]);
        }
        if (STATS) obsmask |= m ? ((mask_t)1 << j) : (mask_t)0;
        if ((STATS && C == 1) || RECOMP) ov[j] = o;
#pragma unroll
        for (int k = 1; k < K; ++k) {
          if (!RECOMP) rt[RECOMP ? 0 : j][k - 1] = c[k];
        }
        if (WALK) {
          // Nothing consumes the observation before the barrier below when the ratios are
          // recomputed later, and a barrier does not wait for shared-memory loads still in
          // flight: the prefetch (async proxy) could overwrite the buffer under them.  Fold the
          // values into a word that is tested before the barrier so the loads must have landed.
          if (RECOMP) in_dep ^= __float_as_uint(o);
          continue;
        }
        if (j == 0) {
          // P = I so far: P <- Ap diag(c), or diag(c) for the first position of the read
#pragma unroll
          for (int i = 0; i < K; ++i)
#pragma unroll
            for (int jj = 0; jj < K; ++jj) {
              const float aij = (i == jj) ? 1.0f : (first ? 0.0f : Am[i][jj]);
              P[i][jj] = (jj == 0) ? aij : aij * c[jj];
            }
        } else {
#pragma unroll
          for (int i = 0; i < K; ++i) {
            float q[K];
            vecA(Am, P[i], q);
            vec_mul_tail<K, PK>(P[i], q, c);
          }
        }
        if ((j & 3) == 3 || j == CH - 1) mat_scale<K, PK>(P, pow2_rescale_noexp(mat_max<K>(P)));
      }
    }

    float v[K], w[K];
    if constexpr (WALK) {
      if (RECOMP && in_dep == 0x7fa5a5a5u) __nanosleep(1);  // (never-taken in practice) consumer of the loads
      if (G > 1) group_sync(bar_id, T); else __syncwarp();  // B1: chunk registers loaded by everyone
      if (gt == 0 && !a.alias_in && !a.late_load) {
        const long long nn = n + read_stride;
        if (nn < a.N) issue_load(nn);
      }
#pragma unroll
      for (int k = 0; k < K; ++k) {
        v[k] = bvv[k];
        w[k] = bww[k];
      }
    } else {

    // ---------------- phase 2: warp-shuffle scans of the chunk matrices ----------------
    // Kogge-Stone over the 32 lanes, inclusive prefix (pre) and suffix (suf); branch-free, with a
    // power-of-two rescale every second level (entries stay far inside the fp32 range).
    float pre[K][K], suf[K][K];
#pragma unroll
    for (int i = 0; i < K; ++i)
#pragma unroll
      for (int j = 0; j < K; ++j) {
        pre[i][j] = P[i][j];
        suf[i][j] = P[i][j];
      }
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      float q[K][K], r[K][K];
#pragma unroll
      for (int i = 0; i < K; ++i)
#pragma unroll
        for (int j = 0; j < K; ++j) q[i][j] = __shfl_up_sync(kFull, pre[i][j], d);
      mat_mul<K, PK>(q, pre, r);
      const bool up_ok = lane >= d;
#pragma unroll
      for (int i = 0; i < K; ++i)
#pragma unroll
        for (int j = 0; j < K; ++j) pre[i][j] = up_ok ? r[i][j] : pre[i][j];
#pragma unroll
      for (int i = 0; i < K; ++i)
#pragma unroll
        for (int j = 0; j < K; ++j) q[i][j] = __shfl_down_sync(kFull, suf[i][j], d);
      mat_mul<K, PK>(suf, q, r);
      const bool dn_ok = lane + d < 32;
#pragma unroll
      for (int i = 0; i < K; ++i)
#pragma unroll
        for (int j = 0; j < K; ++j) suf[i][j] = dn_ok ? r[i][j] : suf[i][j];
      if (d == 2 || d == 8) {
        mat_scale<K, PK>(pre, pow2_rescale_noexp(mat_max<K>(pre)));
        mat_scale<K, PK>(suf, pow2_rescale_noexp(mat_max<K>(suf)));
      }
    }

    float u[K], z[K];
    if (G > 1) {
      if (lane == 31) {
        // warp total (normalised)
        const float sc = pow2_rescale_noexp(mat_max<K>(pre));
#pragma unroll
        for (int i = 0; i < K; ++i)
#pragma unroll
          for (int j = 0; j < K; ++j) s_wtot[gw * KK + i * K + j] = pre[i][j] * sc;
      }
      group_sync(bar_id, T);  // B1: chunk registers loaded by everyone, warp totals visible
    } else {
      __syncwarp();
    }
    if (gt == 0 && !a.alias_in && !a.late_load) {
      // the observation buffer is free: prefetch the next read of this group
      const long long nn = n + read_stride;
      if (nn < a.N) issue_load(nn);
    }
    if (G > 1) {
      // cross-warp boundary vectors by two threads: u_w = pi2 T_0..T_{w-1} (thread 0) and
      // z_w = T_{w+1}..T_{G-1} 1 (thread 32); G <= 32 short dependent steps each
      // (the next warp total is fetched while the current step computes; rescale every 4th step)
      if (gt == 0) {
        float x[K];
#pragma unroll
        for (int k = 0; k < K; ++k) x[k] = tab.pi2[k];
        float tn[KK];
#pragma unroll
        for (int e2 = 0; e2 < KK; ++e2) tn[e2] = s_wtot[e2];
        for (int wv = 0; wv < G; ++wv) {
          float tc[KK];
#pragma unroll
          for (int e2 = 0; e2 < KK; ++e2) tc[e2] = tn[e2];
          if (wv + 1 < G) {
#pragma unroll
            for (int e2 = 0; e2 < KK; ++e2) tn[e2] = s_wtot[(wv + 1) * KK + e2];
          }
#pragma unroll
          for (int k = 0; k < K; ++k) s_u[wv * K + k] = x[k];
          float y[K];
#pragma unroll
          for (int j = 0; j < K; ++j) {
            float sacc = x[0] * tc[j];
#pragma unroll
            for (int i = 1; i < K; ++i) sacc = fmaf(x[i], tc[i * K + j], sacc);
            y[j] = sacc;
          }
          if ((wv & 3) == 3) {
            const float sc = pow2_rescale_noexp(vec_max<K>(y));
#pragma unroll
            for (int k = 0; k < K; ++k) y[k] *= sc;
          }
#pragma unroll
          for (int k = 0; k < K; ++k) x[k] = y[k];
        }
      } else if (gt == 32) {
        float x[K];
#pragma unroll
        for (int k = 0; k < K; ++k) x[k] = 1.0f;
        float tn[KK];
#pragma unroll
        for (int e2 = 0; e2 < KK; ++e2) tn[e2] = s_wtot[(G - 1) * KK + e2];
        for (int wv = G - 1; wv >= 0; --wv) {
          float tc[KK];
#pragma unroll
          for (int e2 = 0; e2 < KK; ++e2) tc[e2] = tn[e2];
          if (wv > 0) {
#pragma unroll
            for (int e2 = 0; e2 < KK; ++e2) tn[e2] = s_wtot[(wv - 1) * KK + e2];
          }
#pragma unroll
          for (int k = 0; k < K; ++k) s_z[wv * K + k] = x[k];
          float y[K];
#pragma unroll
          for (int i = 0; i < K; ++i) {
            float sacc = tc[i * K] * x[0];
#pragma unroll
            for (int j = 1; j < K; ++j) sacc = fmaf(tc[i * K + j], x[j], sacc);
            y[i] = sacc;
          }
          if (((G - 1 - wv) & 3) == 3) {
            const float sc = pow2_rescale_noexp(vec_max<K>(y));
#pragma unroll
            for (int k = 0; k < K; ++k) y[k] *= sc;
          }
#pragma unroll
          for (int k = 0; k < K; ++k) x[k] = y[k];
        }
      }
      group_sync(bar_id, T);  // B2: boundary vectors visible
#pragma unroll
      for (int k = 0; k < K; ++k) {
        u[k] = s_u[gw * K + k];
        z[k] = s_z[gw * K + k];
      }
    } else {
      __syncwarp();
#pragma unroll
      for (int k = 0; k < K; ++k) {
        u[k] = tab.pi2[k];
        z[k] = 1.0f;
      }
    }

    {
      float vo[K], wo[K];
#pragma unroll
      for (int j = 0; j < K; ++j) {
        float s = u[0] * pre[0][j];
#pragma unroll
        for (int i = 1; i < K; ++i) s = fmaf(u[i], pre[i][j], s);
        vo[j] = s;
      }
#pragma unroll
      for (int i = 0; i < K; ++i) {
        float s = suf[i][0] * z[0];
#pragma unroll
        for (int j = 1; j < K; ++j) s = fmaf(suf[i][j], z[j], s);
        wo[i] = s;
      }
#pragma unroll
      for (int k = 0; k < K; ++k) {
        v[k] = __shfl_up_sync(kFull, vo[k], 1);
        w[k] = __shfl_down_sync(kFull, wo[k], 1);
        if (lane == 0) v[k] = u[k];
        if (lane == 31) w[k] = z[k];
      }
    }
    }  // !WALK
    {
      float sv = v[0], sw = w[0];
#pragma unroll
      for (int k = 1; k < K; ++k) {
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

    // this warp's previous bulk stores out of the staging set it is about to overwrite are done
    if (!STATS) {
      if (lane == 0) {
        if (a.nbuf > 1) bulk_wait_read_1(); else bulk_wait_read_all();
      }
      __syncwarp();
    }

    // ---------------- phase 3: local forward / backward sweeps (fully unrolled) ----------------
    double ll_thread = 0.0;
    if (active) {
      float al[K];
      int e = 0;
#pragma unroll
      for (int k = 0; k < K; ++k) al[k] = v[k];
      float* stg = s_stage + (size_t)t0 * (GC ? 1 : K);
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        float nx[K];
        float cf[K];
        ratios(j, cf);
        float Ab[K][K], rsb[K];
        if (BINNED) bin_row(t0 + j, Ab, rsb);
        const float(&Am)[K][K] = BINNED ? Ab : A;
        if (j == 0) {
          float y[K];
          vecA(Am, al, y);
#pragma unroll
          for (int jj = 0; jj < K; ++jj) {
            const float s = first ? al[jj] : y[jj];  // alpha_0 = pi * b_0: no transition
            nx[jj] = (jj == 0) ? s : s * cf[jj];
          }
        } else {
          float y[K];
          vecA(Am, al, y);
          vec_mul_tail<K, PK>(nx, y, cf);
        }
        if ((j & 3) == 3) {
          const float sc = want_ll ? pow2_rescale(vec_max<K>(nx), e) : pow2_rescale_noexp(vec_max<K>(nx));
          vec_scale<K, PK>(nx, sc);
        }
#pragma unroll
        for (int k = 0; k < K; ++k) al[k] = nx[k];
        if (GC) {
          const float lo = fminf(nx[0], nx[1 % K]), hi = fmaxf(nx[0], nx[1 % K]);
          const float rho = lo * rcp_approx(hi);
          stg[j] = (nx[0] <= nx[1 % K]) ? rho : -rho;  // >= 0: alpha ~ (rho, 1); < 0: alpha ~ (1, rho)
        } else if (K == 2) {
          *reinterpret_cast<float2*>(stg + j * 2) = make_float2(nx[0], nx[1 % K]);
        } else if (K == 4) {
          *reinterpret_cast<float4*>(stg + j * 4) = make_float4(nx[0], nx[1 % K], nx[2 % K], nx[3 % K]);
        } else {
#pragma unroll
          for (int k = 0; k < K; ++k) stg[j * K + k] = nx[k];
        }
      }
      if (want_ll) {
        float s = al[0];
#pragma unroll
        for (int k = 1; k < K; ++k) s += al[k];
        // log P(chunk | past) with the dropped scalars (s_0 per transition, b_0 per observed site) restored
        ll_thread = (double)logf(s) + (double)e * 0.6931471805599453094;
        if (BINNED) {
          // the dropped scalar of every transition into this chunk's positions: s_b[0] of that transition's bin
          for (int j = first ? 1 : 0; j < CH; ++j) ll_thread += s_ls0[(int)s_bins[t0 + j]];
        } else {
          ll_thread += (double)(first ? CH - 1 : CH) * tab.log_s0;
        }
#pragma unroll
        for (int c = 0; c < C; ++c) ll_thread += (double)ll_cnt[c] * tab.l00[c] + (double)ll_sumo[c] * tab.dl0[c];
        if (first) ll_thread += tab.log_pi2_sum;
      }

      float bv[K];
#pragma unroll
      for (int k = 0; k < K; ++k) bv[k] = w[k];
      float alc[K];
#pragma unroll
      for (int k = 0; k < K; ++k) alc[k] = al[k];  // alpha of the chunk's last position
#pragma unroll
      for (int j = CH - 1; j >= 0; --j) {
        float gk[K];
        vec_mul<K, PK>(gk, alc, bv);
        float s = 0.0f;
#pragma unroll
        for (int k = 0; k < K; ++k) s += gk[k];
        const float r = rcp_approx(s);
        vec_scale<K, PK>(gk, r);
        // alpha of the previous position (this thread's chunk, or the incoming prefix vector)
        float alp[K];
        if (j > 0) {
          if (GC) {
            const float e1 = stg[j - 1];
            const float rho = fabsf(e1);
            alp[0] = (e1 >= 0.0f) ? rho : 1.0f;
            alp[1 % K] = (e1 >= 0.0f) ? 1.0f : rho;
          } else if (K == 2) {
            const float2 t2 = *reinterpret_cast<const float2*>(stg + (j - 1) * 2);
            alp[0] = t2.x;
            alp[1 % K] = t2.y;
          } else if (K == 4) {
            const float4 t4 = *reinterpret_cast<const float4*>(stg + (j - 1) * 4);
            alp[0] = t4.x;
            alp[1 % K] = t4.y;
            alp[2 % K] = t4.z;
            alp[3 % K] = t4.w;
          } else {
#pragma unroll
            for (int k = 0; k < K; ++k) alp[k] = stg[(j - 1) * K + k];
          }
        } else {
#pragma unroll
          for (int k = 0; k < K; ++k) alp[k] = v[k];
        }
        if (!STATS) {
          if (GC) {
            stg[j] = gk[0];
          } else if (w_gamma_all) {
            if (K == 2) {
              *reinterpret_cast<float2*>(stg + j * 2) = make_float2(gk[0], gk[1 % K]);
            } else if (K == 4) {
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
            s_out1[t0 + j] = gsel;
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
            s_st[t0 + j] = (int8_t)best;
          }
        }

        if (STATS) {
          if (j == 0 && first) {
#pragma unroll
            for (int k = 0; k < K; ++k) acc_start[k] += gk[k];
          }
          if (C == 1) {
            if ((obsmask >> j) & (mask_t)1) {
              if constexpr (STATS && C == 1) {
                vec_fma_vs<K, PK>(acc_num, gk, ov[(STATS && C == 1) ? j : 0]);
                vec_add<K, PK>(acc_den, gk);
              }
            }
          } else {
            // per-channel emission statistics (HMM.py:1874-1880): the observations are still in shared memory
#pragma unroll
            for (int ch = 0; ch < C; ++ch) {
              const float oc = obs_to_float<ObsT>(s_in[(t0 + j) * C + ch]);
              const float mc = (oc == oc) ? 1.0f : 0.0f;
              const float ov1 = (oc == oc) ? oc : 0.0f;
#pragma unroll
              for (int k = 0; k < K; ++k) {
                const float gm = gk[k] * mc;
                acc_num[(STATS ? k * C + ch : 0)] = fmaf(gm, ov1, acc_num[(STATS ? k * C + ch : 0)]);
                acc_den[(STATS ? k * C + ch : 0)] += gm;
              }
            }
          }
        }

        const bool need_step = (j > 0) || STATS;
        if (need_step) {
          float cb[K], nb[K], cr[K];
          ratios(j, cr);
          vec_mul_tail<K, PK>(cb, bv, cr);
          float Ab[K][K], rsb[K];
          if (BINNED) bin_row(t0 + j, Ab, rsb);
          const float(&Am)[K][K] = BINNED ? Ab : A;
          Avec(Am, cb, nb);
          if (STATS) {
            // xi_{t-1}(i,jj) ~ alpha_{t-1}(i) Ap[i][jj] cb[jj], t = t0+j > 0, counted when both
            // endpoints are observed (HMM.py:1583-1591)
            const bool prev_obs = (j > 0) ? (((obsmask >> (j > 0 ? j - 1 : 0)) & (mask_t)1) != 0) : prev_obs0;
            const bool valid = ((obsmask >> j) & (mask_t)1) && prev_obs;
            float s2 = 0.0f;
#pragma unroll
            for (int i = 0; i < K; ++i) s2 = fmaf(alp[i], nb[i], s2);
            const float r2 = valid ? rcp_approx(s2) : 0.0f;
            // the constant factor Ap[i][jj] is applied once per flush (flush_stats)
            if (BINNED) {
              // K = 2: one 128-bit read-modify-write of this thread's slot of the transition's bin
              float4* slot = s_xi + (int)s_bins[t0 + j];
              float4 q = *slot;
              const float a0 = alp[0] * r2, a1 = alp[1 % K] * r2;
              q.x = fmaf(a0, cb[0], q.x);
              q.y = fmaf(a0, cb[1 % K], q.y);
              q.z = fmaf(a1, cb[0], q.z);
              q.w = fmaf(a1, cb[1 % K], q.w);
              *slot = q;
            } else {
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
          if ((((CH - 1) - j) & 3) == 3) {
            const float sc = pow2_rescale_noexp(vec_max<K>(nb));
            vec_scale<K, PK>(nb, sc);
          }
#pragma unroll
          for (int k = 0; k < K; ++k) bv[k] = nb[k];
        }
#pragma unroll
        for (int k = 0; k < K; ++k) alc[k] = alp[k];
      }
    }

    if (STATS) acc_ll += (WALK ? ((active && first) ? a.ll_in[n] : 0.0) : ll_thread);

    // ---------------- store: every warp bulk-stores its own slice (no CTA barrier) ----------------
    if (!STATS) {
      fence_proxy_async();  // make this lane's staging writes visible to the bulk-copy engine
      __syncwarp();
      if (w0 < L) {
        const uint32_t npos = (uint32_t)(w1 - w0);
        if (GC) {
          // expand (gamma_0, 1 - gamma_0) with coalesced 64-bit stores (global rows are 8-byte aligned)
          float2* grow = reinterpret_cast<float2*>(a.gamma) + (size_t)n * L + w0;
          const float* srow = s_stage + w0;
          for (uint32_t i = lane; i < npos; i += 32) {
            const float g0 = srow[i];
            grow[i] = make_float2(g0, 1.0f - g0);
          }
        } else if (w_gamma_all)
          store_row_bulk<float>(a.gamma + ((size_t)n * L + w0) * K, s_stage + (size_t)w0 * K, npos * K * 4u, lane);
        if (w_gamma_one) store_row_bulk<float>(a.gamma + (size_t)n * L + w0, s_out1 + w0, npos * 4u, lane);
        if (w_states) store_row_bulk<int8_t>(a.states + (size_t)n * L + w0, s_st + w0, npos, lane);
      }
      if (lane == 0) bulk_commit();  // one group per iteration (possibly empty) keeps the counts aligned
    }
    if (!WALK && a.loglik != nullptr) {
      const double wsum = warp_sum_f64(ll_thread);
      if (lane == 0) s_ll[gw] = wsum;
      group_sync(bar_id, T);
      if (gt == 0) {
        double s = 0.0;
        for (int i = 0; i < G; ++i) s += s_ll[i];
        a.loglik[n] = s;
      }
    }
    if (STATS && ((iter + 1) % a.flush_every) == 0) flush_stats();
    if (a.alias_in || a.late_load) {
      // the staging buffer doubles as the observation buffer (alias_in), or the backward sweep was still reading
      // the observations (late_load): the next read may only be loaded once every warp is done with the buffer
      if (!STATS && a.alias_in && lane == 0) bulk_wait_read_all();
      group_sync(bar_id, T);
      if (gt == 0) {
        const long long nn = n + read_stride;
        if (nn < a.N) issue_load(nn);
      }
    }
  }

  if (!STATS && lane == 0) bulk_wait_all();  // all bulk stores complete before the CTA retires
  if (STATS) {
    flush_stats();
    const double wl = warp_sum_f64(acc_ll);
    if (lane == 0) wacc[(size_t)wcta * a.S + (a.S - 1)] += wl;
    __syncthreads();
    const int nw = blockDim.x >> 5;
    for (int s = tid; s < a.S; s += blockDim.x) {
      double tot = 0.0;
      for (int wv = 0; wv < nw; ++wv) tot += wacc[(size_t)wv * a.S + s];
      a.partials[(size_t)blockIdx.x * a.S + s] = tot;
    }
  }
  (void)Lpad;
}

}  // namespace smf
