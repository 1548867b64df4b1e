// C-ABI of libsmfhmm.so (see include/smf_hmm.h): argument checks, launch-shape selection and
// dispatch to the per-K launchers.  No torch, no device allocation.  Process state (all of it thread-safe,
// none of it affects results): cached device attributes, the per-kernel shared-memory opt-ins, the two
// path-selection knobs of smf_hmm_set_option, and environment overrides read once at load.
#include <cuda_runtime.h>
#include <limits.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <map>
#include <mutex>
#include <tuple>

#include "features_kernel.cuh"
#include "launch.h"
#include "smf_hmm.h"
#include "post_sweep_kernel.cuh"
#include "seq_kernel.cuh"
#include <algorithm>
#include <vector>
#include "short_kernel.cuh"
#include "sweep_kernel.cuh"
#include "walk_kernel.cuh"

namespace smf {

// Fixed-order reduction of the per-CTA partial statistics.
__global__ void reduce_partials_kernel(const double* __restrict__ partials, int nblocks, int S,
                                       double* __restrict__ stats) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  double tot = 0.0;
  for (int b = 0; b < nblocks; ++b) tot += partials[(size_t)b * S + s];
  stats[s] = tot;
}

// E-step workspace: per-CTA partial statistics, for at most this many CTAs.
static const int kMaxEstepBlocks = 148 * 8;

static std::mutex g_state_mutex;  // guards the caches below (short critical sections, host only)

static int get_device_info(DeviceInfo* out) {
  static DeviceInfo cache[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return SMF_ERR_CUDA;
  std::lock_guard<std::mutex> lock(g_state_mutex);
  if (cache[dev].device != dev) {
    DeviceInfo d;
    if (cudaDeviceGetAttribute(&d.num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
      return SMF_ERR_CUDA;
    if (cudaDeviceGetAttribute(&d.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) !=
        cudaSuccess)
      return SMF_ERR_CUDA;
    d.device = dev;
    cache[dev] = d;
  }
  *out = cache[dev];
  return SMF_OK;
}

struct KernelState {
  int smem_set = -1;      // largest dynamic shared-memory size opted in so far
  bool carveout = false;  // max-shared carve-out already requested
  std::map<std::pair<int, int>, int> occ;  // (block, smem) -> resident CTAs per SM
};
static std::map<std::pair<const void*, int>, KernelState> g_kernels;

int prepare_kernel(const void* kern, int smem_bytes, bool max_carveout) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return SMF_ERR_CUDA;
  std::lock_guard<std::mutex> lock(g_state_mutex);
  KernelState& ks = g_kernels[std::make_pair(kern, dev)];
  if (smem_bytes > ks.smem_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) != cudaSuccess)
      return SMF_ERR_CUDA;
    ks.smem_set = smem_bytes;
  }
  if (max_carveout && !ks.carveout) {
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    ks.carveout = true;
  }
  return SMF_OK;
}

int cached_occupancy(const void* kern, int block, int smem_bytes, int* occ_out) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return SMF_ERR_CUDA;
  std::lock_guard<std::mutex> lock(g_state_mutex);
  KernelState& ks = g_kernels[std::make_pair(kern, dev)];
  auto key = std::make_pair(block, smem_bytes);
  auto it = ks.occ.find(key);
  if (it == ks.occ.end()) {
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, block, (size_t)smem_bytes) != cudaSuccess)
      return SMF_ERR_CUDA;
    it = ks.occ.emplace(key, occ).first;
  }
  *occ_out = it->second;
  return SMF_OK;
}

// SMF_HMM_FORCE_GENERIC=1 routes everything through the generic kernel (testing / A-B timing).
// smf_hmm_set_option("generic", 0 / 1) does the same at run time.
static std::atomic<bool> g_force_generic{[] {
  const char* e = getenv("SMF_HMM_FORCE_GENERIC");
  return e != nullptr && e[0] == '1';
}()};

static int check_tables(const smf_hmm_tables* t) {
  if (t == nullptr) return SMF_ERR_BAD_ARG;
  if (t->n_states < 2 || t->n_states > kMaxK) return SMF_ERR_BAD_ARG;
  if (t->n_channels < 1 || t->n_channels > kMaxC) return SMF_ERR_BAD_ARG;
  if (t->n_bins < 1 || t->n_bins > kMaxB) return SMF_ERR_BAD_ARG;
  if (!t->log_start || !t->log_trans || !t->log_emit1 || !t->log_emit0) return SMF_ERR_BAD_ARG;
  return SMF_OK;
}

// two-kernel path selection: -1 = measured policy (walk_eligible), 0 = never, 1 = whenever possible.
// Initialised from SMF_FB2_WALK, changed at run time with smf_hmm_set_option("walk", v).
static std::atomic<int> g_walk_mode{[] { const char* e = getenv("SMF_FB2_WALK"); return e ? atoi(e) : -1; }()};
static int rescale_cadence(const smf_hmm_tables* t);

static int fb_max_threads(int K, bool stats) { return fb_max_threads_c(K, stats); }
static int fb2_max_threads(int K, bool stats) { return fb2_max_threads_c(K, stats); }

// Launch-shape selection for the generic forward-backward kernel.  A read that does not fit the
// on-chip staging at once is cut into nseg <= kMaxSeg segments (two passes, see fb_kernel.cuh).
static int plan_fb(int K, int C, int nb, int64_t L64, bool stats, int S, int max_smem, FbArgs* a,
                   FbPlan* plan) {
  if (L64 < 1 || L64 > (1 << 24)) return SMF_ERR_TOO_LONG;
  const int L = (int)L64;
  const int maxT = fb_max_threads(K, stats);
  const int KK = K * K;
  const bool binned = nb > 1;
  const int scan_b = align_up(32 * (KK + 2 * K) * 4 + 32 * 8 + kMaxSeg * (KK + 2 * K) * 4, 16);
  const int sA_b = binned ? align_up(nb * KK * 4, 16) : 0;
  auto bytes_for = [&](int Lseg) {
    return align_up(Lseg * C * 4, 16) + align_up(Lseg * K * 4, 16) + align_up(Lseg + 16, 16) + scan_b;
  };
  // single segment: the layout used since v1
  {
    int T = align_up((L + 16) / 17, 32);
    if (T > maxT) T = maxT;
    if (T < 32) T = 32;
    int chunk = (L + T - 1) / T;
    if ((chunk & 1) == 0) ++chunk;
    T = align_up((L + chunk - 1) / chunk, 32);
    const int group_bytes = bytes_for(L);
    int groups = 256 / T;
    if (groups < 1) groups = 1;
    if (groups > 8) groups = 8;
    for (; groups >= 1; --groups) {
      const int warps = groups * T / 32;
      const int bins_b = binned ? align_up(L, 16) : 0;
      const int wacc_b = stats ? align_up(warps * S * 8, 16) : 0;
      const int total = bins_b + sA_b + wacc_b + groups * group_bytes;
      if (total <= max_smem) {
        a->off_bins = 0;
        a->off_sA = bins_b;
        a->off_wacc = bins_b + sA_b;
        a->off_group0 = bins_b + sA_b + wacc_b;
        a->group_bytes = group_bytes;
        a->goff_in = 0;
        a->goff_stage = align_up(L * C * 4, 16);
        a->goff_st = a->goff_stage + align_up(L * K * 4, 16);
        a->goff_scan = a->goff_st + align_up(L + 16, 16);
        a->chunk = chunk;
        a->T = T;
        a->groups = groups;
        a->nseg = 1;
        plan->T = T;
        plan->chunk = chunk;
        plan->groups = groups;
        plan->smem = total;
        plan->block = T * groups;
        return SMF_OK;
      }
    }
  }
  // segmented: all threads, the longest odd chunk whose segment fits
  {
    const int T = maxT;
    const int wacc_b = stats ? align_up((T / 32) * S * 8, 16) : 0;
    int chunk = 1;
    for (int c = 3; c < 4096; c += 2) {
      if (sA_b + wacc_b + bytes_for(T * c) > max_smem) break;
      chunk = c;
    }
    const int Lseg = T * chunk;
    if (sA_b + wacc_b + bytes_for(Lseg) > max_smem) return SMF_ERR_TOO_LONG;
    const int nseg = (L + Lseg - 1) / Lseg;
    if (nseg > kMaxSeg) return SMF_ERR_TOO_LONG;
    a->off_bins = 0;
    a->off_sA = 0;
    a->off_wacc = sA_b;
    a->off_group0 = sA_b + wacc_b;
    a->group_bytes = bytes_for(Lseg);
    a->goff_in = 0;
    a->goff_stage = align_up(Lseg * C * 4, 16);
    a->goff_st = a->goff_stage + align_up(Lseg * K * 4, 16);
    a->goff_scan = a->goff_st + align_up(Lseg + 16, 16);
    a->chunk = chunk;
    a->T = T;
    a->groups = 1;
    a->nseg = nseg;
    plan->T = T;
    plan->chunk = chunk;
    plan->groups = 1;
    plan->smem = sA_b + wacc_b + bytes_for(Lseg);
    plan->block = T;
    return SMF_OK;
  }
}

// ---------------------------------------------------------------------------------------
// fast path (fb2_kernel): single channel, homogeneous transitions, compile-time chunk (17 or 33)
// ---------------------------------------------------------------------------------------
// Resident reads per SM for a candidate layout: shared memory (228 KB per SM, 1 KB reserved per CTA),
// registers (64 K per SM) and the 32-CTA limit.
constexpr int kWalkChunkDefault = 17;
static int fb2_resident(int total_smem, int block, int groups, int regs) {
  const int by_smem = (228 * 1024) / (total_smem + 1024);
  const int by_regs = 65536 / (regs * block);
  int ctas = by_smem < by_regs ? by_smem : by_regs;
  if (ctas > 32) ctas = 32;
  return ctas * groups;
}

static int plan_fb2_chunk(int K, int CH, int64_t L64, bool stats, int S, int es, bool want_gamma_one,
                          bool want_states, int max_smem, Fb2Args* a, Fb2Plan* plan, bool walk = false,
                          bool gc_ok = false /* caller produces the one output set the compact staging serves */,
                          int cta_extra = 0 /* CTA-wide bytes placed after the statistics accumulators */) {
  if (L64 < 1) return SMF_ERR_TOO_LONG;
  if (L64 > (int64_t)CH * 1024) return SMF_ERR_TOO_LONG;
  const int L = (int)L64;
  const int Tact = (L + CH - 1) / CH;
  const int T = align_up(Tact, 32);
  if (T > (walk ? fb2w_max_threads_c(K, stats) : fb2_max_threads(K, stats))) return SMF_ERR_TOO_LONG;
  const int Lpad = Tact * CH;
  const int KK = K * K;
  // registers/thread of the instantiation (ptxas)
  const int regs = walk ? (stats ? 128 : 96) : (((stats && K >= 3) || K == 4) ? 128 : 96);
  const int bar_b = 16;
  const int in_b = align_up((Lpad + 16) * es + 16, 16);
  // compact staging (fb2_kernel OUT == 3, opt-in): one float per position instead of K
  static const int env_gc = [] { const char* e = getenv("SMF_FB2_GC"); return e ? atoi(e) : 0; }();
  // measured: the extra encode / decode / expand instructions cost more than the residency brings
  // (C2 0.64 vs 0.80, 10 kb 0.55 vs 0.66) except where the regular layout no longer fits with a prefetch
  // buffer (>= ~18 kb: 0.65 vs 0.53 at 20 kb); SMF_FB2_GC=1 forces it for A/B runs
  const bool gc = (env_gc == 1 || (env_gc == 0 && L64 >= 18000 && CH == 33)) && gc_ok && K == 2 && !stats && !walk &&
                  !want_gamma_one && es == 4 && (CH == 33 || CH == 29);
  a->gc = gc ? 1 : 0;
  const int stage_b = align_up((Lpad * (gc ? 1 : K) + 4) * 4, 16);
  const int out1_b = want_gamma_one ? align_up((Lpad + 4) * 4, 16) : 0;
  const int st_b = want_states ? align_up(Lpad + 32, 16) : 0;
  const int scan_b = align_up(32 * (KK + 2 * K) * 4 + 32 * 8, 16);
  int groups = 256 / T;
  if (groups < 1) groups = 1;
  if (groups > 8) groups = 8;
  static const int env_nbuf = [] { const char* e = getenv("SMF_FB2_NBUF"); return e ? atoi(e) : 0; }();
  static const int env_alias = [] { const char* e = getenv("SMF_FB2_ALIAS"); return e ? atoi(e) : -1; }();
  // Candidate layouts, in order of preference at equal residency:
  //   double-buffered staging + prefetch  <  single-buffered + prefetch  <  single-buffered, obs aliased
  // The layout that keeps the most reads resident per SM wins (measured: profiles/r01_chunk_sweep.txt).
  int best_res = -1;
  for (int cand = 0; cand < 3; ++cand) {
    const int nbuf = (cand == 0 && !stats) ? 2 : 1;
    const bool alias = (cand == 2);
    if (cand == 0 && stats) continue;
    if (env_nbuf == 1 && nbuf == 2) continue;
    if (env_nbuf == 2 && nbuf == 1) continue;
    // measured: losing the prefetch costs more than the extra resident reads bring (C2 0.71 vs 0.80,
    // 10 kb 0.60 vs 0.64 of HBM peak), so the aliased layout is opt-in (SMF_FB2_ALIAS=1) for now
    // ... it is still the only fast-path layout for the longest reads (e.g. 20 kb, K = 2)
    if (env_alias != 1 && alias && best_res > 0) continue;
    if (env_alias == 1 && !alias) continue;
    const int stage_eff = alias ? (stage_b > in_b ? stage_b : in_b) : stage_b;
    const int obuf_b = stage_eff + out1_b + st_b;  // one output staging set
    const int group_bytes = bar_b + (alias ? 0 : in_b) + nbuf * obuf_b + scan_b;
    for (int gr = groups; gr >= 1; --gr) {
      const int warps = gr * T / 32;
      const int wacc_b = stats ? align_up(warps * S * 8, 16) : 0;
      const int total = wacc_b + cta_extra + gr * group_bytes;
      if (total > max_smem) continue;
      const int res = fb2_resident(total, T * gr, gr, regs);
      if (res > best_res) {
        best_res = res;
        a->off_wacc = 0;
        a->off_bins = wacc_b;  // distance-binned variant: the caller lays [bins | table | log s0] out inside cta_extra
        a->off_group0 = wacc_b + cta_extra;
        a->group_bytes = group_bytes;
        a->goff_bar = 0;
        a->goff_in = bar_b;
        a->goff_stage = alias ? bar_b : bar_b + in_b;
        a->goff_out1 = a->goff_stage + stage_eff;
        a->goff_st = a->goff_out1 + out1_b;
        a->goff_scan = a->goff_stage + nbuf * obuf_b;
        a->nbuf = nbuf;
        a->alias_in = alias ? 1 : 0;
        a->obuf_bytes = obuf_b;
        a->T = T;
        a->Tact = Tact;
        a->groups = gr;
        plan->block = T * gr;
        plan->smem = total;
        plan->chunk = CH;
      }
      break;  // the largest group count that fits is the candidate for this layout
    }
  }
  return best_res > 0 ? SMF_OK : SMF_ERR_TOO_LONG;
}

// chunk 17 whenever the read fits its thread budget; chunk 33 (K = 2, 4) for reads up to twice as long
static int plan_fb2(int K, int64_t L64, bool stats, int S, int es, bool want_gamma_one, bool want_states,
                    int max_smem, Fb2Args* a, Fb2Plan* plan, bool walk = false, bool gc_ok = false) {
  // Measured on B200 (profiles/r01_chunk_sweep.txt): for K = 2 posterior decode the longer chunk wins
  // whenever it keeps the lanes about as busy (half the scan / barrier overhead per position outweighs
  // the halved warp count); for K = 4 (register pressure) and for the E-step (more live registers)
  // chunk 17 is faster.  SMF_FB2_CHUNK overrides.
  static const int env_ch = [] { const char* e = getenv("SMF_FB2_CHUNK"); return e ? atoi(e) : 0; }();
  if (walk) {
    // two-kernel path: no chunk matrices, so the chunk length only trades lane utilisation and
    // boundary-vector traffic against registers
    const int first = (env_ch == 17 || env_ch == 33) ? env_ch : kWalkChunkDefault;
    const int second = first == 33 ? 17 : 33;
    if (plan_fb2_chunk(K, first, L64, stats, S, es, want_gamma_one, want_states, max_smem, a, plan, true) == SMF_OK)
      return SMF_OK;
    return plan_fb2_chunk(K, second, L64, stats, S, es, want_gamma_one, want_states, max_smem, a, plan, true);
  }
  // K = 3 with chunk 33 spills heavily; so does the K = 4 E-step (3 KB of stack: 12-16e9 read-positions/s at 9-12 kb
  // against 22-41e9 for the generic kernel and 44-65e9 for the two-kernel path, profiles/r02_k4_long_estep.txt)
  const bool can33 = (K != 3) && !(K == 4 && stats);
  // lane utilisation of the two chunk lengths (threads are allocated in whole warps)
  auto lane_eff = [&](int ch) {
    const long long tact = (L64 + ch - 1) / ch;
    const long long t = (tact + 31) / 32 * 32;
    return (double)L64 / (double)(t * ch);
  };
  // K = 2 posterior decode of float / uint8 observations has three more chunk lengths compiled (21, 25,
  // 29: fb2x_k2.cu), which smooths the lane-utilisation sawtooth of medium reads (e.g. 600 positions:
  // 19 + 13 idle lanes at chunk 17, one full warp at chunk 21).  Score = lane utilisation, mildly
  // favouring longer chunks (less scan work per position), as measured for 17 vs 33.
  if (K == 2 && !stats && (es == 4 || es == 1) &&
      (env_ch == 9 || env_ch == 11 || env_ch == 13 || env_ch == 15 || env_ch == 21 || env_ch == 25 || env_ch == 29) &&
      plan_fb2_chunk(K, env_ch, L64, stats, S, es, want_gamma_one, want_states, max_smem, a, plan, false, gc_ok) == SMF_OK)
    return SMF_OK;
  if (K == 3 && !stats && (es == 4 || es == 1) && (env_ch == 13 || env_ch == 21) &&
      plan_fb2_chunk(K, env_ch, L64, stats, S, es, want_gamma_one, want_states, max_smem, a, plan, false, gc_ok) == SMF_OK)
    return SMF_OK;
  if (K == 2 && !stats && env_ch == 0 && (es == 4 || es == 1)) {
    // (9..15 only matter for reads of one warp: with several warps the longer chunks always win)
    static const int cand[9] = {33, 29, 25, 21, 17, 15, 13, 11, 9};
    const int ncand = (L64 <= 32 * 17) ? 9 : 5;
    bool tried[9] = {false, false, false, false, false, false, false, false, false};
    for (int round = 0; round < ncand; ++round) {
      int bi = -1;
      double bs = -1.0;
      for (int i = 0; i < ncand; ++i) {
        if (tried[i]) continue;
        double sc = lane_eff(cand[i]) * (0.8 + 0.2 * (cand[i] - 17) / 16.0);
        // one resident read per SM around 9-10 kb: the extra warp of chunk 29 beats the longer chunk
        // (measured 0.59 / 0.66 vs 0.55 / 0.64 of HBM peak at 9 / 10 kb; the other way round at 11 kb)
        // (float rows only: with the uint8 code two reads stay resident and chunk 33 wins, 1.65 vs 2.73 ms at 10 kb)
        if (cand[i] == 29 && L64 >= 8500 && L64 <= 10500 && es == 4) sc *= 1.06;
        if (sc > bs) {
          bs = sc;
          bi = i;
        }
      }
      tried[bi] = true;
      if (plan_fb2_chunk(K, cand[bi], L64, stats, S, es, want_gamma_one, want_states, max_smem, a, plan, false, gc_ok) == SMF_OK)
        return SMF_OK;
    }
    return SMF_ERR_TOO_LONG;
  }
  // K = 3 posterior decode: chunks 13 and 21 (fb2x_k3.cu) fill the troughs of the same sawtooth; at equal
  // lane utilisation 17 is the sweet spot (measured weights 0.88 / 1 / 0.90 for 13 / 17 / 21).
  if (K == 3 && !stats && env_ch == 0 && (es == 4 || es == 1)) {
    static const int cand3[3] = {17, 21, 13};
    static const double w3[3] = {1.0, 0.90, 0.88};
    bool tried[3] = {false, false, false};
    for (int round = 0; round < 3; ++round) {
      int bi = -1;
      double bs = -1.0;
      for (int i = 0; i < 3; ++i) {
        if (tried[i]) continue;
        const double sc = lane_eff(cand3[i]) * w3[i];
        if (sc > bs) {
          bs = sc;
          bi = i;
        }
      }
      tried[bi] = true;
      if (plan_fb2_chunk(K, cand3[bi], L64, stats, S, es, want_gamma_one, want_states, max_smem, a, plan, false, gc_ok) == SMF_OK)
        return SMF_OK;
    }
    return SMF_ERR_TOO_LONG;
  }
  const bool prefer33 = can33 && (env_ch == 33 || (env_ch == 0 && K == 2 && !stats && lane_eff(33) >= 0.9 * lane_eff(17)));
  if (prefer33 &&
      plan_fb2_chunk(K, 33, L64, stats, S, es, want_gamma_one, want_states, max_smem, a, plan, false, gc_ok) == SMF_OK)
    return SMF_OK;
  if (plan_fb2_chunk(K, 17, L64, stats, S, es, want_gamma_one, want_states, max_smem, a, plan, false, gc_ok) == SMF_OK)
    return SMF_OK;
  if (can33 && plan_fb2_chunk(K, 33, L64, stats, S, es, want_gamma_one, want_states, max_smem, a, plan, false, gc_ok) == SMF_OK)
    return SMF_OK;
  return SMF_ERR_TOO_LONG;
}

// ---------------------------------------------------------------------------------------
// two-kernel path (K >= 3): walk_kernel (boundary vectors, log-likelihood) + fb2_kernel<WALK>
// ---------------------------------------------------------------------------------------
constexpr int64_t kWalkMinReads = 8192;  // below this the sequential walk has too little parallelism

// Measured on B200 (profiles/r01_walk.txt): the extra pass over the observations pays off where the
// single-kernel path is FMA/register bound -- K = 4 E-steps at every length (x1.2 at 500 positions,
// x3 at 8 kb where the single kernel has left its chunk-17 range), K = 4 posterior decode up to
// ~6 kb (x2 at 170 positions, x1.3 at 500, break-even near 8 kb), and K = 3 E-steps beyond the
// chunk-17 reach (x1.3 at 10 kb).  K = 3 elsewhere is faster in one kernel.  SMF_FB2_WALK=0/1 or
// smf_hmm_set_option("walk", 0/1) force it off/on.
static bool walk_eligible(const smf_hmm_tables* t, int64_t N, int64_t L, bool stats) {
  const int env_walk = g_walk_mode;
  if (env_walk == 0 || g_force_generic) return false;
  const int K = t->n_states;
  if (t->n_channels != 1 || t->n_bins != 1) return false;
  if (L < 2 || L > (1 << 24)) return false;
  if (rescale_cadence(t) != 4) return false;
  if (env_walk == 2) return !stats || K >= 3;  // any-length sweep kernels wherever they exist
  // reads that do not fit on chip: boundary walk + warp-independent sweeps (post_sweep_kernel), any K
  const bool off_chip = L > (int64_t)33 * fb2w_max_threads_c(K, false);
  // reads that do not fit on chip have only the segmented generic kernel as the alternative: the walk pays off from
  // ~3,000 reads (8,000 x 20 kb: 2.1 vs 3.2 ms for K = 3, 2.4 vs 5.8 ms for K = 4; 1,000 reads: 1.1 vs 0.4 ms)
  const int64_t post_reach = (int64_t)(K == 3 ? 21 : 33) * fb2_max_threads_c(K, false);  // longest single-kernel read
  if (!stats && (off_chip || L > post_reach) && N >= 3000) return true;
  if (K < 3) return false;
  if (env_walk == 1) return true;
  // E-step beyond the single kernel's chunk-17 reach (only the generic kernel is left there): the walk already pays
  // off from ~3,000 reads (K = 4, 9-16 kb, 4,000 reads: 44-46e9 read-positions/s vs 22-41e9; 1,000 reads: 17e9 vs 22-41e9)
  const bool beyond17 = L > (int64_t)17 * fb2_max_threads_c(K, true);
  if (N < ((stats && beyond17) ? 3000 : kWalkMinReads)) return false;
  if (stats) return K == 4 || beyond17;
  // posterior: K = 4 up to ~6 kb and beyond the single kernel's chunk-17 range (8.7 kb; 0.31 vs 0.22 of
  // HBM peak at 9 kb); K = 3 beyond the single kernel's reach (13.4 kb with chunk 21): 0.27 vs 0.10
  // (generic kernel) at 16 kb
  if (K == 4) return L <= 6000 || L > (int64_t)17 * fb2_max_threads_c(4, false);
  return L > (int64_t)21 * fb2_max_threads_c(3, false);
}
// boundary vectors for the shorter chunk (an upper bound for the longer one) + per-read log-likelihoods
static size_t walk_ws_bytes(int K, int64_t N, int64_t L) {
  const size_t tact = (size_t)((L + 16) / 17);
  return 2 * tact * (size_t)N * K * sizeof(float) + (size_t)N * sizeof(double);
}

static int launch_walk_path(bool stats, const smf_hmm_tables* t, int obs_dtype, Fb2Args& a, const Fb2Plan& plan,
                            const DeviceInfo& di, void* ws, int* nblocks, cudaStream_t stream) {
  const int K = t->n_states;
  WalkArgs w;
  w.obs = a.obs;
  w.N = a.N;
  w.L = a.L;
  w.CH = plan.chunk;
  w.Tact = a.Tact;
  const size_t nb = (size_t)a.Tact * (size_t)a.N * K;
  w.bnd_v = reinterpret_cast<float*>(ws);
  w.bnd_w = w.bnd_v + nb;
  double* ll = reinterpret_cast<double*>(reinterpret_cast<char*>(ws) + 2 * ((size_t)((a.L + 16) / 17)) * (size_t)a.N * K * sizeof(float));
  w.ll = stats ? ll : a.loglik;
  int rc = (K == 3) ? launch_walk_k3(obs_dtype, t, w, stream) : launch_walk_k4(obs_dtype, t, w, stream);
  if (rc != SMF_OK) return rc;
  a.bnd_v = w.bnd_v;
  a.bnd_w = w.bnd_w;
  a.ll_in = ll;
  const int out = (stats || obs_dtype != SMF_OBS_F32 || a.gamma == nullptr || a.states == nullptr)
                      ? 2
                      : (a.gamma_state < 0 ? 0 : 1);
  return (K == 3) ? launch_fb2w_k3(plan.chunk, stats, obs_dtype, out, t, a, plan, di, nblocks, stream)
                  : launch_fb2w_k4(plan.chunk, stats, obs_dtype, out, t, a, plan, di, nblocks, stream);
}

// ---------------------------------------------------------------------------------------
// sequential E-step (seq_kernel.cuh): one thread per read, three recursions per position
// ---------------------------------------------------------------------------------------
// -1 = measured policy, 0 = never, 1 = whenever the batch is eligible.  SMF_SEQ / smf_hmm_set_option("seq", v).
// Viterbi: -1 = measured policy, 1 = direct kernel wherever it applies, 0 = staged kernel only.  SMF_VIT / smf_hmm_set_option("vit", v).
static std::atomic<int> g_vit_mode{[] { const char* e = getenv("SMF_VIT"); return e ? atoi(e) : -1; }()};
// one-launch Baum-Welch loop of small grouped fits: 1 = whenever every work item is resident at once (default), 0 = never
static std::atomic<int> g_fit_loop_mode{[] { const char* e = getenv("SMF_FIT_LOOP"); return e ? atoi(e) : 1; }()};
static std::atomic<int> g_seq_mode{[] { const char* e = getenv("SMF_SEQ"); return e ? atoi(e) : -1; }()};
// Reads per launch from which the reads alone fill the machine (148 SMs x 4-5 CTAs x 128 threads); measured
// break-even against the chunk-parallel kernels in profiles/r02_seq_estep.txt.
// (profiles/r02_seq_estep.txt, r02_policy_audit.txt: the chunk-parallel kernels lose ground with the read length, so
// long reads switch earlier -- 30,000 x 12 kb: 1.3-1.6x faster for every K)
static int64_t seq_min_reads(int K, int64_t L) {
  static const int env = [] { const char* e = getenv("SMF_SEQ_MIN_READS"); return e ? atoi(e) : 0; }();
  if (env > 0) return env;
  if (L >= 16384 && K >= 3) return 16384;  // 20,000 x 20 kb: K = 3 3.7-4.1 vs 4.5-4.9 ms, K = 4 4.3-4.4 vs 5.0 (r02_policy_audit3.txt)
  if (L >= 8192) return 24576;
  if (K == 2) return L >= 2000 ? 65536 : 98304;
  return 49152;
}
// shape-only test (used for workspace sizing); the launch additionally needs an eligible dtype / alignment
static bool seq_shape_eligible(const smf_hmm_tables* t, int64_t N, int64_t L) {
  const int mode = g_seq_mode;
  if (mode == 0 || g_force_generic) return false;
  if (t->n_channels != 1 || t->n_bins != 1) return false;
  if (L < 2 || L > (1 << 24)) return false;
  if (rescale_cadence(t) != 4) return false;
  if (mode == 1) return true;
  return N >= seq_min_reads(t->n_states, L);
}
// rows on 16-byte boundaries whose 16-element loads stay inside the buffer: vector loads
static bool seq_rows_aligned(int64_t L, int obs_dtype, const void* obs) {
  if (reinterpret_cast<uintptr_t>(obs) & 15u) return false;
  return obs_dtype == SMF_OBS_U8 ? (L & 15) == 0 : (L & 3) == 0;
}
// E-step: any float / uint8-coded rows (unaligned rows are read element-wise); posterior decode additionally needs
// output rows on 16-byte boundaries (need_aligned)
static bool seq_launch_eligible(const smf_hmm_tables* t, int64_t N, int64_t L, int obs_dtype, const void* obs,
                                bool need_aligned) {
  if (!seq_shape_eligible(t, N, L)) return false;
  if (obs_dtype != SMF_OBS_U8 && obs_dtype != SMF_OBS_F32) return false;
  if (need_aligned) return seq_rows_aligned(L, obs_dtype, obs) && (L & 3) == 0;
  return true;
}
// posterior decode on the sequential kernels: K >= 3 by default (K = 2 has the faster chunk-parallel kernel;
// "seq" = 1 forces it for every eligible batch)
static bool seq_post_shape_eligible(const smf_hmm_tables* t, int64_t N, int64_t L) {
  if (!seq_shape_eligible(t, N, L) || (L & 3)) return false;
  if (g_seq_mode == 1) return true;
  // (the E-step switches earlier for the longest reads; posterior decode of 20,000 x 20 kb measured slower here:
  // K = 4 6.6 vs 5.2 ms, profiles/r02_policy_audit4.txt)
  if (L >= 8192 && N < 24576) return false;
  // measured (profiles/r02_seq_post.txt, r02_final_sweep.txt; ~2.4 ms per 4e8 positions for K = 3 at any length):
  // K = 4 wins at every length (0.31-0.44 of HBM peak vs 0.26-0.40); K = 3 where the chunk-parallel kernels are
  // weak -- short reads (<= 400: 0.31-0.41) and reads beyond the single kernel's reach (> 13.4 kb: 0.27); K = 2 only
  // beyond the fast path's reach (> 21 kb: 0.21-0.25)
  if (t->n_states == 4) return true;
  if (t->n_states == 3) return L <= 400 || L > 13440;
  return L > 21120;
}
static size_t seq_ws_bytes(int K, int64_t N, int64_t L) {
  const size_t tact = (size_t)((L + kSeqCH - 1) / kSeqCH);
  return tact * (size_t)N * SeqVec<4>::KP * sizeof(float) + (size_t)N * sizeof(double) + 64;
}

static inline int obs_elem_size(int dt) { return dt == SMF_OBS_F32 ? 4 : (dt == SMF_OBS_F64 ? 8 : 1); }

// distance-binned models on the fast-path kernel (fb2_kernel<..., BINNED>, posterior decode): K <= 3, float /
// uint8-coded rows.  SMF_FB2_BINNED=0 keeps them on the generic kernel.
static bool fb2_binned_eligible(const smf_hmm_tables* t, int obs_dtype) {
  static const int env = [] { const char* e = getenv("SMF_FB2_BINNED"); return e ? atoi(e) : 1; }();
  return env != 0 && t->n_bins > 1 && t->n_channels == 1 && t->n_states <= 3 &&
         (obs_dtype == SMF_OBS_F32 || obs_dtype == SMF_OBS_U8);
}
// multi-channel models on the fast-path kernel (fb2_kernel<..., C>): K <= 3, C = 2 or 3, float / uint8-coded rows.
// SMF_FB2_MULTI=0 keeps them on the generic kernel.
static bool fb2_multi_eligible(const smf_hmm_tables* t, int obs_dtype) {
  static const int env = [] { const char* e = getenv("SMF_FB2_MULTI"); return e ? atoi(e) : 1; }();
  return env != 0 && t->n_bins == 1 && t->n_states <= 3 && (t->n_channels == 2 || t->n_channels == 3) &&
         (obs_dtype == SMF_OBS_F32 || obs_dtype == SMF_OBS_U8);
}

// ---------------------------------------------------------------------------------------
// short reads (short_kernel): several reads per warp
// ---------------------------------------------------------------------------------------
// -1 = measured policy, 0 = never, 1 = whenever it fits.  SMF_SHORT / smf_hmm_set_option("short", v).
static std::atomic<int> g_short_mode{[] { const char* e = getenv("SMF_SHORT"); return e ? atoi(e) : -1; }()};
// measured (profiles/r01_short.txt): faster than one read per warp up to ~275 positions for K = 2
// (beyond that fb2_kernel's short chunks 9..15 keep a warp's lanes busy),
// ~400 for K = 3, ~200 for K = 4
// (posterior); the E-step variant wins up to ~450 (K = 2, 3) / ~300 (K = 4)
static int short_max_len(int K, bool stats = false) {
  if (stats) return K >= 4 ? 300 : 450;
  return K >= 4 ? 200 : (K == 3 ? 400 : 275);
}

// Picks the odd chunk length that keeps the most lanes busy (weighted by how well it amortises the
// scan; one or more reads per warp) and lays out the warp's shared-memory region.
static int plan_short(int K, int64_t L64, int obs_dtype, bool gamma_all, bool gamma_one, bool want_states,
                      int max_smem, ShortArgs* a, bool stats = false) {
  const int es = obs_elem_size(obs_dtype);
  if (L64 < 1 || L64 > 32 * 33) return SMF_ERR_TOO_LONG;
  const int L = (int)L64;
  double best = -1.0;
  int bch = 0;
  static const int env_ch = [] { const char* e = getenv("SMF_SHORT_CHUNK"); return e ? atoi(e) : 0; }();
  for (int ch = 5; ch <= 33; ch += 2) {
    if (env_ch && ch != env_ch) continue;
    const int tg = (L + ch - 1) / ch;
    if (tg > 32) continue;
    if (stats && !short_stats_chunk_compiled(K, obs_dtype, ch)) continue;
    const int r = 32 / tg;
    const double eff = (double)(r * L) / (32.0 * ch);
    // shared memory per warp decides how many warps hide each other's recurrence latency
    const int kout_c = gamma_all ? K : (gamma_one ? 1 : 0);
    const double wb = 16.0 + 2.0 * (r * L * es + 32) + 32.0 * ch * K * 4 + r * L * kout_c * 4.0 + (want_states ? r * L : 0) + 64;
    const double warps = kShortWarps * floor(227.0 * 1024.0 / (kShortWarps * wb + 1024.0));  // whole CTAs
    if (kShortWarps * wb > max_smem) continue;
    double score = eff * ch / (ch + 6.0) * (warps >= 16.0 ? 1.0 : warps / 16.0);
    // fully unrolled instantiations run ~2.5x fewer instructions per position than the rolled variant
    if (short_chunk_compiled(K, obs_dtype, ch)) score *= 2.5;
    if (score > best) {
      best = score;
      bch = ch;
    }
  }
  if (bch == 0) return SMF_ERR_TOO_LONG;
  const int Tg = (L + bch - 1) / bch, R = 32 / Tg;
  const int kout = gamma_all ? K : (gamma_one ? 1 : 0);
  a->L = L;
  a->CH = bch;
  a->Tg = Tg;
  a->R = R;
  a->in_bytes = align_up(R * L * es + 32, 16);
  a->stack_off = 16 + 2 * a->in_bytes;
  a->gam_off = a->stack_off + align_up(32 * bch * K * 4, 16);
  a->st_off = a->gam_off + (kout ? align_up(R * L * kout * 4 + 16, 16) : 0);
  a->warp_bytes = a->st_off + (want_states ? align_up(R * L + 32, 16) : 0);
  if (kShortWarps * a->warp_bytes > max_smem) return SMF_ERR_TOO_LONG;
  return SMF_OK;
}

static int launch_short(const smf_hmm_tables* t, int obs_dtype, const ShortArgs& a, const DeviceInfo& di,
                        cudaStream_t s) {
  switch (t->n_states) {
    case 2: return launch_short_k2(obs_dtype, t, a, di, s);
    case 3: return launch_short_k3(obs_dtype, t, a, di, s);
    case 4: return launch_short_k4(obs_dtype, t, a, di, s);
  }
  return SMF_ERR_BAD_ARG;
}

static int launch_fb2(bool stats, const smf_hmm_tables* t, int obs_dtype, Fb2Args& a, const Fb2Plan& plan,
                      const DeviceInfo& di, int* nblocks, cudaStream_t stream) {
  // output-set specialisations exist for float and uint8-coded observations; everything else decides at run time
  int out = (stats || (obs_dtype != SMF_OBS_F32 && obs_dtype != SMF_OBS_U8) || a.gamma == nullptr || a.states == nullptr ||
             a.loglik != nullptr)
                ? 2
                : (a.gamma_state < 0 ? 0 : 1);
  if (a.gc && out == 0 && t->n_states == 2 && (plan.chunk == 33 || plan.chunk == 29)) out = 3;
  if (a.gc && out != 3) return SMF_ERR_BAD_ARG;  // the staging tile was sized for the compact layout
  static const int env_minb = [] { const char* e = getenv("SMF_FB2_MINB"); return e ? atoi(e) : 0; }();
  if (t->n_states == 2 && (plan.chunk == 21 || plan.chunk == 25 || plan.chunk == 29))
    return launch_fb2x_k2(plan.chunk, obs_dtype, out, t, a, plan, di, nblocks, stream);
  if (t->n_states == 2 && plan.chunk < 17)
    return launch_fb2y_k2(plan.chunk, obs_dtype, out, t, a, plan, di, nblocks, stream);
  if (t->n_states == 3 && (plan.chunk == 13 || plan.chunk == 21))
    return launch_fb2x_k3(plan.chunk, obs_dtype, out, t, a, plan, di, nblocks, stream);
  switch (t->n_states) {
    case 2: return launch_fb2_k2(plan.chunk, stats, obs_dtype, out, env_minb, t, a, plan, di, nblocks, stream);
    case 3: return launch_fb2_k3(plan.chunk, stats, obs_dtype, out, env_minb, t, a, plan, di, nblocks, stream);
    case 4: return launch_fb2_k4(plan.chunk, stats, obs_dtype, out, env_minb, t, a, plan, di, nblocks, stream);
  }
  return SMF_ERR_BAD_ARG;
}

static int launch_fb(bool stats, const smf_hmm_tables* t, FbArgs& a, const FbPlan& plan, const DeviceInfo& di,
                     int* nblocks, cudaStream_t stream) {
  switch (t->n_states) {
    case 2: return launch_fb_k2(stats, t, a, plan, di, nblocks, stream);
    case 3: return launch_fb_k3(stats, t, a, plan, di, nblocks, stream);
    case 4: return launch_fb_k4(stats, t, a, plan, di, nblocks, stream);
  }
  return SMF_ERR_BAD_ARG;
}


// The fast path needs natural alignment of the row pointers it vectorises.
static bool fb2_aligned(const void* obs, int obs_dtype, const float* gamma, int K, int gamma_state) {
  if (reinterpret_cast<uintptr_t>(obs) % obs_elem_size(obs_dtype)) return false;
  if (gamma != nullptr) {
    const uintptr_t g = reinterpret_cast<uintptr_t>(gamma);
    if (gamma_state < 0 && K == 2 && (g & 7u)) return false;
    if (gamma_state < 0 && K == 4 && (g & 15u)) return false;
    if (g & 3u) return false;
  }
  return true;
}

// Rescale cadence the scaled fp32 recursions need for this model.  Between two power-of-two
// rescales an entry that still matters may shrink by f^R (f = worst per-step emission contrast
// min_k b_k / max_k b_k over o in {0,1}) on top of the g^2 contrast a state switch at both chunk
// ends costs (g = smallest / largest transition weight in a row); the product has to stay inside
// the fp32 normal range with margin.  Typical fitted models need R = 4; models with emissions or
// transitions at the eps clamp need R = 2 or 1 and are routed to the generic kernel.
static int rescale_cadence(const smf_hmm_tables* t) {
  const int K = t->n_states, C = t->n_channels;
  double logf = 0.0;  // log of the worst emission contrast, summed over channels
  for (int c = 0; c < C; ++c) {
    double worst = 0.0;
    for (int o = 0; o < 2; ++o) {
      double lo = 1e300, hi = -1e300;
      for (int k = 0; k < K; ++k) {
        const double v = o ? t->log_emit1[k * C + c] : t->log_emit0[k * C + c];
        lo = v < lo ? v : lo;
        hi = v > hi ? v : hi;
      }
      if (lo - hi < worst) worst = lo - hi;
    }
    logf += worst;
  }
  double logg = 0.0;
  for (int b = 0; b < t->n_bins; ++b)
    for (int i = 0; i < K; ++i) {
      double lo = 1e300, hi = -1e300;
      for (int j = 0; j < K; ++j) {
        const double v = t->log_trans[(b * K + i) * K + j];
        lo = v < lo ? v : lo;
        hi = v > hi ? v : hi;
      }
      if (lo - hi < logg) logg = lo - hi;
    }
  const double budget = -64.0;  // ln(1e-28): stay well inside fp32's normal range
  for (int R = 4; R >= 2; R >>= 1)
    if (2.0 * logg + (R + 1) * logf >= budget) return R;
  return 1;
}

static int stats_len(const smf_hmm_tables* t) {
  const int K = t->n_states, C = t->n_channels;
  return K + t->n_bins * K * K + 2 * K * C + 1;
}

// Integer form of the size classes for integer lengths: lo <= len < hi  <=>  ceil(lo) <= len < ceil(hi).
static void to_int_classes(const ClassTable& ct, IntClassTable* it) {
  memset(it, 0, sizeof(*it));
  it->n = ct.n;
  auto cl = [](double x) -> long long {
    if (x != x) return LLONG_MAX;  // NaN bound: never matches (lo) / never bounds (hi) like the comparison
    if (x >= 9.2e18) return LLONG_MAX;
    if (x <= -9.2e18) return LLONG_MIN;
    return (long long)ceil(x);
  };
  for (int i = 0; i < ct.n; ++i) {
    it->lo[i] = cl(ct.lo[i]);
    it->hi[i] = cl(ct.hi[i]);
    if (ct.lo[i] != ct.lo[i] || ct.hi[i] != ct.hi[i]) {  // NaN bounds compare false: class never matches
      it->lo[i] = LLONG_MAX;
      it->hi[i] = LLONG_MIN;
    }
  }
}

static int fill_class_table(const double* lo, const double* hi, int n, ClassTable* ct) {
  if (n < 0 || n > kMaxClasses) return SMF_ERR_BAD_ARG;
  if (n > 0 && (!lo || !hi)) return SMF_ERR_BAD_ARG;
  memset(ct, 0, sizeof(*ct));
  for (int i = 0; i < n; ++i) {
    ct->lo[i] = lo[i];
    ct->hi[i] = hi[i];
  }
  ct->n = n;
  return SMF_OK;
}

}  // namespace smf

using namespace smf;

extern "C" {

int smf_hmm_abi_version(void) { return SMF_HMM_ABI_VERSION; }

const char* smf_hmm_strerror(int code) {
  switch (code) {
    case SMF_OK: return "ok";
    case SMF_ERR_BAD_ARG: return "bad argument (null pointer, negative size, or unsupported K/C/bins/dtype)";
    case SMF_ERR_TOO_LONG: return "read too long for the on-chip staging of this kernel";
    case SMF_ERR_WORKSPACE: return "workspace too small";
    case SMF_ERR_CUDA: return "CUDA runtime error";
    case SMF_ERR_ALIGN: return "pointer alignment";
  }
  return "unknown error";
}

int smf_hmm_set_option(const char* key, int value) {
  if (key == nullptr) return SMF_ERR_BAD_ARG;
  if (strcmp(key, "walk") == 0) {
    if (value < -1 || value > 2) return SMF_ERR_BAD_ARG;
    g_walk_mode = value;
    return SMF_OK;
  }
  if (strcmp(key, "short") == 0) {
    if (value < -1 || value > 1) return SMF_ERR_BAD_ARG;
    g_short_mode = value;
    return SMF_OK;
  }
  if (strcmp(key, "generic") == 0) {
    if (value < 0 || value > 1) return SMF_ERR_BAD_ARG;
    g_force_generic = value != 0;
    return SMF_OK;
  }
  if (strcmp(key, "vit") == 0) {
    if (value < -1 || value > 1) return SMF_ERR_BAD_ARG;
    g_vit_mode = value;
    return SMF_OK;
  }
  if (strcmp(key, "seq") == 0) {
    if (value < -1 || value > 1) return SMF_ERR_BAD_ARG;
    g_seq_mode = value;
    return SMF_OK;
  }
  if (strcmp(key, "fitloop") == 0) {
    if (value < 0 || value > 1) return SMF_ERR_BAD_ARG;
    g_fit_loop_mode = value;
    return SMF_OK;
  }
  return SMF_ERR_BAD_ARG;
}

int64_t smf_hmm_stats_len(const smf_hmm_tables* tab) {
  if (check_tables(tab) != SMF_OK) return -1;
  return stats_len(tab);
}

size_t smf_hmm_workspace_bytes(int op, const smf_hmm_tables* tab, int64_t n_reads, int64_t n_pos) {
  if (n_reads < 0 || n_pos < 0) return 0;
  switch (op) {
    case SMF_OP_VITERBI: return (size_t)n_reads * (size_t)n_pos;
    case SMF_OP_ESTEP: {
      if (check_tables(tab) != SMF_OK) return 0;
      size_t b = (size_t)kMaxEstepBlocks * (size_t)stats_len(tab) * sizeof(double);
      size_t extra = 0;
      if (walk_eligible(tab, n_reads, n_pos, true)) extra = walk_ws_bytes(tab->n_states, n_reads, n_pos);
      if (seq_shape_eligible(tab, n_reads, n_pos)) {
        const size_t sq = seq_ws_bytes(tab->n_states, n_reads, n_pos);
        if (sq > extra) extra = sq;
      }
      return b + extra;
    }
    case SMF_OP_POSTERIOR: {
      if (check_tables(tab) != SMF_OK) return 0;
      size_t b = walk_eligible(tab, n_reads, n_pos, false) ? walk_ws_bytes(tab->n_states, n_reads, n_pos) : 0;
      if (seq_post_shape_eligible(tab, n_reads, n_pos)) {
        const size_t sq = seq_ws_bytes(tab->n_states, n_reads, n_pos);
        if (sq > b) b = sq;
      }
      return b;
    }
    default: return 0;
  }
}

int64_t smf_hmm_max_positions(const smf_hmm_tables* tab, int want_all_states) {
  (void)want_all_states;
  if (check_tables(tab) != SMF_OK) return -1;
  DeviceInfo di;
  int max_smem = 232448;
  if (get_device_info(&di) == SMF_OK) max_smem = di.max_smem_optin;
  int64_t lo = 1, hi = 1 << 20;
  FbArgs a;
  FbPlan p;
  while (lo < hi) {
    const int64_t mid = (lo + hi + 1) / 2;
    if (plan_fb(tab->n_states, tab->n_channels, tab->n_bins, mid, false, stats_len(tab), max_smem, &a, &p) ==
        SMF_OK)
      lo = mid;
    else
      hi = mid - 1;
  }
  return lo;
}

int smf_hmm_posterior(const smf_hmm_tables* tab, const void* obs, int obs_dtype, int64_t n_reads,
                      int64_t n_pos, const int8_t* bins, float* gamma, int gamma_state,
                      int8_t* states, double* loglik, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_tables(tab);
  if (rc != SMF_OK) return rc;
  if (n_reads < 0 || n_pos < 0) return SMF_ERR_BAD_ARG;
  if (n_reads == 0 || n_pos == 0) return SMF_OK;
  if (obs == nullptr) return SMF_ERR_BAD_ARG;
  if (obs_dtype < SMF_OBS_F32 || obs_dtype > SMF_OBS_U8) return SMF_ERR_BAD_ARG;
  if (tab->n_bins > 1 && n_pos > 1 && bins == nullptr) return SMF_ERR_BAD_ARG;
  if (gamma_state >= tab->n_states) return SMF_ERR_BAD_ARG;
  DeviceInfo di;
  rc = get_device_info(&di);
  if (rc != SMF_OK) return rc;
  const int cadence = rescale_cadence(tab);
  if (tab->n_channels == 1 && tab->n_bins == 1 && !g_force_generic && cadence == 4 &&
      fb2_aligned(obs, obs_dtype, gamma, tab->n_states, gamma_state)) {
    // large K = 4 batches: the sequential kernels also beat the bundled short-read kernel (10^6 x 100: 0.90 vs 1.20 ms)
    const bool seq_first = tab->n_states == 4 && g_short_mode != 1 && workspace != nullptr &&
                           seq_post_shape_eligible(tab, n_reads, n_pos) &&
                           seq_launch_eligible(tab, n_reads, n_pos, obs_dtype, obs, true) &&
                           workspace_bytes >= seq_ws_bytes(tab->n_states, n_reads, n_pos);
    // short reads: several reads per warp
    if (!seq_first && g_short_mode != 0 && (g_short_mode == 1 || n_pos <= short_max_len(tab->n_states))) {
      ShortArgs sa;
      memset(&sa, 0, sizeof(sa));
      if (plan_short(tab->n_states, n_pos, obs_dtype, gamma != nullptr && gamma_state < 0,
                     gamma != nullptr && gamma_state >= 0, states != nullptr, di.max_smem_optin, &sa) == SMF_OK) {
        sa.obs = obs;
        sa.N = n_reads;
        sa.gamma = gamma;
        sa.gamma_state = gamma_state;
        sa.states = states;
        sa.loglik = loglik;
        return launch_short(tab, obs_dtype, sa, di, (cudaStream_t)stream);
      }
    }
    // large batches of K >= 3: one thread per read (checkpointed forward pass, one backward pass, posteriors
    // staged in shared memory and written as whole sectors) when the caller provided the checkpoint workspace
    if (workspace != nullptr && seq_post_shape_eligible(tab, n_reads, n_pos) &&
        seq_launch_eligible(tab, n_reads, n_pos, obs_dtype, obs, true) && (reinterpret_cast<uintptr_t>(workspace) & 15u) == 0 &&
        (reinterpret_cast<uintptr_t>(states) & 3u) == 0 && (reinterpret_cast<uintptr_t>(gamma) & 15u) == 0) {
      if (workspace_bytes < seq_ws_bytes(tab->n_states, n_reads, n_pos)) return SMF_ERR_WORKSPACE;
      char* wws = reinterpret_cast<char*>(workspace);
      SeqArgs q;
      memset(&q, 0, sizeof(q));
      q.g.obs = obs;
      q.g.N = n_reads;
      q.g.L = (int)n_pos;
      q.g.row_stride = (int)n_pos;
      q.g.Tact = (int)((n_pos + kSeqCH - 1) / kSeqCH);
      q.g.ll = loglik != nullptr ? loglik : reinterpret_cast<double*>(wws);
      q.g.ckpt = reinterpret_cast<float*>(wws + (((size_t)n_reads * sizeof(double) + 63) & ~(size_t)63));
      q.g.gamma = gamma;
      q.g.states = states;
      q.g.gamma_state = gamma_state;
      q.g.out_stride = (int)n_pos;
      q.n_items = (n_reads + kSeqThreads - 1) / kSeqThreads;
      q.flush_chunks = 8;
      const int mode = (gamma != nullptr || states != nullptr) ? MODE_POST : -1;
      cudaStream_t st = (cudaStream_t)stream;
      return (tab->n_states == 2)   ? launch_seq_k2(obs_dtype, mode, tab, q, di, nullptr, st)
             : (tab->n_states == 3) ? launch_seq_k3(obs_dtype, mode, tab, q, di, nullptr, st)
                                    : launch_seq_k4(obs_dtype, mode, tab, q, di, nullptr, st);
    }
    Fb2Args b;
    memset(&b, 0, sizeof(b));
    Fb2Plan p2;
    // two-kernel path when the caller provided the boundary-vector workspace
    if (workspace != nullptr && walk_eligible(tab, n_reads, n_pos, false)) {
      if (workspace_bytes < walk_ws_bytes(tab->n_states, n_reads, n_pos)) return SMF_ERR_WORKSPACE;
      const bool on_chip = tab->n_states >= 3 && g_walk_mode != 2 &&
                           plan_fb2(tab->n_states, n_pos, false, stats_len(tab), obs_elem_size(obs_dtype),
                                    gamma != nullptr && gamma_state >= 0, states != nullptr, di.max_smem_optin, &b,
                                    &p2, true) == SMF_OK;
      // K = 4 beyond the chunk-17 range would run the spilling chunk-33 variant on chip: sweep instead
      const bool prefer_sweep = tab->n_states == 4 && n_pos > (int64_t)17 * fb2w_max_threads_c(4, false);
      if (!on_chip || prefer_sweep) {
        // any read length: boundary walk + warp-independent posterior sweeps
        const int K = tab->n_states;
        const int CH = 17;
        const size_t cap = (size_t)((n_pos + 16) / 17) * (size_t)n_reads * K;
        char* wws = reinterpret_cast<char*>(workspace);
        WalkArgs w;
        w.obs = obs;
        w.N = n_reads;
        w.L = (int)n_pos;
        w.CH = CH;
        w.Tact = (int)((n_pos + CH - 1) / CH);
        w.bnd_v = reinterpret_cast<float*>(wws);
        w.bnd_w = w.bnd_v + cap;
        w.ll = loglik;
        cudaStream_t st = (cudaStream_t)stream;
        rc = (K == 2) ? launch_walk_k2(obs_dtype, tab, w, st)
                      : (K == 3 ? launch_walk_k3(obs_dtype, tab, w, st) : launch_walk_k4(obs_dtype, tab, w, st));
        if (rc != SMF_OK) return rc;
        if (gamma == nullptr && states == nullptr) return SMF_OK;  // log-likelihood only
        PostSweepArgs ps;
        memset(&ps, 0, sizeof(ps));
        ps.obs = obs;
        ps.N = n_reads;
        ps.L = (int)n_pos;
        ps.CH = CH;
        ps.Tact = w.Tact;
        ps.slices = (w.Tact + 31) / 32;
        ps.bnd_v = w.bnd_v;
        ps.bnd_w = w.bnd_w;
        ps.gamma = gamma;
        ps.gamma_state = gamma_state;
        ps.states = states;
        return (K == 2) ? launch_post_sweep_k2(obs_dtype, tab, ps, di, st)
                        : (K == 3 ? launch_post_sweep_k3(obs_dtype, tab, ps, di, st)
                                  : launch_post_sweep_k4(obs_dtype, tab, ps, di, st));
      }
      {
        b.obs = obs;
        b.N = n_reads;
        b.L = (int)n_pos;
        b.gamma = gamma;
        b.gamma_state = gamma_state;
        b.states = states;
        b.loglik = loglik;
        b.S = stats_len(tab);
        b.flush_every = 1;
        return launch_walk_path(false, tab, obs_dtype, b, p2, di, workspace, nullptr, (cudaStream_t)stream);
      }
      memset(&b, 0, sizeof(b));
    }
    const bool gc_ok = gamma != nullptr && gamma_state < 0 && states != nullptr && loglik == nullptr &&
                       obs_dtype == SMF_OBS_F32;
    const int prc = plan_fb2(tab->n_states, n_pos, false, stats_len(tab), obs_elem_size(obs_dtype),
                             gamma != nullptr && gamma_state >= 0, states != nullptr, di.max_smem_optin, &b, &p2,
                             false, gc_ok);
    if (prc == SMF_OK) {
      b.obs = obs;
      b.N = n_reads;
      b.L = (int)n_pos;
      b.gamma = gamma;
      b.gamma_state = gamma_state;
      b.states = states;
      b.loglik = loglik;
      b.partials = nullptr;
      b.S = stats_len(tab);
      b.flush_every = 1;
      return launch_fb2(false, tab, obs_dtype, b, p2, di, nullptr, (cudaStream_t)stream);
    }
  }
  if (fb2_binned_eligible(tab, obs_dtype) && !g_force_generic && cadence == 4 && n_pos > 1 &&
      fb2_aligned(obs, obs_dtype, gamma, tab->n_states, gamma_state)) {
    Fb2Args b;
    memset(&b, 0, sizeof(b));
    Fb2Plan p2;
    const int lpad = (int)((n_pos + 16) / 17 * 17);
    const int bins_b = align_up(lpad, 16);
    const int btab_b = tab->n_bins * (tab->n_states == 2 ? 4 : 8) * 4;
    const int ls0_b = tab->n_bins * 8;
    if (plan_fb2_chunk(tab->n_states, 17, n_pos, false, stats_len(tab), obs_elem_size(obs_dtype),
                       gamma != nullptr && gamma_state >= 0, states != nullptr, di.max_smem_optin, &b, &p2, false, false,
                       bins_b + align_up(btab_b, 16) + align_up(ls0_b, 16)) == SMF_OK) {
      b.off_btab = b.off_bins + bins_b;
      b.off_ls0 = b.off_btab + align_up(btab_b, 16);
      b.bins = bins;
      b.n_bins = tab->n_bins;
      b.obs = obs;
      b.N = n_reads;
      b.L = (int)n_pos;
      b.gamma = gamma;
      b.gamma_state = gamma_state;
      b.states = states;
      b.loglik = loglik;
      b.partials = nullptr;
      b.S = stats_len(tab);
      b.flush_every = 1;
      return tab->n_states == 2 ? launch_fb2b_k2(obs_dtype, tab, b, p2, di, (cudaStream_t)stream)
                                : launch_fb2b_k3(obs_dtype, tab, b, p2, di, (cudaStream_t)stream);
    }
  }
  if (fb2_multi_eligible(tab, obs_dtype) && !g_force_generic && cadence == 4 &&
      fb2_aligned(obs, obs_dtype, gamma, tab->n_states, gamma_state)) {
    Fb2Args b;
    memset(&b, 0, sizeof(b));
    Fb2Plan p2;
    if (plan_fb2_chunk(tab->n_states, 17, n_pos, false, stats_len(tab), obs_elem_size(obs_dtype) * tab->n_channels,
                       gamma != nullptr && gamma_state >= 0, states != nullptr, di.max_smem_optin, &b, &p2) == SMF_OK) {
      b.obs = obs;
      b.N = n_reads;
      b.L = (int)n_pos;
      b.gamma = gamma;
      b.gamma_state = gamma_state;
      b.states = states;
      b.loglik = loglik;
      b.partials = nullptr;
      b.S = stats_len(tab);
      b.flush_every = 1;
      return tab->n_states == 2
                 ? launch_fb2m_k2(tab->n_channels, false, obs_dtype, tab, b, p2, di, nullptr, (cudaStream_t)stream)
                 : launch_fb2m_k3(tab->n_channels, false, obs_dtype, tab, b, p2, di, nullptr, (cudaStream_t)stream);
    }
  }
  FbArgs a;
  memset(&a, 0, sizeof(a));
  FbPlan plan;
  rc = plan_fb(tab->n_states, tab->n_channels, tab->n_bins, n_pos, false, stats_len(tab),
               di.max_smem_optin, &a, &plan);
  if (rc != SMF_OK) return rc;
  a.obs = obs;
  a.obs_dtype = obs_dtype;
  a.N = n_reads;
  a.L = (int)n_pos;
  a.C = tab->n_channels;
  a.bins = bins;
  a.gamma = gamma;
  a.gamma_state = gamma_state;
  a.states = states;
  a.loglik = loglik;
  a.partials = nullptr;
  a.S = stats_len(tab);
  a.flush_every = 1;
  a.resc_mask = cadence - 1;
  return launch_fb(false, tab, a, plan, di, nullptr, (cudaStream_t)stream);
}

int smf_hmm_estep(const smf_hmm_tables* tab, const void* obs, int obs_dtype, int64_t n_reads,
                  int64_t n_pos, const int8_t* bins, double* stats, void* workspace,
                  size_t workspace_bytes, void* stream) {
  int rc = check_tables(tab);
  if (rc != SMF_OK) return rc;
  if (n_reads < 0 || n_pos < 0 || stats == nullptr) return SMF_ERR_BAD_ARG;
  const int S = stats_len(tab);
  cudaStream_t st = (cudaStream_t)stream;
  if (n_reads == 0 || n_pos == 0) {
    return cudaMemsetAsync(stats, 0, S * sizeof(double), st) == cudaSuccess ? SMF_OK : SMF_ERR_CUDA;
  }
  if (obs == nullptr || workspace == nullptr) return SMF_ERR_BAD_ARG;
  if (obs_dtype < SMF_OBS_F32 || obs_dtype > SMF_OBS_U8) return SMF_ERR_BAD_ARG;
  if (tab->n_bins > 1 && n_pos > 1 && bins == nullptr) return SMF_ERR_BAD_ARG;
  if (workspace_bytes < smf_hmm_workspace_bytes(SMF_OP_ESTEP, tab, n_reads, n_pos))
    return SMF_ERR_WORKSPACE;
  DeviceInfo di;
  rc = get_device_info(&di);
  if (rc != SMF_OK) return rc;
  const int cadence = rescale_cadence(tab);
  if (tab->n_channels == 1 && tab->n_bins == 1 && !g_force_generic && cadence == 4 &&
      fb2_aligned(obs, obs_dtype, nullptr, tab->n_states, -1)) {
    // large batches of K >= 3: the sequential kernels also beat the bundled short-read kernel (K = 4, 10^6 x 304:
    // 1.8e11 vs 7e10 read-positions/s; K = 3: 2.0e11 vs 1.4e11); K = 2 short reads stay bundled (3.0e11 either way)
    // (uint8-coded K = 2 too since the packed fp32x2 rewrite: 10^6 x 100: 0.39 vs 0.52 ms, profiles/r02_policy_audit3.txt)
    const bool seq_first = (tab->n_states >= 3 || obs_dtype == SMF_OBS_U8) && g_short_mode != 1 &&
                           seq_launch_eligible(tab, n_reads, n_pos, obs_dtype, obs, false);
    // short reads: several reads per warp (E-step variant of short_kernel)
    if (!seq_first && g_short_mode != 0 && obs_dtype != SMF_OBS_F64 &&
        (g_short_mode == 1 || n_pos <= short_max_len(tab->n_states, true))) {
      ShortArgs sa;
      memset(&sa, 0, sizeof(sa));
      if (plan_short(tab->n_states, n_pos, obs_dtype, false, false, false, di.max_smem_optin - align_up(kShortWarps * S * 8, 16),
                     &sa, true) == SMF_OK) {
        sa.obs = obs;
        sa.N = n_reads;
        sa.gamma_state = -1;
        sa.partials = reinterpret_cast<double*>(workspace);
        sa.S = S;
        sa.flush_every = 8;
        sa.wacc_bytes = align_up(kShortWarps * S * 8, 16);
        int nb2 = kMaxEstepBlocks;
        rc = (tab->n_states == 2)   ? launch_short_stats_k2(obs_dtype, tab, sa, di, &nb2, st)
             : (tab->n_states == 3) ? launch_short_stats_k3(obs_dtype, tab, sa, di, &nb2, st)
                                    : launch_short_stats_k4(obs_dtype, tab, sa, di, &nb2, st);
        if (rc != SMF_OK) return rc;
        reduce_partials_kernel<<<(S + 127) / 128, 128, 0, st>>>(sa.partials, nb2, S, stats);
        return cudaGetLastError() == cudaSuccess ? SMF_OK : SMF_ERR_CUDA;
      }
    }
    if (seq_launch_eligible(tab, n_reads, n_pos, obs_dtype, obs, false)) {
      // large batch: one thread per read, checkpointed forward pass + one backward pass with statistics
      char* wws = reinterpret_cast<char*>(workspace) + (size_t)kMaxEstepBlocks * (size_t)S * sizeof(double);
      SeqArgs q;
      memset(&q, 0, sizeof(q));
      q.g.obs = obs;
      q.g.N = n_reads;
      q.g.L = (int)n_pos;
      q.g.row_stride = (int)n_pos;
      q.g.unaligned = seq_rows_aligned(n_pos, obs_dtype, obs) ? 0 : 1;
      q.g.Tact = (int)((n_pos + kSeqCH - 1) / kSeqCH);
      q.g.ll = reinterpret_cast<double*>(wws);
      q.g.ckpt = reinterpret_cast<float*>(wws + (((size_t)n_reads * sizeof(double) + 63) & ~(size_t)63));
      q.g.gamma_state = -1;
      q.n_items = (n_reads + kSeqThreads - 1) / kSeqThreads;
      q.partials = reinterpret_cast<double*>(workspace);
      q.S = S;
      q.flush_chunks = 8;
      int nb2 = kMaxEstepBlocks;
      rc = (tab->n_states == 2)   ? launch_seq_k2(obs_dtype, MODE_STATS, tab, q, di, &nb2, st)
           : (tab->n_states == 3) ? launch_seq_k3(obs_dtype, MODE_STATS, tab, q, di, &nb2, st)
                                  : launch_seq_k4(obs_dtype, MODE_STATS, tab, q, di, &nb2, st);
      if (rc != SMF_OK) return rc;
      reduce_partials_kernel<<<(S + 127) / 128, 128, 0, st>>>(q.partials, nb2, S, stats);
      return cudaGetLastError() == cudaSuccess ? SMF_OK : SMF_ERR_CUDA;
    }
    Fb2Args b;
    memset(&b, 0, sizeof(b));
    Fb2Plan p2;
    if (walk_eligible(tab, n_reads, n_pos, true)) {
      // two-kernel path: boundary walk, then warp-independent sweeps
      static const int env_sch = [] { const char* e = getenv("SMF_SWEEP_CHUNK"); return e ? atoi(e) : 0; }();
      const int K = tab->n_states;
      const int CH = (env_sch >= 17 && (env_sch & 1)) ? env_sch : 17;
      char* wws = reinterpret_cast<char*>(workspace) + (size_t)kMaxEstepBlocks * (size_t)S * sizeof(double);
      const size_t cap = (size_t)((n_pos + 16) / 17) * (size_t)n_reads * K;
      WalkArgs w;
      w.obs = obs;
      w.N = n_reads;
      w.L = (int)n_pos;
      w.CH = CH;
      w.Tact = (int)((n_pos + CH - 1) / CH);
      w.bnd_v = reinterpret_cast<float*>(wws);
      w.bnd_w = w.bnd_v + cap;
      w.ll = reinterpret_cast<double*>(wws + 2 * cap * sizeof(float));
      rc = (K == 3) ? launch_walk_k3(obs_dtype, tab, w, st) : launch_walk_k4(obs_dtype, tab, w, st);
      if (rc != SMF_OK) return rc;
      SweepArgs sa;
      memset(&sa, 0, sizeof(sa));
      sa.obs = obs;
      sa.N = n_reads;
      sa.L = (int)n_pos;
      sa.CH = CH;
      sa.Tact = w.Tact;
      sa.slices = (w.Tact + 31) / 32;
      sa.bnd_v = w.bnd_v;
      sa.bnd_w = w.bnd_w;
      sa.ll_in = w.ll;
      sa.partials = reinterpret_cast<double*>(workspace);
      sa.S = S;
      sa.flush_every = 8;
      int nb2 = kMaxEstepBlocks;
      rc = (K == 3) ? launch_sweep_k3(obs_dtype, tab, sa, di, &nb2, st) : launch_sweep_k4(obs_dtype, tab, sa, di, &nb2, st);
      if (rc != SMF_OK) return rc;
      reduce_partials_kernel<<<(S + 127) / 128, 128, 0, st>>>(sa.partials, nb2, S, stats);
      return cudaGetLastError() == cudaSuccess ? SMF_OK : SMF_ERR_CUDA;
    }
    memset(&b, 0, sizeof(b));
    if (plan_fb2(tab->n_states, n_pos, true, S, obs_elem_size(obs_dtype), false, false, di.max_smem_optin, &b,
                 &p2) == SMF_OK) {
      b.obs = obs;
      b.N = n_reads;
      b.L = (int)n_pos;
      b.gamma = nullptr;
      b.gamma_state = -1;
      b.states = nullptr;
      b.loglik = nullptr;
      b.partials = reinterpret_cast<double*>(workspace);
      b.S = S;
      b.flush_every = 8;
      int nb2 = kMaxEstepBlocks;
      rc = launch_fb2(true, tab, obs_dtype, b, p2, di, &nb2, st);
      if (rc != SMF_OK) return rc;
      reduce_partials_kernel<<<(S + 127) / 128, 128, 0, st>>>(b.partials, nb2, S, stats);
      return cudaGetLastError() == cudaSuccess ? SMF_OK : SMF_ERR_CUDA;
    }
  }
  if (fb2_binned_eligible(tab, obs_dtype) && tab->n_states == 2 && !g_force_generic && cadence == 4 && n_pos > 1 &&
      fb2_aligned(obs, obs_dtype, nullptr, tab->n_states, -1)) {
    // distance-binned K = 2: per-thread, per-bin transition statistics in shared memory (fb2_kernel<..., BINNED, STATS>)
    Fb2Args b;
    Fb2Plan p2;
    const int lpad = (int)((n_pos + 16) / 17 * 17);
    const int bins_b = align_up(lpad, 16);
    const int btab_b = align_up(tab->n_bins * 4 * 4, 16);
    const int ls0_b = align_up(tab->n_bins * 8, 16);
    const int xi_stride = tab->n_bins | 1;
    int block = 0, prc = SMF_ERR_TOO_LONG;
    for (int round = 0; round < 3; ++round) {  // the accumulator area depends on the CTA size the planner picks
      memset(&b, 0, sizeof(b));
      const int sxi_b = block * xi_stride * 16;
      prc = plan_fb2_chunk(2, 17, n_pos, true, S, obs_elem_size(obs_dtype), false, false, di.max_smem_optin, &b, &p2, false,
                           false, bins_b + btab_b + ls0_b + sxi_b);
      if (prc != SMF_OK || p2.block == block) break;
      block = p2.block;
      prc = SMF_ERR_TOO_LONG;
    }
    if (prc == SMF_OK && block > 0) {
      b.off_btab = b.off_bins + bins_b;
      b.off_ls0 = b.off_btab + btab_b;
      b.off_sxi = b.off_ls0 + ls0_b;
      b.bins = bins;
      b.n_bins = tab->n_bins;
      b.obs = obs;
      b.N = n_reads;
      b.L = (int)n_pos;
      b.gamma_state = -1;
      b.partials = reinterpret_cast<double*>(workspace);
      b.S = S;
      b.flush_every = 8;
      int nb2 = kMaxEstepBlocks;
      rc = launch_fb2b_stats_k2(obs_dtype, tab, b, p2, di, &nb2, st);
      if (rc != SMF_OK) return rc;
      reduce_partials_kernel<<<(S + 127) / 128, 128, 0, st>>>(b.partials, nb2, S, stats);
      return cudaGetLastError() == cudaSuccess ? SMF_OK : SMF_ERR_CUDA;
    }
  }
  if (fb2_multi_eligible(tab, obs_dtype) && !g_force_generic && cadence == 4 &&
      fb2_aligned(obs, obs_dtype, nullptr, tab->n_states, -1)) {
    Fb2Args b;
    memset(&b, 0, sizeof(b));
    Fb2Plan p2;
    if (plan_fb2_chunk(tab->n_states, 17, n_pos, true, S, obs_elem_size(obs_dtype) * tab->n_channels, false, false,
                       di.max_smem_optin, &b, &p2) == SMF_OK) {
      b.obs = obs;
      b.N = n_reads;
      b.L = (int)n_pos;
      b.gamma_state = -1;
      b.partials = reinterpret_cast<double*>(workspace);
      b.S = S;
      b.flush_every = 8;
      b.late_load = 1;  // the backward sweep re-reads the observations for the per-channel emission statistics
      int nb2 = kMaxEstepBlocks;
      rc = tab->n_states == 2 ? launch_fb2m_k2(tab->n_channels, true, obs_dtype, tab, b, p2, di, &nb2, st)
                              : launch_fb2m_k3(tab->n_channels, true, obs_dtype, tab, b, p2, di, &nb2, st);
      if (rc != SMF_OK) return rc;
      reduce_partials_kernel<<<(S + 127) / 128, 128, 0, st>>>(b.partials, nb2, S, stats);
      return cudaGetLastError() == cudaSuccess ? SMF_OK : SMF_ERR_CUDA;
    }
  }
  FbArgs a;
  memset(&a, 0, sizeof(a));
  FbPlan plan;
  rc = plan_fb(tab->n_states, tab->n_channels, tab->n_bins, n_pos, true, S, di.max_smem_optin, &a, &plan);
  if (rc != SMF_OK) return rc;
  a.obs = obs;
  a.obs_dtype = obs_dtype;
  a.N = n_reads;
  a.L = (int)n_pos;
  a.C = tab->n_channels;
  a.bins = bins;
  a.gamma = nullptr;
  a.gamma_state = -1;
  a.states = nullptr;
  a.loglik = nullptr;
  a.partials = reinterpret_cast<double*>(workspace);
  a.S = S;
  a.flush_every = 8;
  a.resc_mask = cadence - 1;
  int nblocks = kMaxEstepBlocks;
  rc = launch_fb(true, tab, a, plan, di, &nblocks, st);
  if (rc != SMF_OK) return rc;
  reduce_partials_kernel<<<(S + 127) / 128, 128, 0, st>>>(a.partials, nblocks, S, stats);
  return cudaGetLastError() == cudaSuccess ? SMF_OK : SMF_ERR_CUDA;
}

int smf_hmm_viterbi(const smf_hmm_tables* tab, const void* obs, int obs_dtype, int64_t n_reads,
                    int64_t n_pos, const int8_t* bins, int8_t* states, void* workspace,
                    size_t workspace_bytes, void* stream) {
  int rc = check_tables(tab);
  if (rc != SMF_OK) return rc;
  if (n_reads < 0 || n_pos < 0) return SMF_ERR_BAD_ARG;
  if (n_reads == 0 || n_pos == 0) return SMF_OK;
  if (obs == nullptr || states == nullptr || workspace == nullptr) return SMF_ERR_BAD_ARG;
  if (obs_dtype < SMF_OBS_F32 || obs_dtype > SMF_OBS_U8) return SMF_ERR_BAD_ARG;
  if (n_pos > 0x7fffffff) return SMF_ERR_TOO_LONG;
  if (tab->n_bins > 1 && n_pos > 1 && bins == nullptr) return SMF_ERR_BAD_ARG;
  if (workspace_bytes < (size_t)n_reads * (size_t)n_pos) return SMF_ERR_WORKSPACE;
  DeviceInfo di;
  rc = get_device_info(&di);
  if (rc != SMF_OK) return rc;
  VitArgs a;
  a.obs = obs;
  a.obs_dtype = obs_dtype;
  a.N = n_reads;
  a.L = (int)n_pos;
  a.C = tab->n_channels;
  a.bins = bins;
  a.states = states;
  a.bp = reinterpret_cast<uint8_t*>(workspace);
  cudaStream_t st = (cudaStream_t)stream;
  // single channel, homogeneous transitions, float / coded rows of whole 4-position words: the direct kernel
  // (no shared-memory staging); "vit" = 0 keeps the staged kernel, which handles every other case
  const bool rows_ok = (n_pos & 3) == 0 && (reinterpret_cast<uintptr_t>(obs) & 15u) == 0 &&
                       (reinterpret_cast<uintptr_t>(states) & 3u) == 0 && (reinterpret_cast<uintptr_t>(workspace) & 3u) == 0 &&
                       (obs_dtype == SMF_OBS_F32 || (obs_dtype == SMF_OBS_U8 && (n_pos & 15) == 0 &&
                                                     (reinterpret_cast<uintptr_t>(states) & 15u) == 0 &&
                                                     (reinterpret_cast<uintptr_t>(workspace) & 15u) == 0));
  // measured (profiles/r02_viterbi.txt): 2-2.8x faster for the uint8 code, 1.4x for K = 4, on par for K = 2 and for
  // short K = 3 reads; long float K = 3 reads stay on the staged kernel (9.1 vs 9.8 ms on 262,144 x 10 kb: both are
  // bound by the fp64 pipe there, and the staged kernel's 128-byte row segments are kinder to DRAM)
  const bool direct_wins = obs_dtype == SMF_OBS_U8 || tab->n_states != 3 || n_pos <= 2048;
  if (g_vit_mode != 0 && (g_vit_mode == 1 || direct_wins) && tab->n_channels == 1 && tab->n_bins == 1 && rows_ok &&
      !g_force_generic) {
    switch (tab->n_states) {
      case 2: return launch_vit_direct_k2(tab, a, di, st);
      case 3: return launch_vit_direct_k3(tab, a, di, st);
      case 4: return launch_vit_direct_k4(tab, a, di, st);
    }
  }
  switch (tab->n_states) {
    case 2: return launch_vit_k2(tab, a, di, st);
    case 3: return launch_vit_k3(tab, a, di, st);
    case 4: return launch_vit_k4(tab, a, di, st);
  }
  return SMF_ERR_BAD_ARG;
}

int smf_hmm_call_features(const int8_t* states, const float* post, int target_state,
                          double threshold, int64_t n_reads, int64_t n_pos,
                          const int64_t* coords, const int32_t* grid_left,
                          const int32_t* grid_right, int64_t n_grid, const double* class_lo_host,
                          const double* class_hi_host, int n_classes, uint8_t* class_out,
                          int32_t* run_len, int32_t* intervals, int64_t max_intervals,
                          unsigned long long* n_intervals, void* stream) {
  if (n_reads < 0 || n_pos < 0 || n_grid < 0) return SMF_ERR_BAD_ARG;
  if (n_reads == 0 || n_grid == 0) return SMF_OK;
  if ((states == nullptr) == (post == nullptr)) return SMF_ERR_BAD_ARG;  // exactly one of them
  if (!coords || !grid_left || !grid_right) return SMF_ERR_BAD_ARG;
  if (class_out == nullptr && (run_len != nullptr || intervals == nullptr)) return SMF_ERR_BAD_ARG;
  if (intervals != nullptr && n_intervals == nullptr) return SMF_ERR_BAD_ARG;
  if (n_pos > 0x3fffffff || n_grid > 0x3fffffff) return SMF_ERR_TOO_LONG;
  ClassTable ct;
  int rc = fill_class_table(class_lo_host, class_hi_host, n_classes, &ct);
  if (rc != SMF_OK) return rc;
  DeviceInfo di;
  rc = get_device_info(&di);
  if (rc != SMF_OK) return rc;
  const int maxruns = ((int)n_pos + 1) / 2 + 1;
  const size_t smem = (size_t)maxruns * 8 + (2 * kFeatWarps + 2) * 4 + (size_t)align_up(maxruns, 16);
  if ((long long)smem > di.max_smem_optin) return SMF_ERR_TOO_LONG;
  if (prepare_kernel((const void*)call_features_kernel, (int)smem, false) != SMF_OK) return SMF_ERR_CUDA;
  int occ = 0;
  if (cached_occupancy((const void*)call_features_kernel, kFeatThreads, (int)smem, &occ) != SMF_OK || occ < 1)
    return SMF_ERR_CUDA;
  FeatArgs a;
  a.states = states;
  a.post = post;
  a.target = target_state;
  // (double)post >= threshold  <=>  post >= the smallest float that is >= threshold: the kernels compare
  // floats and still decide exactly like the reference's fp64 comparison (HMM.py:1318-1322)
  float thr_up = (float)threshold;
  if ((double)thr_up < threshold) thr_up = nextafterf(thr_up, INFINITY);
  a.thr = thr_up;
  a.N = n_reads;
  a.L = (int)n_pos;
  a.coords = reinterpret_cast<const long long*>(coords);
  a.grid_left = grid_left;
  a.grid_right = grid_right;
  a.V = (int)n_grid;
  a.class_out = class_out;
  a.run_len = run_len;
  a.intervals = intervals;
  a.max_intervals = max_intervals;
  a.n_intervals = n_intervals;
  if (class_out == nullptr) {
    // interval records only: warp-per-read kernel without shared-memory run tables
    if (reinterpret_cast<uintptr_t>(intervals) & 15u) return SMF_ERR_ALIGN;
    const long long warps_needed = n_reads;
    long long blocks = (warps_needed + kFeatWarps - 1) / kFeatWarps;
    const long long capb = (long long)di.num_sms * 8;
    if (blocks > capb) blocks = capb;
    // rows that start on 16-byte boundaries take the vectorised variant (128-bit loads, bit-mask runs)
    static const int env_vec = [] { const char* e = getenv("SMF_INTERVALS_VEC"); return e ? atoi(e) : 1; }();
    const bool vec_states = states != nullptr && (n_pos % 16) == 0 && (reinterpret_cast<uintptr_t>(states) & 15u) == 0;
    const bool vec_post = post != nullptr && (n_pos % 4) == 0 && (reinterpret_cast<uintptr_t>(post) & 15u) == 0;
    if (env_vec && vec_states)
      call_intervals_vec_kernel<false><<<(int)blocks, kFeatThreads, 0, (cudaStream_t)stream>>>(ct, a);
    else if (env_vec && vec_post)
      call_intervals_vec_kernel<true><<<(int)blocks, kFeatThreads, 0, (cudaStream_t)stream>>>(ct, a);
    else
      call_intervals_kernel<<<(int)blocks, kFeatThreads, 0, (cudaStream_t)stream>>>(ct, a);
    return cudaGetLastError() == cudaSuccess ? SMF_OK : SMF_ERR_CUDA;
  }
  long long cap = (long long)di.num_sms * occ;
  const int nblocks = (int)(n_reads < cap ? n_reads : cap);
  call_features_kernel<<<nblocks, kFeatThreads, smem, (cudaStream_t)stream>>>(ct, a);
  return cudaGetLastError() == cudaSuccess ? SMF_OK : SMF_ERR_CUDA;
}

int smf_hmm_merge_runs(const uint8_t* layer, int64_t n_reads, int64_t n_grid,
                       const int64_t* coords, int coords_are_ints, int64_t distance_threshold,
                       uint8_t* merged, int32_t* merged_len, const double* class_lo_host,
                       const double* class_hi_host, int n_classes, uint8_t* class_out,
                       int32_t* len_cls, void* stream) {
  if (n_reads < 0 || n_grid < 0) return SMF_ERR_BAD_ARG;
  if (n_reads == 0 || n_grid == 0) return SMF_OK;
  if (!layer || !merged || !merged_len) return SMF_ERR_BAD_ARG;
  if (coords_are_ints && !coords) return SMF_ERR_BAD_ARG;
  if (n_grid > 0x3fffffff) return SMF_ERR_TOO_LONG;
  ClassTable ct;
  int rc = fill_class_table(class_lo_host, class_hi_host, n_classes, &ct);
  if (rc != SMF_OK) return rc;
  DeviceInfo di;
  rc = get_device_info(&di);
  if (rc != SMF_OK) return rc;
  // five word arrays (bit masks + word scans) + one staging tile per warp
  const size_t smem = (size_t)(((((int)n_grid + 31) / 32) + 3) & ~3) * 20 + (size_t)kMcWarps * kMcTileBytes;
  if ((long long)smem > di.max_smem_optin) return SMF_ERR_TOO_LONG;
  if (prepare_kernel((const void*)merge_runs_kernel, (int)smem, false) != SMF_OK) return SMF_ERR_CUDA;
  int occ = 0;
  if (cached_occupancy((const void*)merge_runs_kernel, kMcThreads, (int)smem, &occ) != SMF_OK || occ < 1)
    return SMF_ERR_CUDA;
  MergeArgs a;
  memset(&a, 0, sizeof(a));
  a.layer = layer;
  a.N = n_reads;
  a.V = (int)n_grid;
  a.coords = reinterpret_cast<const long long*>(coords);
  a.coords_are_ints = coords_are_ints;
  a.dt = distance_threshold;
  a.merged = merged;
  a.merged_len = merged_len;
  a.class_out = (n_classes > 0) ? class_out : nullptr;
  a.len_cls = (n_classes > 0) ? len_cls : nullptr;
  long long cap = (long long)di.num_sms * occ;
  const int nblocks = (int)(n_reads < cap ? n_reads : cap);
  IntClassTable ict;
  to_int_classes(ct, &ict);
  merge_runs_kernel<<<nblocks, kMcThreads, smem, (cudaStream_t)stream>>>(ict, a);
  return cudaGetLastError() == cudaSuccess ? SMF_OK : SMF_ERR_CUDA;
}

int smf_hmm_merge_chain(const uint8_t* layer, int64_t n_reads, int64_t n_grid, const int64_t* coords,
                        int coords_are_ints, int64_t distance_threshold, const double* class_lo_host,
                        const double* class_hi_host, int n_classes, const uint8_t* covered,
                        const int64_t* span_coords, const double* read_start, const double* read_end,
                        uint8_t* code_out, int32_t* run_len, void* stream) {
  if (n_reads < 0 || n_grid < 0) return SMF_ERR_BAD_ARG;
  if (n_reads == 0 || n_grid == 0) return SMF_OK;
  if (!layer || !code_out || !run_len) return SMF_ERR_BAD_ARG;
  if (coords_are_ints && !coords) return SMF_ERR_BAD_ARG;
  if (!covered && (!span_coords || !read_start || !read_end)) return SMF_ERR_BAD_ARG;
  if (n_classes > 15) return SMF_ERR_BAD_ARG;  // four code bits
  if (n_grid > 0x3fffffff) return SMF_ERR_TOO_LONG;
  ClassTable ct;
  int rc = fill_class_table(class_lo_host, class_hi_host, n_classes, &ct);
  if (rc != SMF_OK) return rc;
  DeviceInfo di;
  rc = get_device_info(&di);
  if (rc != SMF_OK) return rc;
  // five word arrays (bit masks + word scans) + one staging tile per warp
  const size_t smem = (size_t)(((((int)n_grid + 31) / 32) + 3) & ~3) * 20 + (size_t)kMcWarps * kMcTileBytes;
  if ((long long)smem > di.max_smem_optin) return SMF_ERR_TOO_LONG;
  if (prepare_kernel((const void*)merge_runs_kernel, (int)smem, false) != SMF_OK) return SMF_ERR_CUDA;
  int occ = 0;
  if (cached_occupancy((const void*)merge_runs_kernel, kMcThreads, (int)smem, &occ) != SMF_OK || occ < 1)
    return SMF_ERR_CUDA;
  MergeArgs a;
  memset(&a, 0, sizeof(a));
  a.layer = layer;
  a.N = n_reads;
  a.V = (int)n_grid;
  a.coords = reinterpret_cast<const long long*>(coords);
  a.coords_are_ints = coords_are_ints;
  a.dt = distance_threshold;
  a.merged_len = run_len;
  a.packed = code_out;
  a.covered = covered;
  a.span_coords = reinterpret_cast<const long long*>(span_coords);
  a.read_start = read_start;
  a.read_end = read_end;
  long long cap = (long long)di.num_sms * occ;
  const int nblocks = (int)(n_reads < cap ? n_reads : cap);
  IntClassTable ict;
  to_int_classes(ct, &ict);
  merge_runs_kernel<<<nblocks, kMcThreads, smem, (cudaStream_t)stream>>>(ict, a);
  return cudaGetLastError() == cudaSuccess ? SMF_OK : SMF_ERR_CUDA;
}

int smf_hmm_span_mask(const uint8_t* covered, const int64_t* coords, const double* read_start,
                      const double* read_end, int64_t n_reads, int64_t n_grid, uint8_t* valid_out,
                      void* stream) {
  if (n_reads < 0 || n_grid < 0) return SMF_ERR_BAD_ARG;
  if (n_reads == 0 || n_grid == 0) return SMF_OK;
  if (!valid_out) return SMF_ERR_BAD_ARG;
  if (!covered && (!coords || !read_start || !read_end)) return SMF_ERR_BAD_ARG;
  if (n_grid > 0x3fffffff) return SMF_ERR_TOO_LONG;
  SpanArgs a;
  a.covered = covered;
  a.coords = reinterpret_cast<const long long*>(coords);
  a.start = read_start;
  a.end = read_end;
  a.N = n_reads;
  a.V = (int)n_grid;
  a.valid = valid_out;
  dim3 grid((unsigned)((n_grid + 255) / 256 > 64 ? 64 : (n_grid + 255) / 256),
            (unsigned)(n_reads > 32768 ? 32768 : n_reads));
  span_mask_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  return cudaGetLastError() == cudaSuccess ? SMF_OK : SMF_ERR_CUDA;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------
// grouped launches (seq_kernel.cuh, GROUPED = true)
// ---------------------------------------------------------------------------------------
namespace {

struct GroupedPlan {
  size_t off_groups, off_item_group, off_tabs, off_active, off_stats, off_partials, off_scratch, total;
  long long n_items;
  int S;
};

inline size_t up256(size_t x) { return (x + 255) & ~(size_t)255; }
inline size_t seq_tables_bytes(int K) {
  return K == 2 ? sizeof(FbTables<2>) : (K == 3 ? sizeof(FbTables<3>) : sizeof(FbTables<4>));
}

// Validates the group table and lays the workspace out:
//   SeqGroup[G] | item_group[n_items] | FbTables<K>[G] | active[G] | stats[G][S] | partials[n_items][S] (fit)
//   | per group: log-likelihood [N] (when the caller gave none) and alpha checkpoints [Tact][N][KP]
// With `out` != nullptr the device group table and the item table are also written (host copies).
// obs_dtype < 0: sizing only (the workspace size does not depend on the element type; strides are not checked).
int grouped_plan(int K, int obs_dtype, const smf_hmm_group* gh, int G, bool fit, GroupedPlan* pl, char* ws_base,
                 SeqGroup* out_groups, int* out_item_group) {
  if (K < 2 || K > SMF_MAX_STATES || gh == nullptr || G < 1) return SMF_ERR_BAD_ARG;
  const bool sizing_only = obs_dtype < 0;
  if (!sizing_only && obs_dtype != SMF_OBS_U8 && obs_dtype != SMF_OBS_F32) return SMF_ERR_BAD_ARG;
  const int KP = (K == 2) ? 2 : 4;
  const int S = K + K * K + 2 * K + 1;
  long long n_items = 0;
  for (int g = 0; g < G; ++g) {
    const smf_hmm_group& q = gh[g];
    if (q.n_reads < 0 || q.n_pos < 1 || q.n_pos > (1 << 24) || (q.n_reads > 0 && q.obs == nullptr)) return SMF_ERR_BAD_ARG;
    if (q.row_stride < q.n_pos) return SMF_ERR_BAD_ARG;
    if (!sizing_only && (q.row_stride % (obs_dtype == SMF_OBS_U8 ? 16 : 4)) != 0) return SMF_ERR_ALIGN;
    if (!sizing_only && (reinterpret_cast<uintptr_t>(q.obs) & 15u)) return SMF_ERR_ALIGN;
    if (!fit && (q.gamma != nullptr || q.states != nullptr)) {
      if (q.out_stride < q.n_pos || (q.out_stride & 3)) return SMF_ERR_ALIGN;
      if ((reinterpret_cast<uintptr_t>(q.gamma) & 15u) || (reinterpret_cast<uintptr_t>(q.states) & 3u)) return SMF_ERR_ALIGN;
      if (q.gamma_state >= K) return SMF_ERR_BAD_ARG;
    }
    n_items += (q.n_reads + kSeqThreads - 1) / kSeqThreads;
    if (n_items > 0x7fffffff) return SMF_ERR_BAD_ARG;
  }
  size_t off = 0;
  pl->S = S;
  pl->n_items = n_items;
  pl->off_groups = off;
  off = up256(off + (size_t)G * sizeof(SeqGroup));
  pl->off_item_group = off;
  off = up256(off + (size_t)n_items * sizeof(int));
  pl->off_tabs = off;
  off = up256(off + (size_t)G * seq_tables_bytes(K));
  pl->off_active = off;
  off = up256(off + (size_t)(G + 1) * sizeof(int));  // + the grid-barrier counter of the one-launch fit loop
  pl->off_stats = off;
  off = up256(off + (fit ? (size_t)G * S * sizeof(double) : 0));
  pl->off_partials = off;
  off = up256(off + (fit ? (size_t)n_items * S * sizeof(double) : 0));
  pl->off_scratch = off;
  // Work items are dealt to the CTAs in table order: the groups with the longest reads come first in the item
  // table so that the tail of the launch is made of short items (outputs keep the caller's group order).
  std::vector<int> order(G);
  for (int g = 0; g < G; ++g) order[g] = g;
  std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return gh[x].n_pos > gh[y].n_pos; });
  long long item0 = 0;
  for (int oi = 0; oi < G; ++oi) {
    const int g = order[oi];
    const smf_hmm_group& q = gh[g];
    const int tact = (q.n_pos + kSeqCH - 1) / kSeqCH;
    size_t off_ll = 0;
    const bool own_ll = fit || q.loglik == nullptr;
    if (own_ll) {
      off_ll = off;
      off = up256(off + (size_t)q.n_reads * sizeof(double));
    }
    const size_t off_ck = off;
    off = up256(off + (size_t)tact * (size_t)q.n_reads * KP * sizeof(float));
    const long long items = (q.n_reads + kSeqThreads - 1) / kSeqThreads;
    if (out_groups != nullptr) {
      SeqGroup& d = out_groups[g];
      memset(&d, 0, sizeof(d));
      d.obs = q.obs;
      d.N = q.n_reads;
      d.L = q.n_pos;
      d.row_stride = q.row_stride;
      d.Tact = tact;
      d.item0 = (int)item0;
      d.ckpt = reinterpret_cast<float*>(ws_base + off_ck);
      d.ll = own_ll ? reinterpret_cast<double*>(ws_base + off_ll) : q.loglik;
      d.gamma = fit ? nullptr : q.gamma;
      d.states = fit ? nullptr : q.states;
      d.gamma_state = q.gamma_state;
      d.out_stride = q.out_stride;
      for (long long i = 0; i < items; ++i) out_item_group[item0 + i] = g;
    }
    item0 += items;
  }
  pl->total = off;
  return SMF_OK;
}

int grouped_upload(const GroupedPlan& pl, int G, const std::vector<SeqGroup>& groups, const std::vector<int>& item_group,
                   char* ws, cudaStream_t st) {
  if (cudaMemcpyAsync(ws + pl.off_groups, groups.data(), (size_t)G * sizeof(SeqGroup), cudaMemcpyHostToDevice, st) !=
      cudaSuccess)
    return SMF_ERR_CUDA;
  if (cudaMemcpyAsync(ws + pl.off_item_group, item_group.data(), (size_t)pl.n_items * sizeof(int),
                      cudaMemcpyHostToDevice, st) != cudaSuccess)
    return SMF_ERR_CUDA;
  return SMF_OK;
}

}  // namespace

extern "C" {

size_t smf_hmm_grouped_workspace_bytes(int n_states, const smf_hmm_group* groups_host, int n_groups, int fit) {
  GroupedPlan pl;
  const int rc = grouped_plan(n_states, -1, groups_host, n_groups, fit != 0, &pl, nullptr, nullptr, nullptr);
  return rc == SMF_OK ? pl.total : 0;
}

int smf_hmm_grouped_posterior(int obs_dtype, const smf_hmm_group* groups_host, int n_groups,
                              const smf_hmm_tables* tabs_host, void* workspace, size_t workspace_bytes,
                              void* stream) {
  if (tabs_host == nullptr || workspace == nullptr || n_groups < 1) return SMF_ERR_BAD_ARG;
  const int K = tabs_host[0].n_states;
  for (int g = 0; g < n_groups; ++g) {
    int rc = check_tables(&tabs_host[g]);
    if (rc != SMF_OK) return rc;
    if (tabs_host[g].n_states != K || tabs_host[g].n_channels != 1 || tabs_host[g].n_bins != 1) return SMF_ERR_BAD_ARG;
    if (rescale_cadence(&tabs_host[g]) != 4) return SMF_ERR_BAD_ARG;
  }
  GroupedPlan pl;
  std::vector<SeqGroup> groups(n_groups);
  int rc = grouped_plan(K, obs_dtype, groups_host, n_groups, false, &pl, nullptr, nullptr, nullptr);
  if (rc != SMF_OK) return rc;
  if (workspace_bytes < pl.total) return SMF_ERR_WORKSPACE;
  if (reinterpret_cast<uintptr_t>(workspace) & 255u) return SMF_ERR_ALIGN;
  std::vector<int> item_group((size_t)pl.n_items);
  char* ws = reinterpret_cast<char*>(workspace);
  rc = grouped_plan(K, obs_dtype, groups_host, n_groups, false, &pl, ws, groups.data(), item_group.data());
  if (rc != SMF_OK) return rc;
  DeviceInfo di;
  rc = get_device_info(&di);
  if (rc != SMF_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  rc = grouped_upload(pl, n_groups, groups, item_group, ws, st);
  if (rc != SMF_OK) return rc;
  // per-group kernel tables, built on the host like for every other launch
  const size_t tb = seq_tables_bytes(K);
  std::vector<unsigned char> tabs((size_t)n_groups * tb);
  for (int g = 0; g < n_groups; ++g) {
    void* dst = tabs.data() + (size_t)g * tb;
    if (K == 2) build_fb_tables<2>(&tabs_host[g], reinterpret_cast<FbTables<2>*>(dst));
    else if (K == 3) build_fb_tables<3>(&tabs_host[g], reinterpret_cast<FbTables<3>*>(dst));
    else build_fb_tables<4>(&tabs_host[g], reinterpret_cast<FbTables<4>*>(dst));
  }
  if (cudaMemcpyAsync(ws + pl.off_tabs, tabs.data(), tabs.size(), cudaMemcpyHostToDevice, st) != cudaSuccess)
    return SMF_ERR_CUDA;
  bool any_out = false;
  for (int g = 0; g < n_groups; ++g) any_out = any_out || groups_host[g].gamma != nullptr || groups_host[g].states != nullptr;
  SeqArgs a;
  memset(&a, 0, sizeof(a));
  a.groups = reinterpret_cast<const SeqGroup*>(ws + pl.off_groups);
  a.tabs = ws + pl.off_tabs;
  a.item_group = reinterpret_cast<const int*>(ws + pl.off_item_group);
  a.active = nullptr;
  a.n_items = pl.n_items;
  a.S = 0;
  a.flush_chunks = 8;
  const int mode = any_out ? MODE_POST : -1;
  return K == 2 ? launch_seq_grouped_k2(obs_dtype, mode, a, di, st)
                : (K == 3 ? launch_seq_grouped_k3(obs_dtype, mode, a, di, st) : launch_seq_grouped_k4(obs_dtype, mode, a, di, st));
}

// phases of a grouped fit: SMF_FIT_PHASE_ALL runs begin + max_iter x (E-step, M-step) in one call
static int grouped_fit_run(int phase, int K, int obs_dtype, const smf_hmm_group* groups_host, int n_groups, double* params,
                           int iter, int max_iter, double tol, double eps, int update_mask, double* hist, int32_t* n_iter,
                           int32_t* status, double* stats, void* workspace, size_t workspace_bytes, void* stream) {
  if (params == nullptr || hist == nullptr || n_iter == nullptr || status == nullptr || workspace == nullptr)
    return SMF_ERR_BAD_ARG;
  if (max_iter < 0 || !(eps > 0.0)) return SMF_ERR_BAD_ARG;
  if (phase != SMF_FIT_PHASE_ALL && (iter < 0 || iter >= (max_iter > 0 ? max_iter : 1))) return SMF_ERR_BAD_ARG;
  GroupedPlan pl;
  int rc = grouped_plan(K, obs_dtype, groups_host, n_groups, true, &pl, nullptr, nullptr, nullptr);
  if (rc != SMF_OK) return rc;
  if (workspace_bytes < pl.total) return SMF_ERR_WORKSPACE;
  if (reinterpret_cast<uintptr_t>(workspace) & 255u) return SMF_ERR_ALIGN;
  char* ws = reinterpret_cast<char*>(workspace);
  DeviceInfo di;
  rc = get_device_info(&di);
  if (rc != SMF_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (phase == SMF_FIT_PHASE_ALL || phase == SMF_FIT_PHASE_BEGIN) {
    std::vector<SeqGroup> groups(n_groups);
    std::vector<int> item_group((size_t)pl.n_items);
    rc = grouped_plan(K, obs_dtype, groups_host, n_groups, true, &pl, ws, groups.data(), item_group.data());
    if (rc != SMF_OK) return rc;
    rc = grouped_upload(pl, n_groups, groups, item_group, ws, st);
    if (rc != SMF_OK) return rc;
  }
  SeqFitState f;
  memset(&f, 0, sizeof(f));
  f.params = params;
  f.stats = stats != nullptr ? stats : reinterpret_cast<double*>(ws + pl.off_stats);
  f.hist = hist;
  f.active = reinterpret_cast<int*>(ws + pl.off_active);
  f.n_iter = n_iter;
  f.status = status;
  f.tabs = ws + pl.off_tabs;
  f.groups = reinterpret_cast<const SeqGroup*>(ws + pl.off_groups);
  f.partials = reinterpret_cast<const double*>(ws + pl.off_partials);
  f.n_groups = n_groups;
  f.S = pl.S;
  f.max_iter = max_iter > 0 ? max_iter : 1;
  f.eps = eps;
  f.tol = tol;
  f.update_start = (update_mask & 1) != 0;
  f.update_trans = (update_mask & 2) != 0;
  f.update_emission = (update_mask & 4) != 0;
  auto mstep = [&](int it, int what) {
    f.iter = it;
    f.what = what;
    return K == 2 ? launch_seq_mstep_k2(f, st) : (K == 3 ? launch_seq_mstep_k3(f, st) : launch_seq_mstep_k4(f, st));
  };
  SeqArgs a;
  memset(&a, 0, sizeof(a));
  a.groups = reinterpret_cast<const SeqGroup*>(ws + pl.off_groups);
  a.tabs = ws + pl.off_tabs;
  a.item_group = reinterpret_cast<const int*>(ws + pl.off_item_group);
  a.active = f.active;
  a.n_items = pl.n_items;
  a.partials = reinterpret_cast<double*>(ws + pl.off_partials);
  a.S = pl.S;
  a.flush_chunks = 8;
  auto estep = [&]() {
    if (pl.n_items == 0) return (int)SMF_OK;  // this rank holds no reads of any group
    return K == 2 ? launch_seq_grouped_k2(obs_dtype, MODE_STATS, a, di, st)
                  : (K == 3 ? launch_seq_grouped_k3(obs_dtype, MODE_STATS, a, di, st)
                            : launch_seq_grouped_k4(obs_dtype, MODE_STATS, a, di, st));
  };
  switch (phase) {
    case SMF_FIT_PHASE_BEGIN:
      return mstep(-1, SEQ_MSTEP_ALL);  // tables of the initial parameters; every group active
    case SMF_FIT_PHASE_ESTEP:
      rc = estep();
      if (rc != SMF_OK) return rc;
      return mstep(iter, SEQ_MSTEP_REDUCE_ONLY);  // per-item partials -> stats[G][S]
    case SMF_FIT_PHASE_MSTEP:
      return mstep(iter, SEQ_MSTEP_UPDATE_ONLY);
    case SMF_FIT_PHASE_ALL:
      rc = mstep(-1, SEQ_MSTEP_ALL);
      if (rc != SMF_OK) return rc;
      // small fits (every work item resident at once): the whole loop in one cooperative launch
      if (max_iter > 1 && pl.n_items > 0 && g_fit_loop_mode.load() != 0) {
        f.iter = 0;
        f.what = SEQ_MSTEP_ALL;
        unsigned* counter = reinterpret_cast<unsigned*>(ws + pl.off_active) + n_groups;
        rc = K == 2 ? launch_seq_fit_loop_k2(obs_dtype, a, f, counter, di, st)
                    : (K == 3 ? launch_seq_fit_loop_k3(obs_dtype, a, f, counter, di, st)
                              : launch_seq_fit_loop_k4(obs_dtype, a, f, counter, di, st));
        if (rc != SMF_ERR_TOO_LONG) return rc;
      }
      for (int it = 0; it < max_iter; ++it) {
        rc = estep();
        if (rc != SMF_OK) return rc;
        rc = mstep(it, SEQ_MSTEP_ALL);
        if (rc != SMF_OK) return rc;
      }
      return SMF_OK;
  }
  return SMF_ERR_BAD_ARG;
}

int smf_hmm_grouped_fit(int n_states, int obs_dtype, const smf_hmm_group* groups_host, int n_groups,
                        double* params, int max_iter, double tol, double eps, int update_mask, double* hist,
                        int32_t* n_iter, int32_t* status, void* workspace, size_t workspace_bytes, void* stream) {
  return grouped_fit_run(SMF_FIT_PHASE_ALL, n_states, obs_dtype, groups_host, n_groups, params, 0, max_iter, tol, eps,
                         update_mask, hist, n_iter, status, nullptr, workspace, workspace_bytes, stream);
}

int smf_hmm_grouped_fit_step(int phase, int n_states, int obs_dtype, const smf_hmm_group* groups_host, int n_groups,
                             double* params, int iter, int max_iter, double tol, double eps, int update_mask,
                             double* hist, int32_t* n_iter, int32_t* status, double* stats, void* workspace,
                             size_t workspace_bytes, void* stream) {
  if (phase != SMF_FIT_PHASE_BEGIN && phase != SMF_FIT_PHASE_ESTEP && phase != SMF_FIT_PHASE_MSTEP) return SMF_ERR_BAD_ARG;
  if (stats == nullptr) return SMF_ERR_BAD_ARG;
  return grouped_fit_run(phase, n_states, obs_dtype, groups_host, n_groups, params, iter, max_iter, tol, eps, update_mask,
                         hist, n_iter, status, stats, workspace, workspace_bytes, stream);
}

}  // extern "C"
