// Host-side launch plumbing shared by the translation units of libsmfhmm.so.
// Kernels are instantiated in per-K .cu files (compiled in parallel); api.cu only plans and dispatches.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "fb2_kernel.cuh"
#include "fb_kernel.cuh"
#include "smf_hmm.h"
#include "viterbi_kernel.cuh"

namespace smf {

struct DeviceInfo {
  int device = -1;
  int num_sms = 0;
  int max_smem_optin = 0;
};

// Per-(kernel, device) launch state, kept behind a mutex in api.cu: the dynamic shared-memory opt-in is
// raised only when a launch needs more than any earlier one did (cudaFuncSetAttribute is not free), and
// occupancy queries are answered from a small cache.
int prepare_kernel(const void* kern, int smem_bytes, bool max_carveout);
int cached_occupancy(const void* kern, int block, int smem_bytes, int* occ_out);

struct FbPlan {
  int T, chunk, groups, smem, block;
};
struct Fb2Plan {
  int block, smem, chunk;
};

// Threads a group may use for a given model / mode (register budget; equals the kernels' launch bounds).
constexpr int fb_max_threads_c(int K, bool stats) {
  return stats ? (K == 2 ? 640 : (K == 3 ? 512 : 384)) : (K == 4 ? 512 : 640);
}
constexpr int fb2_max_threads_c(int K, bool stats) {
  return stats ? (K == 2 ? 640 : (K == 3 ? 512 : 480)) : (K == 4 ? 512 : 640);
}

// WALK variant (no chunk matrices live): launch bounds of fb2_kernel<..., WALK = true>
constexpr int fb2w_max_threads_c(int K, bool stats) { return stats ? 512 : 640; }

template <int K>
static void build_fb_tables(const smf_hmm_tables* t, FbTables<K>* out) {
  memset(out, 0, sizeof(*out));
  const int C = t->n_channels;
  const double log2e = 1.4426950408889634074;
  double pisum = 0.0;
  for (int k = 0; k < K; ++k) {
    const double p = exp(t->log_start[k]);
    out->pi[k] = (float)p;
    pisum += (double)out->pi[k];
  }
  out->log_pi_sum = log(pisum);
  for (int b = 0; b < t->n_bins; ++b)
    for (int i = 0; i < K; ++i)
      for (int j = 0; j < K; ++j) out->A[b][i][j] = (float)exp(t->log_trans[(b * K + i) * K + j]);
  for (int c = 0; c < C; ++c) {
    const double l0_0 = t->log_emit0[0 * C + c];
    const double d_0 = t->log_emit1[0 * C + c] - l0_0;
    out->l00[c] = l0_0;
    out->dl0[c] = d_0;
    for (int k = 0; k < K; ++k) {
      const double l0 = t->log_emit0[k * C + c];
      const double d = t->log_emit1[k * C + c] - l0;
      out->rl0[k][c] = (float)((l0 - l0_0) * log2e);
      out->rdl[k][c] = (float)((d - d_0) * log2e);
    }
  }
  // fast-path tables (bin 0, channel 0)
  {
    double sdiag[K];
    for (int j = 0; j < K; ++j) sdiag[j] = exp(t->log_trans[j * K + j]);
    for (int i = 0; i < K; ++i)
      for (int j = 0; j < K; ++j) out->Ap[i][j] = (i == j) ? 1.0f : (float)(exp(t->log_trans[i * K + j]) / sdiag[j]);
    double p2sum = 0.0;
    const double l0_0 = t->log_emit0[0];
    const double d_0 = t->log_emit1[0] - l0_0;
    (void)d_0;
    for (int k = 0; k < K; ++k) {
      const double ratio = sdiag[k] / sdiag[0];
      out->pi2[k] = (float)(exp(t->log_start[k]) / ratio);
      p2sum += (double)out->pi2[k];
      const double l0 = t->log_emit0[k * C];
      out->rl0p[k] = (float)((l0 - l0_0) * log2e + log2(ratio));
      out->rmiss[k] = (float)log2(ratio);
    }
    out->log_s0 = t->log_trans[0];
    out->log_pi2_sum = log(p2sum);
    for (int b = 0; b < t->n_bins; ++b) out->log_s0b[b] = t->log_trans[(size_t)b * K * K];
  }
  out->nb = t->n_bins;
  out->C = C;
}

template <int K>
static void build_vit_tables(const smf_hmm_tables* t, VitTables<K>* out) {
  memset(out, 0, sizeof(*out));
  const int C = t->n_channels;
  for (int k = 0; k < K; ++k) out->log_pi[k] = t->log_start[k];
  for (int b = 0; b < t->n_bins; ++b)
    for (int i = 0; i < K; ++i)
      for (int j = 0; j < K; ++j) out->logA[b][i][j] = t->log_trans[(b * K + i) * K + j];
  for (int k = 0; k < K; ++k)
    for (int c = 0; c < C; ++c) {
      out->l1[k][c] = t->log_emit1[k * C + c];
      out->l0[k][c] = t->log_emit0[k * C + c];
    }
  out->nb = t->n_bins;
  out->C = C;
}


// per-K launchers (fb2_k*.cu, fb_k*.cu, vit_k*.cu)
#define SMF_DECLARE_K(KV)                                                                                      \
  int launch_fb2_k##KV(int chunk, bool stats, int obs_dtype, int out_mode, int minb, const smf_hmm_tables* t, \
                       Fb2Args& a, const Fb2Plan& plan, const DeviceInfo& di, int* nblocks, cudaStream_t s);  \
  int launch_fb_k##KV(bool stats, const smf_hmm_tables* t, FbArgs& a, const FbPlan& plan, const DeviceInfo& di, \
                      int* nblocks, cudaStream_t s);                                                           \
  int launch_vit_k##KV(const smf_hmm_tables* t, const VitArgs& a, const DeviceInfo& di, cudaStream_t s);       \
  int launch_vit_direct_k##KV(const smf_hmm_tables* t, const VitArgs& a, const DeviceInfo& di, cudaStream_t s);
SMF_DECLARE_K(2)
SMF_DECLARE_K(3)
SMF_DECLARE_K(4)
#undef SMF_DECLARE_K
// extra chunk lengths (21, 25, 29) of the K = 2 posterior kernel (fb2x_k2.cu)
int launch_fb2x_k2(int chunk, int obs_dtype, int out_mode, const smf_hmm_tables* t, Fb2Args& a, const Fb2Plan& plan,
                   const DeviceInfo& di, int* nblocks, cudaStream_t s);
int launch_fb2y_k2(int chunk, int obs_dtype, int out_mode, const smf_hmm_tables* t, Fb2Args& a, const Fb2Plan& plan,
                   const DeviceInfo& di, int* nblocks, cudaStream_t s);
int launch_fb2x_k3(int chunk, int obs_dtype, int out_mode, const smf_hmm_tables* t, Fb2Args& a, const Fb2Plan& plan,
                   const DeviceInfo& di, int* nblocks, cudaStream_t s);
// multi-channel variant of the fast-path kernel (fb2m_k2.cu, fb2m_k3.cu): C = 2, 3
int launch_fb2m_k2(int n_channels, bool stats, int obs_dtype, const smf_hmm_tables* t, Fb2Args& a, const Fb2Plan& plan,
                   const DeviceInfo& di, int* nblocks, cudaStream_t s);
int launch_fb2m_k3(int n_channels, bool stats, int obs_dtype, const smf_hmm_tables* t, Fb2Args& a, const Fb2Plan& plan,
                   const DeviceInfo& di, int* nblocks, cudaStream_t s);
// distance-binned variant of the fast-path kernel, posterior decode (fb2b_k2.cu, fb2b_k3.cu)
int launch_fb2b_k2(int obs_dtype, const smf_hmm_tables* t, Fb2Args& a, const Fb2Plan& plan, const DeviceInfo& di,
                   cudaStream_t s);
int launch_fb2b_k3(int obs_dtype, const smf_hmm_tables* t, Fb2Args& a, const Fb2Plan& plan, const DeviceInfo& di,
                   cudaStream_t s);
int launch_fb2b_stats_k2(int obs_dtype, const smf_hmm_tables* t, Fb2Args& a, const Fb2Plan& plan, const DeviceInfo& di,
                         int* nblocks, cudaStream_t s);
struct WalkArgs;
struct SeqArgs;
struct SweepArgs;
struct ShortArgs;
struct PostSweepArgs;
struct SeqFitState;
#define SMF_DECLARE_SEQ(KV)                                                                                          \
  int launch_seq_k##KV(int obs_dtype, int mode, const smf_hmm_tables* t, const SeqArgs& a, const DeviceInfo& di,     \
                       int* nblocks, cudaStream_t s);                                                                \
  int launch_seq_grouped_k##KV(int obs_dtype, int mode, const SeqArgs& a, const DeviceInfo& di, cudaStream_t s);     \
  int launch_seq_mstep_k##KV(const SeqFitState& f, cudaStream_t s);                                                 \
  int launch_seq_fit_loop_k##KV(int obs_dtype, const SeqArgs& a, const SeqFitState& f, unsigned* counter,           \
                                const DeviceInfo& di, cudaStream_t s);
SMF_DECLARE_SEQ(2)
SMF_DECLARE_SEQ(3)
SMF_DECLARE_SEQ(4)
#undef SMF_DECLARE_SEQ
int launch_walk_k2(int obs_dtype, const smf_hmm_tables* t, const WalkArgs& a, cudaStream_t s);
int launch_post_sweep_k2(int obs_dtype, const smf_hmm_tables* t, PostSweepArgs& a, const DeviceInfo& di, cudaStream_t s);
int launch_post_sweep_k3(int obs_dtype, const smf_hmm_tables* t, PostSweepArgs& a, const DeviceInfo& di, cudaStream_t s);
int launch_post_sweep_k4(int obs_dtype, const smf_hmm_tables* t, PostSweepArgs& a, const DeviceInfo& di, cudaStream_t s);
// chunk lengths short_kernel is instantiated for with fully unrolled loops (short_launch.inc)
static inline bool short_chunk_compiled(int K, int obs_dtype, int ch) {
  if (K > 4 || obs_dtype == SMF_OBS_F64) return false;
  if (ch == 9 || ch == 11 || ch == 13 || ch == 15 || ch == 17) return true;
  return K == 2 && (ch == 19 || ch == 21 || ch == 23 || ch == 25 || ch == 33);
}
// chunk lengths the E-step variant of short_kernel is instantiated for
static inline bool short_stats_chunk_compiled(int K, int obs_dtype, int ch) {
  return K <= 4 && obs_dtype != SMF_OBS_F64 && (ch == 9 || ch == 13 || ch == 17);
}
int launch_short_stats_k2(int obs_dtype, const smf_hmm_tables* t, const ShortArgs& a, const DeviceInfo& di,
                          int* nblocks, cudaStream_t s);
int launch_short_stats_k3(int obs_dtype, const smf_hmm_tables* t, const ShortArgs& a, const DeviceInfo& di,
                          int* nblocks, cudaStream_t s);
int launch_short_stats_k4(int obs_dtype, const smf_hmm_tables* t, const ShortArgs& a, const DeviceInfo& di,
                          int* nblocks, cudaStream_t s);
int launch_short_k2(int obs_dtype, const smf_hmm_tables* t, const ShortArgs& a, const DeviceInfo& di, cudaStream_t s);
int launch_short_k3(int obs_dtype, const smf_hmm_tables* t, const ShortArgs& a, const DeviceInfo& di, cudaStream_t s);
int launch_short_k4(int obs_dtype, const smf_hmm_tables* t, const ShortArgs& a, const DeviceInfo& di, cudaStream_t s);
#define SMF_DECLARE_WK(KV)                                                                                   \
  int launch_walk_k##KV(int obs_dtype, const smf_hmm_tables* t, const WalkArgs& a, cudaStream_t s);         \
  int launch_sweep_k##KV(int obs_dtype, const smf_hmm_tables* t, SweepArgs& a, const DeviceInfo& di,        \
                         int* nblocks, cudaStream_t s);                                                     \
  int launch_fb2w_k##KV(int chunk, bool stats, int obs_dtype, int out_mode, const smf_hmm_tables* t,       \
                        Fb2Args& a, const Fb2Plan& plan, const DeviceInfo& di, int* nblocks, cudaStream_t s);
SMF_DECLARE_WK(3)
SMF_DECLARE_WK(4)
#undef SMF_DECLARE_WK

static inline int align_up(int x, int a) { return (x + a - 1) / a * a; }

}  // namespace smf
