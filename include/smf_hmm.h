/*
 * smf_hmm.h -- C-ABI of libsmfhmm.so: the B200 (sm_100a) implementation of the smftools
 * HMM footprinting hot path.
 *
 * The reference hot path sits behind a *Python registry* (no FFI):
 *   /root/reference/src/smftools/hmm/HMM.py:27-54   register_hmm / create_hmm
 *   /root/reference/src/smftools/hmm/HMM.py:398     class BaseHMM (fit / decode / annotate_adata)
 * so there is no pre-existing foreign-function table to copy.  Each entry point below
 * states which reference method body it replaces; the Python host mirror
 * (smftools_b200/hmm.py) is the only caller and binds these through ctypes
 * (INTEGRATION.md shows the binding a reference maintainer would add).
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types.
 *   - data buffers (obs, gamma, states, ...) are DEVICE pointers unless a name ends in
 *     `_host`; the small model tables are HOST pointers (they are copied into the kernel
 *     parameter space, no hidden device allocation).
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), allocates
 *     nothing on the device or the host heap, and returns SMF_OK or a negative error code.
 *   - process state, all of it mutex-/atomic-protected and none of it affecting results: cached device
 *     attributes, per-kernel shared-memory opt-ins and occupancy answers, the two path-selection knobs of
 *     smf_hmm_set_option(), and the SMF_* environment overrides (read once when the library is loaded;
 *     listed in DESIGN.md).  Calls may be issued concurrently from several host threads (one per GPU).
 *   - caller-owned workspaces: ask smf_hmm_workspace_bytes() first.
 */
#ifndef SMF_HMM_H
#define SMF_HMM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SMF_HMM_ABI_VERSION 5

#define SMF_MAX_STATES 4
#define SMF_MAX_CHANNELS 4
#define SMF_MAX_BINS 16

/* error codes (mapped to ValueError / RuntimeError by the Python shim) */
#define SMF_OK 0
#define SMF_ERR_BAD_ARG (-1)      /* NULL / negative size / unsupported K, C, dtype   -> ValueError   */
#define SMF_ERR_TOO_LONG (-2)     /* one read does not fit the on-chip staging        -> ValueError   */
#define SMF_ERR_WORKSPACE (-3)    /* workspace too small                              -> ValueError   */
#define SMF_ERR_CUDA (-4)         /* a CUDA runtime call / launch failed              -> RuntimeError */
#define SMF_ERR_ALIGN (-5)        /* pointer not aligned as required                  -> ValueError   */

/* observation element types (obs_dtype) */
#define SMF_OBS_F32 0 /* float,  NaN = missing  (HMM.py:707-708 nan_to_num + ~isnan)       */
#define SMF_OBS_F64 1 /* double, NaN = missing  (the reference's own host dtype)           */
#define SMF_OBS_U8 2  /* uint8 binary code: 0, 1, anything else = missing                  */

/* operations for smf_hmm_workspace_bytes() */
#define SMF_OP_POSTERIOR 0
#define SMF_OP_VITERBI 1
#define SMF_OP_ESTEP 2
#define SMF_OP_FEATURES 3

/*
 * Model tables in the log domain, exactly the quantities the reference forms at the top
 * of every pass (HMM.py:593-595, 1494-1496, 1773-1775, 2149-2150):
 *   log_start[k]        = log(start[k] + eps)
 *   log_trans[b][i][j]  = log(trans_by_bin[b][i][j] + eps)      (n_bins == 1: trans)
 *   log_emit1[k][c]     = log(emission[k][c] + eps)
 *   log_emit0[k][c]     = log1p(-emission[k][c] + eps)
 * The host mirror computes them with the same torch ops as the reference so the Viterbi
 * kernel (which works on these doubles directly) reproduces its decisions bit for bit.
 */
typedef struct smf_hmm_tables {
  int32_t n_states;   /* K in [2, SMF_MAX_STATES]   */
  int32_t n_channels; /* C in [1, SMF_MAX_CHANNELS] */
  int32_t n_bins;     /* 1 (homogeneous) .. SMF_MAX_BINS (distance-binned, HMM.py:2115-2125) */
  int32_t reserved;
  const double* log_start; /* host [K]          */
  const double* log_trans; /* host [n_bins,K,K] */
  const double* log_emit1; /* host [K,C]        */
  const double* log_emit0; /* host [K,C]        */
} smf_hmm_tables;

int smf_hmm_abi_version(void);
const char* smf_hmm_strerror(int code);

/*
 * Process-wide tuning knobs (not needed for correct results).  Keys:
 *   "walk"   -1 (default) let the library choose between the single-kernel and the two-kernel
 *            forward-backward path from its measured policy; 0 never / 1 whenever possible use the
 *            two-kernel path; 2 additionally always takes its any-length sweep kernels (posterior: all K).
 *            smf_hmm_workspace_bytes() follows the setting.
 *   "short"  -1 (default) posterior decode and E-step of short reads (posterior: up to ~275 / 400 / 200
 *            positions for K = 2 / 3 / 4; E-step: ~450 / 450 / 300) run on the several-reads-per-warp
 *            kernel; 0 never / 1 whenever the read fits it (<= 1056 positions).
 *   "seq"    -1 (default) the E-step of large batches (K = 2: >= 98,304 reads, K = 3, 4: >= 49,152; uint8-coded
 *            rows of a multiple of 16 positions, or float rows of a multiple of 4) runs on the
 *            one-thread-per-read kernels (checkpointed forward pass + one backward pass); 0 never /
 *            1 whenever the batch is eligible.  smf_hmm_workspace_bytes() follows the setting.
 *   "vit"    -1 (default) Viterbi decode of single-channel, homogeneous-transition models with float / uint8-coded
 *            rows of a multiple of 4 (uint8: 16) positions runs on the direct kernel (no shared-memory staging) where
 *            it measured faster: uint8 code, K = 2, K = 4, and K = 3 up to 2,048 positions; 0 never / 1 wherever it
 *            applies.  Both kernels make bit-identical decisions.
 *   "fitloop" 1 (default) smf_hmm_grouped_fit runs the whole Baum-Welch loop of a small fit (every 128-read work item
 *            resident at once, i.e. up to a few hundred items) in ONE cooperative launch with grid barriers between
 *            the E-step and the M-step; 0 always enqueues the forward / backward / M-step kernels per iteration.
 *            Identical results (env SMF_FIT_LOOP sets the initial value).
 *   "generic" 0 (default) / 1: route every forward-backward call to the generic kernel (testing and A/B timing; env
 *            SMF_HMM_FORCE_GENERIC=1 sets the initial value).
 * Returns SMF_ERR_BAD_ARG for an unknown key or value.
 */
int smf_hmm_set_option(const char* key, int value);

/* Bytes of caller-owned device workspace an operation needs (0 = none). */
size_t smf_hmm_workspace_bytes(int op, const smf_hmm_tables* tab, int64_t n_reads, int64_t n_pos);

/* Longest read (positions) the fused on-chip posterior / E-step kernel accepts for this model. */
int64_t smf_hmm_max_positions(const smf_hmm_tables* tab, int want_all_states);

/*
 * Posterior decode.  Replaces BaseHMM._alpha_beta + _forward_backward + argmax
 * (HMM.py:572-629, 718; distance-binned variant 2127-2188).
 *   obs         device [N, L, C] of obs_dtype
 *   bins        device int8 [L-1] distance-bin index per step, or NULL when n_bins == 1
 *   gamma       device float [N, L, K] (gamma_state < 0) or [N, L] holding state `gamma_state`; may be NULL
 *   states      device int8 [N, L] = argmax_k gamma (first maximum), may be NULL
 *   loglik      device double [N] = log P(read), may be NULL
 *   workspace   optional device scratch of smf_hmm_workspace_bytes(SMF_OP_POSTERIOR, ...) bytes (16-byte aligned).
 *               That size is non-zero where a path that needs scratch is the faster one for the batch: the
 *               two-kernel path (per-read boundary walk + local sweeps: boundary vectors) for large batches of
 *               K >= 3 models and for reads beyond the single kernel's reach, the sequential kernels (one thread
 *               per read: alpha checkpoints, about 1 byte per read-position) for large K = 4 batches, short / very
 *               long K = 3 reads and K = 2 reads beyond 21 kb (DESIGN.md sections 4.2, 4.7).  NULL always works
 *               (single-kernel or generic path); a non-NULL workspace smaller than required is SMF_ERR_WORKSPACE.
 *   Models:     single-channel homogeneous (all K), multi-channel (C = 2, 3; K <= 3) and distance-binned (K <= 3)
 *               models with float / uint8-coded rows run on the fast-path kernels; everything else (K = 4 or C = 4
 *               multi-channel, float64 multi-channel rows, models whose emissions / transitions sit at the eps
 *               clamp) on the generic kernel.  Results do not depend on the path beyond fp32 rounding.
 */
int smf_hmm_posterior(const smf_hmm_tables* tab, const void* obs, int obs_dtype, int64_t n_reads,
                      int64_t n_pos, const int8_t* bins, float* gamma, int gamma_state,
                      int8_t* states, double* loglik, void* workspace, size_t workspace_bytes,
                      void* stream);

/*
 * Viterbi decode.  Replaces BaseHMM._viterbi (HMM.py:631-671; binned 2190-2236).
 *   states      device int8 [N, L]
 *   workspace   device scratch of smf_hmm_workspace_bytes(SMF_OP_VITERBI, ...) bytes
 *               (packed back-pointers, one byte per read-position)
 */
int smf_hmm_viterbi(const smf_hmm_tables* tab, const void* obs, int obs_dtype, int64_t n_reads,
                    int64_t n_pos, const int8_t* bins, int8_t* states, void* workspace,
                    size_t workspace_bytes, void* stream);

/*
 * Baum-Welch E-step.  Replaces the body of fit_em between the forward-backward pass and the
 * M-step (HMM.py:1566-1596; multi 1849-1882; binned 2289-2320).
 *   stats       device double [smf_hmm_stats_len(tab)] =
 *               [ start_acc K | trans_acc n_bins*K*K | emit_num K*C | emit_den K*C | loglik 1 ]
 *               (overwritten; reduced in a fixed order => run-to-run deterministic)
 *   workspace   device scratch of smf_hmm_workspace_bytes(SMF_OP_ESTEP, ...) bytes
 */
int64_t smf_hmm_stats_len(const smf_hmm_tables* tab);
int smf_hmm_estep(const smf_hmm_tables* tab, const void* obs, int obs_dtype, int64_t n_reads,
                  int64_t n_pos, const int8_t* bins, double* stats, void* workspace,
                  size_t workspace_bytes, void* stream);

/*
 * ---- grouped launches: many (parameters, N_g, L_g) groups per call ------------------------------------
 *
 * The reference fits and applies one HMM per (reference x methylation type [x sample / barcode]) group, one
 * after the other (cli/hmm_adata.py:488-524 HMMTrainer.fit_or_load per task; hmm/fit_plan.py:12-14 group
 * fields; hmm_max_fit_reads = 1000 reads per fit).  A group that small cannot fill the GPU, and an EM
 * iteration per group costs a host round trip.  These entry points take a TABLE of groups -- each with its own
 * observation matrix, read count, read length and parameters -- and run all of them in one launch sequence:
 * one thread per read (checkpointed forward pass + one backward pass), work items of 128
 * reads dealt over the SMs regardless of group boundaries.  smf_hmm_grouped_fit additionally keeps the whole
 * Baum-Welch loop on the device (E-step statistics -> per-group reduction -> M-step + stop rule + table
 * rebuild, HMM.py:1566-1630), so max_iter iterations of every group are enqueued without a host round trip.
 * Single-channel, homogeneous-transition models (arch "single") only.
 */
typedef struct smf_hmm_group {
  const void* obs;     /* device [n_reads, row_stride] of obs_dtype (SMF_OBS_U8 or SMF_OBS_F32), 16-byte aligned   */
  int64_t n_reads;
  int32_t n_pos;       /* L_g: positions per read of this group                                                   */
  int32_t row_stride;  /* elements between rows, >= n_pos; a multiple of 16 (uint8) / 4 (float); the pad is not read
                          as data but must be addressable                                                           */
  float* gamma;        /* posterior: device [n_reads, out_stride, K] (gamma_state < 0) or [n_reads, out_stride]; NULL ok */
  int8_t* states;      /* posterior: device [n_reads, out_stride]; NULL ok                                          */
  double* loglik;      /* device [n_reads] per-read log-likelihood; NULL ok                                         */
  int32_t gamma_state;
  int32_t out_stride;  /* positions between output rows, >= n_pos, a multiple of 4 (columns >= n_pos are scratch)   */
} smf_hmm_group;

#define SMF_FIT_RAN_ALL 0     /* the group ran max_iter iterations                                                 */
#define SMF_FIT_CONVERGED 1   /* |dLL| < tol * max(1, |LL|) (HMM.py:1616-1628) after n_iter iterations             */
#define SMF_FIT_HANDED_BACK 2 /* after n_iter iterations the parameters left the dynamic range these kernels cover
                                 with their fixed rescale cadence: params / hist are valid up to there and the caller
                                 continues the group with smf_hmm_estep (which routes such models itself)          */

/* fit != 0: workspace of smf_hmm_grouped_fit, else of smf_hmm_grouped_posterior.  0 on invalid arguments. */
size_t smf_hmm_grouped_workspace_bytes(int n_states, const smf_hmm_group* groups_host, int n_groups, int fit);

/*
 * Posterior decode of every group with its own model (HMM.py:572-629, 718 per group).  tabs_host[g] are the
 * log tables of group g (host pointers, as everywhere).  Outputs per group as in smf_hmm_posterior.
 * SMF_ERR_BAD_ARG when a model is not single-channel / homogeneous or needs a finer rescale cadence than
 * these kernels use (decode such a group with smf_hmm_posterior).
 */
int smf_hmm_grouped_posterior(int obs_dtype, const smf_hmm_group* groups_host, int n_groups,
                              const smf_hmm_tables* tabs_host, void* workspace, size_t workspace_bytes,
                              void* stream);

/*
 * Baum-Welch fit of every group, entirely on the device (replaces n_groups x the fit_em loop, HMM.py:1518-1630).
 *   params     device double [n_groups, K + K*K + K] = start | trans | emission per group: initial values in,
 *              fitted values out (normalised and clamped like HMM.py:492-511 after every update)
 *   update_mask bit 0 start, bit 1 transitions, bit 2 emissions (the reference's update_* flags)
 *   hist       device double [n_groups, max_iter]: total log-likelihood before each M-step
 *   n_iter     device int32 [n_groups]: iterations done;  status: device int32 [n_groups] SMF_FIT_*
 * Everything is enqueued on `stream`; read params / hist / n_iter / status back after synchronising it.
 */
int smf_hmm_grouped_fit(int n_states, int obs_dtype, const smf_hmm_group* groups_host, int n_groups,
                        double* params, int max_iter, double tol, double eps, int update_mask, double* hist,
                        int32_t* n_iter, int32_t* status, void* workspace, size_t workspace_bytes, void* stream);

/*
 * The same fit in steps, for groups whose reads are SHARDED over several processes (one per GPU): every process
 * passes the same group table (same order, same n_pos) with its own shard of each group's reads (n_reads may be 0),
 * and between the E-step and the M-step of an iteration the caller sum-all-reduces `stats` over the processes
 * (NCCL on the same stream: nothing waits on the host).  The M-step is deterministic, so every process ends each
 * iteration with bit-identical parameters, stop decisions and kernel tables.
 *   SMF_FIT_PHASE_BEGIN   once: uploads the group table, builds the tables of the initial parameters
 *   SMF_FIT_PHASE_ESTEP   iteration `iter`: statistics of this process's reads -> stats (device double [n_groups, S],
 *                         S = K + K*K + 2*K + 1; zeros for groups that have stopped)
 *   SMF_FIT_PHASE_MSTEP   iteration `iter`: M-step, stop rule and table rebuild of every active group from `stats`
 * Same workspace (smf_hmm_grouped_workspace_bytes(..., fit = 1)) and same params / hist / n_iter / status buffers in
 * every call of one fit.
 */
#define SMF_FIT_PHASE_ALL 0
#define SMF_FIT_PHASE_BEGIN 1
#define SMF_FIT_PHASE_ESTEP 2
#define SMF_FIT_PHASE_MSTEP 3
int smf_hmm_grouped_fit_step(int phase, int n_states, int obs_dtype, const smf_hmm_group* groups_host, int n_groups,
                             double* params, int iter, int max_iter, double tol, double eps, int update_mask,
                             double* hist, int32_t* n_iter, int32_t* status, double* stats, void* workspace,
                             size_t workspace_bytes, void* stream);

/*
 * Footprint / feature calling on decoded reads.  Replaces the per-read Python loops of
 * BaseHMM.annotate_adata (HMM.py:1318-1370) for one feature group:
 *   membership[n,t] = states[n,t] == target            (post == NULL, Viterbi mode)
 *                   = (double)post[n,t] >= threshold   (post != NULL, marginal mode; fp64 comparison as in
 *                                                       the reference, for any threshold value)
 *   every run [s,e) of membership gets genomic length coords[e-1]-coords[s]+1, the first
 *   size class with lo <= len < hi, and is span-filled into the full grid columns
 *   [left[s], right[e-1]) (searchsorted left/right of the site coordinate in the full grid,
 *   precomputed on the host: HMM.py:1344-1345).
 * The kernel requires the site coordinates to be strictly increasing and present in the
 * (sorted) full grid, which makes the span-filled runs pairwise disjoint and non-abutting, so
 * the run length of a cell is the same in the all_{group}_features layer and in its class
 * layer (HMM.py:1361-1370); the host mirror checks this and otherwise uses its own slow path.
 * Outputs (all device, row-major [N, V]):
 *   class_out   uint8: 0 = no feature, c+1 = size class c       (dense "which layer is 1"); may be NULL
 *               when only the compacted interval records are wanted (then run_len must be NULL too)
 *   run_len     int32: run length in grid columns of the run covering the cell, else 0; may be NULL
 *   intervals   int32 [max_intervals,4] = (read, grid_start, grid_end, class) compacted runs; may be NULL
 *   n_intervals device uint64 counter (zeroed by the caller); total runs found, may exceed max_intervals
 */
int smf_hmm_call_features(const int8_t* states, const float* post, int target_state,
                          double threshold, int64_t n_reads, int64_t n_pos,
                          const int64_t* coords, const int32_t* grid_left,
                          const int32_t* grid_right, int64_t n_grid, const double* class_lo_host,
                          const double* class_hi_host, int n_classes, uint8_t* class_out,
                          int32_t* run_len, int32_t* intervals, int64_t max_intervals,
                          unsigned long long* n_intervals, void* stream);

/*
 * Merge adjacent runs of a binary layer when the coordinate gap is <= distance_threshold and
 * (optionally) re-bin the merged runs into size classes.  Replaces
 * BaseHMM.merge_intervals_to_new_layer + write_size_class_layers_from_binary
 * (HMM.py:1006-1123).  layer is uint8 [N, V] (non-zero = member); coords int64 [V].
 *   merged      uint8 [N, V]; merged_len int32 [N, V]
 *   class_out / len_cls as in smf_hmm_call_features (may be NULL when n_classes == 0)
 */
int smf_hmm_merge_runs(const uint8_t* layer, int64_t n_reads, int64_t n_grid,
                       const int64_t* coords, int coords_are_ints, int64_t distance_threshold,
                       uint8_t* merged, int32_t* merged_len, const double* class_lo_host,
                       const double* class_hi_host, int n_classes, uint8_t* class_out,
                       int32_t* len_cls, void* stream);

/*
 * Fused merge chain: what tools/partitioned_hmm.py:78-108 (_apply_merges) does to one aggregate
 * feature layer -- merge_intervals_to_new_layer (HMM.py:1006-1067), write_size_class_layers_from_binary
 * on the merged layer (HMM.py:1069-1123) and mask_layers_outside_read_span (HMM.py:148-231) over all
 * layers of the chain -- in ONE pass over the layer, with a packed result the caller expands lazily:
 *   code_out  uint8 [N, V]: bits 0-3 = size class of the merged run + 1 (0 = no class matched),
 *             0x40 = cell belongs to a merged run, 0x80 = cell inside the read span (not NaN-masked)
 *   run_len   int32 [N, V]: length of the merged run in grid columns, 0 outside runs
 * so that  {layer}_merged = (code & 0x40) != 0,  its _lengths = run_len,  class layer c =
 * (code & 15) == c + 1,  its _lengths = run_len there, and every layer is NaN where (code & 0x80) == 0.
 * Read span: `covered` uint8 [N, V] (non-zero = keep) when given, else a cell is kept iff
 * int(read_start[n]) <= span_coords[v] <= int(read_end[n]) (all kept when either bound is not finite).
 * n_classes <= 15.
 */
int smf_hmm_merge_chain(const uint8_t* layer, int64_t n_reads, int64_t n_grid, const int64_t* coords,
                        int coords_are_ints, int64_t distance_threshold, const double* class_lo_host,
                        const double* class_hi_host, int n_classes, const uint8_t* covered,
                        const int64_t* span_coords, const double* read_start, const double* read_end,
                        uint8_t* code_out, int32_t* run_len, void* stream);

/*
 * Read-span masking.  Replaces mask_layers_outside_read_span (HMM.py:148-231):
 *   valid_out[n,v] = covered[n,v] != 0                          (covered != NULL)
 *                  = !(coords[v] < start[n] || coords[v] > end[n])  (non-finite start/end => all valid)
 */
int smf_hmm_span_mask(const uint8_t* covered, const int64_t* coords, const double* read_start,
                      const double* read_end, int64_t n_reads, int64_t n_grid, uint8_t* valid_out,
                      void* stream);

/*
 * ---- boundary to the reference's host-side array contract -------------------------------------------
 *
 * BaseHMM.decode returns int64 states and posteriors in the model dtype (float64 by default) as host
 * numpy arrays (HMM.py:718-720); annotate_adata leaves float64 layers with NaN outside the read span
 * (HMM.py:1256-1377, 148-231).  The kernels above produce int8 / float / packed codes; the entry points
 * below convert on the device so that the wide result crosses the host link exactly once, straight into
 * page-locked destination arrays.
 */

/* int8 states -> int64 and float posteriors -> double, one launch (either pair may be NULL / 0 elements).
 * states8 must be 4-byte aligned, gamma32 / the outputs 16-byte aligned. */
int smf_hmm_widen_decode(const int8_t* states8, int64_t* states64, int64_t n_states_elems,
                         const float* gamma32, double* gamma64, int64_t n_gamma_elems, void* stream);

/* layer kinds of smf_hmm_expand_layers() */
#define SMF_LAYER_STATES 0    /* {prefix}_states: state at site columns, -1 elsewhere     (HMM.py:1256-1262) */
#define SMF_LAYER_POSTERIOR 1 /* {prefix}_posterior_<state>: posterior at site columns, 0 elsewhere (:1264-1273) */
#define SMF_LAYER_ANY 2       /* all_{group}_features (class_out > 0)  /  {layer}_merged (code & 0x40)      */
#define SMF_LAYER_ANY_LEN 3   /* its _lengths layer                                                          */
#define SMF_LAYER_CLASS 4     /* {prefix}_{feature}: size class `cls` of the group                           */
#define SMF_LAYER_CLASS_LEN 5 /* its _lengths layer                                                          */

typedef struct smf_hmm_layer_group {
  const uint8_t* cls_or_code; /* device [N, V]: class_out of smf_hmm_call_features, or code_out of smf_hmm_merge_chain */
  const int32_t* run_len;     /* device [N, V] run lengths (may be NULL when no *_LEN layer is requested)             */
  int32_t packed;             /* 1: cls_or_code is a merge-chain code byte (class bits 0-3, member 0x40, valid 0x80)   */
  int32_t reserved;
} smf_hmm_layer_group;

typedef struct smf_hmm_layer_desc {
  void* out;       /* device [N, V], double (out_f32 == 0) or float (out_f32 == 1) */
  int32_t kind;    /* SMF_LAYER_* */
  int32_t group;   /* index into groups for kinds >= SMF_LAYER_ANY */
  int32_t cls;     /* size class index for SMF_LAYER_CLASS / _CLASS_LEN */
  int32_t out_f32;
} smf_hmm_layer_desc;

/*
 * Expand the packed per-chunk annotation results into the final dense layers in ONE pass: every cell of
 * every requested layer is written once, already in its final dtype and NaN-masked where `valid` is 0 (or,
 * for packed groups, where the code byte's 0x80 bit is clear).  Replaces the host-side layer fills of
 * annotate_adata (HMM.py:1256-1273, 1348-1370), the dtype cast + NaN masking of
 * mask_layers_outside_read_span (HMM.py:227-231) and the layer writes of _apply_merges
 * (tools/partitioned_hmm.py:78-108).
 *   states / post      decoded states int8 [N, L] and one posterior column: element (n, t) is read at
 *                      post[(n * L + t) * post_stride]; either may be NULL when no layer needs it
 *   site_of_col        device int32 [V]: site index of a grid column, -1 where the column is not a site
 *   valid              device uint8 [N, V] from smf_hmm_span_mask, or NULL (keep everything)
 *   groups / layers    HOST arrays (copied into the kernel parameters); <= 8 groups, <= 64 layers
 */
int smf_hmm_expand_layers(int64_t n_reads, int64_t n_grid, int64_t n_pos, const int8_t* states,
                          const float* post, int64_t post_stride, const int32_t* site_of_col,
                          const uint8_t* valid, const smf_hmm_layer_group* groups_host, int n_groups,
                          const smf_hmm_layer_desc* layers_host, int n_layers, void* stream);

/*
 * Model-input builder (the step before the path).  Replaces build_single_channel /
 * build_multi_channel_union (cli/hmm_adata.py:251-319: the site columns of the per-position matrix, on the
 * union grid for several methylation channels, NaN where a channel has no site) and
 * _mask_uncovered_model_input (tools/partitioned_hmm.py:170-188: missing where covered_base_mask is False),
 * fused, producing the kernels' observation layout directly:
 *   layer         device [N, V] of layer_dtype (SMF_OBS_F32 / F64 with NaN = missing, or SMF_OBS_U8 code)
 *   cols          device int32 [L, C]: grid column of (position, channel), -1 = no site of that channel
 *   covered       device uint8 [N, n_cov_grid] or NULL; covered_cols device int32 [L] = its column for position l
 *   out           device [N, L, C]: uint8 code 0 / 1 / 255 (out_dtype SMF_OBS_U8) or float, NaN = missing
 *   flag          device int (zeroed by the caller), may be NULL: set to 1 when the output type cannot carry a
 *                 value exactly -- a call other than 0 / 1 / NaN for the uint8 code (the host mirror then
 *                 asks for float output), or a float64 value that float32 does not represent
 */
int smf_hmm_build_obs(const void* layer, int layer_dtype, int64_t n_reads, int64_t n_grid,
                      const int32_t* cols, int64_t n_pos, int n_channels, const uint8_t* covered,
                      int64_t n_cov_grid, const int32_t* covered_cols, void* out, int out_dtype, int* flag,
                      void* stream);

/*
 * ---- layer reductions from the packed results (what the reference's consumers re-derive from the dense layers)
 *
 * Membership of a cell: class_out > 0 (or == select_class + 1), resp. for packed merge-chain codes the 0x40 bit
 * (or (code & 15) == select_class + 1); a cell counts only where `valid` != 0 (NULL = everywhere) and, for packed
 * codes, where the 0x80 bit is set -- exactly the cells that are finite in the reference's float64 layers.
 */

/* Per-column sums over the reads: colsum[0][v] = cells of the aggregate layer that are > 0, colsum[c + 1][v] = of
 * size class c, colvalid[v] = finite cells (outputs are zeroed by the call).  Their totals are the "modified" /
 * "valid" counts of tools/partitioned_hmm.py:519-529; colsum / colvalid is np.nanmean(axis=0) of
 * hmm/call_hmm_peaks.py:151.  n_classes <= 15. */
int smf_hmm_layer_colsums(const uint8_t* cls_or_code, int packed, const uint8_t* valid, int64_t n_reads,
                          int64_t n_grid, int n_classes, uint64_t* colsum /*[n_classes + 1, V]*/,
                          uint64_t* colvalid /*[V]*/, void* stream);

/* means[v] = colsum[v] / colvalid[v] (0 where no read covers v: nan_to_num) and metric = the 'same'-mode rolling
 * mean np.convolve(means, ones(w) / w, "same") (hmm/call_hmm_peaks.py:151-158).  Like np.convolve, `metric` has
 * max(n_grid, w) elements when w > 1 (a grid narrower than the window yields w values), else n_grid. */
int smf_hmm_column_metric(const uint64_t* colsum, const uint64_t* colvalid, int64_t n_grid, int rolling_window,
                          double* means, double* metric, void* stream);

/* Runs-per-read and run-size histograms of one binary layer (tools/partitioned_hmm.py:572-585, 618-630):
 * count_hist[k] = reads with k runs (k <= V / 2 + 1), size_hist[s] = runs of s grid columns (s <= V); zeroed by the
 * call.  select_class = -1 for the aggregate / merged layer. */
int smf_hmm_run_histograms(const uint8_t* cls_or_code, int packed, int select_class, const uint8_t* valid,
                           int64_t n_reads, int64_t n_grid, uint64_t* count_hist /*[V / 2 + 2]*/,
                           uint64_t* size_hist /*[V + 1]*/, void* stream);

/* Page-lock / release caller-owned host memory (cudaHostRegister, portable) so that copies into it run
 * at link speed and asynchronously; plain async copy on `stream`. */
#define SMF_COPY_H2D 0
#define SMF_COPY_D2H 1
#define SMF_COPY_D2D 2
int smf_hmm_host_register(void* host_ptr, size_t bytes);
int smf_hmm_host_unregister(void* host_ptr);
int smf_hmm_host_is_pinned(const void* host_ptr); /* 1 when the address is page-locked (registered or cudaHostAlloc) */
int smf_hmm_copy_async(void* dst, const void* src, size_t bytes, int kind, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SMF_HMM_H */
