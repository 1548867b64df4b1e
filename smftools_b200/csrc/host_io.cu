// Boundary kernels between the packed device results and the reference's host-side array contract:
// widening of decode outputs (int8 -> int64, float -> double: BaseHMM.decode returns int64 states and
// model-dtype posteriors, HMM.py:718-720), expansion of the packed annotation results into the final
// float64 / NaN-masked AnnData layers (HMM.py:1256-1377 + mask_layers_outside_read_span :148-231), and
// thin wrappers around page-locking / async copies so the Python host mirror needs no second CUDA binding.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "smf_hmm.h"

namespace smf {

constexpr int kMaxLayerGroups = 8;
constexpr int kMaxLayers = 64;

struct ExpandGroup {
  const uint8_t* cls;  // class_out (plain) or merge-chain code byte (packed)
  const int32_t* len;  // run length per cell
  int packed;
};
struct ExpandLayer {
  void* out;
  int kind, group, cls, f32;
};
struct ExpandArgs {
  long long N;
  int V, L;
  const int8_t* states;    // [N, L] or nullptr
  const float* post;       // element (n, t) at post[(n * L + t) * post_stride]
  long long post_stride;
  const int32_t* site_of_col;  // [V] site index of a grid column, -1 = not a site
  const uint8_t* valid;        // [N, V] or nullptr (all cells kept unless a packed group says otherwise)
  int n_groups, n_layers;
  ExpandGroup g[kMaxLayerGroups];
  ExpandLayer l[kMaxLayers];
};

__device__ __forceinline__ void put_cell(const ExpandLayer& ly, long long idx, double v, bool keep) {
  if (ly.f32) {
    reinterpret_cast<float*>(ly.out)[idx] = keep ? (float)v : __int_as_float(0x7fc00000);
  } else {
    reinterpret_cast<double*>(ly.out)[idx] = keep ? v : __longlong_as_double(0x7ff8000000000000LL);
  }
}

// One thread per grid cell; every layer of the cell is produced from one read of the packed inputs.
// Pure streaming: (2 + 5 * groups) bytes in, 8 bytes out per cell and layer, all accesses coalesced.
__global__ void __launch_bounds__(256) expand_layers_kernel(const __grid_constant__ ExpandArgs a) {
  const long long cells = a.N * (long long)a.V;
  for (long long idx = blockIdx.x * 256LL + threadIdx.x; idx < cells; idx += gridDim.x * 256LL) {
    const long long n = idx / a.V;
    const int v = (int)(idx - n * a.V);
    bool keep = a.valid ? (a.valid[idx] != 0) : true;
    // validity bits of packed (merge-chain) groups apply to every layer of the cell
    for (int gi = 0; gi < a.n_groups; ++gi)
      if (a.g[gi].packed) keep = keep && (a.g[gi].cls[idx] & 0x80u);
    int cur_g = -1, cls = 0, any = 0, len_any = 0, len_raw = 0;  // group values, loaded when the group changes
    int site = -1;
    if (a.site_of_col) site = a.site_of_col[v];
    for (int li = 0; li < a.n_layers; ++li) {
      const ExpandLayer& ly = a.l[li];
      if (ly.kind >= SMF_LAYER_ANY && ly.group != cur_g) {
        cur_g = ly.group;
        const unsigned c = a.g[cur_g].cls[idx];
        len_raw = a.g[cur_g].len ? a.g[cur_g].len[idx] : 0;
        if (a.g[cur_g].packed) {
          cls = (int)(c & 15u);
          any = (c & 0x40u) ? 1 : 0;
          len_any = len_raw;  // the merged layer's lengths are the run lengths themselves
        } else {
          cls = (int)c;
          any = c > 0u ? 1 : 0;
          len_any = c > 0u ? len_raw : 0;
        }
      }
      double val = 0.0;
      switch (ly.kind) {
        case SMF_LAYER_STATES: val = site >= 0 ? (double)a.states[n * a.L + site] : -1.0; break;
        case SMF_LAYER_POSTERIOR: val = site >= 0 ? (double)a.post[(n * a.L + site) * a.post_stride] : 0.0; break;
        case SMF_LAYER_ANY: val = (double)any; break;
        case SMF_LAYER_ANY_LEN: val = (double)len_any; break;
        case SMF_LAYER_CLASS: val = cls == ly.cls + 1 ? 1.0 : 0.0; break;
        case SMF_LAYER_CLASS_LEN: val = cls == ly.cls + 1 ? (double)len_raw : 0.0; break;
      }
      put_cell(ly, idx, val, keep);
    }
  }
}

// int8 states -> int64, float posteriors -> double (two independent streams in one launch).
__global__ void __launch_bounds__(256) widen_decode_kernel(const int8_t* __restrict__ s8, long long* __restrict__ s64,
                                                           long long ns, const float* __restrict__ g32,
                                                           double* __restrict__ g64, long long ng) {
  const long long stride = gridDim.x * 256LL;
  const long long t0 = blockIdx.x * 256LL + threadIdx.x;
  // four elements per thread and step: 32-bit loads of states, 128-bit loads of posteriors
  if (s8 != nullptr) {
    const long long n4 = ns >> 2;
    const int* s32 = reinterpret_cast<const int*>(s8);
    longlong2* o2 = reinterpret_cast<longlong2*>(s64);
    for (long long i = t0; i < n4; i += stride) {
      const int w = __ldg(s32 + i);
      longlong2 a, b;
      a.x = (long long)(signed char)(w & 0xff);
      a.y = (long long)(signed char)((w >> 8) & 0xff);
      b.x = (long long)(signed char)((w >> 16) & 0xff);
      b.y = (long long)(signed char)((w >> 24) & 0xff);
      o2[2 * i] = a;
      o2[2 * i + 1] = b;
    }
    for (long long i = (n4 << 2) + t0; i < ns; i += stride) s64[i] = (long long)s8[i];
  }
  if (g32 != nullptr) {
    const long long n4 = ng >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(g32);
    double2* o2 = reinterpret_cast<double2*>(g64);
    for (long long i = t0; i < n4; i += stride) {
      const float4 x = __ldg(g4 + i);
      o2[2 * i] = make_double2((double)x.x, (double)x.y);
      o2[2 * i + 1] = make_double2((double)x.z, (double)x.w);
    }
    for (long long i = (n4 << 2) + t0; i < ng; i += stride) g64[i] = (double)g32[i];
  }
}


// ---- model-input builder (N2): gather the site columns of a per-position layer into the kernels' -----
// observation layout, apply the coverage mask, and code binary calls as uint8 (0 / 1 / 255 = missing).
struct BuildArgs {
  const void* layer;   // [N, V]
  int layer_dtype;     // SMF_OBS_F32 / F64 (NaN = missing) / U8 (0, 1, other = missing)
  long long N;
  int V, L, C;
  const int32_t* cols;          // [L, C] grid column of (position, channel), -1 = the channel has no site there
  const uint8_t* covered;       // [N, VC] or nullptr
  int VC;                       // columns of the coverage matrix
  const int32_t* covered_cols;  // [L] grid column whose coverage applies to position l
  void* out;                    // [N, L, C] uint8 code or float
  int out_dtype;                // SMF_OBS_U8 / SMF_OBS_F32
  int* nonbinary;               // set to 1 when a value other than 0 / 1 / NaN was met (uint8 output only)
};

__global__ void __launch_bounds__(256) build_obs_kernel(const __grid_constant__ BuildArgs a) {
  const long long per_read = (long long)a.L * a.C;
  const long long total = a.N * per_read;
  bool odd = false;
  for (long long idx = blockIdx.x * 256LL + threadIdx.x; idx < total; idx += gridDim.x * 256LL) {
    const long long n = idx / per_read;
    const int lc = (int)(idx - n * per_read);
    const int l = lc / a.C;
    const int col = a.cols[lc];
    float v = __int_as_float(0x7fc00000);
    bool exact = true;  // value representable in the output without changing its meaning
    if (col >= 0 && !(a.covered != nullptr && a.covered[n * a.VC + a.covered_cols[l]] == 0)) {
      const long long src = n * a.V + col;
      if (a.layer_dtype == SMF_OBS_F32) {
        v = reinterpret_cast<const float*>(a.layer)[src];
      } else if (a.layer_dtype == SMF_OBS_F64) {
        const double d = reinterpret_cast<const double*>(a.layer)[src];
        v = (float)d;
        exact = (d != d) || ((double)v == d);
      } else {
        const unsigned b = reinterpret_cast<const unsigned char*>(a.layer)[src];
        v = b == 0u ? 0.0f : (b == 1u ? 1.0f : __int_as_float(0x7fc00000));
      }
    }
    if (a.out_dtype == SMF_OBS_U8) {
      unsigned char code = 255;
      if (v == 0.0f) code = 0;
      else if (v == 1.0f) code = 1;
      else if (v == v) odd = true;  // a probabilistic / non-binary call: the byte code cannot carry it
      reinterpret_cast<unsigned char*>(a.out)[idx] = code;
    } else {
      if (!exact) odd = true;  // float64 input that float32 cannot represent
      reinterpret_cast<float*>(a.out)[idx] = v;
    }
  }
  if (odd && a.nonbinary != nullptr) *a.nonbinary = 1;
}

static int stream_grid() {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms * 8;
}

}  // namespace smf

using namespace smf;

extern "C" {

int smf_hmm_host_register(void* ptr, size_t bytes) {
  if (ptr == nullptr || bytes == 0) return SMF_ERR_BAD_ARG;
  return cudaHostRegister(ptr, bytes, cudaHostRegisterPortable) == cudaSuccess ? SMF_OK : (cudaGetLastError(), SMF_ERR_CUDA);
}

int smf_hmm_host_unregister(void* ptr) {
  if (ptr == nullptr) return SMF_ERR_BAD_ARG;
  return cudaHostUnregister(ptr) == cudaSuccess ? SMF_OK : (cudaGetLastError(), SMF_ERR_CUDA);
}

int smf_hmm_host_is_pinned(const void* ptr) {
  if (ptr == nullptr) return 0;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return at.type == cudaMemoryTypeHost ? 1 : 0;
}

int smf_hmm_copy_async(void* dst, const void* src, size_t bytes, int kind, void* stream) {
  if (bytes == 0) return SMF_OK;
  if (dst == nullptr || src == nullptr) return SMF_ERR_BAD_ARG;
  cudaMemcpyKind k;
  switch (kind) {
    case SMF_COPY_H2D: k = cudaMemcpyHostToDevice; break;
    case SMF_COPY_D2H: k = cudaMemcpyDeviceToHost; break;
    case SMF_COPY_D2D: k = cudaMemcpyDeviceToDevice; break;
    default: return SMF_ERR_BAD_ARG;
  }
  return cudaMemcpyAsync(dst, src, bytes, k, (cudaStream_t)stream) == cudaSuccess ? SMF_OK : (cudaGetLastError(), SMF_ERR_CUDA);
}

int smf_hmm_widen_decode(const int8_t* states8, int64_t* states64, int64_t n_states_elems, const float* gamma32,
                         double* gamma64, int64_t n_gamma_elems, void* stream) {
  if (n_states_elems < 0 || n_gamma_elems < 0) return SMF_ERR_BAD_ARG;
  const bool do_s = states8 != nullptr && n_states_elems > 0;
  const bool do_g = gamma32 != nullptr && n_gamma_elems > 0;
  if (do_s && states64 == nullptr) return SMF_ERR_BAD_ARG;
  if (do_g && gamma64 == nullptr) return SMF_ERR_BAD_ARG;
  if (!do_s && !do_g) return SMF_OK;
  if (do_s && ((reinterpret_cast<uintptr_t>(states8) & 3u) || (reinterpret_cast<uintptr_t>(states64) & 15u)))
    return SMF_ERR_ALIGN;
  if (do_g && ((reinterpret_cast<uintptr_t>(gamma32) & 15u) || (reinterpret_cast<uintptr_t>(gamma64) & 15u)))
    return SMF_ERR_ALIGN;
  widen_decode_kernel<<<stream_grid(), 256, 0, (cudaStream_t)stream>>>(
      do_s ? states8 : nullptr, reinterpret_cast<long long*>(states64), n_states_elems, do_g ? gamma32 : nullptr,
      gamma64, n_gamma_elems);
  return cudaGetLastError() == cudaSuccess ? SMF_OK : SMF_ERR_CUDA;
}

int smf_hmm_expand_layers(int64_t n_reads, int64_t n_grid, int64_t n_pos, const int8_t* states, const float* post,
                          int64_t post_stride, const int32_t* site_of_col, const uint8_t* valid,
                          const smf_hmm_layer_group* groups_host, int n_groups,
                          const smf_hmm_layer_desc* layers_host, int n_layers, void* stream) {
  if (n_reads < 0 || n_grid < 0 || n_pos < 0) return SMF_ERR_BAD_ARG;
  if (n_groups < 0 || n_groups > kMaxLayerGroups || n_layers < 0 || n_layers > kMaxLayers) return SMF_ERR_BAD_ARG;
  if (n_reads == 0 || n_grid == 0 || n_layers == 0) return SMF_OK;
  if (n_grid > 0x3fffffff || n_pos > 0x3fffffff) return SMF_ERR_TOO_LONG;
  if ((n_groups > 0 && groups_host == nullptr) || layers_host == nullptr) return SMF_ERR_BAD_ARG;
  ExpandArgs a;
  memset(&a, 0, sizeof(a));
  a.N = n_reads;
  a.V = (int)n_grid;
  a.L = (int)n_pos;
  a.states = states;
  a.post = post;
  a.post_stride = post_stride > 0 ? post_stride : 1;
  a.site_of_col = site_of_col;
  a.valid = valid;
  a.n_groups = n_groups;
  a.n_layers = n_layers;
  for (int g = 0; g < n_groups; ++g) {
    if (groups_host[g].cls_or_code == nullptr) return SMF_ERR_BAD_ARG;
    a.g[g].cls = groups_host[g].cls_or_code;
    a.g[g].len = groups_host[g].run_len;
    a.g[g].packed = groups_host[g].packed;
  }
  for (int i = 0; i < n_layers; ++i) {
    const smf_hmm_layer_desc& d = layers_host[i];
    if (d.out == nullptr) return SMF_ERR_BAD_ARG;
    switch (d.kind) {
      case SMF_LAYER_STATES:
        if (states == nullptr || site_of_col == nullptr) return SMF_ERR_BAD_ARG;
        break;
      case SMF_LAYER_POSTERIOR:
        if (post == nullptr || site_of_col == nullptr) return SMF_ERR_BAD_ARG;
        break;
      case SMF_LAYER_ANY:
      case SMF_LAYER_ANY_LEN:
      case SMF_LAYER_CLASS:
      case SMF_LAYER_CLASS_LEN:
        if (d.group < 0 || d.group >= n_groups) return SMF_ERR_BAD_ARG;
        if ((d.kind == SMF_LAYER_ANY_LEN || d.kind == SMF_LAYER_CLASS_LEN) && groups_host[d.group].run_len == nullptr)
          return SMF_ERR_BAD_ARG;
        break;
      default: return SMF_ERR_BAD_ARG;
    }
    a.l[i].out = d.out;
    a.l[i].kind = d.kind;
    a.l[i].group = d.group;
    a.l[i].cls = d.cls;
    a.l[i].f32 = d.out_f32;
  }
  const long long cells = n_reads * n_grid;
  long long blocks = (cells + 255) / 256;
  const int cap = stream_grid();
  if (blocks > cap) blocks = cap;
  expand_layers_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(a);
  return cudaGetLastError() == cudaSuccess ? SMF_OK : SMF_ERR_CUDA;
}

int smf_hmm_build_obs(const void* layer, int layer_dtype, int64_t n_reads, int64_t n_grid, const int32_t* cols,
                      int64_t n_pos, int n_channels, const uint8_t* covered, int64_t n_cov_grid,
                      const int32_t* covered_cols, void* out, int out_dtype, int* flag, void* stream) {
  if (n_reads < 0 || n_grid < 0 || n_pos < 0 || n_channels < 1 || n_channels > SMF_MAX_CHANNELS) return SMF_ERR_BAD_ARG;
  if (n_reads == 0 || n_pos == 0) return SMF_OK;
  if (!layer || !cols || !out) return SMF_ERR_BAD_ARG;
  if (layer_dtype < SMF_OBS_F32 || layer_dtype > SMF_OBS_U8) return SMF_ERR_BAD_ARG;
  if (out_dtype != SMF_OBS_U8 && out_dtype != SMF_OBS_F32) return SMF_ERR_BAD_ARG;
  if (covered != nullptr && covered_cols == nullptr) return SMF_ERR_BAD_ARG;
  if (n_grid > 0x3fffffff || n_cov_grid > 0x3fffffff || n_pos * n_channels > 0x3fffffff) return SMF_ERR_TOO_LONG;
  BuildArgs a;
  a.layer = layer;
  a.layer_dtype = layer_dtype;
  a.N = n_reads;
  a.V = (int)n_grid;
  a.L = (int)n_pos;
  a.C = n_channels;
  a.cols = cols;
  a.covered = covered;
  a.VC = (int)n_cov_grid;
  a.covered_cols = covered_cols;
  a.out = out;
  a.out_dtype = out_dtype;
  a.nonbinary = flag;
  const long long total = n_reads * n_pos * n_channels;
  long long blocks = (total + 255) / 256;
  const int cap = stream_grid();
  if (blocks > cap) blocks = cap;
  build_obs_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(a);
  return cudaGetLastError() == cudaSuccess ? SMF_OK : SMF_ERR_CUDA;
}

}  // extern "C"
