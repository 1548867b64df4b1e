// Viterbi decode (K2): max-plus forward sweep with packed back-pointers + back-trace.
//
// One thread owns one read and walks it sequentially in fp64, repeating the reference's
// operation order exactly (HMM.py:647-669):
//     logB = o*log(p+eps) + (1-o)*log1p(-p+eps)      (two products, one sum; 0 when masked)
//     cand = delta[t-1,i] + logA[i,j];  best over i with FIRST maximum;  delta[t,j] = best + logB
// so delta -- and therefore every arg-max decision, including exact ties -- is bit-identical
// to torch's fp64 CPU result when the log tables come from the same torch ops.
//
// Data movement: a CTA of kVitReads reads stages a [reads x kVitTile positions] observation
// tile through shared memory: each warp fetches whole 128-byte row segments (coalesced) one tile
// AHEAD into registers, so the loads of tile i+1 are in flight while tile i is being decoded;
// threads then read their own row (odd row stride: conflict free).  The K two-bit back-pointers
// of a position are packed into one int8 in a shared-memory tile, which is flushed as 32-bit
// words to a [N, L] byte scratch in HBM (one read's back-pointers are L bytes; 128 resident
// reads x 10 kb do not fit on chip) and streamed back tile by tile, in reverse, for the
// back-trace.  States leave through the same tile.
#pragma once
#include <type_traits>

#include "hmm_common.cuh"

namespace smf {

constexpr int kVitReads = 128;  // reads (threads) per CTA
constexpr int kVitTile = 32;    // positions per staged tile
constexpr int kVitWarps = kVitReads / 32;

struct VitArgs {
  const void* obs;
  int obs_dtype;
  long long N;
  int L;
  int C;
  const int8_t* bins;  // device [L-1] or nullptr
  int8_t* states;      // [N, L]
  uint8_t* bp;         // [N, L] packed back-pointers (scratch)
};

// typed observation load: OBSK = SMF_OBS_F32 / F64 / U8 known at compile time
template <int OBSK>
__device__ __forceinline__ typename std::conditional<OBSK == SMF_OBS_F64, double, float>::type vit_load(
    const void* base, long long idx) {
  if (OBSK == SMF_OBS_F32) return __ldg(reinterpret_cast<const float*>(base) + idx);
  if (OBSK == SMF_OBS_F64) return __ldg(reinterpret_cast<const double*>(base) + idx);
  const unsigned v = __ldg(reinterpret_cast<const unsigned char*>(base) + idx);
  return v == 0u ? 0.0f : (v == 1u ? 1.0f : __int_as_float(0x7fc00000));
}

template <int K, bool MULTI, bool BINNED, int OBSK>
__global__ void __launch_bounds__(kVitReads, 4) viterbi_kernel(const __grid_constant__ VitTables<K> tab,
                                                            const __grid_constant__ VitArgs a) {
  constexpr bool WIDE = (OBSK == SMF_OBS_F64);
  using obs_t = typename std::conditional<WIDE, double, float>::type;
  extern __shared__ __align__(16) unsigned char smem[];
  const int C = a.C;
  const int L = a.L;
  const int row_elems = kVitTile * C;
  const int row_stride = row_elems | 1;  // odd stride (in elements): conflict-free per-thread rows
  obs_t* tile = reinterpret_cast<obs_t*>(smem);
  constexpr int kBpStride = kVitTile + 4;  // bytes; 9 words -> conflict-free rows, 4-byte aligned
  uint8_t* bpt = smem + (((size_t)kVitReads * row_stride * sizeof(obs_t) + 15) & ~(size_t)15);
  int8_t* sbins = reinterpret_cast<int8_t*>(bpt + kVitReads * kBpStride);

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int wbase = (tid >> 5) * 32;  // first row of this warp
  const long long n0 = (long long)blockIdx.x * kVitReads;
  const int nrows = (int)min((long long)kVitReads, a.N - n0);
  const bool live = tid < nrows;
  // 32-bit access to the byte matrices needs every row start 4-byte aligned
  const bool words_ok = ((L & 3) == 0) && ((reinterpret_cast<uintptr_t>(a.bp) & 3) == 0) &&
                        ((reinterpret_cast<uintptr_t>(a.states) & 3) == 0);

  if (BINNED) {
    for (int i = tid; i < L - 1; i += kVitReads) sbins[i] = a.bins[i];
  }

  // ---- observation tile fetch (single channel: register prefetch, one row segment per step) ----
  obs_t nxt[MULTI ? 1 : 32];
  auto fetch_tile = [&](int tile0) {
    if (MULTI) return;
    const int tw = min(kVitTile, L - tile0);
    long long gidx = (n0 + wbase) * (long long)L + tile0 + lane;
    const bool col_ok = lane < tw;
#pragma unroll
    for (int r = 0; r < 32; ++r) {
      obs_t v = 0;
      if (col_ok && wbase + r < nrows)
        v = vit_load<OBSK>(a.obs, gidx);
      nxt[r] = v;
      gidx += L;
    }
  };
  auto commit_tile = [&]() {
    if (MULTI) return;
#pragma unroll
    for (int r = 0; r < 32; ++r) tile[(wbase + r) * row_stride + lane] = nxt[r];
  };
  auto load_tile_multi = [&](int tile0) {
    const int tw = min(kVitTile, L - tile0);
    const int n_el = tw * C;
    for (int r = 0; r < 32; ++r) {
      const int row = wbase + r;
      if (row >= nrows) break;
      const long long gbase = ((n0 + row) * (long long)L + tile0) * C;
      for (int c = lane; c < n_el; c += 32)
        tile[row * row_stride + c] = vit_load<OBSK>(a.obs, gbase + c);
    }
  };
  // byte tile <-> global: 32-bit words when aligned (8 words per row, 4 rows per warp step)
  auto tile_to_global = [&](uint8_t* gmat, int tile0, int tw) {
    if (words_ok && tw == kVitTile) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = wbase + i * 4 + (lane >> 3);
        if (row < nrows)
          *reinterpret_cast<uint32_t*>(gmat + (n0 + row) * (long long)L + tile0 + (lane & 7) * 4) =
              *reinterpret_cast<const uint32_t*>(bpt + row * kBpStride + (lane & 7) * 4);
      }
    } else {
      for (int r = 0; r < 32; ++r) {
        const int row = wbase + r;
        if (row < nrows && lane < tw) gmat[(n0 + row) * (long long)L + tile0 + lane] = bpt[row * kBpStride + lane];
      }
    }
  };
  auto global_to_tile = [&](const uint8_t* gmat, int tile0, int tw) {
    if (words_ok && tw == kVitTile) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = wbase + i * 4 + (lane >> 3);
        if (row < nrows)
          *reinterpret_cast<uint32_t*>(bpt + row * kBpStride + (lane & 7) * 4) =
              *reinterpret_cast<const uint32_t*>(gmat + (n0 + row) * (long long)L + tile0 + (lane & 7) * 4);
      }
    } else {
      for (int r = 0; r < 32; ++r) {
        const int row = wbase + r;
        if (row < nrows && lane < tw) bpt[row * kBpStride + lane] = gmat[(n0 + row) * (long long)L + tile0 + lane];
      }
    }
  };

  double delta[K];
#pragma unroll
  for (int k = 0; k < K; ++k) delta[k] = 0.0;

  // ---------------- forward max-plus sweep ----------------
  fetch_tile(0);
  for (int tile0 = 0; tile0 < L; tile0 += kVitTile) {
    const int tw = min(kVitTile, L - tile0);
    __syncthreads();  // previous tile fully consumed / flushed (and sbins visible)
    if (MULTI) load_tile_multi(tile0); else commit_tile();
    __syncthreads();
    if (tile0 + kVitTile < L) fetch_tile(tile0 + kVitTile);  // loads stay in flight during the decode below
    if (live) {
      const obs_t* my = tile + tid * row_stride;
      uint8_t* mybp = bpt + tid * kBpStride;
      // one position: emission, max-plus step, packed back-pointers.  FIRST_TILE keeps the t == 0 test.
      auto step = [&](int j, bool first_tile) {
        const int t = tile0 + j;
        double logB[K];
#pragma unroll
        for (int k = 0; k < K; ++k) logB[k] = 0.0;
        if (!MULTI) {
          const double o = (double)my[j];
          const bool m = (o == o);
          if (OBSK == SMF_OBS_U8) {
            // binary code: o is exactly 0 or 1, and 0*l1 + 1*l0 == l0, 1*l1 + 0*l0 == l1 exactly
#pragma unroll
            for (int k = 0; k < K; ++k) logB[k] = m ? (o != 0.0 ? tab.l1[k][0] : tab.l0[k][0]) : 0.0;
          } else {
            // branch-free (lanes of a warp hold different reads): the three-rounding expression is
            // also exact for o in {0, 1}.  A missing site uses os = om = 0: both products are (signed)
            // zeros, their sum is a zero, and delta + (+-0) == delta bit for bit -- the same as adding
            // the reference's masked 0 -- so no per-state select is needed.
            const double os = m ? o : 0.0;
            const double om = m ? __dsub_rn(1.0, os) : 0.0;
#pragma unroll
            for (int k = 0; k < K; ++k)
              logB[k] = __dadd_rn(__dmul_rn(os, tab.l1[k][0]), __dmul_rn(om, tab.l0[k][0]));
          }
        } else {
          for (int c = 0; c < C; ++c) {
            const double o = (double)my[j * C + c];
            if (o == o) {
              const double om = __dsub_rn(1.0, o);
#pragma unroll
              for (int k = 0; k < K; ++k)
                logB[k] = __dadd_rn(logB[k], __dadd_rn(__dmul_rn(o, tab.l1[k][c]), __dmul_rn(om, tab.l0[k][c])));
            }
          }
        }
        unsigned packed = 0;
        if (first_tile && t == 0) {
#pragma unroll
          for (int k = 0; k < K; ++k) delta[k] = __dadd_rn(tab.log_pi[k], logB[k]);
        } else {
          const int bin = BINNED ? (int)sbins[t - 1] : 0;
          double nd[K];
#pragma unroll
          for (int jn = 0; jn < K; ++jn) {
            double best = __dadd_rn(delta[0], tab.logA[bin][0][jn]);
            unsigned arg = 0;
#pragma unroll
            for (int i = 1; i < K; ++i) {
              const double cnd = __dadd_rn(delta[i], tab.logA[bin][i][jn]);
              if (cnd > best) {
                best = cnd;
                arg = i;
              }
            }
            nd[jn] = __dadd_rn(best, logB[jn]);
            packed |= arg << (2 * jn);
          }
#pragma unroll
          for (int k = 0; k < K; ++k) delta[k] = nd[k];
        }
        mybp[j] = (uint8_t)packed;
      };
      if (tile0 > 0 && tw == kVitTile && !MULTI) {
        // full interior tile: fully unrolled (no loop counter, immediate shared-memory offsets)
#pragma unroll
        for (int j = 0; j < kVitTile; ++j) step(j, false);
      } else {
#pragma unroll 4
        for (int j = 0; j < tw; ++j) step(j, true);
      }
    }
    __syncthreads();
    tile_to_global(a.bp, tile0, tw);
  }

  // last state: first maximum of delta[L-1] (torch.argmax, HMM.py:664)
  int s = 0;
  {
    double best = delta[0];
#pragma unroll
    for (int k = 1; k < K; ++k)
      if (delta[k] > best) {
        best = delta[k];
        s = k;
      }
  }

  // ---------------- back-trace, tiles in reverse ----------------
  const int ntiles = (L + kVitTile - 1) / kVitTile;
  for (int ti = ntiles - 1; ti >= 0; --ti) {
    const int tile0 = ti * kVitTile;
    const int tw = min(kVitTile, L - tile0);
    __syncthreads();  // this CTA's global bp writes visible; previous tile flushed
    global_to_tile(a.bp, tile0, tw);
    __syncthreads();
    if (live) {
      uint8_t* my = bpt + tid * kBpStride;
#pragma unroll 4
      for (int j = tw - 1; j >= 0; --j) {
        const unsigned packed = my[j];
        my[j] = (uint8_t)s;  // state at position tile0 + j
        if (tile0 + j > 0) s = (int)((packed >> (2 * s)) & 3u);  // s[t-1] = psi[t, s[t]]
      }
    }
    __syncthreads();
    tile_to_global(reinterpret_cast<uint8_t*>(a.states), tile0, tw);
  }
}

// -----------------------------------------------------------------------------------------------------
// viterbi_direct_kernel: single channel, homogeneous transitions, rows of a multiple of 4 positions.
//
// Same arithmetic as viterbi_kernel (fp64, the reference's operation order, first-index ties: bit-identical
// decisions), different data movement: NOTHING is staged.  A thread reads its own read 16 positions at a time
// straight from global memory (4 x LDG.128 of floats / one LDG.128 of the uint8 code, the next 16 positions in
// flight while the current ones are decoded: every sector fetched is used whole), keeps the 16 packed
// back-pointer bytes of the chunk in four registers and writes them with one 16-byte store (four 4-byte stores
// when the rows are not 16-byte aligned); the back-trace streams them back the same way, one chunk ahead, and
// writes 16 states per store.  No shared memory, no barrier, no transposition: ~25 of viterbi_kernel's 93
// thread-instructions per position were tile traffic (profiles/r01_c3v_viterbi_ncu.txt).
// -----------------------------------------------------------------------------------------------------
constexpr int kVitCH = 16;

template <int OBSK>
struct VitChunk;
template <>
struct VitChunk<SMF_OBS_F32> {
  float4 q[4];
  __device__ __forceinline__ void load(const void* row, int t0, int len) {
    const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(row) + t0);
#pragma unroll
    for (int g = 0; g < 4; ++g) q[g] = (4 * g < len) ? __ldg(p + g) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __device__ __forceinline__ double value(int j) const {
    const float4 v = q[j >> 2];
    return (double)((j & 3) == 0 ? v.x : ((j & 3) == 1 ? v.y : ((j & 3) == 2 ? v.z : v.w)));
  }
};
template <>
struct VitChunk<SMF_OBS_U8> {
  uint4 w;
  __device__ __forceinline__ void load(const void* row, int t0, int len) {
    const unsigned char* p = reinterpret_cast<const unsigned char*>(row) + t0;
    if (len == kVitCH) {
      w = __ldg(reinterpret_cast<const uint4*>(p));  // rows are 16-byte aligned in this variant (L % 16 == 0)
    } else {
      unsigned v[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
#pragma unroll
      for (int g = 0; g < 4; ++g)
        if (4 * g < len) v[g] = __ldg(reinterpret_cast<const unsigned*>(p) + g);
      w = make_uint4(v[0], v[1], v[2], v[3]);
    }
  }
  __device__ __forceinline__ unsigned code(int j) const {
    const unsigned word = (j >> 2) == 0 ? w.x : ((j >> 2) == 1 ? w.y : ((j >> 2) == 2 ? w.z : w.w));
    return (word >> (8 * (j & 3))) & 0xffu;
  }
};

template <bool A16>
__device__ __forceinline__ void vit_store16(uint8_t* p, const unsigned (&v)[4], int len) {
  if (A16 && len == kVitCH) {
    *reinterpret_cast<uint4*>(p) = make_uint4(v[0], v[1], v[2], v[3]);
  } else {
#pragma unroll
    for (int g = 0; g < 4; ++g)
      if (4 * g < len) *reinterpret_cast<unsigned*>(p + 4 * g) = v[g];
  }
}
template <bool A16>
__device__ __forceinline__ void vit_load16(const uint8_t* p, unsigned (&v)[4], int len) {
  if (A16 && len == kVitCH) {
    const uint4 t = *reinterpret_cast<const uint4*>(p);
    v[0] = t.x;
    v[1] = t.y;
    v[2] = t.z;
    v[3] = t.w;
  } else {
#pragma unroll
    for (int g = 0; g < 4; ++g) v[g] = (4 * g < len) ? *reinterpret_cast<const unsigned*>(p + 4 * g) : 0u;
  }
}

// A16: every row of obs (uint8) / bp / states starts on a 16-byte boundary (L % 16 == 0)
template <int K, int OBSK, bool A16>
__global__ void __launch_bounds__(128) viterbi_direct_kernel(const __grid_constant__ VitTables<K> tab,
                                                             const __grid_constant__ VitArgs a) {
  const int L = a.L;
  const int nch = (L + kVitCH - 1) / kVitCH;
  // transition / emission tables in registers (constant-bank reads inside the unrolled steps otherwise)
  double lA[K][K], l1[K], l0[K];
#pragma unroll
  for (int i = 0; i < K; ++i) {
    l1[i] = tab.l1[i][0];
    l0[i] = tab.l0[i][0];
#pragma unroll
    for (int j = 0; j < K; ++j) lA[i][j] = tab.logA[0][i][j];
  }
  for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < a.N; n += (long long)gridDim.x * blockDim.x) {
    const void* row = (OBSK == SMF_OBS_U8) ? (const void*)(reinterpret_cast<const unsigned char*>(a.obs) + n * (long long)L)
                                           : (const void*)(reinterpret_cast<const float*>(a.obs) + n * (long long)L);
    uint8_t* bprow = a.bp + n * (long long)L;
    uint8_t* strow = reinterpret_cast<uint8_t*>(a.states) + n * (long long)L;
    double delta[K];
#pragma unroll
    for (int k = 0; k < K; ++k) delta[k] = 0.0;
    VitChunk<OBSK> cur, nxt;
    cur.load(row, 0, min(kVitCH, L));
    // ---------------- forward max-plus sweep ----------------
    for (int c = 0; c < nch; ++c) {
      const int t0 = c * kVitCH;
      const int len = min(kVitCH, L - t0);
      if (c + 1 < nch) nxt.load(row, t0 + kVitCH, min(kVitCH, L - t0 - kVitCH));
      unsigned bpw[4] = {0u, 0u, 0u, 0u};
      auto step = [&](int j, bool maybe_first) {
        double logB[K];
        if constexpr (OBSK == SMF_OBS_U8) {
          const unsigned code = cur.code(j);
#pragma unroll
          for (int k = 0; k < K; ++k) logB[k] = code < 2u ? (code != 0u ? l1[k] : l0[k]) : 0.0;
        } else {
          // branch-free; exact for o in {0, 1}; a missing site adds a (signed) zero: see viterbi_kernel
          const double o = cur.value(j);
          const bool m = (o == o);
          const double os = m ? o : 0.0;
          const double om = m ? __dsub_rn(1.0, os) : 0.0;
#pragma unroll
          for (int k = 0; k < K; ++k) logB[k] = __dadd_rn(__dmul_rn(os, l1[k]), __dmul_rn(om, l0[k]));
        }
        unsigned packed = 0;
        if (maybe_first && t0 + j == 0) {
#pragma unroll
          for (int k = 0; k < K; ++k) delta[k] = __dadd_rn(tab.log_pi[k], logB[k]);
        } else {
          double nd[K];
#pragma unroll
          for (int jn = 0; jn < K; ++jn) {
            // first-index maximum; the back-pointer is selected as an already shifted constant (j and jn are
            // compile-time here), so a decision costs one fp64 compare, two 32-bit selects of the value and one of the code
            double best = __dadd_rn(delta[0], lA[0][jn]);
            unsigned code = 0u;
#pragma unroll
            for (int i = 1; i < K; ++i) {
              const double cnd = __dadd_rn(delta[i], lA[i][jn]);
              const bool p = cnd > best;
              best = p ? cnd : best;
              code = p ? ((unsigned)i << (2 * jn + 8 * (j & 3))) : code;
            }
            nd[jn] = __dadd_rn(best, logB[jn]);
            packed |= code;
          }
#pragma unroll
          for (int k = 0; k < K; ++k) delta[k] = nd[k];
        }
        bpw[j >> 2] |= packed;
      };
      if (c > 0 && len == kVitCH) {
#pragma unroll
        for (int j = 0; j < kVitCH; ++j) step(j, false);
      } else {
#pragma unroll
        for (int j = 0; j < kVitCH; ++j)
          if (j < len) step(j, true);
      }
      vit_store16<A16>(bprow + t0, bpw, len);
      cur = nxt;
    }
    // last state: first maximum of delta[L-1] (torch.argmax, HMM.py:664)
    unsigned s = 0;
    {
      double best = delta[0];
#pragma unroll
      for (int k = 1; k < K; ++k)
        if (delta[k] > best) {
          best = delta[k];
          s = k;
        }
    }
    // ---------------- back-trace, chunks in reverse (this thread re-reads its own stores) ----------------
    unsigned bw[4], bn[4];
    {
      const int t0 = (nch - 1) * kVitCH;
      vit_load16<A16>(bprow + t0, bw, L - t0);
    }
    for (int c = nch - 1; c >= 0; --c) {
      const int t0 = c * kVitCH;
      const int len = min(kVitCH, L - t0);
      if (c > 0) vit_load16<A16>(bprow + t0 - kVitCH, bn, kVitCH);
      unsigned sw[4] = {0u, 0u, 0u, 0u};
#pragma unroll
      for (int j = kVitCH - 1; j >= 0; --j) {
        if (j < len) {
          const unsigned packed = (bw[j >> 2] >> (8 * (j & 3))) & 0xffu;
          sw[j >> 2] |= s << (8 * (j & 3));       // state at position t0 + j
          if (t0 + j > 0) s = (packed >> (2 * s)) & 3u;  // s[t-1] = psi[t, s[t]]
        }
      }
      vit_store16<A16>(strow + t0, sw, len);
#pragma unroll
      for (int g = 0; g < 4; ++g) bw[g] = bn[g];
    }
  }
}


}  // namespace smf
