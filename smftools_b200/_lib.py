"""ctypes binding of ``libsmfhmm.so`` (the C-ABI declared in ``include/smf_hmm.h``).

The product path has no CPU fallback: if the shared library is missing, or a compute entry
point is called without a CUDA device, an exception is raised.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsmfhmm.so")

SMF_OK = 0
SMF_ERR_BAD_ARG = -1
SMF_ERR_TOO_LONG = -2
SMF_ERR_WORKSPACE = -3
SMF_ERR_CUDA = -4
SMF_ERR_ALIGN = -5

OBS_F32, OBS_F64, OBS_U8 = 0, 1, 2
OP_POSTERIOR, OP_VITERBI, OP_ESTEP, OP_FEATURES = 0, 1, 2, 3
FIT_PHASE_ALL, FIT_PHASE_BEGIN, FIT_PHASE_ESTEP, FIT_PHASE_MSTEP = 0, 1, 2, 3
MAX_STATES, MAX_CHANNELS, MAX_BINS = 4, 4, 16

EXPORTS = (
    "smf_hmm_abi_version",
    "smf_hmm_set_option",
    "smf_hmm_strerror",
    "smf_hmm_workspace_bytes",
    "smf_hmm_max_positions",
    "smf_hmm_posterior",
    "smf_hmm_viterbi",
    "smf_hmm_stats_len",
    "smf_hmm_estep",
    "smf_hmm_grouped_workspace_bytes",
    "smf_hmm_grouped_posterior",
    "smf_hmm_grouped_fit",
    "smf_hmm_grouped_fit_step",
    "smf_hmm_call_features",
    "smf_hmm_merge_runs",
    "smf_hmm_merge_chain",
    "smf_hmm_span_mask",
    "smf_hmm_widen_decode",
    "smf_hmm_expand_layers",
    "smf_hmm_build_obs",
    "smf_hmm_layer_colsums",
    "smf_hmm_column_metric",
    "smf_hmm_run_histograms",
    "smf_hmm_host_register",
    "smf_hmm_host_unregister",
    "smf_hmm_host_is_pinned",
    "smf_hmm_copy_async",
)

LAYER_STATES, LAYER_POSTERIOR, LAYER_ANY, LAYER_ANY_LEN, LAYER_CLASS, LAYER_CLASS_LEN = range(6)
COPY_H2D, COPY_D2H, COPY_D2D = 0, 1, 2


class LayerGroup(ctypes.Structure):
    """Mirror of ``smf_hmm_layer_group``."""

    _fields_ = [("cls_or_code", c_void_p), ("run_len", c_void_p), ("packed", c_int32), ("reserved", c_int32)]


class LayerDesc(ctypes.Structure):
    """Mirror of ``smf_hmm_layer_desc``."""

    _fields_ = [("out", c_void_p), ("kind", c_int32), ("group", c_int32), ("cls", c_int32), ("out_f32", c_int32)]


class Tables(ctypes.Structure):
    """Mirror of ``smf_hmm_tables``."""

    _fields_ = [
        ("n_states", c_int32),
        ("n_channels", c_int32),
        ("n_bins", c_int32),
        ("reserved", c_int32),
        ("log_start", POINTER(c_double)),
        ("log_trans", POINTER(c_double)),
        ("log_emit1", POINTER(c_double)),
        ("log_emit0", POINTER(c_double)),
    ]


class Group(ctypes.Structure):
    """Mirror of ``smf_hmm_group``."""

    _fields_ = [
        ("obs", c_void_p),
        ("n_reads", c_int64),
        ("n_pos", c_int32),
        ("row_stride", c_int32),
        ("gamma", c_void_p),
        ("states", c_void_p),
        ("loglik", c_void_p),
        ("gamma_state", c_int32),
        ("out_stride", c_int32),
    ]


class HostTables:
    """Owns the fp64 host arrays a ``Tables`` struct points at."""

    def __init__(self, log_start, log_trans, log_emit1, log_emit0):
        self.log_start = np.ascontiguousarray(log_start, dtype=np.float64).reshape(-1)
        K = self.log_start.shape[0]
        self.log_trans = np.ascontiguousarray(log_trans, dtype=np.float64).reshape(-1, K, K)
        self.log_emit1 = np.ascontiguousarray(log_emit1, dtype=np.float64).reshape(K, -1)
        self.log_emit0 = np.ascontiguousarray(log_emit0, dtype=np.float64).reshape(K, -1)
        self.K = K
        self.C = self.log_emit1.shape[1]
        self.n_bins = self.log_trans.shape[0]
        dp = POINTER(c_double)
        self.struct = Tables(
            self.K,
            self.C,
            self.n_bins,
            0,
            self.log_start.ctypes.data_as(dp),
            self.log_trans.ctypes.data_as(dp),
            self.log_emit1.ctypes.data_as(dp),
            self.log_emit0.ctypes.data_as(dp),
        )

    @property
    def ref(self):
        return ctypes.byref(self.struct)


_lib = None


def library_built() -> bool:
    return os.path.isfile(LIB_PATH)


ABI_VERSION = 5


def load():
    """Load libsmfhmm.so (once).  Raises if it has not been built -- there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not library_built():
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `python -m smftools_b200.build`). smftools_b200 has no CPU fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    T = POINTER(Tables)
    lib.smf_hmm_abi_version.restype = c_int
    lib.smf_hmm_abi_version.argtypes = []
    lib.smf_hmm_strerror.restype = c_char_p
    lib.smf_hmm_strerror.argtypes = [c_int]
    lib.smf_hmm_set_option.restype = c_int
    lib.smf_hmm_set_option.argtypes = [c_char_p, c_int]
    lib.smf_hmm_workspace_bytes.restype = c_size_t
    lib.smf_hmm_workspace_bytes.argtypes = [c_int, T, c_int64, c_int64]
    lib.smf_hmm_max_positions.restype = c_int64
    lib.smf_hmm_max_positions.argtypes = [T, c_int]
    lib.smf_hmm_stats_len.restype = c_int64
    lib.smf_hmm_stats_len.argtypes = [T]
    lib.smf_hmm_posterior.restype = c_int
    lib.smf_hmm_posterior.argtypes = [T, c_void_p, c_int, c_int64, c_int64, c_void_p, c_void_p, c_int,
                                      c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]
    lib.smf_hmm_viterbi.restype = c_int
    lib.smf_hmm_viterbi.argtypes = [T, c_void_p, c_int, c_int64, c_int64, c_void_p, c_void_p, c_void_p,
                                    c_size_t, c_void_p]
    lib.smf_hmm_estep.restype = c_int
    lib.smf_hmm_estep.argtypes = [T, c_void_p, c_int, c_int64, c_int64, c_void_p, c_void_p, c_void_p,
                                  c_size_t, c_void_p]
    GP = POINTER(Group)
    lib.smf_hmm_grouped_workspace_bytes.restype = c_size_t
    lib.smf_hmm_grouped_workspace_bytes.argtypes = [c_int, GP, c_int, c_int]
    lib.smf_hmm_grouped_posterior.restype = c_int
    lib.smf_hmm_grouped_posterior.argtypes = [c_int, GP, c_int, T, c_void_p, c_size_t, c_void_p]
    lib.smf_hmm_grouped_fit.restype = c_int
    lib.smf_hmm_grouped_fit.argtypes = [c_int, c_int, GP, c_int, c_void_p, c_int, c_double, c_double, c_int, c_void_p,
                                        c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]
    lib.smf_hmm_grouped_fit_step.restype = c_int
    lib.smf_hmm_grouped_fit_step.argtypes = [c_int, c_int, c_int, GP, c_int, c_void_p, c_int, c_int, c_double, c_double,
                                             c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]
    lib.smf_hmm_call_features.restype = c_int
    lib.smf_hmm_call_features.argtypes = [c_void_p, c_void_p, c_int, c_double, c_int64, c_int64, c_void_p,
                                          c_void_p, c_void_p, c_int64, POINTER(c_double), POINTER(c_double),
                                          c_int, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p]
    lib.smf_hmm_merge_runs.restype = c_int
    lib.smf_hmm_merge_runs.argtypes = [c_void_p, c_int64, c_int64, c_void_p, c_int, c_int64, c_void_p,
                                       c_void_p, POINTER(c_double), POINTER(c_double), c_int, c_void_p,
                                       c_void_p, c_void_p]
    lib.smf_hmm_merge_chain.restype = c_int
    lib.smf_hmm_merge_chain.argtypes = [c_void_p, c_int64, c_int64, c_void_p, c_int, c_int64, POINTER(c_double),
                                        POINTER(c_double), c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_void_p, c_void_p]
    lib.smf_hmm_span_mask.restype = c_int
    lib.smf_hmm_span_mask.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p,
                                      c_void_p]
    lib.smf_hmm_widen_decode.restype = c_int
    lib.smf_hmm_widen_decode.argtypes = [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p]
    lib.smf_hmm_expand_layers.restype = c_int
    lib.smf_hmm_expand_layers.argtypes = [c_int64, c_int64, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_void_p,
                                          POINTER(LayerGroup), c_int, POINTER(LayerDesc), c_int, c_void_p]
    lib.smf_hmm_layer_colsums.restype = c_int
    lib.smf_hmm_layer_colsums.argtypes = [c_void_p, c_int, c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p, c_void_p]
    lib.smf_hmm_column_metric.restype = c_int
    lib.smf_hmm_column_metric.argtypes = [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p]
    lib.smf_hmm_run_histograms.restype = c_int
    lib.smf_hmm_run_histograms.argtypes = [c_void_p, c_int, c_int, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p]
    lib.smf_hmm_build_obs.restype = c_int
    lib.smf_hmm_build_obs.argtypes = [c_void_p, c_int, c_int64, c_int64, c_void_p, c_int64, c_int, c_void_p, c_int64,
                                      c_void_p, c_void_p, c_int, c_void_p, c_void_p]
    lib.smf_hmm_host_register.restype = c_int
    lib.smf_hmm_host_register.argtypes = [c_void_p, c_size_t]
    lib.smf_hmm_host_unregister.restype = c_int
    lib.smf_hmm_host_unregister.argtypes = [c_void_p]
    lib.smf_hmm_host_is_pinned.restype = c_int
    lib.smf_hmm_host_is_pinned.argtypes = [c_void_p]
    lib.smf_hmm_copy_async.restype = c_int
    lib.smf_hmm_copy_async.argtypes = [c_void_p, c_void_p, c_size_t, c_int, c_void_p]
    if lib.smf_hmm_abi_version() != ABI_VERSION:
        raise RuntimeError("libsmfhmm.so ABI version mismatch; rebuild the extension")
    _lib = lib
    return lib


def check(code: int, what: str = "") -> None:
    """Map a C-ABI status to the reference's exception conventions."""
    if code == SMF_OK:
        return
    msg = load().smf_hmm_strerror(code).decode()
    if what:
        msg = f"{what}: {msg}"
    if code in (SMF_ERR_BAD_ARG, SMF_ERR_TOO_LONG, SMF_ERR_WORKSPACE, SMF_ERR_ALIGN):
        raise ValueError(msg)
    raise RuntimeError(msg)
