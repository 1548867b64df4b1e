"""Load the UNMODIFIED reference HMM module from /root/reference (container only).

TEST INFRASTRUCTURE -- never imported by the product package.

``import smftools`` fails in this image (``smftools/__init__.py`` pulls in anndata), but
``smftools/hmm/HMM.py`` itself only needs numpy / scipy.sparse / torch and three light
helper modules (``HMM.py:8-13``).  Registering empty parent packages whose ``__path__``
points at the reference tree lets ``import smftools.hmm.HMM`` run the reference class
unmodified.

``/root/reference`` does not exist on the GPU box: only ``oracle/make_golden.py`` and
the container-only cross-check tests call :func:`load_reference_hmm`.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_SRC = os.environ.get("SMF_REFERENCE_SRC", "/root/reference/src")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_SRC, "smftools", "hmm", "HMM.py"))


def load_reference_hmm():
    """Return the reference ``smftools.hmm.HMM`` module (imported once, cached)."""
    if "smftools.hmm.HMM" in sys.modules and getattr(
        sys.modules["smftools.hmm.HMM"], "__smf_reference__", False
    ):
        return sys.modules["smftools.hmm.HMM"]
    if not reference_available():
        raise FileNotFoundError(f"reference tree not found under {REFERENCE_SRC}")
    root = os.path.join(REFERENCE_SRC, "smftools")
    pkg = types.ModuleType("smftools")
    pkg.__path__ = [root]
    sub = types.ModuleType("smftools.hmm")
    sub.__path__ = [os.path.join(root, "hmm")]
    sys.modules["smftools"] = pkg
    sys.modules["smftools.hmm"] = sub
    mod = importlib.import_module("smftools.hmm.HMM")
    mod.__smf_reference__ = True
    return mod


class _PermissiveStub(types.ModuleType):
    """Stand-in for a module the reference imports at load time but this image lacks (anndata-backed
    ``smftools.readwrite``, matplotlib): attribute lookups succeed, calling anything raises."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)

        def _stub(*args, **kwargs):
            raise RuntimeError(f"stubbed {self.__name__}.{name} was called")

        _stub.__name__ = name
        return _stub


def load_reference_cli():
    """Return the reference's ``(smftools.cli.hmm_adata, smftools.tools.partitioned_hmm)`` modules,
    unmodified, with the absent I/O / plotting dependencies stubbed.  The functions on and
    next to the hot path (``HMMTrainer``, ``build_single_channel``, ``build_multi_channel_union``,
    ``_mask_uncovered_model_input``, ``_apply_merges``, the layer reductions) only need numpy / pandas."""
    load_reference_hmm()
    root = os.path.join(REFERENCE_SRC, "smftools")
    if "smftools.readwrite" not in sys.modules:
        sys.modules["smftools.readwrite"] = _PermissiveStub("smftools.readwrite")
    for nm in ("matplotlib", "matplotlib.colors", "matplotlib.pyplot"):
        try:
            importlib.import_module(nm)
        except Exception:
            sys.modules[nm] = _PermissiveStub(nm)
    for sub in ("cli", "tools", "informatics", "plotting", "analysis", "preprocessing"):
        name = f"smftools.{sub}"
        if name not in sys.modules:
            pkg = types.ModuleType(name)
            pkg.__path__ = [os.path.join(root, sub)]
            sys.modules[name] = pkg
    return (importlib.import_module("smftools.cli.hmm_adata"),
            importlib.import_module("smftools.tools.partitioned_hmm"))


class DuckAnnData:
    """Minimal stand-in for AnnData (the reference only touches these attributes)."""

    def __init__(self, n_obs, var_names, obs=None, var=None, X=None):
        import pandas as pd

        self.var_names = pd.Index([str(v) for v in var_names])
        self.n_obs = int(n_obs)
        self.n_vars = len(self.var_names)
        self.shape = (self.n_obs, self.n_vars)
        self.layers = {}
        self.uns = {}
        self.obs = obs if obs is not None else pd.DataFrame(index=[f"r{i}" for i in range(n_obs)])
        self.var = var if var is not None else pd.DataFrame(index=self.var_names)
        self.obs_names = self.obs.index
        self.X = X

    def __getitem__(self, key):
        """``adata[:, column_mask]`` -- the only slicing the input builders use (cli/hmm_adata.py:184, 273)."""
        import numpy as np

        rows, cols = key
        if not (isinstance(rows, slice) and rows == slice(None)):
            raise NotImplementedError("DuckAnnData only slices columns")
        cols = np.asarray(cols)
        idx = np.flatnonzero(cols) if cols.dtype == bool else cols.astype(int)
        sub = DuckAnnData(self.n_obs, self.var_names[idx], obs=self.obs, var=self.var.iloc[idx],
                          X=None if self.X is None else np.asarray(self.X)[:, idx])
        sub.layers = {k: np.asarray(v)[:, idx] for k, v in self.layers.items()}
        sub.uns = self.uns
        return sub
