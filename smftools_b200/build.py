"""Build ``libsmfhmm.so`` in-tree with nvcc for sm_100a (cross-compiles without a GPU).

The kernels are heavily templated (fully unrolled per-chunk loops), so the library is split into
one translation unit per state count and kernel family; they are compiled in parallel and linked
into a single shared object.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
OUT = os.path.join(HERE, "libsmfhmm.so")
SOURCES = (["api.cu", "host_io.cu", "reduce_kernels.cu"] + [f"{fam}_k{k}.cu" for fam in ("fb2", "fb", "vit") for k in (2, 3, 4)]
           + [f"fb2w_k{k}.cu" for k in (2, 3, 4)] + [f"short_k{k}.cu" for k in (2, 3, 4)] + [f"seq_k{k}.cu" for k in (2, 3, 4)] + ["fb2x_k2.cu", "fb2x_k3.cu", "fb2y_k2.cu", "fb2m_k2.cu", "fb2m_k3.cu", "fb2b_k2.cu", "fb2b_k3.cu"])
HEADERS = ["f32x2.cuh", "seq_kernel.cuh", "seq_launch.inc", "post_sweep_kernel.cuh", "short_kernel.cuh", "short_launch.inc", "sweep_kernel.cuh", "hmm_common.cuh", "fb_kernel.cuh", "fb2_kernel.cuh", "viterbi_kernel.cuh", "features_kernel.cuh",
           "walk_kernel.cuh", "launch.h", "fb2_launch.inc", "fb2m_launch.inc", "fb2b_launch.inc", "fb2w_launch.inc", "fb_launch.inc", "vit_launch.inc"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _deps():
    return [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(ROOT, "include", "smf_hmm.h")]


def needs_build() -> bool:
    if not os.path.isfile(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in _deps())


def _obj_stale(src: str, obj: str) -> bool:
    """True when the object is missing or older than the source or any header it included last time
    (dependency file written by ``nvcc -MD``)."""
    dep = obj[:-2] + ".d"
    if not (os.path.isfile(obj) and os.path.isfile(dep)):
        return True
    t = os.path.getmtime(obj)
    try:
        with open(dep) as fh:
            words = fh.read().replace("\\\n", " ").split()
    except OSError:
        return True
    for w in words[1:]:
        if w.startswith("/usr/") or w.endswith(":"):
            continue
        if not os.path.isfile(w) or os.path.getmtime(w) > t:
            return True
    return os.path.getmtime(os.path.join(CSRC, src)) > t


def _compile(src: str, verbose: bool, force: bool = False):
    obj = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
    if not force and not verbose and not _obj_stale(src, obj):
        return src, obj, subprocess.CompletedProcess([], 0, "", "")
    cmd = [_nvcc(), "-O3", "-std=c++17", *ARCH, "-lineinfo", "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"),
           "-I", CSRC, "-MD", "-MF", obj[:-2] + ".d", "-c", os.path.join(CSRC, src), "-o", obj]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    return src, obj, res


def build(force: bool = False, verbose: bool = False, jobs: int | None = None) -> str:
    if not force and not needs_build():
        return OUT
    os.makedirs(OBJ, exist_ok=True)
    jobs = jobs or min(len(SOURCES), max(1, (os.cpu_count() or 2)))
    objs = []
    with ThreadPoolExecutor(max_workers=jobs) as pool:
        for src, obj, res in pool.map(lambda s: _compile(s, verbose, force), SOURCES):
            if res.returncode != 0:
                sys.stderr.write(res.stdout + res.stderr)
                raise RuntimeError(f"nvcc failed on {src}")
            if verbose:
                sys.stderr.write(res.stderr)
            objs.append(obj)
    link = [_nvcc(), "-shared", *ARCH, "-Xcompiler", "-fPIC", "-o", OUT] + objs
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed linking libsmfhmm.so")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
