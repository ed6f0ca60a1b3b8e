"""Builds ``lib/libmlstm_b200.so`` (the C-ABI library of include/mlstm_b200.h) in-tree with
nvcc for sm_100a.  nvcc cross-compiles without a GPU; the built .so travels to the GPU box.

    python -m xlstm_yolo_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
LIB = os.path.join(LIBDIR, "libmlstm_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = _sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(PKG, "..", "include", "mlstm_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ into one shared library. Returns its path."""
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(LIBDIR, exist_ok=True)
    objs = []
    procs = []
    for src in _sources():
        obj = os.path.join(LIBDIR, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [nvcc, *NVCC_FLAGS, "-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, pr in procs:
        out, _ = pr.communicate()
        log.append(f"== {os.path.basename(src)}\n{out}")
        if pr.returncode != 0:
            if os.path.exists(LIB):
                os.remove(LIB)   # never leave a stale library behind a failed build: tests would silently run the old kernels
            raise RuntimeError(f"nvcc failed for {src}:\n{out}")
    cmd = [nvcc, "-shared", "-o", LIB, *objs, "-lcudart_static", "-ldl", "-lrt", "-lpthread"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    with open(os.path.join(LIBDIR, "build.log"), "w") as fh:
        fh.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return LIB


def build_variant(tag: str, defines=(), sources=("mlstm_tc_bwd_fused128.cu",)) -> str:
    """Developer build: ``lib/libmlstm_b200_<tag>.so`` = the normal objects with the named sources recompiled under the
    given -D defines (A/B experiments; select it with MLSTM_B200_LIB=<path>)."""
    build()
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    for src in _sources():
        base = os.path.basename(src)
        obj = os.path.join(LIBDIR, base[:-3] + ".o")
        if base in sources:
            obj = os.path.join(LIBDIR, base[:-3] + f"_{tag}.o")
            subprocess.run([nvcc, *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-c", src, "-o", obj], check=True,
                           stdout=subprocess.DEVNULL, stderr=subprocess.STDOUT)
        objs.append(obj)
    out = os.path.join(LIBDIR, f"libmlstm_b200_{tag}.so")
    subprocess.run([nvcc, "-shared", "-o", out, *objs, "-lcudart_static", "-ldl", "-lrt", "-lpthread"], check=True)
    return out


def build_timeline(sources=("mlstm_tc_bwd_fused128.cu",)) -> str:
    """``lib/libmlstm_b200_tl.so``: the named sources under -DMLSTM_TIMELINE (CTA 0 stamps clock64() per phase into the
    workspace; tests/gpu_tools/timeline_*.py read it)."""
    return build_variant("tl", ("MLSTM_TIMELINE",), sources)


if __name__ == "__main__":
    if "--timeline" in sys.argv:
        print(build_timeline())
        sys.exit(0)
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
