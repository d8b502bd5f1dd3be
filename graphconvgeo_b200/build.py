"""In-tree build of libgcg.so (sm_100a only) with nvcc.

``python -m graphconvgeo_b200.build`` or ``__graft_entry__.build()``.  The
shared object lands next to this file so that it travels with the repo
snapshot to the GPU box; it is git-ignored.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(ROOT, "include")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "libgcg.so")

SOURCES = ["gcg_core.cu", "gcg_spmm.cu", "gcg_spmm_stream.cu", "gcg_gemm.cu", "gcg_gemm_tc.cu", "gcg_elementwise.cu",
           "gcg_graph.cu", "gcg_host.cpp", "gcg_peer.cu", "gcg_spgemm.cu", "gcg_comm.cu"]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-O3,-Wall,-fopenmp", "--expt-relaxed-constexpr",
              "-I", INCLUDE, "-I", CSRC]
# nccl.h (types only: gcg_comm.cu binds NCCL with dlopen at run time): the system header, else the wheel's copy
for _d in ("/usr/include", os.path.join(os.path.dirname(os.path.dirname(sys.executable)), "lib",
                                        "python%d.%d" % sys.version_info[:2], "site-packages", "nvidia", "nccl", "include")):
    if os.path.exists(os.path.join(_d, "nccl.h")):
        if _d != "/usr/include":
            NVCC_FLAGS += ["-I", _d]
        break


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; libgcg.so cannot be built")
    return exe


def _digest(paths):
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(os.path.relpath(p, ROOT).encode())
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _sources():
    return [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def build(force: bool = False, verbose: bool = False) -> str:
    srcs = [os.path.join(CSRC, s) for s in _sources()]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    deps.append(os.path.join(INCLUDE, "gcg.h"))
    stamp = os.path.join(OBJ, "stamp")
    dig = _digest(deps)
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(OBJ, os.path.basename(src) + ".o")
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-x", "cu", "-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-Xcompiler", "-fopenmp", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    with open(stamp, "w") as f:
        f.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
