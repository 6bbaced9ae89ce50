"""Builds libggnn_b200.so IN-TREE with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
from __future__ import annotations

import os
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libggnn_b200.so")
NVCC_FLAGS = ["-shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
              "-std=c++17", "-Xcompiler", "-fopenmp", "-lgomp"]   # OpenMP: the host-side scan of a dense adjacency (ggnn_set_graph_dense)
if os.environ.get("GGNN_TC_TRACE") == "1":   # trace build: phase stamps / event log of the tile-local tcgen05 kernel (tools/tc_trace.py)
    NVCC_FLAGS.append("-DGGNN_TC_TRACE")


def sources():
    out = []
    for root, _, files in os.walk(CSRC):
        out += [os.path.join(root, f) for f in files if f.endswith((".cu", ".cuh", ".h"))]
    out.append(os.path.join(os.path.dirname(PKG_DIR), "include", "ggnn_b200.h"))
    return sorted(out)


HASH_PATH = LIB_PATH + ".srchash"   # content hash of the sources the .so was built from (mtimes do not survive snapshots)


def source_hash() -> str:
    import hashlib
    h = hashlib.sha256()
    for s in sources():
        if os.path.exists(s):
            h.update(os.path.relpath(s, PKG_DIR).encode())
            with open(s, "rb") as f:
                h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH) or not os.path.exists(HASH_PATH):
        return True
    with open(HASH_PATH) as f:
        return f.read().strip() != source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ... -> libggnn_b200.so (cross-compiles without a GPU)."""
    if not force and not is_stale():
        return LIB_PATH
    nvcc = os.environ.get("NVCC") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        nvcc = "nvcc"
    cus = [s for s in sources() if s.endswith(".cu")]
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH + ".tmp"] + cus
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), res.stderr[-4000:]))
    os.replace(LIB_PATH + ".tmp", LIB_PATH)
    with open(HASH_PATH, "w") as f:
        f.write(source_hash())
    if verbose:
        print(res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
