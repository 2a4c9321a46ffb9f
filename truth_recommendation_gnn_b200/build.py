"""In-tree nvcc build of the sm_100a kernel library (``lib/libtrg_b200.so``).

The built ``.so`` is git-ignored but travels to the GPU box with the working tree.  nvcc
cross-compiles without a GPU.  ``python -m truth_recommendation_gnn_b200.build`` rebuilds.
"""
from __future__ import annotations

import concurrent.futures
import hashlib
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_DIR = os.path.join(_HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libtrg_b200.so")
OBJ_DIR = os.path.join(_HERE, "build")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_digest():
    h = hashlib.sha256()
    hdrs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(_HERE, "..", "include", "trg_b200.h"))
    for p in hdrs:
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()[:16]


def _compile_one(src, obj, log):
    cmd = [NVCC, *NVCC_FLAGS, "-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIB_DIR, exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    digest = _deps_digest()
    jobs, objs = [], []
    for src in sources():
        with open(src, "rb") as f:
            tag = hashlib.sha256(f.read() + digest.encode()).hexdigest()[:16]
        base = os.path.splitext(os.path.basename(src))[0]
        obj = os.path.join(OBJ_DIR, f"{base}.{tag}.o")
        objs.append(obj)
        if force or not os.path.exists(obj):
            jobs.append((src, obj, os.path.join(OBJ_DIR, f"{base}.log")))
    if jobs:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for fut in [ex.submit(_compile_one, *j) for j in jobs]:
                obj = fut.result()
                if verbose:
                    print("compiled", obj)
    stamp = os.path.join(OBJ_DIR, "link.stamp")
    want = "\n".join(objs)
    have = open(stamp).read() if os.path.exists(stamp) else ""
    if jobs or force or not os.path.exists(LIB_PATH) or have != want:
        cmd = [NVCC, "-shared", "-o", LIB_PATH, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
               "-lcuda"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        with open(stamp, "w") as f:
            f.write(want)
        # drop stale objects
        keep = set(objs)
        for f in os.listdir(OBJ_DIR):
            p = os.path.join(OBJ_DIR, f)
            if f.endswith(".o") and p not in keep:
                os.remove(p)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
