"""Build libb4cp.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc.

nvcc cross-compiles for sm_100a without a GPU; the built .so travels to the GPU box with the
repo snapshot.  `python -m bert4clickpath_b200.build` rebuilds what is stale.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJDIR = os.path.join(CSRC, "build")
LIB = os.path.join(HERE, "libb4cp.so")
HEADER = os.path.join(os.path.dirname(HERE), "include", "b4cp.h")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime():
    m = os.path.getmtime(HEADER)
    for f in os.listdir(CSRC):
        if f.endswith(".cuh") or f.endswith(".h"):
            m = max(m, os.path.getmtime(os.path.join(CSRC, f)))
    return m


def _compile(src, verbose):
    obj = os.path.join(OBJDIR, src[:-3] + ".o")
    cmd = [NVCC] + FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    log = os.path.join(OBJDIR, src[:-3] + ".ptxas.log")
    with open(log, "w") as f:
        f.write(r.stderr)
    if verbose:
        print(f"[b4cp build] compiled {src}")
    return obj


def build_lib(force=False, verbose=True):
    os.makedirs(OBJDIR, exist_ok=True)
    dep_m = _deps_mtime()
    todo, objs = [], []
    for src in _sources():
        obj = os.path.join(OBJDIR, src[:-3] + ".o")
        objs.append(obj)
        src_m = max(os.path.getmtime(os.path.join(CSRC, src)), dep_m)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < src_m:
            todo.append(src)
    if todo:
        with ThreadPoolExecutor(max_workers=min(8, len(todo))) as ex:
            list(ex.map(lambda s: _compile(s, verbose), todo))
    if todo or not os.path.exists(LIB):
        # shared cudart: the library binds to the CUDA runtime the process already has (torch's)
        # instead of embedding a second, static copy of it; rpath covers processes without torch
        cmd = [NVCC, "-shared", "-cudart", "shared", "-o", LIB] + objs + [
            "-gencode", "arch=compute_100a,code=sm_100a",
            "-Xlinker", "-rpath,/usr/local/cuda/lib64"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(f"[b4cp build] linked {LIB}")
    return LIB


if __name__ == "__main__":
    build_lib(force="--force" in sys.argv)
