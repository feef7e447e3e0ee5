"""Build recipe of libcnvetti_b200.so (hand-written CUDA for sm_100a + the C ABI).

    python -m cnvetti_b200.build [--force] [--verbose]

The library is built IN-TREE (cnvetti_b200/libcnvetti_b200.so) so that it travels to the GPU box
with the repo snapshot.  nvcc cross-compiles for sm_100a without a GPU.
"""
import concurrent.futures
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "libcnvetti_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    # IEEE arithmetic everywhere: no fast-math, no flush-to-zero, exact div/sqrt
    "--ftz=false", "--prec-div=true", "--prec-sqrt=true",
    "-I", INCLUDE,
]


def nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; the CUDA extension cannot be built")
    return exe


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cpp")))  # (.cpp: host-only, nvcc hands it to g++)


def _deps_mtime():
    m = 0.0
    for root in (CSRC, INCLUDE):
        for f in os.listdir(root):
            if f.endswith((".cu", ".cuh", ".h", ".cpp")):
                m = max(m, os.path.getmtime(os.path.join(root, f)))
    return max(m, os.path.getmtime(__file__))


def build(force=False, verbose=False):
    newest = _deps_mtime()
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= newest:
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    exe = nvcc()

    def compile_one(src):
        obj = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
        cmd = [exe] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
            ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with concurrent.futures.ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(compile_one, sources()))
    cmd = [exe, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lz", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
