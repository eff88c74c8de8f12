"""In-tree build of libhavatar_b200.so (sm_100a only) with plain nvcc -- no JIT cache, no torch headers.

    python -m havatar_b200.build [--force] [--verbose]

The library has a pure C ABI (include/havatar_b200.h) and links only against the CUDA runtime, so
it is loaded with ctypes (havatar_b200/_lib.py) and travels to the GPU box with the repo snapshot.
"""
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "csrc", "build")
LIB = os.path.join(HERE, "libhavatar_b200.so")
SOURCES = ["render_api.cu", "render_simt.cu", "render_tc.cu", "render_tc2.cu", "render_tc3.cu", "render_bwd.cu", "ops.cu", "fir_cl.cu", "conv_tc.cu", "style.cu", "conv_wgrad.cu", "optim.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC"] + os.environ.get("HAV_NVCC_DEFS", "").split()


def _nvcc():
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found: libhavatar_b200.so cannot be built (there is no CPU fallback)")


def _digest(paths):
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode() + b"\0" + f.read())
    return h.hexdigest()


def _sources():
    return [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _deps():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    hdrs.append(os.path.join(ROOT, "include", "havatar_b200.h"))
    return hdrs + [os.path.join(CSRC, s) for s in _sources()] + [os.path.abspath(__file__)]


def _flags_tag():
    return os.environ.get("HAV_NVCC_DEFS", "")


def is_fresh():
    stamp = LIB + ".stamp"
    if not (os.path.exists(LIB) and os.path.exists(stamp)):
        return False
    with open(stamp) as f:
        return f.read().strip() == _digest(_deps()) + _flags_tag()


def build_library(force=False, verbose=False):
    """Compile every .cu for sm_100a and link libhavatar_b200.so.  Returns the library path."""
    if not force and is_fresh():
        return LIB
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    flags = FLAGS + (["-Xptxas", "-v"] if verbose else [])

    def compile_one(src):
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        cmd = [nvcc] + ARCH + flags + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, _sources()))
    cmd = [nvcc] + ARCH + ["-shared", "-o", LIB] + objs   # static cudart: no dependency on torch's copy
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    with open(LIB + ".stamp", "w") as f:
        f.write(_digest(_deps()) + _flags_tag())
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
