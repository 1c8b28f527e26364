"""Build libguidegen_sm100.so (sm_100a only) from csrc/*.cu with nvcc, in-tree.

    python -m jointimagegeneration_b200.build [--force]

The shared library is the C-ABI boundary declared in include/guidegen_sm100.h.  It is
built into jointimagegeneration_b200/lib/ so that it travels with the repo snapshot to
the GPU box (the .so is git-ignored, not gpurun-ignored).
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "lib")
# tuning builds: GG_BUILD_TAG=<tag> GG_BUILD_DEFS="-DFOO=1 ..." -> lib/libguidegen_sm100_<tag>.so (loaded with GG_LIB=<path>)
TAG = os.environ.get("GG_BUILD_TAG", "")
OBJ_DIR = os.path.join(OUT_DIR, "obj" + ("_" + TAG if TAG else ""))
LIB = os.path.join(OUT_DIR, "libguidegen_sm100%s.so" % ("_" + TAG if TAG else ""))
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-I", INCLUDE, "--expt-relaxed-constexpr"] + os.environ.get("GG_BUILD_DEFS", "").split()


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs += [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE)]
    return max(os.path.getmtime(h) for h in hdrs)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ_DIR, exist_ok=True)
    srcs = _sources()
    hdr_m = _deps_mtime()
    todo, objs = [], []
    for s in srcs:
        o = os.path.join(OBJ_DIR, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), hdr_m):
            todo.append((s, o))

    def cc(so):
        s, o = so
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (s, r.stdout, r.stderr))
        return r.stderr

    if todo:
        with ThreadPoolExecutor(max_workers=min(8, len(todo))) as ex:
            for log in ex.map(cc, todo):
                if verbose and log:
                    sys.stderr.write(log)
    if todo or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart_static",
                                                       "-Xcompiler", "-fPIC"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
