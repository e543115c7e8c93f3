"""Build libta_b200.so (sm_100a only) in-tree with nvcc. No torch headers, no JIT cache: the .so lives next to
this file so that it travels to the GPU box with the repo snapshot.

    python -m transferattack_b200._build [--force] [--verbose]
"""
import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
SO = os.path.join(HERE, "libta_b200.so")
SOURCES = ["lib.cu", "elementwise.cu", "reduce.cu", "aten_mean.cu", "fused_update.cu", "dim.cu", "dim_direct.cu", "dwconv.cu", "philox.cu", "longtail.cu", "spectrum.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-fmad=false",              # no implicit FMA contraction: one rounding per reference op (csrc/common.cuh)
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def _deps():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(os.path.dirname(HERE), "include", "ta_b200.h"))
    return hdrs


def _stale(target, srcs):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in srcs)


def _compile(src, verbose):
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    path = os.path.join(CSRC, src)
    if not _stale(obj, [path] + _deps()):
        return obj, ""
    cmd = [NVCC] + FLAGS + ["-c", path, "-o", obj]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, p.stdout, p.stderr))
    with open(obj + ".ptxas.log", "w") as f:
        f.write(p.stderr)
    return obj, p.stderr if verbose else ""


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    with cf.ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        res = list(ex.map(lambda s: _compile(s, verbose), SOURCES))
    objs = [r[0] for r in res]
    if verbose:
        for _, log in res:
            if log:
                print(log)
    if force or _stale(SO, objs):
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", SO] + objs
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (p.stdout, p.stderr))
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
