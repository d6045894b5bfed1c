"""Builds rumi_slam_b200/librumi_orb.so (hand-written CUDA for sm_100a + the C ABI of include/rumi_orb.h).

In-tree build with plain nvcc (cross-compiles without a GPU): one object per .cu in parallel, then one shared
library.  `python -m rumi_slam_b200.build [--force] [--verbose]`.
"""
import concurrent.futures as cf
import fcntl
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, "librumi_orb.so")
SOURCES = ["pyramid.cu", "pyramid_strip.cu", "blur.cu", "fast.cu", "octree.cu", "describe.cu", "match.cu", "match_umma.cu", "bow.cu", "stereo.cu", "flow.cu", "api.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--fmad=false",
         "-Xcompiler", "-fPIC,-ffp-contract=off,-O2", "-Xptxas", "-v", "-I", os.path.join(HERE, "..", "include")]


def _headers():
    out = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh", ".inc"))]
    out.append(os.path.join(HERE, "..", "include", "rumi_orb.h"))
    return out


def _compile(src, verbose):
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    r = subprocess.run([NVCC, *FLAGS, "-c", os.path.join(CSRC, src), "-o", obj], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s" % (src, r.stderr[-6000:]))
    if verbose:
        sys.stderr.write(r.stderr)
    return obj, r.stderr


STAMP = LIB + ".stamp"


def _fingerprint(deps):
    """Content hash of every source / header / flag: file times do not survive the copy to the GPU box."""
    h = hashlib.sha256(" ".join(FLAGS[:-1]).encode())
    for d in sorted(deps):
        h.update(os.path.basename(d).encode())
        with open(d, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _fresh(fp):
    try:
        return os.path.exists(LIB) and open(STAMP).read().strip() == fp
    except OSError:
        return False


def build(force=False, verbose=False):
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    deps = [os.path.join(CSRC, s) for s in srcs] + _headers()
    fp = _fingerprint(deps)
    if not force and _fresh(fp):
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    with open(os.path.join(OBJ, ".lock"), "w") as lock:         # ranks of one torchrun job: one builds, the rest wait
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and _fresh(fp):
            return LIB
        return _build_locked(srcs, fp, verbose)


def _build_locked(srcs, fp, verbose):
    with cf.ThreadPoolExecutor(max_workers=8) as ex:
        res = list(ex.map(lambda s: _compile(s, verbose), srcs))
    log = "".join(r[1] for r in res)
    with open(os.path.join(OBJ, "ptxas.log"), "w") as f:
        f.write(log)
    r = subprocess.run([NVCC, "-shared", "-o", LIB, *[o for o, _ in res], "-lcudart", "-ldl"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stderr[-4000:])
    with open(STAMP, "w") as f:
        f.write(fp + "\n")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
