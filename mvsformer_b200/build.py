"""Build libmvs_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m mvsformer_b200.build [--force] [--verbose]

Every ``csrc/*.cu`` is compiled to ``build/obj/*.o`` in parallel (skipped when the object is
newer than the source and the shared headers), then linked to ``mvsformer_b200/lib/libmvs_b200.so``.
The ``.so`` is git-ignored but travels to the GPU box with the repo snapshot.
"""
import argparse
import concurrent.futures as cf
import glob
import os
import shutil
import subprocess
import sys
import time

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(ROOT, "build", "obj")
LIB = os.path.join(PKG, "lib", "libmvs_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-I", os.path.join(ROOT, "include"),
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; cannot build libmvs_b200.so")
    return exe


def _newest_header():
    hdrs = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.inl")) + glob.glob(os.path.join(ROOT, "include", "*.h")) + [__file__]
    return max(os.path.getmtime(h) for h in hdrs)


def _compile(src, force, verbose):
    obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
    if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), _newest_header()):
        return obj, 0.0, ""
    t0 = time.time()
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, res.stdout, res.stderr))
    return obj, time.time() - t0, res.stderr


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    if not srcs:
        raise RuntimeError("no CUDA sources under %s" % CSRC)
    objs, rebuilt = [], False
    with cf.ThreadPoolExecutor(max_workers=min(len(srcs), os.cpu_count() or 4)) as pool:
        for obj, dt, log in pool.map(lambda s: _compile(s, force, verbose), srcs):
            objs.append(obj)
            if dt > 0:
                rebuilt = True
                print("[mvs_b200 build] %-18s %.1fs" % (os.path.basename(obj), dt))
                if verbose and log:
                    print(log)
    if rebuilt or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        cmd = [_nvcc(), "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (res.stdout, res.stderr))
        print("[mvs_b200 build] linked %s" % LIB)
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    sys.stdout.write(build(a.force, a.verbose) + "\n")
