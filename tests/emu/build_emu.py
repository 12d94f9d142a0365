"""Builds tests/emu/libmvs_emu.so (CPU emulation of the training-path kernels; test-only)."""
import glob
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB = os.path.join(HERE, "libmvs_emu.so")


def build(force=False):
    deps = [os.path.join(HERE, "emu.cpp"), os.path.join(ROOT, "include", "mvs_b200.h")]
    deps += glob.glob(os.path.join(ROOT, "mvsformer_b200", "csrc", "train_*")) + \
        glob.glob(os.path.join(ROOT, "mvsformer_b200", "csrc", "fusion_*")) + \
        [os.path.join(ROOT, "mvsformer_b200", "csrc", "geometry.cuh")]
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) > max(os.path.getmtime(d) for d in deps):
        return LIB
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-x", "c++", deps[0], "-o", LIB, "-lm"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ failed:\n%s\n%s" % (res.stdout, res.stderr))
    return LIB


if __name__ == "__main__":
    print(build(force=True))
