"""One-process GPU check of the training path: the parity tests that exercise every training kernel, then the
per-entry-point CUDA-event profile of a cfg-5-shaped step (scripts/train_profile.py).  Usage on the GPU box:
    python scripts/gpu_train_check.py gpurun_out/t"""
import contextlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.chdir(ROOT)
out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/t"
os.makedirs(out, exist_ok=True)

import pytest  # noqa: E402

with open(os.path.join(out, "train_tests_final.log"), "w") as f, contextlib.redirect_stdout(f), contextlib.redirect_stderr(f):
    rc = pytest.main(["tests/test_gpu_train.py", "-q", "-k", "conv_bn_act or stagenet_training or aggregate or group_corr",
                      "-p", "no:cacheprovider"])
print("pytest rc", int(rc))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import train_profile  # noqa: E402

with open(os.path.join(out, "train_profile_final.json"), "w") as f, contextlib.redirect_stdout(f):
    train_profile.main()
print(open(os.path.join(out, "train_profile_final.json")).read()[:1800])
