"""Recipe for ``oracle/_ref/``: the UNMODIFIED reference files of the hot path, staged where the GPU box can
see them.  TEST / BENCH INFRASTRUCTURE — the product package never imports anything under ``oracle/``.

``/root/reference`` exists only in the build container.  ``oracle/_ref/`` is git-ignored (reference sources
never enter this repository's history) but NOT gpurun-ignored, so the staged copy travels to the GPU box with
the snapshot, like a compiled ``.so`` would for a C reference.  Staged: ``models/*.py`` (warping.py, module.py,
mvsformer_model.py and the two backbone files mvsformer_model.py imports at module level) and ``utils.py``
(imported by models/vision_transformer.py:25).  Nothing is edited; ``oracle/ref_import.py`` loads them with the
same two inert stand-ins (``timm``, ``omegaconf``) it uses for ``/root/reference``.

    python -m oracle.build_ref          # called by __graft_entry__.build() when /root/reference is present

Used by: ``bench.py --impl reference`` (CPU arm, kind "reference"), ``bench.py``'s ``gpu_eager_baseline`` leg
(the same reference modules moved to cuda:0 = "PyTorch eager + cuDNN on the same box") and the
cross-checks in ``tests/``.
"""
import filecmp
import glob
import os
import shutil

SRC = "/root/reference"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
FILES = ["utils.py"] + ["models/" + n for n in ("__init__.py", "warping.py", "module.py", "mvsformer_model.py", "gvt.py",
                                                "vision_transformer.py")]


def staged():
    return os.path.isfile(os.path.join(DST, "models", "mvsformer_model.py"))


def build(verbose=True):
    if not os.path.isfile(os.path.join(SRC, "models", "mvsformer_model.py")):
        if verbose:
            print("[oracle/_ref] %s absent; keeping the staged copy (%s)" % (SRC, "present" if staged() else "missing"))
        return staged()
    os.makedirs(os.path.join(DST, "models"), exist_ok=True)
    for rel in FILES:
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        if not (os.path.exists(dst) and filecmp.cmp(src, dst, shallow=False)):
            shutil.copyfile(src, dst)
    stale = [p for p in glob.glob(os.path.join(DST, "**", "*.py"), recursive=True)
             if os.path.relpath(p, DST).replace(os.sep, "/") not in FILES]
    for p in stale:
        os.remove(p)
    if verbose:
        print("[oracle/_ref] staged %d unmodified reference files from %s" % (len(FILES), SRC))
    return True


if __name__ == "__main__":
    build()
