"""mvsformer_b200 — B200-native (sm_100a) plane-sweep MVS depth engine.

Drop-in for the per-reference-view cascade of ewrfcas/MVSFormer (StageNet, the CostRegNet family,
the warping and scheduling functions); everything runs in libmvs_b200.so (see include/mvs_b200.h).
Importing the package does not load the library; the first kernel call does, and raises if the
library is not built.  There is no CPU or PyTorch fallback.
"""
__version__ = "0.1.0"
