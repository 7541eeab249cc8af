"""kuberay_b200 — Blackwell-native batched reconcile engine for KubeRay's RayCluster controller hot path.

Only what the path needs lives here:
  csrc/        CUDA kernels (sm_100a) + the C-ABI library (libkrengine.so, include/kr_engine.h)
  abi.py       ctypes/numpy mirror of include/kr_engine.h
  snapshot.py  columnar snapshot container + string interner + object->column packer
  synthetic.py deterministic synthetic snapshots
  engine.py    ctypes binding of the C ABI (the product path; raises if the CUDA library is missing)
  reconciler.py host-side mirror of the reference RayClusterReconciler that consumes engine records

The CPU oracle (oracle/) is test infrastructure and is never imported from this package.
"""
__all__ = ["abi", "snapshot", "synthetic", "engine"]
