#!/usr/bin/env python
"""A few small passes over every kernel family, for `compute-sanitizer --tool {memcheck,racecheck,synccheck,initcheck}`
(development aid).  Usage on the GPU box:  compute-sanitizer --tool racecheck python tools/sanitize_run.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from kuberay_b200 import abi, synthetic  # noqa: E402
from kuberay_b200.engine import Engine  # noqa: E402


def one(snap, flags, label):
    eng = Engine.for_snapshot(snap)
    try:
        views = eng.load(snap)
        res = eng.reconcile(flags)
        rows = np.arange(0, snap.dims["pods"], 7, dtype=np.uint32)
        views["p_packed"][rows] ^= np.uint32(1 << 5)
        eng.commit_pod_rows(rows)
        eng.reconcile(flags)
        vals = np.stack([views[c][rows].view(np.uint32) for c, _d, _m, dim in abi.COLUMNS if dim == "pods"], axis=1)
        eng.commit(abi.PART_OBJECTS)
        eng.commit_pod_values(rows, vals)
        res = eng.reconcile(flags)
        print(label, "ok:", res.n_actions, "actions", res.n_create_total, "creates", flush=True)
    finally:
        eng.close()


def main():
    one(*synthetic.generate(synthetic.config("C2", n_clusters=200, jobs=True)), "fast pipeline")
    one(*synthetic.generate(synthetic.SynthParams(n_clusters=60, pods_per_cluster=41, groups=2, multihost_frac=0.5)), "multi-host")
    one(*synthetic.generate(synthetic.SynthParams(n_clusters=20, pods_per_cluster=200, groups=40)), "many groups")
    one(*synthetic.generate(synthetic.SynthParams(n_clusters=3, pods_per_cluster=1500, groups=2)), "big bucket -> radix")
    os.environ["KR_NO_FUSE"] = "1"
    one(*synthetic.generate(synthetic.config("C2", n_clusters=200)), "unfused scans")
    import fuzz_objects
    for seed in range(6):
        snap, flags = fuzz_objects.snapshot(seed, big=True)
        one(snap, flags, f"fuzz {seed}")
    eng = Engine(0, 1, 1, 1, 1, 1, 1, 16, 4096)
    msgs = [bytes([65 + i % 26]) * n for i, n in enumerate((0, 1, 55, 56, 63, 64, 65, 119, 120, 128, 1000, 4097))]
    print("hash_batch", len(eng.hash_batch(msgs)), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
