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


def incremental(snap, flags, label):
    """Bucket pipeline + device-side incremental epochs (kr_incr.cuh): pod rows, object rows, a structural change, unfetched passes."""
    flags.fetch_pod_lists = 0
    eng = Engine.for_snapshot(snap, slack=1.2)
    eng.set_fixed_layout(True)
    try:
        views = eng.begin(snap.sizes())
        eng.fill(views, snap)
        eng.commit()
        eng.reconcile(flags)
        pod_cols = [c for c, _d, _m, dim in abi.COLUMNS if dim == "pods"]
        rng = np.random.default_rng(1)
        changed = 0
        for epoch in range(4):
            rows = np.unique(rng.integers(0, snap.dims["pods"], 60)).astype(np.uint32)
            views["p_packed"][rows] ^= np.uint32(1 << 5)
            gone = rows[:5]
            for c in pod_cols:
                views[c][gone] = 0
            views["p_packed"][gone] = np.uint32(abi.PP_TOMBSTONE)
            if snap.dims["groups"]:
                views["g_replicas"][epoch % snap.dims["groups"]] += 1
            eng.commit(abi.PART_OBJECTS)
            if epoch % 2:
                eng.commit_pod_rows(rows)
            else:
                eng.commit_pod_values(rows, np.stack([views[c][rows].view(np.uint32) for c in pod_cols], axis=1))
            if epoch == 2:
                eng.reconcile_device_only(flags)
                res = eng.fetch()
            else:
                res = eng.reconcile(flags)
            changed += int(res.n_changed) if res.changed_clusters is not None else 0
        if snap.dims["groups"]:
            views["g_name_id"][0] += 12345   # a table key: the next pass is a full one
            eng.commit(abi.PART_OBJECTS)
            res = eng.reconcile(flags)
        print(label, "ok:", res.n_actions, "actions;", changed, "records recomputed incrementally", flush=True)
    finally:
        eng.close()


def main():
    incremental(*synthetic.generate(synthetic.config("C2", n_clusters=300, pods_per_cluster=20, groups=2, jobs=True, wtd_group_frac=0.3)), "incremental epochs")
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
