#!/usr/bin/env python
"""A longer run of the native packer's random informer-event streams than tests/test_packer.py affords (4 seeds x 10 epochs there): per seed a
fuzz-generated object set, then epochs of mixed Pod / RayCluster / RayJob events — structural ones included — through kr_packer_*, every
epoch compared with the oracle on an independently re-packed snapshot (the test's own Mirror / check).  usage (GPU box):
python tools/packer_soak.py [first_seed] [seeds] [epochs]"""
import copy
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fuzz_objects  # noqa: E402
import test_packer as tp  # noqa: E402
from kuberay_b200 import abi  # noqa: E402
from kuberay_b200.packer import Packer  # noqa: E402
from oracle import oracle  # noqa: E402

first = int(sys.argv[1]) if len(sys.argv) > 1 else 10
seeds = int(sys.argv[2]) if len(sys.argv) > 2 else 40
epochs = int(sys.argv[3]) if len(sys.argv) > 3 else 30
oracle.lib()
total = inc = 0
for seed in range(first, first + seeds):
    rng = np.random.default_rng(seed)
    clusters, pods, jobs = fuzz_objects.generate(seed, big=True)
    for i, c in enumerate(clusters):
        c["generation"], c["resourceVersion"] = 1, 100 + i
    for i, j in enumerate(jobs):
        j.setdefault("name", f"rayjob-{i}")
    pk = Packer(max_clusters=64, max_groups=512, max_wtd=512, max_pods=8192, max_heads=256, max_jobs=64, max_creates=1 << 16, max_json_bytes=4 << 20)
    try:
        m = tp.Mirror(copy.deepcopy(clusters), copy.deepcopy(pods), jobs, pk)
        assert pk.flush() == abi.PACK_FULL
        tp.check(m, oracle, lean=True)
        counter = [0]
        for epoch in range(epochs):
            tp._events(rng, m, counter, structural=True)
            mode = pk.flush()
            assert not mode & abi.PACK_FULL, (seed, epoch)
            tp.check(m, oracle, lean=bool(epoch % 3))
            total += 1
            inc += bool(mode & abi.PACK_POD_ROWS)
    finally:
        pk.close()
print(f"packer soak ok: seeds {first}..{first + seeds - 1} x {epochs} epochs = {total} epochs ({inc} with pod-row commits), every one equal to the oracle")
