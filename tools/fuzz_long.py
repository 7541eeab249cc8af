#!/usr/bin/env python
"""A longer run of the adversarial differential fuzz than the test suite affords (tests/test_gpu_parity.py::test_fuzz_adversarial_snapshots
covers seeds 0-399): engine vs oracle, byte for byte, full and compact results, plus one incremental epoch per snapshot.
usage (GPU box): python tools/fuzz_long.py [first_seed] [count]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fuzz_objects  # noqa: E402
from kuberay_b200 import abi  # noqa: E402
from kuberay_b200.engine import Engine  # noqa: E402
from oracle import oracle  # noqa: E402

first = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
count = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
POD_COLS = [c for c, _d, _m, dim in abi.COLUMNS if dim == "pods"]
n_inc = 0
for seed in range(first, first + count):
    snap, flags = fuzz_objects.snapshot(seed, big=(seed % 10 == 0))
    want = oracle.run(snap, flags, threads=1)
    eng = Engine.for_snapshot(snap, slack=1.5)
    try:
        eng.set_fixed_layout(True)
        views = eng.begin(snap.sizes())
        eng.fill(views, snap)
        eng.commit()
        d = want.diff(eng.reconcile(flags))
        assert not d, (seed, d[:6])
        lean = abi.kr_flags.from_buffer_copy(flags)
        lean.fetch_pod_lists = 0
        d = want.diff(eng.reconcile(lean))
        assert not d, ("compact", seed, d[:6])
        if snap.dims["pods"]:  # one incremental epoch: flip PodReady on a few rows
            rows = np.unique(np.random.default_rng(seed).integers(0, snap.dims["pods"], 3)).astype(np.uint32)
            snap.cols["p_packed"][rows] ^= np.uint32(1 << abi.PP_READY_SHIFT)
            for c in POD_COLS:
                views[c][rows] = snap.cols[c][rows]
            eng.commit_pod_values(rows, np.stack([snap.cols[c][rows].view(np.uint32) for c in POD_COLS], axis=1))
            got = eng.reconcile(lean)
            d = oracle.run(snap, lean, threads=1).diff(got)
            assert not d, ("incremental", seed, d[:6])
            n_inc += got.changed_clusters is not None
    finally:
        eng.close()
print(f"fuzz ok: seeds {first}..{first + count - 1}, {n_inc} of the follow-up epochs were incremental on the device")
