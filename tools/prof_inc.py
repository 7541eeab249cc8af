#!/usr/bin/env python
"""A few device-side incremental epochs over one workload, for `ncu` (development aid) — the pod churn of bench.py's
e2e_incremental_1pct_pod_churn leg (1 % of the pods touched per epoch) on top of a resident full pass:
  ncu --set full --clock-control none -k regex:'k_inc|k_decide2' -o /tmp/inc python tools/prof_inc.py C3"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kuberay_b200 import abi, synthetic  # noqa: E402
from kuberay_b200.engine import Engine  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "C3"
snap, flags = synthetic.generate(synthetic.config(wl))
flags.fetch_pod_lists = 0
eng = Engine.for_snapshot(snap)
views = eng.load(snap)
eng.reconcile(flags, copy=False)
rng = np.random.default_rng(5)
npods = snap.dims["pods"]
pod_cols = [c for c, _d, _m, dim in abi.COLUMNS if dim == "pods"]
for epoch in range(int(os.environ.get("EPOCHS", "3"))):
    rows = np.unique(rng.integers(0, npods, npods // 100)).astype(np.uint32)
    views["p_packed"][rows] ^= np.uint32(1 << 5)
    eng.commit(abi.PART_OBJECTS)
    eng.commit_pod_values(rows, np.stack([views[c][rows].view(np.uint32) for c in pod_cols], axis=1))
    prof = eng.reconcile_profiled(flags)
    res = eng.fetch(copy=False)
    print(epoch, "changed", res.n_changed, [(k, round(v * 1e3, 1)) for k, v in prof["kernels"]], flush=True)
eng.close()
