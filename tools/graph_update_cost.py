#!/usr/bin/env python
"""Cost of an epoch whose row counts differ from the previous one under KR_OPT_FIXED_LAYOUT (development aid):
kr_snapshot_begin(new counts) + stream re-capture + cudaGraphExecUpdate, against an epoch with unchanged counts."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kuberay_b200 import abi, synthetic  # noqa: E402
from kuberay_b200.engine import Engine  # noqa: E402


def main():
    snap, flags = synthetic.generate(synthetic.config(sys.argv[1] if len(sys.argv) > 1 else "C3"))
    flags.fetch_pod_lists = 0
    eng = Engine.for_snapshot(snap, slack=1.1)
    eng.set_fixed_layout(True)
    eng.load(snap)
    eng.reconcile(flags, copy=False)
    sizes = snap.sizes()
    for label, vary in (("same counts", False), ("n_pods changes every epoch", True)):
        ts = []
        for i in range(40):
            if vary:
                sizes.n_pods = snap.dims["pods"] - (i & 1)  # drop / restore the last pod row
            t0 = time.perf_counter()
            eng.begin(sizes)
            eng.commit(abi.PART_OBJECTS)
            eng.reconcile(flags, copy=False)
            ts.append(time.perf_counter() - t0)
        print(f"{label:32s} {1e3 * np.median(ts[5:]):.3f} ms per epoch (begin + KR_PART_OBJECTS + reconcile_batch)")
    eng.close()


if __name__ == "__main__":
    main()
