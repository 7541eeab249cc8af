#!/usr/bin/env python
"""A few serialised passes over one workload, for `ncu` (development aid):
  ncu --set full --clock-control none --import-source on -k regex:'k_match2|k_decide2' -s 6 -c 6 -o gpurun_out/prof python tools/prof_pass.py C3"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from kuberay_b200 import synthetic  # noqa: E402
from kuberay_b200.engine import Engine  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "C3"
snap, flags = synthetic.generate(synthetic.config(wl))
flags.fetch_pod_lists = int(os.environ.get("TL_POD_LISTS", "0"))
eng = Engine.for_snapshot(snap)
eng.set_incremental(False)  # every pass here is the FULL pass
eng.load(snap)
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for _ in range(int(os.environ.get("PASSES", "4"))):
    flush.zero_(); torch.cuda.synchronize()
    print(eng.reconcile_profiled(flags)["kernels"])
eng.close()
