#!/usr/bin/env python
"""Timeline of one graph replay of the pass (development aid).

Build:  make -C kuberay_b200/csrc ../../tools/libkrengine_tl.so      (adds -DKR_TIMELINE: %globaltimer stamps per kernel)
Run  :  KR_ENGINE_LIB=tools/libkrengine_tl.so python tools/timeline.py [workload]   (on the GPU box)
Prints, for each kernel of the pass, first-block start and last-block end relative to the earliest stamp, averaged over
the replays, with the L2 flushed before each replay.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from kuberay_b200 import synthetic  # noqa: E402
from kuberay_b200.engine import Engine, lib  # noqa: E402

NAMES = {10: "  creates: scans done", 11: "  creates: fills done", 9: "k_clear", 0: "k_build_tables", 1: "k_match / k_match2", 2: "k_place_fused", 3: "k_decide_small / k_decide2", 4: "k_decide<general>", 5: "k_decide<phase 1>",
         12: "decide phase 1 (small / k_decide2)", 6: "k_creates_fused", 7: "k_hash", 8: "k_jobs"}


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "C3"
    snap, flags = synthetic.generate(synthetic.config(wl))
    flags.fetch_pod_lists = int(os.environ.get("TL_POD_LISTS", "0"))  # 0: production configuration (bucket pipeline)
    flags.skip_hash = int(os.environ.get("TL_SKIP_HASH", "0"))
    eng = Engine.for_snapshot(snap)
    eng.set_incremental(False)  # every pass here is the FULL pass
    eng.load(snap)
    L = lib()
    L.kr_debug_timeline.argtypes = [C.c_void_p, C.c_void_p]
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    acc = np.zeros((32, 2))
    reps, tot = 20, 0.0
    for i in range(reps + 3):
        flush.zero_(); torch.cuda.synchronize()
        eng.reconcile_device_only(flags)
        out = np.zeros(64, dtype=np.uint64)
        L.kr_debug_timeline(eng._h, out.ctypes.data)
        t = out.reshape(32, 2).astype(np.float64)
        used = t[:, 1] > 0
        t0 = t[used, 0].min()
        if i >= 3:
            acc[used] += (t[used] - t0) / 1e3
            tot += eng.last_profile()["kernels_ms"] * 1e3
    acc /= reps
    print(f"{wl}: pass {tot / reps:.1f} us (CUDA events)")
    for k in sorted(NAMES, key=lambda k: acc[k, 0] if acc[k, 1] else 1e18):
        if acc[k, 1]:
            print(f"  {NAMES[k]:36s} start {acc[k, 0]:7.1f} us   end {acc[k, 1]:7.1f} us   ({acc[k, 1] - acc[k, 0]:6.1f} us)")
    eng.close()


if __name__ == "__main__":
    main()
