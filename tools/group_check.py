#!/usr/bin/env python
"""kr_group on N real GPUs (run under `gpurun --gpus N`): C5 — 1 000 autoscaling RayClusters x 100 pods (100 k pods), UID-hash
sharded over every visible device by the native coordinator, each shard checked against the oracle, the per-group delta records
all-gathered over NCCL; prints one JSON line with the parallel epoch time."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kuberay_b200 import abi, synthetic  # noqa: E402
from kuberay_b200.engine import Group, lib  # noqa: E402
from oracle import oracle  # noqa: E402

n = lib().kr_device_count()
wl = sys.argv[1] if len(sys.argv) > 1 else "C5"
snap, flags = synthetic.generate(synthetic.config(wl, wtd_group_frac=0.3) if wl == "C5" else synthetic.config(wl))
flags.fetch_pod_lists = 0
d = snap.dims
cap = abi.kr_config(0, d["clusters"] + 1, d["groups"] + 1, d["wtd"] + 1, d["pods"] + 1, d["heads"] + 1, d["jobs"] + 1, max(1024, d["pods"]), d["json"] + 64)
grp = Group(cap, list(range(n)))
t0 = time.perf_counter(); sizes, *_ = grp.route(snap); route_ms = 1e3 * (time.perf_counter() - t0)
grp.commit(); res = grp.reconcile(flags)
ok = True
for r in range(n):
    sh = synthetic.shard_by_uid(snap, r, n)
    ok &= not oracle.run(sh, flags, threads=8).diff(res[r])
gathered, slot, used_nccl = grp.allgather_group_results()
steps = 20
t0 = time.perf_counter()
for _ in range(steps):
    grp.commit(); grp.reconcile(flags, copy=False)
epoch_ms = 1e3 * (time.perf_counter() - t0) / steps
t0 = time.perf_counter()
for _ in range(steps):
    grp.allgather_group_results()
gather_ms = 1e3 * (time.perf_counter() - t0) / steps
grp.close()
print(json.dumps({"workload": wl, "n_gpus": n, "clusters": d["clusters"], "pods": d["pods"], "parity_every_shard": bool(ok), "route_ms": route_ms,
                  "epoch_ms_commit_plus_reconcile": epoch_ms, "reconciles_per_s_e2e": d["clusters"] / (epoch_ms / 1e3),
                  "allgather_ms": gather_ms, "allgather_used_nccl": bool(used_nccl), "allgather_slot_bytes": slot,
                  "max_over_mean_shard_pods": max(s.n_pods for s in sizes) / (d["pods"] / n)}))
sys.exit(0 if ok else 1)
