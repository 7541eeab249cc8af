"""Incrementally maintained arena: the host-side mirror of what the shim's informer handlers do between epochs.

SURVEY §8(f) rank 1.  The device keeps the last snapshot; an epoch uploads only what changed:
  * Pod Add / Update / Delete events rewrite single rows of the seven per-pod columns (a deleted Pod's row becomes a
    KR_PP_TOMBSTONE row, an added Pod takes the lowest free row) -> kr_snapshot_commit_pod_rows(rows);
  * RayCluster / worker-group / workersToDelete / head-aux / RayJob rows are small -> kr_snapshot_commit_parts(KR_PART_OBJECTS);
  * spec changes -> KR_PART_JSON as well.
A change that moves a table's row count (a RayCluster or a head Pod appears / disappears, the workersToDelete lists change
length, the arena runs out of free rows) changes the column layout and takes the full begin + commit path ("rebase").

This class re-packs the objects on the host with the ordinary packer and diffs the columns: it is the executable
statement of the protocol for the tests, not a fast packer (that is the Go shim's job).
"""
from __future__ import annotations

import heapq

import numpy as np

from . import abi
from . import snapshot as snp
from .engine import Engine

_POD_COLS = [name for name, _dt, _m, dim in abi.COLUMNS if dim == "pods"]
_OBJ_COLS = [name for name, _dt, _m, dim in abi.COLUMNS if dim not in ("pods", "json")]


def _is_head(pod: dict) -> bool:
    return (pod.get("labels") or {}).get(snp.RAY_NODE_TYPE_LABEL) == "head"


class LiveArena:
    def __init__(self, clusters: list[dict], pods: list[dict], jobs: list[dict] | None = None, spare_rows: int = 64, device: int = 0,
                 engine: bool = True):
        self.clusters = {(c.get("namespace", "default"), c["name"]): c for c in clusters}
        self.jobs = list(jobs or [])
        self.rows: list[dict | None] = list(pods) + [None] * spare_rows
        self.interner = snp.Interner()
        self.device = device
        self.use_engine = engine
        self.engine: Engine | None = None
        self.stats = {"rebase": 0, "incremental": 0, "rows": 0}
        self._need_rebase = True
        self._dirty: set[int] = set()
        self._index()

    # ------------------------------------------------------------------ events
    def _index(self):
        self.row_of = {(p.get("namespace", "default"), p["name"]): i for i, p in enumerate(self.rows) if p is not None}
        self.free = [i for i, p in enumerate(self.rows) if p is None]
        heapq.heapify(self.free)

    def upsert_pod(self, pod: dict):
        key = (pod.get("namespace", "default"), pod["name"])
        row = self.row_of.get(key)
        if row is None:
            if _is_head(pod) or not self.free:
                if not self.free:
                    self.rows.extend([None] * max(64, len(self.rows) // 8))  # grow the arena: layout change
                    self._index()
                self._need_rebase = True  # a new head-aux row (n_heads) or new capacity (n_pods)
            row = heapq.heappop(self.free)
            self.row_of[key] = row
        elif _is_head(pod) != _is_head(self.rows[row]):
            self._need_rebase = True
        self.rows[row] = pod
        self._dirty.add(row)

    def delete_pod(self, namespace: str, name: str) -> bool:
        row = self.row_of.pop((namespace, name), None)
        if row is None:
            return False
        if _is_head(self.rows[row]):
            self._need_rebase = True
        self.rows[row] = None
        heapq.heappush(self.free, row)
        self._dirty.add(row)
        return True

    def upsert_cluster(self, cluster: dict):
        key = (cluster.get("namespace", "default"), cluster["name"])
        if key not in self.clusters:
            self._need_rebase = True
        self.clusters[key] = cluster

    def delete_cluster(self, namespace: str, name: str):
        if self.clusters.pop((namespace, name), None) is not None:
            self._need_rebase = True

    # ------------------------------------------------------------------ epoch
    def pack(self) -> tuple[snp.Snapshot, snp.PackMeta]:
        pods = [p if p is not None else snp.TOMBSTONE for p in self.rows]
        return snp.pack_objects([self.clusters[k] for k in sorted(self.clusters)], pods, self.jobs, interner=self.interner)

    def fresh_pack(self) -> tuple[snp.Snapshot, snp.PackMeta]:
        """The same objects packed from scratch without free rows (same relative List order): the semantic reference."""
        pods = [p for p in self.rows if p is not None]
        return snp.pack_objects([self.clusters[k] for k in sorted(self.clusters)], pods, self.jobs, interner=self.interner)  # same ids

    def flush(self) -> str:
        """Bring the device copy up to date; returns "rebase" or "incremental"."""
        snap, meta = self.pack()
        same_layout = (not self._need_rebase) and getattr(self, "snap", None) is not None and bytes(snap.sizes()) == bytes(self.snap.sizes())
        if not self.use_engine:
            mode = "incremental" if same_layout else "rebase"
        elif not same_layout:
            if self.engine is None or not self._fits(snap):
                if self.engine is not None:
                    self.engine.close()
                self.engine = Engine.for_snapshot(snap, device=self.device, slack=1.5)
            self.views = self.engine.begin(snap.sizes())
            self.engine.fill(self.views, snap)
            self.engine.commit()
            mode = "rebase"
        else:
            parts = 0
            if any(not np.array_equal(self.views[c], snap.cols[c]) for c in _OBJ_COLS):
                for c in _OBJ_COLS:
                    np.copyto(self.views[c], snap.cols[c])
                parts |= abi.PART_OBJECTS
            if not np.array_equal(self.views["json"], snap.cols["json"]):
                np.copyto(self.views["json"], snap.cols["json"])
                parts |= abi.PART_JSON
            if parts:
                self.engine.commit(parts)
            rows = np.array(sorted(self._dirty), dtype=np.uint32)
            self._epoch = getattr(self, "_epoch", 0) + 1
            if self._epoch % 2 and rows.size:
                # journal style: hand the rows over (and keep the pinned arenas current for a later full commit)
                vals = np.stack([snap.cols[c][rows].view(np.uint32) for c in _POD_COLS], axis=1)
                for c in _POD_COLS:
                    self.views[c][rows] = snap.cols[c][rows]
                self.engine.commit_pod_values(rows, vals)
            else:
                for c in _POD_COLS:
                    self.views[c][rows] = snap.cols[c][rows]
                for c in _POD_COLS:  # every row that was not reported dirty must already be identical
                    assert np.array_equal(self.views[c], snap.cols[c]), c
                if rows.size:
                    self.engine.commit_pod_rows(rows)
            self.stats["rows"] += int(rows.size)
            mode = "incremental"
        self.snap, self.meta = snap, meta
        self._need_rebase = False
        self._dirty.clear()
        self.stats[mode] += 1
        return mode

    def _fits(self, snap) -> bool:
        d, c = snap.dims, self.engine.cfg
        return (d["clusters"] <= c.max_clusters and d["groups"] <= c.max_groups and d["wtd"] <= c.max_wtd and d["pods"] <= c.max_pods and
                d["heads"] <= c.max_heads and d["jobs"] <= c.max_jobs and d["json"] <= c.max_json_bytes)

    def reconcile(self, flags: abi.kr_flags | None = None) -> abi.Results:
        return self.engine.reconcile(flags or self.meta.flags)

    def close(self):
        if self.engine is not None:
            self.engine.close()
            self.engine = None
