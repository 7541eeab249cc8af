"""ctypes binding of the native event-driven packer (kr_packer_*, kuberay_b200/csrc/kr_packer.cpp) + the adapter that turns the
dict objects of the test fixtures into the plain C structs a Go informer handler would fill from *corev1.Pod / *rayv1.RayCluster.

The adapter only extracts fields (the same cheap per-object derivations pack_objects does: flag bits, condition codes); interning,
row management, the CSR tables, the spec JSON and the incremental commits are the native packer's."""
from __future__ import annotations

import ctypes as C
import json

from . import abi
from . import snapshot as snp
from .engine import Engine, EngineError, lib


def _s(v) -> abi.kr_str:
    if v is None:
        return abi.kr_str(None, 0)
    b = v if isinstance(v, bytes) else str(v).encode("utf-8")
    return abi.kr_str(b, len(b))


class Packer:
    def __init__(self, device=0, max_clusters=1024, max_groups=4096, max_wtd=4096, max_pods=65536, max_heads=2048, max_jobs=1024,
                 max_creates=65536, max_json_bytes=64 << 20):
        L = self._L = lib()
        P = C.POINTER
        L.kr_packer_create.argtypes = [P(abi.kr_config), P(C.c_void_p)]
        L.kr_packer_destroy.argtypes = [C.c_void_p]; L.kr_packer_destroy.restype = None
        L.kr_packer_engine.argtypes = [C.c_void_p]; L.kr_packer_engine.restype = C.c_void_p
        L.kr_packer_pod_upsert.argtypes = [C.c_void_p, P(abi.kr_pod_obj)]
        L.kr_packer_pod_delete.argtypes = [C.c_void_p, abi.kr_str, abi.kr_str]
        L.kr_packer_cluster_upsert.argtypes = [C.c_void_p, P(abi.kr_cluster_obj)]
        L.kr_packer_cluster_delete.argtypes = [C.c_void_p, abi.kr_str, abi.kr_str]
        L.kr_packer_job_upsert.argtypes = [C.c_void_p, P(abi.kr_job_obj)]
        L.kr_packer_job_delete.argtypes = [C.c_void_p, abi.kr_str, abi.kr_str]
        L.kr_packer_flush.argtypes = [C.c_void_p, P(C.c_uint32)]
        L.kr_packer_sizes.argtypes = [C.c_void_p, P(abi.kr_sizes)]
        L.kr_packer_bufs.argtypes = [C.c_void_p, P(abi.kr_snapshot_bufs)]
        L.kr_packer_intern.argtypes = [C.c_void_p, abi.kr_str]; L.kr_packer_intern.restype = C.c_uint32
        L.kr_packer_string.argtypes = [C.c_void_p, C.c_uint32, P(abi.kr_str)]
        L.kr_packer_cluster_row.argtypes = [C.c_void_p, abi.kr_str, abi.kr_str]; L.kr_packer_cluster_row.restype = C.c_int64
        L.kr_packer_pod_row.argtypes = [C.c_void_p, abi.kr_str, abi.kr_str]; L.kr_packer_pod_row.restype = C.c_int64
        L.kr_packer_pod_key.argtypes = [C.c_void_p, C.c_uint32, P(abi.kr_str), P(abi.kr_str)]
        L.kr_packer_epoch.argtypes = [C.c_void_p, P(C.c_uint64), P(C.c_uint64)]
        L.kr_packer_cluster_epoch.argtypes = [C.c_void_p, C.c_uint32, P(C.c_uint64), P(C.c_uint64)]
        L.kr_packer_last_error.argtypes = [C.c_void_p]; L.kr_packer_last_error.restype = C.c_char_p
        if L.kr_device_count() <= 0:
            raise EngineError(abi.KR_E_NO_DEVICE, "no CUDA device visible (this engine has no CPU fallback)")
        cfg = abi.kr_config(device, max_clusters, max_groups, max_wtd, max_pods, max_heads, max_jobs, max_creates, max_json_bytes)
        self._h = C.c_void_p()
        rc = L.kr_packer_create(C.byref(cfg), C.byref(self._h))
        if rc != 0:
            raise EngineError(rc, "kr_packer_create failed")
        self.engine = Engine.__new__(Engine)  # a view over the packer's engine (not owned)
        self.engine._L, self.engine._h, self.engine.sizes, self.engine.cfg = L, C.c_void_p(L.kr_packer_engine(self._h)), abi.kr_sizes(), cfg

    def _check(self, rc):
        if rc != 0:
            raise EngineError(rc, self._L.kr_packer_last_error(self._h).decode())

    def close(self):
        if self._h:
            self.engine._h = C.c_void_p()
            self._L.kr_packer_destroy(self._h)
            self._h = C.c_void_p()

    # ------------------------------------------------------------------ events (dict objects: tests/golden/README.md)
    def upsert_pod(self, pod: dict):
        labels, ann = pod.get("labels") or {}, pod.get("annotations") or {}
        o = abi.kr_pod_obj()
        o.ns, o.name = _s(pod.get("namespace", "default")), _s(pod["name"])
        o.cluster, o.group = _s(labels.get(snp.RAY_CLUSTER_LABEL)), _s(labels.get(snp.RAY_NODE_GROUP_LABEL))
        o.replica_name, o.replica_index = _s(labels.get(snp.REPLICA_NAME_LABEL)), _s(labels.get(snp.REPLICA_INDEX_LABEL))
        o.node_type = snp._NODE_TYPE.get(labels.get(snp.RAY_NODE_TYPE_LABEL, ""), abi.NT_NONE)
        o.phase = snp._PHASE.get(pod.get("phase", ""), abi.PHASE_UNKNOWN)
        o.ready_cond = snp.pod_ready_code(pod)
        o.restart_never = 1 if pod.get("restartPolicy") == "Never" else 0
        o.ray_terminated = 1 if snp.ray_container_terminated(pod) else 0
        o.has_deletion_ts = 1 if pod.get("deletionTimestamp") else 0
        if o.node_type == abi.NT_HEAD:
            st, reason, msg = snp.head_pod_ready_condition(pod)
            o.head_ready_status = snp._COND.get(st, abi.COND_UNKNOWN) if st != "" else abi.COND_UNKNOWN
            o.head_ready_reason, o.head_ready_msg = _s(reason), _s(msg)
            o.pod_ip = _s(pod.get("podIP"))
            o.recreate_hash, o.kuberay_version = _s(ann.get(snp.RECREATE_HASH_ANNOT, "")), _s(ann.get(snp.KUBERAY_VERSION_ANNOT, ""))
        self._check(self._L.kr_packer_pod_upsert(self._h, C.byref(o)))

    def delete_pod(self, ns: str, name: str):
        self._check(self._L.kr_packer_pod_delete(self._h, _s(ns), _s(name)))

    def upsert_cluster(self, c: dict):
        spec, status = c.get("spec") or {}, c.get("status") or {}
        ns, name = c.get("namespace", "default"), c["name"]
        o = abi.kr_cluster_obj()
        o.ns, o.name, o.uid = _s(ns), _s(name), _s(c.get("uid"))
        o.resource_version, o.generation = int(c.get("resourceVersion", 0)), int(c.get("generation", 0))
        fl = 0
        if spec.get("suspend") is True:
            fl |= abi.CF_SUSPEND
        if spec.get("suspend") is False:
            fl |= abi.CF_SUSPEND_SET_FALSE
        if spec.get("enableInTreeAutoscaling") is True:
            fl |= abi.CF_AUTOSCALING
        us = spec.get("upgradeStrategy")
        if (us.get("type") if isinstance(us, dict) else us) == "Recreate":
            fl |= abi.CF_UPGRADE_RECREATE
        if (c.get("annotations") or {}).get(snp.SKIP_HEAD_RESTART_ANNOT) == "true":
            fl |= abi.CF_SKIP_HEAD_RESTART
        exp = c.get("expectations") or {}
        if exp.get("head", True):
            fl |= abi.CF_HEAD_EXPECT_OK
        if c.get("deletionTimestamp") or c.get("skip"):
            fl |= abi.CF_SKIP
        if status.get("reason", "") != "":
            fl |= abi.CF_OLD_REASON_NONEMPTY
        svc = c.get("headService", {"count": 1, "clusterIP": "10.0.0.1", "name": f"{name}-head-svc"})
        if snp.compute_endpoints(status.get("endpoints"), svc) != status.get("endpoints"):
            fl |= abi.CF_ENDPOINTS_CHANGED
        o.flags = fl
        o.suspend_status = snp.find_suspend_status(status.get("conditions"))
        ext = c.get("extErr") or {}
        o.ext_err_kind = ext.get("kind", 0)
        o.ext_err_msg = _s(ext["message"]) if "message" in ext else _s(None)
        o.old_state = snp._STATE.get(status.get("state", ""), abi.STATE_OTHER)
        for k, key in enumerate(("readyWorkerReplicas", "availableWorkerReplicas", "desiredWorkerReplicas", "minWorkerReplicas", "maxWorkerReplicas")):
            o.old_counts[k] = status.get(key, 0)
        hr_reason = hr_msg = rf_msg = None
        for cond in status.get("conditions") or []:
            slot = snp._COND_SLOT.get(cond.get("type"))
            if slot is None:
                continue
            o.old_cond_status[slot] = snp._COND.get(cond.get("status", ""), abi.COND_UNKNOWN)
            reason, msg = cond.get("reason", ""), cond.get("message", "")
            if slot == abi.COND_PROVISIONED:
                var = snp._PROV_VARIANTS.get((reason, msg), abi.CV_OTHER)
            elif slot in (abi.COND_SUSPENDING, abi.COND_SUSPENDED):
                var = abi.CV_CANONICAL if (reason == cond["type"] and msg == "") else abi.CV_OTHER
            elif slot == abi.COND_HEAD_POD_READY:
                var = abi.CV_HEAD_NOT_FOUND if (reason, msg) == (snp.HEAD_NOT_FOUND_REASON, snp.HEAD_NOT_FOUND_MSG) else abi.CV_HEAD_FROM_POD
                hr_reason, hr_msg = reason, msg
            else:
                var = snp._REPLICA_FAILURE_KIND.get(reason, abi.CV_OTHER)
                rf_msg = msg
            o.old_cond_variant[slot] = var
        o.old_head_ready_reason, o.old_head_ready_msg, o.old_replica_failure_msg = _s(hr_reason), _s(hr_msg), _s(rf_msg)
        head = status.get("head") or {}
        for k, key in enumerate(("podIP", "serviceIP", "podName", "serviceName")):
            o.old_head[k] = _s(head.get(key))
        o.svc_count = min(svc.get("count", 1), 2)
        ip = svc.get("clusterIP", "")
        o.svc_ip_kind = abi.SVCIP_EMPTY if ip == "" else (abi.SVCIP_NONE if ip == "None" else abi.SVCIP_NORMAL)
        o.svc_ip = _s(ip) if ip not in ("", "None") else _s(None)
        o.svc_name = _s(svc.get("name", ""))
        o.status_summary = _s(snp.status_summary_key(status))
        groups = spec.get("workerGroupSpecs") or []
        garr = (abi.kr_group_obj * max(len(groups), 1))()
        keep = []
        for gi, grp in enumerate(groups):
            g = garr[gi]
            g.name = _s(grp["groupName"])
            gf = 0
            for key, fld, nil in (("replicas", "replicas", abi.GF_REPLICAS_NIL), ("minReplicas", "min_replicas", abi.GF_MIN_NIL), ("maxReplicas", "max_replicas", abi.GF_MAX_NIL)):
                v = grp.get(key)
                if v is None:
                    gf |= nil
                else:
                    setattr(g, fld, v)
            g.num_hosts = grp.get("numOfHosts", 1)
            if grp.get("suspend") is True:
                gf |= abi.GF_SUSPEND
            if exp.get(grp["groupName"], True):
                gf |= abi.GF_EXPECT_OK
            g.flags = gf
            names = grp.get("workersToDelete") or (grp.get("scaleStrategy") or {}).get("workersToDelete") or []
            warr = (abi.kr_str * max(len(names), 1))(*[_s(n) for n in names])
            keep.append(warr)
            g.workers_to_delete, g.n_workers_to_delete = warr, len(names)
        o.groups, o.n_groups = garr, len(groups)
        if "specJson" in c:  # bytes marshalled by the Go side: taken verbatim
            sj = c["specJson"].encode() if isinstance(c["specJson"], str) else bytes(c["specJson"])
            o.spec_json_verbatim = 1
        else:
            sj = json.dumps(spec).encode("utf-8")
        o.spec_json, o.spec_json_len = sj, len(sj)
        self._check(self._L.kr_packer_cluster_upsert(self._h, C.byref(o)))

    def delete_cluster(self, ns: str, name: str):
        self._check(self._L.kr_packer_cluster_delete(self._h, _s(ns), _s(name)))

    def upsert_job(self, j: dict):
        o = abi.kr_job_obj(_s(j.get("namespace", "default")), _s(j["name"]), _s((j.get("status") or {}).get("rayClusterName") or None),
                           _s(snp.status_summary_key((j.get("status") or {}).get("rayClusterStatus"))))
        self._check(self._L.kr_packer_job_upsert(self._h, C.byref(o)))

    # ------------------------------------------------------------------ epoch
    def flush(self) -> int:
        mode = C.c_uint32()
        self._check(self._L.kr_packer_flush(self._h, C.byref(mode)))
        self._check(self._L.kr_packer_sizes(self._h, C.byref(self.engine.sizes)))
        return mode.value

    def flags(self, **kw) -> abi.kr_flags:
        return abi.default_flags(id_head_not_found_reason=self._L.kr_packer_intern(self._h, _s(snp.HEAD_NOT_FOUND_REASON)),
                                 id_head_not_found_msg=self._L.kr_packer_intern(self._h, _s(snp.HEAD_NOT_FOUND_MSG)), **kw)

    def column(self, name: str):
        """Read-only numpy view of one arena column the packer maintains (live rows only)."""
        import numpy as np
        bufs = abi.kr_snapshot_bufs()
        self._check(self._L.kr_packer_bufs(self._h, C.byref(bufs)))
        dt, mult, dim = next((d, m, dm) for n, d, m, dm in abi.COLUMNS if n == name)
        s = self.engine.sizes
        count = {"clusters": s.n_clusters, "groups": s.n_groups, "wtd": s.n_wtd, "pods": s.n_pods, "heads": s.n_heads, "jobs": s.n_jobs, "json": s.json_bytes}[dim] * mult
        ptr = C.cast(getattr(bufs, name), C.c_void_p).value
        if not count:
            return np.zeros(0, dtype=dt)
        return np.frombuffer((C.c_uint8 * (np.dtype(dt).itemsize * count)).from_address(ptr), dtype=dt, count=count)

    def string(self, i: int):
        out = abi.kr_str()
        self._check(self._L.kr_packer_string(self._h, int(i), C.byref(out)))
        return None if not out.p and out.n == 0 and i == 0 else C.string_at(out.p, out.n).decode()

    def cluster_row(self, ns, name) -> int:
        return int(self._L.kr_packer_cluster_row(self._h, _s(ns), _s(name)))

    def pod_row(self, ns, name) -> int:
        return int(self._L.kr_packer_pod_row(self._h, _s(ns), _s(name)))

    def pod_key(self, row: int):
        a, b = abi.kr_str(), abi.kr_str()
        self._check(self._L.kr_packer_pod_key(self._h, row, C.byref(a), C.byref(b)))
        return (None, None) if not a.p else (C.string_at(a.p, a.n).decode(), C.string_at(b.p, b.n).decode())

    def epoch(self) -> tuple[int, int]:
        e, v = C.c_uint64(), C.c_uint64()
        self._check(self._L.kr_packer_epoch(self._h, C.byref(e), C.byref(v)))
        return e.value, v.value

    def cluster_epoch(self, row: int) -> tuple[int, int]:
        rv, gen = C.c_uint64(), C.c_uint64()
        self._check(self._L.kr_packer_cluster_epoch(self._h, row, C.byref(rv), C.byref(gen)))
        return rv.value, gen.value
