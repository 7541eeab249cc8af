"""Adversarial object-level generator for differential tests (oracle vs engine, oracle list modes against each other).

Unlike kuberay_b200.synthetic (the SURVEY §8d benchmark distributions) this one draws every field the path reads
uniformly from its whole domain, on tiny snapshots with heavy name reuse: same cluster name in several namespaces, pods
labelled for clusters / groups that do not exist, zero or several head pods, workersToDelete naming pods of other groups,
nil / negative / int32-wrapping replica numbers, numOfHosts <= 0, unparsable replica-index labels, every phase and
condition value, every old-status variant.  The objects go through snapshot.pack_objects like the golden scenarios.
"""
import base64
import hashlib

import numpy as np

from kuberay_b200 import snapshot as snp

L_CLUSTER, L_TYPE, L_GROUP = "ray.io/cluster", "ray.io/node-type", "ray.io/group"
L_RIDX, L_RNAME = "ray.io/worker-group-replica-index", "ray.io/worker-group-replica-name"
COND_TYPES = ["HeadPodReady", "RayClusterProvisioned", "RayClusterSuspending", "RayClusterSuspended", "RayClusterReplicaFailure"]
PHASES = ["", "Pending", "Running", "Running", "Running", "Succeeded", "Failed", "Unknown", "Bogus"]


def _hash32(b: bytes) -> str:
    return base64.b32hexencode(hashlib.sha1(b).digest()).decode()


def _pick(rng, seq):
    return seq[int(rng.integers(len(seq)))]


def _replica_number(rng):
    r = rng.random()
    if r < 0.2:
        return None
    if r < 0.85:
        return int(rng.integers(-2, 7))
    return _pick(rng, [2 ** 31 - 1, -2 ** 31, 2 ** 30, 2 ** 30 + 1, 2 ** 29 + 2, 65536, 65537])


def _expected(g) -> int:
    from oracle import oracle
    return oracle.desired_replicas(g["replicas"], g["minReplicas"], g["maxReplicas"], g["numOfHosts"], bool(g.get("suspend")))


def _conditions(rng):
    out = []
    for t in COND_TYPES:
        if rng.random() < 0.5:
            continue
        status = _pick(rng, ["True", "False", "Unknown", ""])
        if t == "RayClusterProvisioned":
            reason, msg = _pick(rng, [("AllPodRunningAndReadyFirstTime", "All Ray Pods are ready for the first time"),
                                      ("RayClusterPodsProvisioning", "RayCluster Pods are provisioning"), ("Other", "x")])
        elif t in ("RayClusterSuspending", "RayClusterSuspended"):
            reason, msg = _pick(rng, [(t, ""), (t, "msg"), ("UserRequest", "")])
        elif t == "HeadPodReady":
            reason, msg = _pick(rng, [("HeadPodNotFound", "Head Pod not found"), ("HeadPodRunningAndReady", ""), ("ContainersNotReady", "ray-head: boom"),
                                      ("Unknown", "")])
        else:
            reason, msg = _pick(rng, [("FailedDeleteAllPods", "e1"), ("FailedDeleteHeadPod", "e2"), ("FailedCreateHeadPod", "e3"),
                                      ("FailedDeleteWorkerPod", "e4"), ("FailedCreateWorkerPod", "e5"), ("Whatever", "e6")])
        out.append({"type": t, "status": status, "reason": reason, "message": msg})
    rng.shuffle(out)
    return out


def _old_status(rng, name):
    if rng.random() < 0.15:
        return {}
    st = {"state": _pick(rng, ["", "ready", "suspended", "failed", "unhealthy"]), "conditions": _conditions(rng)}
    if rng.random() < 0.2:
        st["reason"] = "old reason"
    for k in ("readyWorkerReplicas", "availableWorkerReplicas", "desiredWorkerReplicas", "minWorkerReplicas", "maxWorkerReplicas"):
        if rng.random() < 0.7:
            st[k] = int(rng.integers(0, 6))
    if rng.random() < 0.6:
        st["head"] = {"podIP": _pick(rng, ["", "10.1.0.1", "10.1.0.2"]), "serviceIP": _pick(rng, ["", "10.0.0.1", "10.0.0.9"]),
                      "podName": _pick(rng, ["", "p0", "p1"]), "serviceName": _pick(rng, ["", f"{name}-head-svc"])}
    if rng.random() < 0.5:
        st["endpoints"] = _pick(rng, [{}, {"dashboard": "8265"}, {"client": "10001", "dashboard": "8265"}])
    return st


def generate(seed: int, max_clusters: int = 10, big: bool = False):
    """-> (clusters, pods, jobs) as the dict objects snapshot.pack_objects takes."""
    rng = np.random.default_rng(seed)
    namespaces = [f"ns{i}" for i in range(int(rng.integers(1, 4)))]
    cluster_pool = [f"c{i}" for i in range(6)]
    group_pool = [f"g{i}" for i in range(6)] + ["headgroup"]
    clusters, pods, jobs = [], [], []
    seen = set()
    pod_no = 0
    for _ in range(int(rng.integers(1, max_clusters + 1))):
        ns, name = _pick(rng, namespaces), _pick(rng, cluster_pool)
        if (ns, name) in seen:
            continue
        seen.add((ns, name))
        spec_json = bytes(rng.integers(32, 127, size=int(rng.integers(0, 400)), dtype=np.uint8))
        spec = {}
        for key in ("suspend", "enableInTreeAutoscaling"):
            v = _pick(rng, [None, None, None, True, False])
            if v is not None:
                spec[key] = v
        us = _pick(rng, [None, None, "Recreate", "Recreate", "None"])
        if us:
            spec["upgradeStrategy"] = {"type": us}
        gnames = list(rng.permutation(group_pool[:6])[: int(rng.integers(0, 5))])
        groups = []
        for gn in gnames:
            g = {"groupName": str(gn), "replicas": _replica_number(rng), "minReplicas": _replica_number(rng), "maxReplicas": _replica_number(rng),
                 "numOfHosts": _pick(rng, [1, 1, 1, 1, 2, 4, 0, -1, 3, 65536])}
            if rng.random() < 0.15:
                g["suspend"] = True
            # keep the create lists small: |expected| in the billions means (after the int32 wrap of expected - running,
            # raycluster_controller.go:836) a billion creates — a capacity error on both sides, not a parity case
            if abs(_expected(g)) > 300:
                g["replicas"], g["maxReplicas"] = int(rng.integers(0, 7)), _pick(rng, [None, 2 ** 31 - 1, 5])
                if abs(_expected(g)) > 300:
                    g["minReplicas"] = None
                if abs(_expected(g)) > 300:
                    g["numOfHosts"] = 1
            groups.append(g)
        spec["workerGroupSpecs"] = groups
        c = {"namespace": ns, "name": name, "uid": f"uid-{seed}-{ns}-{name}", "spec": spec, "specJson": spec_json, "status": _old_status(rng, name)}
        if rng.random() < 0.3:
            c["annotations"] = {snp.SKIP_HEAD_RESTART_ANNOT: _pick(rng, ["true", "false"])}
        exp = {"head": rng.random() < 0.85}
        for g in groups:
            exp[g["groupName"]] = rng.random() < 0.85
        c["expectations"] = exp
        if rng.random() < 0.05:
            c["deletionTimestamp"] = "2026-01-01T00:00:00Z"
        c["headService"] = {"count": _pick(rng, [0, 1, 1, 1, 2]), "clusterIP": _pick(rng, ["", "None", "10.0.0.1", "10.0.0.9"]), "name": f"{name}-head-svc"}
        if rng.random() < 0.25:
            c["extErr"] = {"kind": int(rng.integers(0, 8)), "message": _pick(rng, ["e1", "e2", "boom"])}
        clusters.append(c)

        # pods of this cluster (plus strays labelled for it)
        mine = []
        n_heads = _pick(rng, [0, 1, 1, 1, 1, 1, 1, 1, 2, 3])
        for _h in range(n_heads):
            ann = {}
            if rng.random() < 0.7:
                ann[snp.RECREATE_HASH_ANNOT] = _pick(rng, [_hash32(spec_json), _hash32(spec_json), _hash32(b"other"), "short", ""])
            if rng.random() < 0.7:
                ann[snp.KUBERAY_VERSION_ANNOT] = _pick(rng, [snp.KUBERAY_VERSION, snp.KUBERAY_VERSION, "v0.0.1", ""])
            mine.append({"labels": {L_CLUSTER: name, L_TYPE: "head", L_GROUP: _pick(rng, ["headgroup", "headgroup", str(_pick(rng, group_pool))])},
                         "annotations": ann, "podIP": _pick(rng, ["", "10.1.0.1", "10.1.0.2"])})
        for g in groups:
            hosts = g["numOfHosts"] if 0 < g["numOfHosts"] <= 4 else 1
            n = int(rng.integers(0, 9 if not big else 40)) * (hosts if rng.random() < 0.7 else 1)
            for k in range(n):
                labels = {L_CLUSTER: name, L_GROUP: g["groupName"]}
                t = _pick(rng, ["worker", "worker", "worker", None, "redis-cleanup", "head"] if rng.random() < 0.15 else ["worker"])
                if t:
                    labels[L_TYPE] = t
                if rng.random() < 0.8:
                    labels[L_RIDX] = _pick(rng, [str(k // hosts), str(int(rng.integers(0, 6))), "-1", "abc", "007", "+3", "99999999999", "9223372036854775808", ""])
                if hosts > 1 or rng.random() < 0.2:
                    if rng.random() < 0.9:
                        labels[L_RNAME] = f"{g['groupName']}-r{k // hosts if rng.random() < 0.85 else int(rng.integers(0, 4))}"
                mine.append({"labels": labels})
        for _s in range(int(rng.integers(0, 3))):  # strays: right cluster label, group that is not in the spec / no group label / no type
            labels = {L_CLUSTER: name}
            if rng.random() < 0.6:
                labels[L_GROUP] = str(_pick(rng, group_pool))
            if rng.random() < 0.6:
                labels[L_TYPE] = _pick(rng, ["worker", "head", "redis-cleanup", "bogus"])
            mine.append({"labels": labels})
        for p in mine:
            p["namespace"] = ns if rng.random() < 0.95 else _pick(rng, namespaces)
            p["name"] = f"p{pod_no}"
            pod_no += 1
            p["phase"] = "Running" if (p["labels"].get(L_TYPE) == "head" and rng.random() < 0.6) else _pick(rng, PHASES)
            if rng.random() < 0.8:
                p["conditions"] = [{"type": "Ready", "status": _pick(rng, ["True", "True", "True", "False", "Unknown", ""]),
                                    "reason": _pick(rng, ["", "", "ContainersNotReady", "PodCompleted"]), "message": _pick(rng, ["", "m"])}]
            p["restartPolicy"] = _pick(rng, ["Always", "Always", "Never", "OnFailure"])
            if rng.random() < 0.15:
                p["rayContainerTerminated"] = True
            if rng.random() < 0.1:
                p["deletionTimestamp"] = "2026-01-01T00:00:00Z"
        # workersToDelete: own pods, pods of other groups / clusters, names that do not exist
        names_here = [p["name"] for p in mine] or ["p0"]
        for g in groups:
            if rng.random() < 0.4:
                g["workersToDelete"] = [_pick(rng, names_here + ["ghost", f"p{max(0, pod_no - 30)}"]) for _ in range(int(rng.integers(1, 5)))]
        pods.extend(mine)
    for _ in range(int(rng.integers(0, 4))):  # pods of no known cluster
        labels = {L_GROUP: "g0", L_TYPE: "worker"}
        if rng.random() < 0.7:
            labels[L_CLUSTER] = _pick(rng, ["nope", "c0", "c5"])
        pods.append({"namespace": _pick(rng, namespaces + ["elsewhere"]), "name": f"p{pod_no}", "labels": labels, "phase": _pick(rng, PHASES)})
        pod_no += 1
    order = rng.permutation(len(pods))
    pods = [pods[i] for i in order]
    for _ in range(int(rng.integers(0, 6))):
        c = _pick(rng, clusters)
        st = dict(c["status"]) if rng.random() < 0.6 else _old_status(rng, c["name"])
        jobs.append({"namespace": _pick(rng, [c["namespace"], c["namespace"], "elsewhere"]),
                     "status": {"rayClusterName": _pick(rng, [c["name"], c["name"], c["name"], "missing", ""]), "rayClusterStatus": st}})
    return clusters, pods, jobs


def snapshot(seed: int, **kw):
    """-> (Snapshot, kr_flags) with the process flags also drawn from the seed."""
    clusters, pods, jobs = generate(seed, **kw)
    snap, meta = snp.pack_objects(clusters, pods, jobs)
    rng = np.random.default_rng(seed ^ 0x5EED)
    f = meta.flags
    f.gate_status_conditions = int(rng.random() < 0.8)
    f.gate_multihost_indexing = int(rng.random() < 0.8)
    f.env_random_pod_delete = int(rng.random() < 0.4)
    return snap, f
