"""CPU restatement of the reference's Pod-metadata step — TEST INFRASTRUCTURE ONLY.

Only tests/ may import this; the product path is kuberay_b200/csrc/kr_podmeta.cpp behind the C ABI.  Each function follows the
reference lines it cites (ray-operator/controllers/ray/...); nothing is copied.  Pinned by tests/test_podmeta.py against the
reference's own unit-test vectors (utils/util_test.go:110-248, common/pod_test.go:1328-1350, 2188-2227)."""
from __future__ import annotations

import json

# unicode.IsPunct(rune(b)) for a byte widened to a rune: ASCII category P, plus the Latin-1 members
_PUNCT = set(b"!\"#%&'()*,-./:;?@[\\]_{}") | {0xA1, 0xA7, 0xAB, 0xB6, 0xB7, 0xBB, 0xBF}


def _enc(s: str) -> bytes:
    return s.encode("utf-8", "surrogateescape")


def _dec(b: bytes) -> str:
    return b.decode("utf-8", "surrogateescape")


def check_name(s: bytes) -> bytes:
    """utils/util.go:217-240."""
    if len(s) > 50:
        s = s[len(s) - 50:]
    if 0x30 <= s[0] <= 0x39:
        s = b"r" + s[1:]
    if s[0] in _PUNCT:
        s = b"r" + s[1:]
    return s


def check_label(s: bytes) -> bytes:
    """utils/util.go:247-265."""
    if len(s) > 63:
        s = s[len(s) - 63:]
    if s[0] in _PUNCT:
        s = b"r" + s[1:]
    return s


def pod_name(prefix: bytes, node_type: str, is_generate_name: bool) -> bytes:
    """utils/util.go:198-215 (ToLower on ASCII letters: prefixes are RFC 1123 names)."""
    r = bytes((c + 32 if 0x41 <= c <= 0x5A else c) for c in prefix[:50] + b"-" + node_type.encode())
    return r + (b"-" if is_generate_name else b"")


def label_pod(node_type: str, cluster: str, group: str, override: dict) -> dict:
    """common/pod.go:775-799."""
    labels = {
        "ray.io/is-ray-node": "yes", "ray.io/cluster": cluster, "ray.io/node-type": node_type, "ray.io/group": group,
        "ray.io/identifier": _dec(check_label(_enc(f"{cluster}-{node_type}"))),
        "app.kubernetes.io/name": "kuberay", "app.kubernetes.io/created-by": "kuberay-operator",
    }
    for k, v in override.items():
        if k in ("ray.io/node-type", "ray.io/group", "ray.io/cluster"):
            continue
        labels[k] = v
    return labels


def merge_labels(template_labels: dict | None, group_labels: dict | None) -> dict:
    """common/pod.go:1276-1283."""
    m = dict(template_labels or {})
    m.update(group_labels or {})
    return m


def pod_meta(cluster: dict, create: tuple, *, kuberay_version="v1.5.0", deterministic_head_name=False,
             multihost_indexing_gate=True, cluster_hash: str | None = None) -> dict:
    """ObjectMeta of one Pod: DefaultHeadPodTemplate / DefaultWorkerPodTemplate metadata (common/pod.go:166-190, 352-357, 420-443),
    BuildPod's serve label (:583-588), createHeadPod's annotations (raycluster_controller.go:1313-1316), SetControllerReference."""
    spec = cluster.get("spec") or {}
    annots = cluster.get("annotations") or {}
    name, ns = cluster["name"], cluster.get("namespace", "default")
    g, replica_index, host_index, replica_name = create
    head = g < 0
    grp = (spec.get("headGroupSpec") or {}) if head else spec["workerGroupSpecs"][g]
    tmeta = (grp.get("template") or {}).get("metadata") or {}
    node_type = "head" if head else "worker"
    gname = "headgroup" if head else grp.get("groupName", "")
    labels = label_pod(node_type, name, gname, merge_labels(tmeta.get("labels"), grp.get("labels")))
    if not head and multihost_indexing_gate:                                              # common/pod.go:430-439
        labels["ray.io/worker-group-replica-index"] = str(replica_index)
        if int(grp.get("numOfHosts", 1)) > 1:
            labels["ray.io/worker-group-replica-name"] = replica_name
            labels["ray.io/replica-host-index"] = str(host_index)
    crd = (cluster.get("labels") or {}).get("ray.io/originated-from-crd")
    if crd == "RayService":                                                               # common/pod.go:583-588
        labels["ray.io/serve"] = "false" if head else "true"
    a = dict(tmeta.get("annotations") or {})
    if "ray.io/overwrite-container-cmd" in annots and annots["ray.io/overwrite-container-cmd"].lower() == "true":   # :62-75
        a["ray.io/overwrite-container-cmd"] = "true"
    ft_opts = spec.get("gcsFaultToleranceOptions")
    ft = ("ray.io/ft-enabled" in annots and annots["ray.io/ft-enabled"].lower() == "true") or ft_opts is not None  # util.go:753-756
    if head:
        a["ray.io/ft-enabled"] = "true" if ft else "false"                                # :85-87
        if ft:                                                                            # :105-114
            sns = cluster.get("uid", "")
            if "ray.io/external-storage-namespace" in annots:
                sns = annots["ray.io/external-storage-namespace"]
            if ft_opts and ft_opts.get("externalStorageNamespace"):
                sns = ft_opts["externalStorageNamespace"]
            a["ray.io/external-storage-namespace"] = sns
        if cluster_hash:                                                                  # raycluster_controller.go:1313-1316
            a["ray.io/upgrade-strategy-recreate-hash"] = cluster_hash
            a["ray.io/kuberay-version"] = kuberay_version
    if head:
        nm = _dec(pod_name(_enc(name), "head", not deterministic_head_name))
        key = "name" if deterministic_head_name else "generateName"
    else:
        nm = _dec(pod_name(_enc(f"{name}-{gname}"), "worker", True))
        key = "generateName"
    return {key: nm, "namespace": ns, "labels": labels, "annotations": a,
            "ownerReferences": [{"apiVersion": "ray.io/v1", "kind": "RayCluster", "name": name, "uid": cluster.get("uid", ""),
                                 "controller": True, "blockOwnerDeletion": True}]}


def go_string(s: str) -> str:
    """encoding/json's string encoding with HTML escaping, for valid UTF-8 input."""
    out = ['"']
    for ch in s:
        o = ord(ch)
        if ch == '"':
            out.append('\\"')
        elif ch == "\\":
            out.append("\\\\")
        elif ch == "\b":
            out.append("\\b")
        elif ch == "\f":
            out.append("\\f")
        elif ch == "\n":
            out.append("\\n")
        elif ch == "\r":
            out.append("\\r")
        elif ch == "\t":
            out.append("\\t")
        elif o < 0x20 or ch in "<>&":
            out.append("\\u%04x" % o)
        elif o in (0x2028, 0x2029):
            out.append("\\u%04x" % o)
        else:
            out.append(ch)
    out.append('"')
    return "".join(out)


def go_marshal(meta: dict) -> bytes:
    """The bytes encoding/json writes for the patch: fields in the order ObjectMeta declares them, map keys sorted by bytes."""
    def m(d: dict) -> str:
        return "{" + ",".join(f"{go_string(k)}:{go_string(v)}" for k, v in sorted(d.items(), key=lambda kv: _enc(kv[0]))) + "}"
    key = "name" if "name" in meta else "generateName"
    o = meta["ownerReferences"][0]
    return _enc(f'{{"{key}":{go_string(meta[key])},"namespace":{go_string(meta["namespace"])},"labels":{m(meta["labels"])},'
                f'"annotations":{m(meta["annotations"])},"ownerReferences":[{{"apiVersion":"ray.io/v1","kind":"RayCluster","name":{go_string(o["name"])},'
                f'"uid":{go_string(o["uid"])},"controller":true,"blockOwnerDeletion":true}}]}}')


def expand_creates(group_results, create_idx, groups: list[dict], head_create: bool, multihost_indexing_gate=True) -> list[tuple]:
    """Create order of reconcilePods (:692-735 head; :869-889 workers; :1081-1094 multi-host) as (group, replicaIndex, hostIndex);
    the generated replica name is random in the reference, so it is not part of the restated tuple."""
    MULTIHOST = 1 << 0
    out = [(-1, 0, 0)] if head_create else []
    for g, (gr, grp) in enumerate(zip(group_results, groups)):
        mh = multihost_indexing_gate and int(grp.get("numOfHosts", 1)) > 1
        for k in range(int(gr["n_create"])):
            idx = int(create_idx[int(gr["create_off"]) + k]) if multihost_indexing_gate else 0
            if mh:
                out += [(g, idx, j) for j in range(int(grp["numOfHosts"]))]
            else:
                out.append((g, idx, 0))
    return out
