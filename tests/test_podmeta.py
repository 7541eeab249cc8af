"""Pod metadata builder (kr_pod_*, kuberay_b200/csrc/kr_podmeta.cpp) — host-side, runs without a GPU.

1. the reference's own vectors for the name helpers, transcribed: utils/util_test.go:110-143 (TestWorkerPodName), :145-213
   (TestHeadPodName), :215-248 (TestCheckName); common/pod_test.go:2188-2227 (TestMergeLabels), :1328-1350
   (TestDeafultWorkerPodTemplateWithReplicaGrpAndIndex), :296-475 (FT annotations), :996-1010 (serve label);
   raycluster_controller_unit_test.go:1025-1045 (the labels the selectors read)
2. the native builder against the CPU restatement (oracle/podmeta.py) on fuzzed RayClusters, as parsed JSON AND byte for byte
   against Go's encoding (field order, sorted maps, escapes)
3. engine results -> create tuples (kr_pod_creates_expand)"""
import json
import random
import re

import numpy as np
import pytest

from kuberay_b200 import abi
from kuberay_b200 import podmeta as pm
from kuberay_b200.engine import EngineError
from oracle import podmeta as ref

# ----------------------------------------------------------------------------------------------------------- 1. reference vectors
WORKER_NAMES = [  # util_test.go:116-126
    ("ray-cluster-group-name-01", "ray-cluster-group-name-01-worker-"),
    ("ray-cluster-0000000000000000000000011111111122222233333333333333-group-name", "ray-cluster-00000000000000000000000111111111222222-worker-"),
]
HEAD_NAMES = [  # util_test.go:154-190: (prefix, deterministic, expected)
    ("ray-cluster-01", True, "ray-cluster-01-head"),
    ("ray-cluster-01", False, "ray-cluster-01-head-"),
    ("ray-cluster-0000000000000000000000011111111122222233333333333333", True, "ray-cluster-00000000000000000000000111111111222222-head"),
    ("ray-cluster-0000000000000000000000011111111122222233333333333333", False, "ray-cluster-00000000000000000000000111111111222222-head-"),
]
CHECK_NAMES = [  # util_test.go:221-235
    ("72fbcc7e-a661-4b18e-ca41-e903-fc3ae634b18e-lazer090scholar-director-s", "rca41-e903-fc3ae634b18e-lazer090scholar-director-s"),
    ("--------566666--------444433-----------222222----------4444", "r6666--------444433-----------222222----------4444"),
    ("acceptable-name-head-12345", "acceptable-name-head-12345"),
]
MERGE_LABELS = [  # pod_test.go:2194-2218
    ({"pod-label-key": "pod-label-value"}, {"ray/io:some-label": "ray-node-label-value"}, {"pod-label-key": "pod-label-value", "ray/io:some-label": "ray-node-label-value"}),
    ({"accelerator-type": "GPU", "market-type": "spot"}, {"accelerator-type": "TPU-V6E"}, {"accelerator-type": "TPU-V6E", "market-type": "spot"}),
    ({}, {"group-labels": "group-label-value"}, {"group-labels": "group-label-value"}),
    ({"pod-label": "pod-label-value"}, {}, {"pod-label": "pod-label-value"}),
    (None, None, {}),
]
FIXED = {"ray.io/is-ray-node", "ray.io/cluster", "ray.io/node-type", "ray.io/group", "ray.io/identifier", "app.kubernetes.io/name", "app.kubernetes.io/created-by"}


def _cluster(name="raycluster-sample", ns="default", **kw):
    c = {"name": name, "namespace": ns, "uid": "0f5b5f0c-7c1e-4f0e-9a0a-3f1d8a4c2b11",
         "spec": {"headGroupSpec": {"template": {"metadata": {}}},
                  "workerGroupSpecs": [{"groupName": "small-group", "numOfHosts": 1, "template": {"metadata": {}}}]}}
    c.update(kw)
    return c


@pytest.mark.parametrize("prefix,want", WORKER_NAMES)
def test_worker_pod_name_vectors(prefix, want):
    got = pm.pod_name(prefix, "worker", True)
    assert got == want == ref.pod_name(prefix.encode(), "worker", True).decode()
    assert len(got) <= 58  # 63 - the 5 generated characters (util_test.go:137-140)


@pytest.mark.parametrize("prefix,det,want", HEAD_NAMES)
def test_head_pod_name_vectors(prefix, det, want):
    got = pm.pod_name(prefix, "head", not det)
    assert got == want == ref.pod_name(prefix.encode(), "head", not det).decode()
    assert len(got) <= 58
    meta = pm.build_pod_meta(_cluster(prefix), [(-1, 0, 0, "")], pm.PodMetaEnv(deterministic_head_name=det))[0]
    assert meta == ({"name": want} if det else {"generateName": want}) | {k: meta[k] for k in ("namespace", "labels", "annotations", "ownerReferences")}


@pytest.mark.parametrize("s,want", CHECK_NAMES)
def test_check_name_vectors(s, want):
    assert pm.check_name(s) == want == ref.check_name(s.encode()).decode()


def test_check_label_and_name_edges():
    long = "x" * 30 + "-" + "y" * 40
    assert pm.check_label(long) == long[-63:] and len(pm.check_label(long)) == 63
    assert pm.check_label("-abc") == "rabc" and pm.check_label("_abc") == "rabc" and pm.check_label("9abc") == "9abc"
    # category S is not punctuation for unicode.IsPunct: $ + < = > ^ ` | ~ stay
    for ch in "$+<=>^`|~":
        assert pm.check_label(ch + "abc") == ch + "abc" and pm.check_name(ch + "abc") == ch + "abc"
    for ch in "!\"#%&'()*,-./:;?@[\\]_{}":
        assert pm.check_label(ch + "abc") == "rabc" and pm.check_name(ch + "abc") == "rabc"
    assert pm.check_name("0") == "r" and pm.check_name("a") == "a"
    # a cut that lands on a digit / dash is fixed up AFTER the cut (util.go:221-237)
    assert pm.check_name("a" * 10 + "7" + "b" * 49) == "r" + "b" * 49
    # the byte is widened to a rune: 0xB7 (as a Latin-1 code point: the middle dot) is punctuation, 0xB5 is not
    assert ref.check_label(b"\xb7abc") == b"rabc" and ref.check_label(b"\xb5abc") == b"\xb5abc"
    assert pm.check_label(b"\xb7abc".decode("utf-8", "surrogateescape")).encode("utf-8", "surrogateescape") == b"rabc"
    assert pm.check_label(b"\xb5abc".decode("utf-8", "surrogateescape")).encode("utf-8", "surrogateescape") == b"\xb5abc"
    for f in (pm.check_name, pm.check_label):
        with pytest.raises(EngineError):
            f("")  # the reference indexes s[0] and panics


@pytest.mark.parametrize("tl,gl,want", MERGE_LABELS)
def test_merge_labels_vectors(tl, gl, want):
    c = _cluster()
    g = c["spec"]["workerGroupSpecs"][0]
    if tl is not None:
        g["template"]["metadata"]["labels"] = tl
    if gl is not None:
        g["labels"] = gl
    assert ref.merge_labels(tl, gl) == want
    meta = pm.build_pod_meta(c, [(0, 0, 0, "")], pm.PodMetaEnv(multihost_indexing_gate=False))[0]
    assert {k: v for k, v in meta["labels"].items() if k not in FIXED} == want


def test_selector_labels_and_protected_keys():
    """raycluster_controller_unit_test.go:1025-1045: the labels the List selectors read; common/pod.go:786-793: node-type, group and
    cluster cannot be overridden by the template, the others can."""
    c = _cluster()
    c["spec"]["headGroupSpec"]["template"]["metadata"]["labels"] = {"ray.io/cluster": "evil", "ray.io/node-type": "worker", "ray.io/group": "x",
                                                                    "ray.io/identifier": "mine", "app.kubernetes.io/name": "other", "team": "a<b"}
    meta = pm.build_pod_meta(c, [(-1, 0, 0, "")])[0]
    assert meta["labels"] == {"ray.io/is-ray-node": "yes", "ray.io/cluster": "raycluster-sample", "ray.io/node-type": "head", "ray.io/group": "headgroup",
                              "ray.io/identifier": "mine", "app.kubernetes.io/name": "other", "app.kubernetes.io/created-by": "kuberay-operator", "team": "a<b"}
    w = pm.build_pod_meta(c, [(0, 3, 0, "")])[0]
    assert w["labels"]["ray.io/identifier"] == "raycluster-sample-worker" and w["labels"]["ray.io/group"] == "small-group"
    assert w["generateName"] == "raycluster-sample-small-group-worker-" and w["namespace"] == "default"
    assert w["ownerReferences"] == [{"apiVersion": "ray.io/v1", "kind": "RayCluster", "name": "raycluster-sample", "uid": c["uid"], "controller": True, "blockOwnerDeletion": True}]


def test_replica_group_and_index_labels():
    """pod_test.go:1328-1350: NumOfHosts = 4, replica name, index 0, host 2; and the single-host case only gets the index."""
    c = _cluster()
    c["spec"]["workerGroupSpecs"][0]["numOfHosts"] = 4
    m = pm.build_pod_meta(c, [(0, 0, 2, "small-group-abcde")])[0]
    assert "name" not in m
    assert m["labels"]["ray.io/worker-group-replica-name"] == "small-group-abcde"
    assert m["labels"]["ray.io/worker-group-replica-index"] == "0" and m["labels"]["ray.io/replica-host-index"] == "2"
    c["spec"]["workerGroupSpecs"][0]["numOfHosts"] = 1
    m = pm.build_pod_meta(c, [(0, 7, 0, "")])[0]
    assert m["labels"]["ray.io/worker-group-replica-index"] == "7"
    assert "ray.io/worker-group-replica-name" not in m["labels"] and "ray.io/replica-host-index" not in m["labels"]
    m = pm.build_pod_meta(c, [(0, 7, 0, "")], pm.PodMetaEnv(multihost_indexing_gate=False))[0]
    assert not any(k.startswith("ray.io/worker-group-replica") for k in m["labels"])


@pytest.mark.parametrize("annots,ft_opts,want_ft,want_ns", [
    ({}, None, "false", None),
    ({"ray.io/ft-enabled": "true"}, None, "true", "UID"),                                          # pod_test.go:306-310
    ({"ray.io/ft-enabled": "TRUE", "ray.io/external-storage-namespace": "test-ns"}, None, "true", "test-ns"),   # :311-316
    ({"ray.io/ft-enabled": "false"}, {"redisAddress": "redis:6379"}, "true", "UID"),               # options alone enable it (util.go:755)
    ({"ray.io/external-storage-namespace": "a"}, {"redisAddress": "r", "externalStorageNamespace": "opt-ns"}, "true", "opt-ns"),  # :620-631
])
def test_ft_annotations(annots, ft_opts, want_ft, want_ns):
    c = _cluster(annotations=annots)
    if ft_opts is not None:
        c["spec"]["gcsFaultToleranceOptions"] = ft_opts
    head, worker = pm.build_pod_meta(c, [(-1, 0, 0, ""), (0, 0, 0, "")])
    assert head["annotations"]["ray.io/ft-enabled"] == want_ft
    assert head["annotations"].get("ray.io/external-storage-namespace") == (c["uid"] if want_ns == "UID" else want_ns)
    assert worker["annotations"] == {}  # pod_test.go:290-292: neither annotation on a worker
    assert head == ref.pod_meta(c, (-1, 0, 0, "")) and worker == ref.pod_meta(c, (0, 0, 0, ""))


def test_overwrite_cmd_hash_stamps_and_serve_label():
    c = _cluster(annotations={"ray.io/overwrite-container-cmd": "True"}, labels={"ray.io/originated-from-crd": "RayService"})
    c["spec"]["headGroupSpec"]["template"]["metadata"]["annotations"] = {"keep": "me"}
    head, worker = pm.build_pod_meta(c, [(-1, 0, 0, ""), (0, 0, 0, "")], pm.PodMetaEnv(kuberay_version="v9.9.9"), cluster_hash="ABCDEF0123")
    assert head["annotations"] == {"keep": "me", "ray.io/overwrite-container-cmd": "true", "ray.io/ft-enabled": "false",
                                   "ray.io/upgrade-strategy-recreate-hash": "ABCDEF0123", "ray.io/kuberay-version": "v9.9.9"}
    assert worker["annotations"] == {"ray.io/overwrite-container-cmd": "true"}
    assert head["labels"]["ray.io/serve"] == "false" and worker["labels"]["ray.io/serve"] == "true"   # pod_test.go:996-1010
    # no hash (clusterHash == ""): no stamps (raycluster_controller.go:1313)
    head = pm.build_pod_meta(c, [(-1, 0, 0, "")])[0]
    assert "ray.io/kuberay-version" not in head["annotations"] and "ray.io/upgrade-strategy-recreate-hash" not in head["annotations"]
    c["labels"] = {"ray.io/originated-from-crd": "RayJob"}
    assert "ray.io/serve" not in pm.build_pod_meta(c, [(-1, 0, 0, "")])[0]["labels"]


def test_invalid_inputs_fail_loudly():
    c = _cluster()
    with pytest.raises(EngineError):
        pm.build_pod_meta(c, [(5, 0, 0, "")])
    with pytest.raises(EngineError):
        pm.build_pod_meta(_cluster(name=""), [(-1, 0, 0, "")])


# ----------------------------------------------------------------------------------------------------------- 2. fuzz vs restatement
def _rand_text(rng, alphabet, lo, hi):
    return "".join(rng.choice(alphabet) for _ in range(rng.randint(lo, hi)))


NAMEC = "abcdefghijklmnopqrstuvwxyz0123456789-"
ANY = NAMEC + "ABCXYZ_./<>&\"\\ \t\né 中\U0001f600:"


def _rand_map(rng, n_max):
    keys = ["ray.io/cluster", "ray.io/group", "ray.io/node-type", "ray.io/identifier", "ray.io/is-ray-node", "app.kubernetes.io/name", "team", "zone",
            "ray.io/worker-group-replica-index", "ray.io/serve", "a", "Z", "été"]
    return {(rng.choice(keys) if rng.random() < 0.7 else _rand_text(rng, ANY, 1, 12)): _rand_text(rng, ANY, 0, 20) for _ in range(rng.randint(0, n_max))}


def _rand_cluster(rng):
    name = _rand_text(rng, NAMEC, 1, rng.choice([8, 30, 70]))
    c = {"name": name, "namespace": _rand_text(rng, NAMEC, 1, 12), "uid": _rand_text(rng, "0123456789abcdef-", 0, 36), "spec": {}}
    annots = {}
    if rng.random() < 0.4:
        annots["ray.io/overwrite-container-cmd"] = rng.choice(["true", "True", "false", ""])
    if rng.random() < 0.4:
        annots["ray.io/ft-enabled"] = rng.choice(["true", "TRUE", "false", "yes"])
    if rng.random() < 0.3:
        annots["ray.io/external-storage-namespace"] = _rand_text(rng, ANY, 0, 10)
    if annots or rng.random() < 0.5:
        c["annotations"] = annots
    if rng.random() < 0.5:
        c["labels"] = {"ray.io/originated-from-crd": rng.choice(["RayService", "RayJob", "RayCluster", "bogus"])}
    if rng.random() < 0.3:
        c["spec"]["gcsFaultToleranceOptions"] = {"redisAddress": "r:6379", **({"externalStorageNamespace": _rand_text(rng, ANY, 0, 8)} if rng.random() < 0.6 else {})}

    def grp(head):
        g = {"template": {"metadata": {}}}
        if rng.random() < 0.7:
            g["template"]["metadata"]["labels"] = _rand_map(rng, 5)
        if rng.random() < 0.5:
            g["template"]["metadata"]["annotations"] = _rand_map(rng, 4)
        if rng.random() < 0.5:
            g["labels"] = _rand_map(rng, 4)
        if not head:
            g["groupName"] = _rand_text(rng, NAMEC + "ABC", 1, rng.choice([6, 20, 60]))
            g["numOfHosts"] = rng.choice([1, 1, 2, 4])
        return g
    c["spec"]["headGroupSpec"] = grp(True)
    c["spec"]["workerGroupSpecs"] = [grp(False) for _ in range(rng.randint(0, 4))]
    return c


@pytest.mark.parametrize("seed", range(8))
def test_native_builder_matches_restatement_byte_for_byte(seed):
    rng = random.Random(seed)
    for _ in range(60):
        c = _rand_cluster(rng)
        ng = len(c["spec"]["workerGroupSpecs"])
        creates = [(-1, 0, 0, "")] if rng.random() < 0.6 else []
        for _ in range(rng.randint(0, 12)):
            if ng:
                g = rng.randrange(ng)
                creates.append((g, rng.choice([0, 1, 5, 2 ** 31 - 1, -3]), rng.randrange(4), c["spec"]["workerGroupSpecs"][g]["groupName"] + "-" + _rand_text(rng, "bcdfghjklmnpqrstvwxz2456789", 5, 5)))
        rng.shuffle(creates)
        env = pm.PodMetaEnv(kuberay_version=rng.choice(["v1.5.0", "nightly"]), deterministic_head_name=rng.random() < 0.3, multihost_indexing_gate=rng.random() < 0.8)
        h = rng.choice([None, "", "0123456789ABCDEFGHIJKLMNOPQRSTUV"])
        raw = pm.build_pod_meta(c, creates, env, cluster_hash=h, raw=True)
        assert len(raw) == len(creates)
        for t, b in zip(creates, raw):
            want = ref.pod_meta(c, t, kuberay_version=env.kuberay_version, deterministic_head_name=env.deterministic_head_name,
                                multihost_indexing_gate=env.multihost_indexing_gate, cluster_hash=h)
            assert json.loads(b) == want, (c, t)
            assert b == ref.go_marshal(want), (b, ref.go_marshal(want))
            nm = want.get("name") or want["generateName"]
            assert len(nm) <= 58 and nm == nm.lower()


# ----------------------------------------------------------------------------------------------------------- 3. results -> tuples
def _gr(rows):
    a = np.zeros(len(rows), dtype=abi.group_result_dtype)
    for i, (n_create, off, flags) in enumerate(rows):
        a[i]["n_create"], a[i]["create_off"], a[i]["flags"] = n_create, off, flags
    return a


def test_creates_expand_order_names_and_gate():
    groups = [{"groupName": "cpu", "numOfHosts": 1}, {"groupName": "tpu-slice", "numOfHosts": 4}, {"groupName": "idle", "numOfHosts": 1}]
    arena = np.array([99, 0, 2, 5, 98, 1, 3], dtype=np.int32)   # cpu owns [1,4), tpu-slice owns [5,7); the rest is other clusters' / reserved
    gr = _gr([(3, 1, 0), (2, 5, abi.GR_MULTIHOST), (0, 7, 0)])
    got = pm.expand_creates(gr, arena, groups, head_create=True, seed=42)
    assert [t[:3] for t in got] == ref.expand_creates(gr, arena, groups, True)
    assert got[0] == (-1, 0, 0, "") and [t[:3] for t in got[1:4]] == [(0, 0, 0), (0, 2, 0), (0, 5, 0)] and all(t[3] == "" for t in got[1:4])
    mh = got[4:]
    assert [t[:3] for t in mh] == [(1, 1, j) for j in range(4)] + [(1, 3, j) for j in range(4)]
    names = [t[3] for t in mh]
    assert len(set(names[:4])) == 1 and len(set(names[4:])) == 1 and names[0] != names[4]
    for n in (names[0], names[4]):   # util.go:377-379 + the apimachinery alphabet
        assert re.fullmatch(r"tpu-slice-[bcdfghjklmnpqrstvwxz2456789]{5}", n)
    assert pm.expand_creates(gr, arena, groups, True, seed=42) == got and pm.expand_creates(gr, arena, groups, True, seed=43) != got
    # gate off: createWorkerPod(..., "", 0, 0) for every create (:887); no multi-host path at all (the engine never sets the flag then)
    gr0 = _gr([(3, 1, 0), (2, 5, 0), (0, 7, 0)])
    off = pm.expand_creates(gr0, arena, groups, False, pm.PodMetaEnv(multihost_indexing_gate=False))
    assert off == [(0, 0, 0, "")] * 3 + [(1, 0, 0, "")] * 2
    assert [t[:3] for t in off] == ref.expand_creates(gr0, arena, groups, False, multihost_indexing_gate=False)
    assert pm.expand_creates(_gr([]), arena, [], head_create=False) == []
    # tuples feed straight into the builder
    c = _cluster()
    c["spec"]["workerGroupSpecs"] = [{"groupName": g["groupName"], "numOfHosts": g["numOfHosts"], "template": {"metadata": {}}} for g in groups]
    metas = pm.build_pod_meta(c, got)
    assert metas[0]["labels"]["ray.io/node-type"] == "head" and metas[5]["labels"]["ray.io/replica-host-index"] == "1"
    assert metas[5]["labels"]["ray.io/worker-group-replica-name"] == names[0] and metas[5]["generateName"] == "raycluster-sample-tpu-slice-worker-"
