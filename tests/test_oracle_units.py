"""CPU tests: the oracle's scalar pieces against the reference's known-answer tables (tests/golden, transcribed from the Go
tests cited in each fixture), SHA-1 / base32hex standards vectors, and the host-only evaluations of the packer."""
import base64
import hashlib
import json
import os

import numpy as np
import pytest

from kuberay_b200 import abi, snapshot as snapmod, specjson
from kuberay_b200.reconciler import FakeClient, RayClusterReconciler

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    with open(os.path.join(GOLD, name + ".json")) as f:
        return json.load(f)


class OracleBackend:
    def run(self, snap, flags):
        from oracle import oracle
        return oracle.run(snap, flags)


ARITH = load("replica_arithmetic")


@pytest.mark.parametrize("c", ARITH["desired_replicas"]["cases"])
def test_get_worker_group_desired_replicas(c, oracle_mod):
    """utils/util_test.go:555-601"""
    assert oracle_mod.desired_replicas(c["replicas"], c["min"], c["max"], c["hosts"], c["suspend"]) == c["want"]


def test_desired_replicas_int32_wrap(oracle_mod):
    """the multiply is an int32 multiply in Go (utils/util.go:403): 2^30 * 4 wraps to 0, (2^30+1)*4 wraps to 4"""
    assert oracle_mod.desired_replicas(2 ** 30, 0, None, 4) == 0
    assert oracle_mod.desired_replicas(2 ** 30 + 1, 0, None, 4) == 4
    assert oracle_mod.desired_replicas(2 ** 31 - 1, 0, None, 2) == -2


def _cluster_with_groups(groups):
    wg = []
    for i, g in enumerate(groups):
        wg.append({"groupName": f"g{i}", "replicas": g.get("replicas", g.get("min")), "minReplicas": g.get("min"), "maxReplicas": g.get("max"),
                   "numOfHosts": g.get("hosts", 1), "suspend": g.get("suspend")})
        if "replicas" in g and g["replicas"] is None:
            wg[-1]["replicas"] = None
    return {"namespace": "default", "name": "c", "spec": {"headGroupSpec": {"rayStartParams": {}}, "workerGroupSpecs": wg}, "specJson": "{}"}


def _status_counts(groups, pods=()):
    client = FakeClient([_cluster_with_groups(groups)], pods)
    pr = RayClusterReconciler(client, OracleBackend())._pass()
    return pr.res.clusters[0]


@pytest.mark.parametrize("c", ARITH["min_max"]["cases"], ids=lambda c: c["name"])
def test_calculate_min_and_max_replicas(c):
    """utils/util_test.go:603-710"""
    cr = _status_counts(c["groups"])
    assert (int(cr["counts"][3]), int(cr["counts"][4])) == (c["want_min"], c["want_max"])


@pytest.mark.parametrize("c", ARITH["desired_cluster"]["cases"], ids=lambda c: c["name"])
def test_calculate_desired_replicas(c):
    """utils/util_test.go:712-800"""
    assert int(_status_counts(c["groups"])["counts"][2]) == c["want"]


@pytest.mark.parametrize("c", ARITH["max_overflow"]["cases"], ids=lambda c: c["name"])
def test_calculate_max_replicas_overflow(c):
    """utils/util_test.go:802-894"""
    assert int(_status_counts(c["groups"])["counts"][4]) == c["want_max"]


def _pod(name, node_type=None, phase="Running", ready=None, **kw):
    labels = {"ray.io/cluster": "c"}
    if node_type:
        labels["ray.io/node-type"] = node_type
    p = {"namespace": "default", "name": name, "labels": labels, "phase": phase}
    if ready is not None:
        p["conditions"] = [{"type": "Ready", "status": ready}]
    p.update(kw)
    return p


def test_calculate_available_and_ready_replicas():
    """utils/util_test.go:408-475"""
    g = ARITH["available_ready"]
    pods = [_pod(p["name"], p.get("nodeType"), p["phase"], p.get("ready")) for p in g["pods"]]
    cr = _status_counts([{"min": 0, "max": 5, "replicas": 3}], pods)
    assert int(cr["counts"][1]) == g["want_available"] and int(cr["counts"][0]) == g["want_ready"]


@pytest.mark.parametrize("c", ARITH["check_all_pods_running"]["cases"], ids=lambda c: c["name"])
def test_check_all_pods_running(c):
    """utils/util_test.go:58-130"""
    pods = [_pod(f"p{i}", "worker", p["phase"], p.get("ready")) for i, p in enumerate(c["pods"])]
    cr = _status_counts([{"min": 0, "max": 5, "replicas": 0}], pods)
    assert bool(int(cr["status_flags"]) & abi.SF_ALL_PODS_RUNNING) == c["want"]


@pytest.mark.parametrize("c", load("should_delete_pod")["cases"])
def test_should_delete_pod(c, oracle_mod):
    """raycluster_controller_unit_test.go:2380-2503"""
    pod = {"name": "p", "phase": c["phase"], "restartPolicy": c["restartPolicy"], "rayContainerTerminated": c["terminated"]}
    pk, _ = snapmod.pack_pod_word(pod)
    assert bool(oracle_mod.lib().kr_oracle_should_delete(pk)) == c["want"]


HPR = load("head_pod_ready")


@pytest.mark.parametrize("c", HPR["status_cases"])
def test_find_head_pod_ready_condition_status(c):
    """utils/util_test.go:934-970"""
    pod = {"phase": c["phase"], "conditions": [{"type": "Ready", "status": c["ready"], "reason": "ContainersNotReady"}]}
    assert snapmod.head_pod_ready_condition(pod)[0] == c["want_status"]


@pytest.mark.parametrize("c", HPR["message_cases"], ids=lambda c: c["name"])
def test_find_head_pod_ready_message(c):
    """utils/util_test.go:972-1038"""
    pod = {"phase": "Pending", "conditions": [{"type": "Ready", "status": "False", "reason": "ContainersNotReady", "message": c["message"]}],
           "containerStatuses": c["containerStatuses"]}
    _, reason, message = snapmod.head_pod_ready_condition(pod)
    assert (reason, message) == (c["want_reason"], c["want_message"])


def test_sha1_and_base32hex_vectors(oracle_mod):
    v = load("sha1_vectors")
    for c in v["sha1"]:
        assert oracle_mod.sha1(c["msg"].encode()).hex() == c["hex"]
    for msg, enc in v["base32hex"]:
        assert base64.b32hexencode(msg.encode()).decode() == enc  # pins the python encoder the next loop leans on
    rng = np.random.default_rng(7)
    for n in [0, 1, 54, 55, 56, 57, 63, 64, 65, 119, 120, 121, 127, 128, 1000, 6144, 70000]:
        m = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert oracle_mod.sha1(m) == hashlib.sha1(m).digest()
        assert oracle_mod.hash32(m) == base64.b32hexencode(hashlib.sha1(m).digest()).decode()  # 20 bytes -> 32 chars, no padding


def test_hash_relations_of_the_muted_spec():
    """TestGenerateHashWithoutReplicasAndWorkersToDelete rayservice_controller_unit_test.go:39-97 (relational: no literal digest
    exists anywhere in the reference)."""
    sc = load("reconcile_scenarios")["base"]["cluster"]["spec"]

    def h(spec):
        return base64.b32hexencode(hashlib.sha1(specjson.muted_spec_json(spec)).digest()).decode()

    import copy
    base = h(sc)
    s2 = copy.deepcopy(sc); s2["workerGroupSpecs"][0]["replicas"] += 1
    assert h(s2) == base
    s3 = copy.deepcopy(sc); s3["rayVersion"] = "2.100.0"
    assert h(s3) != base
    s4 = copy.deepcopy(sc); s4["headGroupSpec"]["template"]["spec"]["tolerations"] = [{"key": "k", "operator": "Exists"}]
    s4["workerGroupSpecs"][0]["template"]["spec"]["tolerations"] = [{"key": "k", "operator": "Exists"}]
    assert h(s4) == base
    s5 = copy.deepcopy(sc); s5["headGroupSpec"]["template"]["spec"]["schedulingGates"] = [{"name": "kueue.x-k8s.io/admission"}]
    assert h(s5) == base
    s6 = copy.deepcopy(sc); s6["workerGroupSpecs"][0]["workersToDelete"] = ["a", "b"]; s6["workerGroupSpecs"][0]["minReplicas"] = 7
    s6["upgradeStrategy"] = {"type": "Recreate"}
    assert h(s6) == base
    assert len(base) == 32 and set(base) <= set("0123456789ABCDEFGHIJKLMNOPQRSTUV")


def test_muted_spec_json_follows_go_encoding_rules():
    """omitempty, null for nil'ed non-omitempty pointers, sorted map keys, HTML escaping."""
    spec = {"rayVersion": "2.9<&>", "headGroupSpec": {"rayStartParams": {"b": "2", "a": "1"}, "template": {"spec": {"containers": [{"name": "h"}]}}},
            "workerGroupSpecs": [{"groupName": "g", "replicas": 3, "minReplicas": 0, "maxReplicas": 5, "rayStartParams": {}, "numOfHosts": 0,
                                  "template": {"spec": {"containers": [{"name": "w"}]}}}]}
    got = specjson.muted_spec_json(spec).decode()
    assert got == ('{"headGroupSpec":{"template":{"spec":{"containers":[{"name":"h"}]}},"rayStartParams":{"a":"1","b":"2"}},'
                   '"rayVersion":"2.9\\u003c\\u0026\\u003e","workerGroupSpecs":[{"groupName":"g","minReplicas":null,"maxReplicas":null,'
                   '"rayStartParams":{},"template":{"spec":{"containers":[{"name":"w"}]}},"scaleStrategy":{}}]}')


def test_find_suspend_status_order():
    """utils/util.go:153-162: first True among Suspending / Suspended in slice order"""
    conds = [{"type": "RayClusterSuspended", "status": "True"}, {"type": "RayClusterSuspending", "status": "True"}]
    assert snapmod.find_suspend_status(conds) == abi.SUSPEND_SUSPENDED
    assert snapmod.find_suspend_status(list(reversed(conds))) == abi.SUSPEND_SUSPENDING
    assert snapmod.find_suspend_status([{"type": "RayClusterSuspending", "status": "False"}]) == abi.SUSPEND_NONE


def test_atoi_semantics_of_replica_index_label():
    """strconv.Atoi (raycluster_controller.go:857-860): sign allowed, no spaces, invalid => label ignored"""
    def idx(v):
        pk, r = snapmod.pack_pod_word({"name": "p", "labels": {snapmod.REPLICA_INDEX_LABEL: v}})
        return (bool(pk & abi.PP_HAS_REPLICA_IDX), r)
    assert idx("7") == (True, 7) and idx("+7") == (True, 7) and idx("-3") == (True, -3) and idx("007") == (True, 7)
    assert idx(" 7")[0] is False and idx("7a")[0] is False and idx("")[0] is False and idx("1_0")[0] is False


@pytest.mark.parametrize("seed0", [0, 100, 200])
def test_fuzz_list_modes_agree(seed0, oracle_mod):
    """Adversarial snapshots (tests/fuzz_objects.py): the indexed checker and the namespace-scan List mode (the CPU
    baseline's cost structure) must produce byte-identical records, single- and multi-threaded."""
    import fuzz_objects
    for seed in range(seed0, seed0 + 100):
        snap, flags = fuzz_objects.snapshot(seed)
        a = oracle_mod.run(snap, flags, list_mode=oracle_mod.INDEXED, threads=1)
        b = oracle_mod.run(snap, flags, list_mode=oracle_mod.NS_SCAN, threads=3)
        d = a.diff(b)
        assert not d, (seed, d[:5])


def _i32(x: int) -> int:
    x &= 0xFFFFFFFF
    return x - (1 << 32) if x & 0x80000000 else x


def _desired_replicas_model(replicas, min_, max_, hosts, suspend):
    """GetWorkerGroupDesiredReplicas (utils/util.go:386-404) in Go's int32 arithmetic, written out independently."""
    mn = 0 if min_ is None else min_
    mx = 2 ** 31 - 1 if max_ is None else max_
    if suspend:
        return 0
    if replicas is None or replicas < mn:
        w = mn
    elif replicas > mx:
        w = mx
    else:
        w = replicas
    return _i32(w * hosts)


def test_desired_replicas_property(oracle_mod):
    """Property test over the whole int32 domain (nil / extreme / wrapping values included) against the independent model."""
    hyp = pytest.importorskip("hypothesis")
    st = hyp.strategies
    i32 = st.one_of(st.none(), st.integers(-2 ** 31, 2 ** 31 - 1), st.sampled_from([0, 1, -1, 2 ** 31 - 1, -2 ** 31, 2 ** 30, 65536]))
    hosts = st.one_of(st.integers(-2 ** 31, 2 ** 31 - 1), st.sampled_from([0, 1, 2, 4, -1, 65536, 2 ** 31 - 1]))

    @hyp.settings(max_examples=3000, deadline=None)
    @hyp.given(i32, i32, i32, hosts, st.booleans())
    def check(replicas, mn, mx, h, suspend):
        assert oracle_mod.desired_replicas(replicas, mn, mx, h, suspend) == _desired_replicas_model(replicas, mn, mx, h, suspend)

    check()


def test_sha1_random_lengths_against_hashlib(oracle_mod):
    rng = np.random.default_rng(7)
    for n in list(range(0, 200)) + [int(x) for x in rng.integers(200, 20000, 40)]:
        m = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert oracle_mod.sha1(m) == hashlib.sha1(m).digest()
        assert oracle_mod.hash32(m) == base64.b32hexencode(hashlib.sha1(m).digest()).decode()


def test_oracle_context_reuse_ranges_and_reps(oracle_mod):
    """The CPU arm's machinery (bench.py CpuArm): one shared index per snapshot, many runs.  A context must give the same records as
    the one-shot entry point — whole snapshot, any cluster range, either List mode, any thread count — and `reps` must not leak state
    from one repetition into the next (the per-pod action scratch is cleared per cluster)."""
    from kuberay_b200 import synthetic
    from oracle import oracle
    snap, flags = synthetic.generate(synthetic.config("C2", n_clusters=120, pods_per_cluster=14, groups=2, jobs=True, wtd_group_frac=0.4, seed=17))
    want = oracle_mod.run(snap, flags)
    cx = oracle.Context(snap)
    try:
        for mode in (oracle.INDEXED, oracle.NS_SCAN):
            for threads in (1, 3):
                got = cx.run(flags, list_mode=mode, threads=threads)
                assert not want.diff(got), (mode, threads)
        # ranges: the records of the clusters inside the range equal the full run's (records outside are whatever the last run left)
        for c0, c1, reps in ((0, 40, 1), (40, 120, 1), (17, 18, 5), (0, 120, 3)):
            got = cx.run_range(flags, c0, c1, list_mode=oracle.NS_SCAN, threads=2, reps=reps)
            for f in want.clusters.dtype.names:
                if f != "pod_start":  # (where the cluster's pods sit in the full per-cluster list: written by whole-snapshot runs only)
                    assert np.array_equal(got.clusters[f][c0:c1], want.clusters[f][c0:c1]), (f, c0, c1, reps)
            assert np.array_equal(got.hash[c0:c1], want.hash[c0:c1])
            g0, g1 = int(snap.c_group_off[c0]), int(snap.c_group_off[c1 - 1] + snap.c_group_cnt[c1 - 1])
            for f in ("expected", "n_list", "n_unhealthy", "n_running", "diff", "n_create", "flags"):
                assert np.array_equal(got.groups[f][g0:g1], want.groups[f][g0:g1]), (f, c0, c1)
        got = cx.run(flags)  # and a full run afterwards is still clean
        assert not want.diff(got)
    finally:
        cx.close()
