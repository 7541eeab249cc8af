"""Replay of the reference's own known-answer tests (tests/golden/*.json, transcribed from the Go test files cited inside)
through the host-side reconciler mirror.  Every scenario runs against the CPU oracle (pins the oracle; `-m "not gpu"`) and
against the CUDA engine through the C ABI (`-m gpu`)."""
import copy
import json
import os

import pytest

from kuberay_b200 import abi, snapshot as snapmod
from kuberay_b200.reconciler import Env, EngineBackend, FakeClient, RayClusterReconciler

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    with open(os.path.join(GOLD, name + ".json")) as f:
        return json.load(f)


class OracleBackend:
    def run(self, snap, flags):
        from oracle import oracle
        return oracle.run(snap, flags)


@pytest.fixture(params=["oracle", pytest.param("engine", marks=pytest.mark.gpu)])
def backend(request):
    return OracleBackend() if request.param == "oracle" else EngineBackend(0)


NS, CN = "default", "raycluster-sample"
SC = load("reconcile_scenarios")


def make_client(sc, cluster_patch=None):
    cluster = copy.deepcopy(SC["base"]["cluster"])
    pods = copy.deepcopy(SC["base"]["pods"])
    for k, v in (sc.get("patch_spec") or {}).items():
        cluster["spec"][k] = v
    for k, v in (sc.get("patch_group") or {}).items():
        cluster["spec"]["workerGroupSpecs"][0][k] = v
    if cluster_patch:
        cluster.update(cluster_patch)
    for p in pods:
        p.update((sc.get("patch_pods") or {}).get(p["name"], {}))
    client = FakeClient([cluster], pods)
    for name in sc.get("pre_delete") or []:
        assert client.delete_pod(NS, name)
    return client


def workers(client):
    return [p for p in client.pods_of(NS, CN, **{"ray.io/group": "small-group"})]


def heads(client):
    return client.pods_of(NS, CN, **{"ray.io/node-type": "head"})


@pytest.mark.parametrize("sc", SC["scenarios"], ids=[s["name"] for s in SC["scenarios"]])
def test_reconcile_pods_scenarios(sc, backend):
    """raycluster_controller_unit_test.go: the scenario tests on fixture F0 (citation in each fixture entry)."""
    client = make_client(sc)
    r = RayClusterReconciler(client, backend, Env(**sc.get("env", {})))
    before = {p["name"] for p in workers(client)}
    err = r.reconcile_pods(NS, CN)
    assert (err is not None) == sc["want_err"], err
    after = {p["name"] for p in workers(client)}
    assert len(after) == sc["want_workers"], sorted(after)
    for gone in sc.get("want_gone", []):
        assert gone not in after
    if "want_heads" in sc:
        assert len(heads(client)) == sc["want_heads"]
    if "want_random_deletes" in sc:
        named = set(sc["patch_group"]["workersToDelete"])
        assert len((before - after) - named) == sc["want_random_deletes"]
    if "want_creates" in sc:
        assert len(after - before) == sc["want_creates"] and len(before - after) == sc["want_deletes"]
    # pods that remain are all non-failed
    assert all(p.get("phase") not in ("Failed", "Succeeded") for p in workers(client))


def test_terminated_workers_multi_pass(backend):
    """Test_TerminatedWorkers_NoAutoscaler raycluster_controller_unit_test.go:2093-2221"""
    tw = SC["terminated_workers"]
    client = make_client(tw)
    r = RayClusterReconciler(client, backend)
    for step in tw["passes"]:
        for p in workers(client):  # envtest-style: newly created pods are patched to Running by the test (:2125-2129)
            if p.get("phase", "") == "":
                p["phase"] = "Running"
        if "set_phase" in step:
            workers(client)[0]["phase"] = step["set_phase"]["first_worker"]
        err = r.reconcile_pods(NS, CN)
        assert (err is not None) == step["want_err"], err
        if err:
            assert err.startswith("delete 1 unhealthy worker Pods")  # :811
        assert len(workers(client)) == step["want_workers"]


def test_terminated_head_restart_policy(backend):
    """Test_TerminatedHead_RestartPolicy :2223-2309 and Test_RunningPods_RayContainerTerminated :2311-2378"""
    th = SC["terminated_head"]
    cluster = copy.deepcopy(SC["base"]["cluster"])
    cluster["spec"]["workerGroupSpecs"] = []
    client = FakeClient([cluster], [copy.deepcopy(SC["base"]["pods"][0])])
    r = RayClusterReconciler(client, backend)
    for step in th["steps"]:
        if "head" in step:
            hp = heads(client)[0]
            hp.update(step["head"])
            hp.pop("containerStatuses", None) if "rayContainerTerminated" in step["head"] else None
        err = r.reconcile_pods(NS, CN)
        assert (err is not None) == step["want_err"], err
        assert len(client.pods) == step["want_pods"]
        for p in client.pods.values():
            if p.get("phase", "") == "":
                p["phase"] = "Running"


@pytest.mark.parametrize("case", load("recreate_upgrade")["cases"], ids=lambda c: c["name"])
def test_should_recreate_pods_for_upgrade(case, backend):
    """TestShouldRecreatePodsForUpgrade raycluster_controller_unit_test.go:3680-3814"""
    import base64
    import hashlib
    cluster = copy.deepcopy(SC["base"]["cluster"])
    cluster["spec"]["upgradeStrategy"] = case["upgradeStrategy"]
    cluster["spec"]["workerGroupSpecs"][0]["workersToDelete"] = []
    current = base64.b32hexencode(hashlib.sha1(snapmod.spec_hash_input(cluster["spec"])).digest()).decode()
    pods = []
    if case["head"] is not None:
        hp = copy.deepcopy(SC["base"]["pods"][0])
        hp["annotations"] = {snapmod.RECREATE_HASH_ANNOT: current if case["head"]["hash"] == "<current>" else case["head"]["hash"],
                             snapmod.KUBERAY_VERSION_ANNOT: snapmod.KUBERAY_VERSION if case["head"]["version"] == "<current>" else case["head"]["version"]}
        pods.append(hp)
    client = FakeClient([cluster], pods)
    r = RayClusterReconciler(client, backend)
    pr = r._pass()
    got = int(pr.res.clusters[0]["path"]) == abi.PATH_RECREATE_DELETE_ALL
    assert got == case["want"]
    assert bytes(pr.res.hash[0]).decode() == current  # the engine's hash is the annotation the controller would write
    if case["head"] and case["head"]["version"] not in ("<current>", ""):
        assert pr.res.clusters[0]["head_update_annotations"] == 1  # :1155-1162


# ------------------------------------------------------------------------------------------------ calculateStatus

def status_fixture(n_workers=3, ready="True", svc_ip="aaa.bbb.ccc.ddd"):
    cluster = copy.deepcopy(SC["base"]["cluster"])
    cluster["spec"]["workerGroupSpecs"][0]["workersToDelete"] = []
    cluster["headService"] = {"count": 1, "clusterIP": svc_ip, "name": "raycluster-sample-head-svc"}
    cond = [{"type": "Ready", "status": ready}]
    pods = [{"namespace": NS, "name": "headNode", "labels": {"ray.io/cluster": CN, "ray.io/node-type": "head"}, "phase": "Running", "podIP": "1.2.3.4", "conditions": copy.deepcopy(cond)}]
    for i in range(n_workers):
        pods.append({"namespace": NS, "name": f"workerNode-{i}", "labels": {"ray.io/cluster": CN, "ray.io/node-type": "worker"}, "phase": "Running", "podIP": "1.2.3.4", "conditions": copy.deepcopy(cond)})
    return cluster, pods


def cond_of(status, t):
    return next((c for c in status.get("conditions") or [] if c["type"] == t), None)


def test_calculate_status(backend):
    """TestCalculateStatus raycluster_controller_unit_test.go:1611-1723"""
    cluster, pods = status_fixture()
    client = FakeClient([cluster], pods)
    r = RayClusterReconciler(client, backend, Env(status_conditions_gate=False))
    new, err = r.calculate_status(NS, CN, None)
    assert err is None
    assert new["head"]["podIP"] == "1.2.3.4" and new["head"]["serviceIP"] == "aaa.bbb.ccc.ddd" and new["head"]["serviceName"] == "raycluster-sample-head-svc"
    assert new["state"] == "ready" and new["_state_changed"]          # StateTransitionTimes[ready] = LastUpdateTime (:1711-1716)
    new, err = r.calculate_status(NS, CN, ("FailedCreateHeadPod", "invalid"))
    assert err is None and not new.get("conditions")                   # gate off => no conditions
    r.env = Env(status_conditions_gate=True)
    new, _ = r.calculate_status(NS, CN, None)
    assert cond_of(new, "HeadPodReady")["status"] == "True"
    client.pods[(NS, "headNode")]["conditions"] = [{"type": "Ready", "status": "False"}]
    for k in [k for k in client.pods if k[1].startswith("workerNode")]:
        client.delete_pod(*k)
    new, _ = r.calculate_status(NS, CN, None)
    assert cond_of(new, "HeadPodReady")["status"] == "False"
    client.pods[(NS, "headNode")]["phase"] = "Failed"
    new, _ = r.calculate_status(NS, CN, None)
    assert cond_of(new, "HeadPodReady")["status"] == "False"
    new, err = r.calculate_status(NS, CN, ("FailedCreateHeadPod", "invalid"))
    assert err is None
    rf = cond_of(new, "ReplicaFailure")
    assert rf["status"] == "True" and rf["reason"] == "FailedCreateHeadPod" and rf["message"] == "invalid"


def test_calculate_status_without_desired_replicas(backend):
    """TestCalculateStatusWithoutDesiredReplicas :1727-1780 — only the head exists, desired 3 => State stays empty"""
    cluster, pods = status_fixture(n_workers=0)
    client = FakeClient([cluster], pods)
    new, err = RayClusterReconciler(client, backend).calculate_status(NS, CN, None)
    assert err is None and new.get("state", "") == "" and new.get("reason", "") == "" and not new["_state_changed"]


def test_calculate_status_with_suspended_worker_groups(backend):
    """TestCalculateStatusWithSuspendedWorkerGroups :1784-1847"""
    cluster, pods = status_fixture(n_workers=0)
    cluster["spec"]["workerGroupSpecs"][0].update({"suspend": True, "minReplicas": 100, "maxReplicas": 100, "replicas": 100})
    client = FakeClient([cluster], pods)
    new, err = RayClusterReconciler(client, backend).calculate_status(NS, CN, None)
    assert err is None
    assert (new["desiredWorkerReplicas"], new["minWorkerReplicas"], new["maxWorkerReplicas"], new["state"]) == (0, 0, 0, "ready")


def test_calculate_status_reconcile_error_back_and_forth(backend):
    """TestCalculateStatusWithReconcileErrorBackAndForth :1851-1944: err -> nil -> err; State "" -> ready -> stays ready"""
    cluster, pods = status_fixture()
    client = FakeClient([cluster], pods)
    r = RayClusterReconciler(client, backend)
    new, _ = r.calculate_status(NS, CN, "invalid")
    assert new.get("state", "") == ""
    r.update_status(NS, CN, new)
    new, _ = r.calculate_status(NS, CN, None)
    assert new["state"] == "ready" and new["_state_changed"]
    r.update_status(NS, CN, new, now="t1")
    new, _ = r.calculate_status(NS, CN, "invalid2")
    assert new["state"] == "ready" and not new["_state_changed"]       # transition time unchanged on the 3rd call
    r.update_status(NS, CN, new, now="t2")
    assert client.clusters[(NS, CN)]["status"]["stateTransitionTimes"]["ready"] == "t1"


def test_rayclusterprovisioned_condition(backend):
    """TestRayClusterProvisionedCondition :1946-2042"""
    steps = load("status_scenarios")["provisioned"]["steps"]
    cluster, pods = status_fixture(n_workers=1, ready="False")
    cluster["spec"]["workerGroupSpecs"][0]["replicas"] = 1
    client = FakeClient([cluster], pods)
    r = RayClusterReconciler(client, backend)
    for st in steps:
        client.pods[(NS, "headNode")]["conditions"] = [{"type": "Ready", "status": st["head_ready"]}]
        client.pods[(NS, "workerNode-0")]["conditions"] = [{"type": "Ready", "status": st["worker_ready"]}]
        new, err = r.calculate_status(NS, CN, None)
        assert err is None
        c = cond_of(new, "RayClusterProvisioned")
        assert [c["status"], c["reason"]] == st["want"]
        r.update_status(NS, CN, new)


def test_state_transition_times_no_state_change(backend):
    """TestStateTransitionTimes_NoStateChange :2044-2091"""
    cluster, pods = status_fixture()
    cluster["status"] = {"state": "ready", "stateTransitionTimes": {"ready": "t-earlier"}}
    client = FakeClient([cluster], pods)
    r = RayClusterReconciler(client, backend)
    new, _ = r.calculate_status(NS, CN, None)
    assert not new["_state_changed"]
    r.update_status(NS, CN, new, now="t-now")
    assert client.clusters[(NS, CN)]["status"]["stateTransitionTimes"]["ready"] == "t-earlier"


@pytest.mark.parametrize("case", load("inconsistent_status")["cases"], ids=lambda c: c["name"])
def test_inconsistent_ray_cluster_status(case, backend):
    """TestInconsistentRayClusterStatus utils/consistency_test.go:16-146 — replayed through the RayJob roll-up join
    (rayjob_controller.go:885): job.status.rayClusterStatus (old) vs the RayCluster's stored status (new)."""
    gold = load("inconsistent_status")
    old = gold["old"]
    new = copy.deepcopy(old)
    new.update(case.get("set", {}))
    new.setdefault("endpoints", {}).update(case.get("set_endpoint", {}))
    new.setdefault("head", {}).update(case.get("set_head", {}))
    cluster = copy.deepcopy(SC["base"]["cluster"])
    cluster["status"] = new
    job = {"namespace": NS, "name": "rayjob-sample", "status": {"rayClusterName": CN, "rayClusterStatus": old}}
    client = FakeClient([cluster], [], [job])
    pr = RayClusterReconciler(client, backend)._pass()
    jr = pr.res.jobs[0]
    assert jr["cluster_idx"] == 0
    assert bool(jr["status_changed"]) == case["want"]
    assert bool(jr["not_ready"]) == (new["state"] != "ready")


def test_full_reconcile_loop_converges(backend):
    """Reconcile() end to end on F0 until quiescent: decisions, status write, requeue policy (raycluster_controller.go:296-355)."""
    sc = {"patch_group": {"workersToDelete": []}, "patch_spec": {"enableInTreeAutoscaling": False}}
    client = make_client(sc)
    for p in client.pods.values():
        p["conditions"] = [{"type": "Ready", "status": "True"}]
        if p["name"] != "headNode":
            p["labels"]["ray.io/node-type"] = "worker"
    r = RayClusterReconciler(client, backend)
    requeue, err = r.reconcile(NS, CN)
    assert err is None and requeue == 2.0             # two prefix deletes + first status write => short requeue
    assert len(workers(client)) == 3
    requeue, err = r.reconcile(NS, CN)
    st = client.clusters[(NS, CN)]["status"]
    assert err is None and st["state"] == "ready" and st["readyWorkerReplicas"] == 3 and st["desiredWorkerReplicas"] == 3
    requeue, err = r.reconcile(NS, CN)
    assert (requeue, err) == (300.0, None)            # nothing left to do: periodic resync only


# ------------------------------------------------------------------------------------------------ multi-host worker groups
def mh_client(replicas=3, hosts=4, autoscaling=True):
    cluster = copy.deepcopy(SC["base"]["cluster"])
    cluster["spec"]["enableInTreeAutoscaling"] = autoscaling
    cluster["spec"]["workerGroupSpecs"][0].update({"replicas": replicas, "minReplicas": 0, "maxReplicas": 4, "numOfHosts": hosts, "workersToDelete": []})
    return FakeClient([cluster], [copy.deepcopy(SC["base"]["pods"][0])])


def run_all(client):
    for p in client.pods.values():
        if p.get("phase", "") == "":
            p["phase"] = "Running"


def replica_groups(client):
    groups = {}
    for p in workers(client):
        groups.setdefault(p["labels"][snapmod.REPLICA_NAME_LABEL], []).append(p)
    return groups


def test_multihost_group_lifecycle(backend):
    """raycluster_controller_test.go:925-1123 (envtest "multi-host" suite) restated on the fake client:
    replicas=3 x numOfHosts=4 -> 12 pods in 3 replica groups with indices 0..2 and host indices 0..3; autoscaler scale-down
    by workersToDelete removes whole replica groups; scale-up reuses the lowest free replica index."""
    client = mh_client()
    r = RayClusterReconciler(client, backend)
    assert r.reconcile_pods(NS, CN) is None
    run_all(client)
    groups = replica_groups(client)
    assert len(workers(client)) == 12 and len(groups) == 3
    assert sorted(g[0]["labels"][snapmod.REPLICA_INDEX_LABEL] for g in groups.values()) == ["0", "1", "2"]
    for g in groups.values():
        assert sorted(p["labels"]["ray.io/replica-host-index"] for p in g) == ["0", "1", "2", "3"]
        assert len({p["labels"][snapmod.REPLICA_INDEX_LABEL] for p in g}) == 1
    assert r.reconcile_pods(NS, CN) is None and len(workers(client)) == 12   # steady
    new, _ = r.calculate_status(NS, CN, None)
    assert (new["desiredWorkerReplicas"], new["minWorkerReplicas"], new["maxWorkerReplicas"]) == (12, 0, 16)

    # autoscaler scale-down: replicas 3 -> 2, one pod of the victim replica group named in workersToDelete (:1010-1038)
    victim = next(g for g in groups.values() if g[0]["labels"][snapmod.REPLICA_INDEX_LABEL] == "1")
    grp = client.clusters[(NS, CN)]["spec"]["workerGroupSpecs"][0]
    grp["replicas"], grp["workersToDelete"] = 2, [victim[2]["name"]]
    err = r.reconcile_pods(NS, CN)
    assert err == "deleted 4 worker Pods based on ScaleStrategy, requeueing"
    assert len(workers(client)) == 8
    grp["workersToDelete"] = []  # cleanUpWorkersToDelete: the autoscaler clears the list (raycluster_controller_test.go:1095-1097)
    assert r.reconcile_pods(NS, CN) is None and len(workers(client)) == 8

    # scale back up: the freed replica index 1 is reused (:1081-1094)
    grp["replicas"] = 3
    assert r.reconcile_pods(NS, CN) is None
    run_all(client)
    groups = replica_groups(client)
    assert len(workers(client)) == 12 and sorted(g[0]["labels"][snapmod.REPLICA_INDEX_LABEL] for g in groups.values()) == ["0", "1", "2"]


def test_multihost_incomplete_and_unhealthy_replicas(backend):
    """:975-1007 — an externally deleted pod makes its replica group incomplete: the remaining pods are deleted and the pass
    aborts; an unhealthy pod takes its whole replica group with it and the replacement is created in the same pass."""
    client = mh_client(replicas=2)
    r = RayClusterReconciler(client, backend)
    assert r.reconcile_pods(NS, CN) is None
    run_all(client)
    g0 = next(iter(replica_groups(client).values()))
    client.delete_pod(NS, g0[0]["name"])
    err = r.reconcile_pods(NS, CN)
    assert err and err.startswith("cleaned up incomplete replica group") and len(workers(client)) == 4
    assert r.reconcile_pods(NS, CN) is None
    run_all(client)
    assert len(workers(client)) == 8 and len(replica_groups(client)) == 2
    # unhealthy pod: whole replica group deleted (:986-1007) and replaced (:1079-1094) without an error
    sick = next(iter(replica_groups(client).values()))
    sick[1]["phase"] = "Failed"
    sick_names = {p["name"] for p in sick}
    assert r.reconcile_pods(NS, CN) is None
    assert not sick_names & {p["name"] for p in workers(client)} and len(workers(client)) == 8


def test_multihost_random_scale_down_only_without_autoscaler(backend):
    """:1095-1121"""
    for autoscaling, want in ((True, 12), (False, 8)):
        client = mh_client(replicas=3, autoscaling=autoscaling)
        r = RayClusterReconciler(client, backend)
        assert r.reconcile_pods(NS, CN) is None
        run_all(client)
        client.clusters[(NS, CN)]["spec"]["workerGroupSpecs"][0]["replicas"] = 2
        assert r.reconcile_pods(NS, CN) is None
        assert len(workers(client)) == want and all(len(g) == 4 for g in replica_groups(client).values())


# ------------------------------------------------------------------ head info of calculateStatus (:1344-1513, :1319-1342)
def _status_client(pods=None, svc=None):
    cluster = copy.deepcopy(SC["base"]["cluster"])
    if svc is not None:
        cluster["headService"] = svc
    return FakeClient([cluster], copy.deepcopy(SC["base"]["pods"]) if pods is None else pods)


def test_head_pod_ip_and_name_table(backend):
    """TestGetHeadPodIPAndNameFromGetRayClusterHeadPod (raycluster_controller_unit_test.go:1344-1415): one head -> its IP and
    name; none -> both empty, no error; two -> error; head without an IP yet -> name only."""
    base = copy.deepcopy(SC["base"]["pods"])
    head = next(p for p in base if (p.get("labels") or {}).get("ray.io/node-type") == "head")
    head["podIP"] = "1.2.3.4"
    new, err = RayClusterReconciler(_status_client(base), backend).calculate_status(NS, CN)
    assert err is None and (new["head"]["podIP"], new["head"]["podName"]) == ("1.2.3.4", head["name"])

    new, err = RayClusterReconciler(_status_client([]), backend).calculate_status(NS, CN)
    assert err is None and (new["head"]["podIP"], new["head"]["podName"]) == ("", "")

    extra = {"namespace": NS, "name": "unexpectedExtraHeadNode", "labels": {"ray.io/cluster": CN, "ray.io/node-type": "head", "ray.io/group": "headgroup"}}
    new, err = RayClusterReconciler(_status_client(base + [extra]), backend).calculate_status(NS, CN)
    assert new is None and err == "found multiple heads"

    no_ip = copy.deepcopy(base)
    next(p for p in no_ip if p["name"] == head["name"]).pop("podIP")
    new, err = RayClusterReconciler(_status_client(no_ip), backend).calculate_status(NS, CN)
    assert err is None and (new["head"]["podIP"], new["head"]["podName"]) == ("", head["name"])


def test_head_service_ip_and_name_table(backend):
    """TestGetHeadServiceIPAndName / ...OnHeadlessService (:1417-1513): one head Service -> its ClusterIP and name; none or two
    -> error; a headless Service (ClusterIP None) -> the head Pod's IP."""
    svc_name = f"{CN}-head-svc"
    new, err = RayClusterReconciler(_status_client(svc={"count": 1, "clusterIP": "1.2.3.4", "name": svc_name}), backend).calculate_status(NS, CN)
    assert err is None and (new["head"]["serviceIP"], new["head"]["serviceName"]) == ("1.2.3.4", svc_name)
    new, err = RayClusterReconciler(_status_client(svc={"count": 0}), backend).calculate_status(NS, CN)
    assert new is None and err == "unable to find head service"
    new, err = RayClusterReconciler(_status_client(svc={"count": 2, "clusterIP": "1.2.3.4", "name": svc_name}), backend).calculate_status(NS, CN)
    assert new is None and err == "found multiple head services"
    pods = copy.deepcopy(SC["base"]["pods"])
    head = next(p for p in pods if (p.get("labels") or {}).get("ray.io/node-type") == "head")
    head["podIP"] = "10.9.8.7"
    new, err = RayClusterReconciler(_status_client(pods, svc={"count": 1, "clusterIP": "None", "name": svc_name}), backend).calculate_status(NS, CN)
    assert err is None and (new["head"]["serviceIP"], new["head"]["serviceName"]) == ("10.9.8.7", svc_name)


def test_update_endpoints_from_the_head_service_ports(backend):
    """TestUpdateEndpoints (:1319-1342): status.endpoints maps every head Service port name to its (node or target) port."""
    ports = [{"name": "client", "targetPort": 10001}, {"name": "dashboard", "targetPort": 8265}, {"name": "metrics", "targetPort": 8080},
             {"name": "gcs-server", "targetPort": 6379}, {"name": "serve", "targetPort": 8000}]
    client = _status_client(svc={"count": 1, "clusterIP": "10.0.0.1", "name": f"{CN}-head-svc", "ports": ports})
    new, err = RayClusterReconciler(client, backend).calculate_status(NS, CN)
    assert err is None
    assert new["endpoints"] == {"client": "10001", "dashboard": "8265", "metrics": "8080", "gcs-server": "6379", "serve": "8000"}
    assert new["_needs_write"]  # the endpoints differ from the (empty) stored status


# ---------------------------------------------------------------------------------------------- API-call failures
# reconcilePods joins one of the five ErrFailed* markers onto the error of a failing Create/Delete
# (raycluster_controller.go:637,663,705,736,770,800,826,879,887,922); calculateStatus receives that error and only then sets
# ReplicaFailure (:1563-1571) and refuses State=ready (:1599).  The engine's record was computed for reconcileErr == nil, so
# the mirror (like the Go shim) discards its status half and re-evaluates status-only with the error kind.

def _ready_fixture():
    cluster = copy.deepcopy(SC["base"]["cluster"])
    cluster["spec"]["enableInTreeAutoscaling"] = False
    cluster["spec"]["workerGroupSpecs"][0].update({"replicas": 5, "workersToDelete": []})
    pods = copy.deepcopy(SC["base"]["pods"])
    for p in pods:
        p["conditions"] = [{"type": "Ready", "status": "True"}]
        if p["name"] != "headNode":
            p.setdefault("labels", {})["ray.io/node-type"] = "worker"
    return cluster, pods


def _cond(status, t):
    return next((c for c in status.get("conditions") or [] if c["type"] == t), None)


def test_failed_create_head_sets_replica_failure(backend):
    cluster, pods = _ready_fixture()
    client = FakeClient([cluster], [p for p in pods if p["name"] != "headNode"])
    client.fail_create.add("head")
    r = RayClusterReconciler(client, backend)
    requeue, err = r.reconcile(NS, CN)
    assert requeue == 2.0 and err.startswith("FailedCreateHeadPod")
    st = client.clusters[(NS, CN)]["status"]
    rf = _cond(st, "ReplicaFailure")
    assert rf and rf["status"] == "True" and rf["reason"] == "FailedCreateHeadPod" and "injected failure" in rf["message"]
    assert st.get("state", "") != "ready"
    assert len(heads(client)) == 0 and len(workers(client)) == 5  # reconcilePods returned at :736: workers untouched
    # the failure clears: next pass creates the head and removes the condition (:1572-1574)
    client.fail_create.clear()
    r.reconcile(NS, CN)
    assert len(heads(client)) == 1
    assert _cond(client.clusters[(NS, CN)]["status"], "ReplicaFailure") is None


def test_failed_delete_worker_blocks_ready_state(backend):
    """All pods Running+Ready and |pods| == desired+1, so a pass WITHOUT the failure would set State=ready; the random
    delete of one surplus... (here: a workersToDelete name) fails at the API server => ReplicaFailure, no ready."""
    cluster, pods = _ready_fixture()
    cluster["spec"]["workerGroupSpecs"][0].update({"replicas": 5, "workersToDelete": ["pod3"]})
    client = FakeClient([cluster], pods)
    client.fail_delete.add((NS, "pod3"))
    r = RayClusterReconciler(client, backend)
    _, err = r.reconcile(NS, CN)
    assert err.startswith("FailedDeleteWorkerPod")
    st = client.clusters[(NS, CN)]["status"]
    assert _cond(st, "ReplicaFailure")["reason"] == "FailedDeleteWorkerPod"
    assert st.get("state", "") != "ready"
    assert len(workers(client)) == 5


def test_workers_to_delete_name_outside_the_snapshot_is_still_deleted(backend):
    """:817-822 issues Delete(ns, name) for every name; a Pod the packed snapshot does not list (not yet in the informer
    cache / another shard) resolves to wtd_pod_idx = -1 in the engine but is deleted at the API server all the same."""
    cluster, pods = _ready_fixture()
    cluster["spec"]["enableInTreeAutoscaling"] = True
    cluster["spec"]["workerGroupSpecs"][0].update({"replicas": 5, "workersToDelete": ["late-pod"]})
    client = FakeClient([cluster], pods)
    r = RayClusterReconciler(client, backend)
    pr = r._pass()
    late = copy.deepcopy(pods[1]); late["name"] = "late-pod"
    client.create_pod(late)  # arrives after the snapshot was packed
    ci = r._cluster_index(pr, NS, CN)
    assert int(pr.res.wtd_pod_idx[int(pr.snap.g_wtd_off[int(pr.snap.c_group_off[ci])])]) == -1
    assert r._apply_decisions(pr, ci) is None
    assert (NS, "late-pod") not in client.pods
    assert ("Normal", "DeletedWorkerPod", f"Deleted pod {NS}/late-pod") in client.events
