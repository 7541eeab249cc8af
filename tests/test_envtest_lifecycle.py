"""The reference's envtest suites for the RayCluster controller (raycluster_controller_test.go), restated as closed loops on
the host-side reconciler mirror: reconcile -> side effects on the fake client -> next reconcile sees them.  envtest's
`Eventually(...)` becomes "reconcile until the predicate holds" (a handful of passes); there is no kubelet, so — as in the
Go suites — the test flips Pods to Running itself.

Every scenario runs against the CPU oracle (`-m "not gpu"`) and against the CUDA engine through the C ABI (`-m gpu`).
"""
import pytest

from kuberay_b200.reconciler import Env, EngineBackend, FakeClient, RayClusterReconciler

NS = "default"
GROUP = "small-group"


class OracleBackend:
    def run(self, snap, flags):
        from oracle import oracle
        return oracle.run(snap, flags)


@pytest.fixture(params=["oracle", pytest.param("engine", marks=pytest.mark.gpu)])
def backend(request):
    return OracleBackend() if request.param == "oracle" else EngineBackend(0)


def cluster_template(name, autoscaling=None):
    """rayClusterTemplate (raycluster_controller_test.go:44-120): one worker group, replicas 3, minReplicas 0, maxReplicas 4."""
    spec = {"headGroupSpec": {"rayStartParams": {}, "template": {"spec": {"containers": [{"name": "ray-head"}]}}},
            "workerGroupSpecs": [{"groupName": GROUP, "replicas": 3, "minReplicas": 0, "maxReplicas": 4, "numOfHosts": 1, "workersToDelete": [],
                                  "template": {"spec": {"containers": [{"name": "ray-worker"}]}}}]}
    if autoscaling is not None:
        spec["enableInTreeAutoscaling"] = autoscaling
    return {"namespace": NS, "name": name, "uid": f"uid-{name}", "spec": spec, "status": {}}


def workers(client, name):
    return client.pods_of(NS, name, **{"ray.io/group": GROUP})


def heads(client, name):
    return client.pods_of(NS, name, **{"ray.io/node-type": "head"})


def everything(client, name):
    return client.pods_of(NS, name)


def set_running(pods):
    # envtest marks PodReady true when a test sets Status.Phase = Running (raycluster_controller_test.go:197-199)
    for p in pods:
        p["phase"] = "Running"
        p["conditions"] = [{"type": "Ready", "status": "True"}]
        p.setdefault("podIP", "10.0.0.7")


def eventually(r, name, pred, passes=8):
    for _ in range(passes):
        if pred():
            return True
        r.reconcile(NS, name)
    return pred()


def consistently(r, name, pred, passes=4):
    for _ in range(passes):
        r.reconcile(NS, name)
        if not pred():
            return False
    return True


def status(client, name):
    return client.clusters[(NS, name)].get("status") or {}


def cond_true(client, name, ctype):
    return any(c.get("type") == ctype and c.get("status") == "True" for c in status(client, name).get("conditions") or [])


def test_basic_lifecycle(backend):
    """raycluster_controller_test.go:132-249: 1 head + 3 workers appear; all Running -> state ready with its transition time;
    a deleted worker is replaced; replicas above maxReplicas is clamped to maxReplicas and stays there."""
    name = "raycluster-basic"
    client = FakeClient([cluster_template(name)], [])
    r = RayClusterReconciler(client, backend)
    assert eventually(r, name, lambda: len(workers(client, name)) == 3)                       # :154-159
    assert len(heads(client, name)) == 1                                                      # :161-169
    set_running(everything(client, name))                                                     # :171-195
    assert eventually(r, name, lambda: status(client, name).get("state") == "ready")          # :197-205
    assert status(client, name)["stateTransitionTimes"].get("ready")                          # :206-212
    assert status(client, name)["readyWorkerReplicas"] == 3 and status(client, name)["availableWorkerReplicas"] == 3

    victim = workers(client, name)[0]                                                         # :215-228
    assert client.delete_pod(NS, victim["name"])
    assert eventually(r, name, lambda: len(workers(client, name)) == 3)
    assert victim["name"] not in {p["name"] for p in workers(client, name)}

    client.clusters[(NS, name)]["spec"]["workerGroupSpecs"][0]["replicas"] = 5                # :230-249 (maxReplicas is 4)
    assert eventually(r, name, lambda: len(workers(client, name)) == 4)
    assert consistently(r, name, lambda: len(workers(client, name)) == 4)
    assert len(heads(client, name)) == 1


def test_autoscaler_scale_down_then_up(backend):
    """raycluster_controller_test.go:426-530: with in-tree autoscaling the autoscaler names its victim in workersToDelete and
    lowers replicas; exactly that Pod goes; after it clears the list a higher replicas count adds Pods back."""
    name = "raycluster-autoscaler"
    client = FakeClient([cluster_template(name, autoscaling=True)], [])
    r = RayClusterReconciler(client, backend)
    assert eventually(r, name, lambda: len(workers(client, name)) == 3)                       # :491-496
    set_running(everything(client, name))
    grp = client.clusters[(NS, name)]["spec"]["workerGroupSpecs"][0]
    victim = workers(client, name)[0]["name"]
    grp["replicas"], grp["workersToDelete"] = 2, [victim]                                     # :498-509
    assert eventually(r, name, lambda: len(workers(client, name)) == 2)                       # :511-514
    assert victim not in {p["name"] for p in workers(client, name)}
    assert consistently(r, name, lambda: len(workers(client, name)) == 2)                     # no random deletes under autoscaling
    grp["workersToDelete"] = []                                                               # cleanUpWorkersToDelete :516-518
    grp["replicas"] = 4                                                                       # :521-534
    assert eventually(r, name, lambda: len(workers(client, name)) == 4)
    assert len(heads(client, name)) == 1


@pytest.mark.parametrize("conditions_gate", [True, False])
def test_suspend_and_resume(backend, conditions_gate):
    """raycluster_controller_test.go:565-735 (testSuspendRayCluster, with and without the RayClusterStatusConditions gate)."""
    name = "raycluster-suspend"
    client = FakeClient([cluster_template(name)], [])
    r = RayClusterReconciler(client, backend, Env(status_conditions_gate=conditions_gate))
    spec = client.clusters[(NS, name)]["spec"]
    assert eventually(r, name, lambda: len(workers(client, name)) == 3)                       # :603-608

    spec["suspend"] = True                                                                    # :610-626
    assert eventually(r, name, lambda: len(everything(client, name)) == 0)
    assert eventually(r, name, lambda: status(client, name).get("state") == "suspended")      # :628-641
    if conditions_gate:
        assert eventually(r, name, lambda: cond_true(client, name, "RayClusterSuspended") and not cond_true(client, name, "RayClusterSuspending"))
        assert not cond_true(client, name, "RayClusterProvisioned")
        assert (status(client, name).get("head") or {}).get("podName", "") == ""

    spec["suspend"] = False                                                                   # :643-680: resume, then suspend again
    assert eventually(r, name, lambda: len(heads(client, name)) == 1 and len(workers(client, name)) == 3)
    set_running(workers(client, name))                                                        # head stays Pending
    spec["suspend"] = True
    assert eventually(r, name, lambda: len(everything(client, name)) == 0)
    assert eventually(r, name, lambda: status(client, name).get("state") == "suspended")
    if conditions_gate:  # (the Go suite waits for the condition with its own Eventually: the deprecated state field never left "suspended")
        assert eventually(r, name, lambda: cond_true(client, name, "RayClusterSuspended") and not cond_true(client, name, "RayClusterSuspending"))

    spec["suspend"] = False                                                                   # :682-704
    assert eventually(r, name, lambda: len(heads(client, name)) == 1 and len(workers(client, name)) == 3)
    set_running(everything(client, name))
    assert eventually(r, name, lambda: status(client, name).get("state") == "ready")          # :706-718
    if conditions_gate:
        assert eventually(r, name, lambda: not cond_true(client, name, "RayClusterSuspended") and not cond_true(client, name, "RayClusterSuspending"))
        assert cond_true(client, name, "RayClusterProvisioned")
        assert (status(client, name).get("head") or {}).get("podName", "") != ""


def test_worker_group_suspend(backend):
    """raycluster_controller_test.go:838-878 (worker-group suspend): the group's Pods go, the head stays; un-suspending brings
    the three workers back."""
    name = "raycluster-group-suspend"
    client = FakeClient([cluster_template(name)], [])
    r = RayClusterReconciler(client, backend)
    grp = client.clusters[(NS, name)]["spec"]["workerGroupSpecs"][0]
    assert eventually(r, name, lambda: len(workers(client, name)) == 3)
    set_running(everything(client, name))
    grp["suspend"] = True
    assert eventually(r, name, lambda: len(workers(client, name)) == 0)
    assert consistently(r, name, lambda: len(workers(client, name)) == 0 and len(heads(client, name)) == 1)
    grp["suspend"] = False
    assert eventually(r, name, lambda: len(workers(client, name)) == 3)
    assert len(heads(client, name)) == 1


class FinalizerClient(FakeClient):
    """Pods that carry a finalizer are not removed by Delete: they get a deletionTimestamp and stay listed until the finalizer is
    removed — what the apiserver does in the "atomically with Condition" suite (raycluster_controller_test.go:760-768)."""

    def delete_pod(self, ns, name):
        pod = self.pods.get((ns, name))
        if pod is not None and pod.get("finalizers"):
            pod["deletionTimestamp"] = "2026-01-01T00:00:00Z"
            return True
        return super().delete_pod(ns, name)

    def remove_finalizers(self):
        for key in list(self.pods):
            self.pods[key].pop("finalizers", None)
            if self.pods[key].get("deletionTimestamp"):
                del self.pods[key]


def suspend_status(client, name):
    suspending, suspended = cond_true(client, name, "RayClusterSuspending"), cond_true(client, name, "RayClusterSuspended")
    assert not (suspending and suspended)
    return "RayClusterSuspending" if suspending else ("RayClusterSuspended" if suspended else "")


def test_suspend_is_atomic_with_the_condition(backend):
    """raycluster_controller_test.go:741-836: once RayClusterSuspending is on, flipping spec.suspend back to false does not stop
    the suspension; only after the blocked Pods are really gone do both conditions clear and fresh Pods get created."""
    name = "raycluster-suspend-atomically"
    client = FinalizerClient([cluster_template(name)], [])
    r = RayClusterReconciler(client, backend)
    spec = client.clusters[(NS, name)]["spec"]
    assert eventually(r, name, lambda: len(everything(client, name)) == 4)                    # :757-768
    for p in everything(client, name):
        p["finalizers"] = ["ray.io/deletion-blocker"]
    old_names = {p["name"] for p in everything(client, name)}

    spec["suspend"] = True                                                                    # :770-777
    assert eventually(r, name, lambda: suspend_status(client, name) == "RayClusterSuspending")
    spec["suspend"] = False                                                                   # :779-785
    assert consistently(r, name, lambda: suspend_status(client, name) == "RayClusterSuspending")
    assert {p["name"] for p in everything(client, name)} == old_names                         # still there, all terminating
    assert all(p.get("deletionTimestamp") for p in everything(client, name))

    client.remove_finalizers()                                                                # :787-817
    assert eventually(r, name, lambda: suspend_status(client, name) == "")
    assert consistently(r, name, lambda: suspend_status(client, name) == "")
    assert eventually(r, name, lambda: len(everything(client, name)) == 4)
    assert not ({p["name"] for p in everything(client, name)} & old_names)

    spec["suspend"] = True                                                                    # :819-830
    assert eventually(r, name, lambda: len(everything(client, name)) == 0)
    assert eventually(r, name, lambda: suspend_status(client, name) == "RayClusterSuspended")
    assert consistently(r, name, lambda: suspend_status(client, name) == "RayClusterSuspended")


def test_group_suspend_with_autoscaler_is_stopped_by_validation(backend):
    """raycluster_controller_test.go:880-920: worker-group suspend together with in-tree autoscaling fails ValidateRayClusterSpec
    (utils/validation.go stays in the Go prelude, raycluster_controller.go:159-185), so reconcilePods never runs for the object.
    The shim hands such an object to the engine with KR_CF_SKIP: no decision is produced and no Pod is touched."""
    name = "raycluster-suspend-workergroup-autoscaler"
    client = FakeClient([cluster_template(name, autoscaling=True)], [])
    r = RayClusterReconciler(client, backend)
    assert eventually(r, name, lambda: len(everything(client, name)) == 4)
    client.clusters[(NS, name)]["spec"]["workerGroupSpecs"][0]["suspend"] = True
    client.clusters[(NS, name)]["skip"] = True  # what the prelude's validation error amounts to for the batch
    assert consistently(r, name, lambda: len(workers(client, name)) == 3 and len(heads(client, name)) == 1)


def test_created_pods_carry_the_built_manifest(backend):
    """What reaches client.Create is the whole Pod of buildHeadPod / buildWorkerPod (raycluster_controller.go:1387-1433), not just the metadata the
    engine's create tuples decide: generated `ray start` command, KubeRay's env, probes, /dev/shm, the worker's wait-gcs-ready init container."""
    name = "raycluster-built"
    tpl = cluster_template(name, autoscaling=True)
    tpl["spec"]["headGroupSpec"]["template"]["spec"]["containers"][0].update(image="rayproject/ray:2.9.0", resources={"limits": {"cpu": "2", "memory": "4Gi"}})
    tpl["spec"]["workerGroupSpecs"][0]["template"]["spec"]["containers"][0].update(image="rayproject/ray:2.9.0", resources={"limits": {"cpu": "4", "memory": "8Gi", "nvidia.com/gpu": "1"}})
    client = FakeClient([tpl], [])
    r = RayClusterReconciler(client, backend)
    assert eventually(r, name, lambda: len(workers(client, name)) == 3 and len(heads(client, name)) == 1)
    head = heads(client, name)[0]["spec"]
    assert [c["name"] for c in head["containers"]] == ["ray-head", "autoscaler"] and head["serviceAccountName"] == name
    assert head["containers"][0]["args"][0].startswith("ulimit -n 65536; ray start --head ") and "--no-monitor" in head["containers"][0]["args"][0]
    assert "--num-cpus=2" in head["containers"][0]["args"][0] and "--memory=4294967296" in head["containers"][0]["args"][0]
    assert {"name": "ray-logs", "emptyDir": {}} in head["volumes"] and "livenessProbe" in head["containers"][0]
    for w in workers(client, name):
        spec = w["spec"]
        ray = spec["containers"][0]
        assert f"--address={name}-head-svc.{NS}.svc.cluster.local:6379" in ray["args"][0] and "--num-gpus=1" in ray["args"][0]
        assert spec["initContainers"][0]["name"] == "wait-gcs-ready" and spec["volumes"] == [{"name": "shared-mem", "emptyDir": {"medium": "Memory", "sizeLimit": "8Gi"}}]
        assert any(e["name"] == "RAY_ADDRESS" and e["value"].endswith(":6379") for e in ray["env"]) and w["labels"]["ray.io/group"] == GROUP
